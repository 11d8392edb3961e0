"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle, bit-exact.
All marked `gpu`; sizes are kept where the oracle finishes in seconds, larger sizes are checked
through size-independent properties (round trips, linearity)."""
import random

import numpy as np
import pytest

import zkdcap_b200 as zk
from oracle import pymodel as pm
from importlib import import_module

shapes = import_module("zk-dcap-verifier_b200.shapes")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return zk.default_context()


def affine(orc, g1):
    return orc.g1_to_affine(np.asarray(g1).reshape(-1, 12))


# ---- field ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("field,mod", [("fr", pm.R_MOD), ("fq", pm.P_MOD)])
def test_field_ops(ctx, orc, field, mod):
    rnd = random.Random(1)
    edge = [0, 1, 2, mod - 1, mod - 2, (1 << 253) + 5, (1 << 32) - 1, 1 << 32, 1 << 224, mod >> 1]
    xs = edge + [rnd.randrange(mod) for _ in range(5000)]
    ys = list(reversed(edge)) + [rnd.randrange(mod) for _ in range(5000)]
    a, b = pm.to_mont_limbs(xs, mod), pm.to_mont_limbs(ys, mod)
    for op in ("add", "sub", "mul"):
        assert np.array_equal(ctx.field_op(field, op, a, b), orc.field_op(field, op, a, b)), op
    for op in ("neg", "sqr", "double"):
        assert np.array_equal(ctx.field_op(field, op, a), orc.field_op(field, op, a)), op
    assert np.array_equal(ctx.field_op(field, "from_mont", a), orc.field_op(field, "to_raw", a))
    assert np.array_equal(ctx.field_op(field, "to_mont", orc.field_op(field, "to_raw", a)), a)
    assert np.array_equal(ctx.field_op(field, "inv", a[:300]), orc.field_op(field, "inv", a[:300]))


# ---- best_fft ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 12, 13, 16, 17])
def test_best_fft_vs_oracle(ctx, orc, k):
    a = orc.gen_scalars(0x177 + k, 1 << k)
    omega = pm.to_mont_limbs([pm.omega_for(k)], pm.R_MOD)[0]
    want = orc.best_fft(a, omega, k)
    got = zk.best_fft(a.copy(), omega, k, ctx)
    assert np.array_equal(got, want)
    # inverse root
    oinv = pm.to_mont_limbs([pow(pm.omega_for(k), -1, pm.R_MOD)], pm.R_MOD)[0]
    assert np.array_equal(zk.best_fft(a.copy(), oinv, k, ctx), orc.best_fft(a, oinv, k))


def test_best_fft_edge_inputs(ctx, orc):
    k = 9
    omega = pm.to_mont_limbs([pm.omega_for(k)], pm.R_MOD)[0]
    zeros = np.zeros((1 << k, 4), dtype=np.uint64)
    assert not zk.best_fft(zeros.copy(), omega, k, ctx).any()
    mx = pm.to_mont_limbs([pm.R_MOD - 1] * (1 << k), pm.R_MOD)
    assert np.array_equal(zk.best_fft(mx.copy(), omega, k, ctx), orc.best_fft(mx, omega, k))
    with pytest.raises(AssertionError):
        zk.best_fft(zeros.copy(), omega, k + 1, ctx)


@pytest.mark.parametrize("k", [20, 22])
def test_best_fft_large_roundtrip(ctx, orc, k):
    n = 1 << k
    a = orc.gen_scalars(0x177 + k, n)
    w = pm.omega_for(k)
    omega = pm.to_mont_limbs([w], pm.R_MOD)[0]
    f = zk.best_fft(a.copy(), omega, k, ctx)
    # spot evaluations: f[i] = poly(omega^i)
    for i in (0, 1, n // 2 + 7, n - 1):
        wi = pm.to_mont_limbs([pow(w, i, pm.R_MOD)], pm.R_MOD)[0]
        assert np.array_equal(orc.eval_polynomial(a, wi), f[i])
    oinv = pm.to_mont_limbs([pow(w, -1, pm.R_MOD)], pm.R_MOD)[0]
    back = zk.best_fft(f.copy(), oinv, k, ctx)
    ninv = pm.to_mont_limbs([pow(n, -1, pm.R_MOD)], pm.R_MOD)
    assert np.array_equal(ctx.field_op("fr", "mul", back, np.tile(ninv, (n, 1))), a)


# ---- EvaluationDomain ----------------------------------------------------------------------------------
@pytest.mark.parametrize("j,k", [(3, 4), (5, 4), (5, 10), (4, 11), (9, 6), (2, 5), (5, 14), (3, 16)])
def test_domain_vs_oracle(ctx, orc, j, k):
    d = zk.EvaluationDomain(j, k, ctx)
    o = orc.Domain(j, k)
    assert d.extended_k == o.extended_k
    for name in ("omega", "omega_inv", "extended_omega", "extended_omega_inv", "g_coset", "g_coset_inv",
                 "ifft_divisor", "extended_ifft_divisor"):
        assert np.array_equal(getattr(d, name), getattr(o, name)), name
    a = orc.gen_scalars(j * 1000 + k, 1 << k)
    assert np.array_equal(d.lagrange_to_coeff(a), o.lagrange_to_coeff(a))
    assert np.array_equal(d.coeff_to_lagrange(a), o.coeff_to_lagrange(a))
    ext = d.coeff_to_extended(a)
    assert np.array_equal(ext, o.coeff_to_extended(a))
    e = orc.gen_scalars(j * 1000 + k + 1, d.ext_n)
    assert np.array_equal(d.extended_to_coeff(e), o.extended_to_coeff(e))
    assert np.array_equal(d.divide_by_vanishing_poly(e), o.divide_by_vanishing_poly(e))
    back = d.extended_to_coeff(ext)
    assert np.array_equal(back[:1 << k], a) and not back[1 << k:].any()


# ---- best_multiexp -------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 31, 32, 33, 100, 257, 1000, 4096, 1 << 14])
def test_best_multiexp_vs_oracle(ctx, orc, n):
    bases = orc.gen_points(0x5CA1A4 + n, max(n, 1))[:n]
    sc = orc.gen_scalars(n + 17, max(n, 1))[:n]
    got = zk.best_multiexp(sc, bases, ctx)
    want = orc.best_multiexp(sc, bases)
    assert np.array_equal(affine(orc, got), affine(orc, want))
    one = pm.to_mont_limbs([1], pm.P_MOD)[0]
    if n == 0:
        assert not got[:4].any() and np.array_equal(got[4:8], one) and not got[8:].any()
    else:
        assert np.array_equal(got[8:], one)          # normalised z = 1


def test_best_multiexp_edge_cases(ctx, orc):
    n = 600
    bases = orc.gen_points(5, n)
    bases[10:40] = bases[0]                                  # repeated base: P + P inside a bucket
    bases[40] = bases[0]; bases[40, 4:] = orc.field_op("fq", "neg", bases[0, 4:])    # -P
    bases[50] = 0                                            # identity
    rnd = random.Random(3)
    ks = [rnd.choice([0, 1, 2, 255, 1 << 88, pm.R_MOD - 1, pm.R_MOD - 2, rnd.randrange(pm.R_MOD)]) for _ in range(n)]
    for i in range(10, 41):
        ks[i] = 5
    sc = pm.to_mont_limbs(ks, pm.R_MOD)
    assert np.array_equal(affine(orc, zk.best_multiexp(sc, bases, ctx)), affine(orc, orc.best_multiexp(sc, bases)))
    # everything cancels -> identity
    b2 = np.concatenate([bases[:1], bases[:1]]); b2[1, 4:] = orc.field_op("fq", "neg", bases[0, 4:])
    s2 = pm.to_mont_limbs([12345, 12345], pm.R_MOD)
    out = zk.best_multiexp(s2, b2, ctx)
    assert not out[8:].any() and not affine(orc, out).any()
    # all-zero and all-one scalars, 88-bit and 8-bit bounded scalars
    for ks in ([0] * n, [1] * n, [rnd.getrandbits(88) for _ in range(n)], [rnd.getrandbits(8) for _ in range(n)]):
        sc = pm.to_mont_limbs(ks, pm.R_MOD)
        assert np.array_equal(affine(orc, zk.best_multiexp(sc, bases, ctx)), affine(orc, orc.best_multiexp(sc, bases)))
    with pytest.raises(AssertionError):
        zk.best_multiexp(sc[:5], bases[:6], ctx)


@pytest.mark.parametrize("k,pre", [(8, True), (8, False), (12, True), (14, True), (15, False), (16, True)])
def test_params_commit_vs_oracle(ctx, orc, k, pre):
    n = 1 << k
    g, gl = orc.gen_points(100 + k, n), orc.gen_points(200 + k, n)
    params = zk.ParamsKZG(k, g, gl, ctx, precompute=pre)
    p = orc.gen_scalars(300 + k, n)
    assert np.array_equal(affine(orc, params.commit(p)), affine(orc, orc.best_multiexp(p, g)))
    assert np.array_equal(affine(orc, params.commit_lagrange(p)), affine(orc, orc.best_multiexp(p, gl)))
    m = n // 2 + 3                                       # shorter polynomial: the first m bases
    assert np.array_equal(affine(orc, params.commit(p[:m])), affine(orc, orc.best_multiexp(p[:m], g[:m])))
    # skewed scalars as real advice columns have them (zeros, bits, bytes, 88-bit limbs)
    rnd = random.Random(k)
    ks = [rnd.choice([0, 0, 0, 1, rnd.getrandbits(8), rnd.getrandbits(88)]) for _ in range(n)]
    sc = pm.to_mont_limbs(ks, pm.R_MOD)
    assert np.array_equal(affine(orc, params.commit_lagrange(sc)), affine(orc, orc.best_multiexp(sc, gl)))


def test_msm_linearity_large(ctx, orc):
    k = 17
    n = 1 << k
    g = orc.gen_points(77, n)
    params = zk.ParamsKZG(k, g, g, ctx, precompute=True)
    a, b = orc.gen_scalars(1, n), orc.gen_scalars(2, n)
    s = ctx.field_op("fr", "add", a, b)
    lhs = params.commit(s)
    rhs = zk.g1_sum(np.stack([params.commit(a), params.commit(b)]), ctx)
    assert np.array_equal(lhs, rhs)
    assert np.array_equal(affine(orc, lhs), affine(orc, orc.best_multiexp(s, g)))


# ---- evaluate_h -------------------------------------------------------------------------------------------
def _evalh_case(orc, cs, k, seed):
    from tests.test_evalh_host import make_case, oracle_h, L
    case = make_case(cs, k, seed, orc)
    want, desc, keep = oracle_h(orc, cs, k, case)
    return case, want, L


@pytest.mark.parametrize("shape,k", [("toy_shape", 3), ("toy_shape", 8), ("qe3_shape", 4), ("qe3_shape", 9)])
def test_evaluate_h_vs_oracle(ctx, orc, shape, k):
    cs = getattr(shapes, shape)()
    case, want, L = _evalh_case(orc, cs, k, 7 + k)
    ch = pm.to_mont_limbs(case["challenges"], pm.R_MOD) if case["challenges"] else []
    for coeff_form in (False, True):
        if coeff_form:
            pk = zk.ProvingKey(cs, k, case["fixed_polys"], case["sigma_polys"], case["l0_poly"], case["l_last_poly"],
                               case["l_blind_poly"], coeff_form=True, ctx=ctx)
        else:
            pk = zk.ProvingKey(cs, k, case["fixed_cosets"], case["sigma_cosets"], case["l0"], case["l_last"],
                               case["l_active"], ctx=ctx)
        got = pk.evaluate_h(case["advice"], case["instance"], ch, L(case["y"]), L(case["beta"]), L(case["gamma"]),
                            L(case["theta"]), case["lookup_polys"], case["z_polys"])
        assert np.array_equal(got, want), "coeff_form=%s" % coeff_form
        st = pk.program_stats()
        assert st["muls_per_row"] > 0 and st["columns_read"] > 0


def test_quotient_pipeline_roundtrip(ctx, orc):
    """h numerator -> divide_by_vanishing_poly -> extended_to_coeff, all on the GPU, equals the oracle pipeline."""
    cs, k = shapes.toy_shape(), 6
    case, want, L = _evalh_case(orc, cs, k, 99)
    d = zk.EvaluationDomain(cs.degree(), k, ctx)
    o = case["dom"]
    got = d.extended_to_coeff(d.divide_by_vanishing_poly(want))
    assert np.array_equal(got, o.extended_to_coeff(o.divide_by_vanishing_poly(want)))


# ---- batched commitments and the batched-affine pairing round ------------------------------------------------
def _commit_batch(ctx, params_n, scalars_list, which):
    """zk_msm_batch_dev through the raw ABI with the library's own device-memory helpers"""
    import ctypes as C
    lib = ctx.lib
    cnt = len(scalars_list)
    n = params_n
    flat = np.ascontiguousarray(np.stack(scalars_list)).reshape(cnt * n, 4)
    d_s, d_o = C.c_void_p(), C.c_void_p()
    ctx.check(lib.zk_dev_alloc(ctx.h, flat.nbytes, C.byref(d_s)))
    ctx.check(lib.zk_dev_alloc(ctx.h, cnt * 96, C.byref(d_o)))
    ctx.check(lib.zk_h2d(ctx.h, d_s, flat.ctypes.data_as(C.c_void_p), flat.nbytes))
    ctx.check(lib.zk_msm_batch_dev(ctx.h, d_s, which, n, cnt, d_o))
    out = np.empty((cnt, 12), dtype=np.uint64)
    ctx.check(lib.zk_d2h(ctx.h, out.ctypes.data_as(C.c_void_p), d_o, out.nbytes))
    ctx.check(lib.zk_dev_free(ctx.h, d_s)); ctx.check(lib.zk_dev_free(ctx.h, d_o))
    return out


@pytest.mark.parametrize("k,cnt", [(10, 5), (13, 3)])
def test_commit_batch_vs_oracle(ctx, orc, k, cnt):
    n = 1 << k
    g, gl = orc.gen_points(900 + k, n), orc.gen_points(901 + k, n)
    gl[7] = gl[6]                                  # equal neighbours in a bucket are likely for equal scalars
    zk.ParamsKZG(k, g, gl, ctx, precompute=True)
    rnd = random.Random(k)
    cols = [orc.gen_scalars(50 + i, n) for i in range(cnt - 2)]
    cols.append(pm.to_mont_limbs([rnd.choice([0, 1, 1, 2, rnd.getrandbits(8)]) for _ in range(n)], pm.R_MOD))
    cols.append(pm.to_mont_limbs([1] * n, pm.R_MOD))
    got = _commit_batch(ctx, n, cols, 2)
    for i in range(cnt):
        assert np.array_equal(affine(orc, got[i]), affine(orc, orc.best_multiexp(cols[i], gl))), i


def test_batched_affine_round_forced():
    """Re-runs the MSM parity tests with the batched-affine pairing round forced on (ZK_MSM_BA=1): it is otherwise
    only taken for >= 2^24 bucket entries.  A fresh process is needed because the switch is read per call but the
    default context is shared."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, ZK_MSM_BA="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "tests/test_gpu_parity.py", "-k",
                        "(multiexp or commit or linearity) and not forced"], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


# ---- ParamsKZG file format, point compression (csrc/params.cu) ------------------------------------------------------
def test_point_compression_round_trip_and_bytes(ctx, orc):
    from oracle import prover as op
    from oracle import pymodel as pm
    pts = orc.gen_points(77, 300)
    pts[5] = 0                                                     # the identity
    comp = zk.g1_compress(pts, ctx)
    want = [op.compress_point(p) for p in pm.limbs_to_affine(pts)]
    assert [bytes(c) for c in comp] == want
    back = zk.g1_decompress(comp, ctx)
    assert np.array_equal(back, pts)
    # negated y flips the sign bit only
    neg = pts.copy()
    neg[:, 4:] = orc.field_op("fq", "neg", pts[:, 4:])
    neg[5] = 0
    comp2 = zk.g1_compress(neg, ctx)
    assert np.array_equal(comp2[:, :31], comp[:, :31]) and np.array_equal(np.delete(comp2[:, 31] ^ comp[:, 31], 5), np.full(299, 0x80, np.uint8))
    assert np.array_equal(zk.g1_decompress(comp2, ctx), neg)
    # invalid encodings: x >= p, and an x whose x^3 + 3 is not a square
    bad = comp.copy()
    bad[0] = np.frombuffer((pm.P_MOD + 1).to_bytes(32, "little"), dtype=np.uint8)
    with pytest.raises(zk.ZkError):
        zk.g1_decompress(bad, ctx)
    x = 4
    while pow((x ** 3 + 3) % pm.P_MOD, (pm.P_MOD - 1) // 2, pm.P_MOD) == 1:
        x += 1
    bad = comp.copy()
    bad[0] = np.frombuffer(x.to_bytes(32, "little"), dtype=np.uint8)
    with pytest.raises(zk.ZkError):
        zk.g1_decompress(bad, ctx)


def test_params_file_round_trip(ctx, orc):
    from oracle import prover as op
    from oracle import pymodel as pm
    k = 6
    n = 1 << k
    g, gl = orc.gen_points(91, n), orc.gen_points(92, n)
    g2 = bytes(range(64)); s_g2 = bytes(range(64, 128))
    data = (k).to_bytes(4, "little") + b"".join(op.compress_point(p) for p in pm.limbs_to_affine(g)) + \
        b"".join(op.compress_point(p) for p in pm.limbs_to_affine(gl)) + g2 + s_g2
    params = zk.ParamsKZG.read(data, ctx)
    assert params.k == k and bytes(params.g2) == g2 and bytes(params.s_g2) == s_g2
    poly = orc.gen_scalars(93, n)
    assert np.array_equal(orc.g1_to_affine(params.commit(poly)), orc.g1_to_affine(orc.best_multiexp(poly, g)))
    assert np.array_equal(orc.g1_to_affine(params.commit_lagrange(poly)), orc.g1_to_affine(orc.best_multiexp(poly, gl)))
    assert params.write() == data
    with pytest.raises(zk.ZkError):
        zk.ParamsKZG.read(data[:-1], ctx)


def test_point_encodings_v04_and_raw_against_golden_proof(ctx, orc):
    """halo2curves >= 0.4 compressed form on the commitments of the reference's golden proof (tests/golden), and
    SerdeFormat::RawBytes"""
    import json
    import os
    from oracle import prover as op
    from oracle import pymodel as pm
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_proof_points.json")) as f:
        g = json.load(f)
    idx = sorted(int(i) for i in g["points"])
    comp = np.stack([np.frombuffer(bytes.fromhex(g["chunks"][i]), dtype=np.uint8) for i in idx])
    pts = zk.g1_decompress(comp, ctx, fmt="processed_v04")
    want = [tuple(int(v, 16) for v in g["points"][str(i)]) for i in idx]
    assert pm.limbs_to_affine(pts) == want
    assert np.array_equal(zk.g1_compress(pts, ctx, fmt="processed_v04"), comp)
    with pytest.raises(zk.ZkError):                                  # flagged chunks are not 0.3.1 encodings
        zk.g1_decompress(comp[[idx.index(i) for i in g["flagged"]]], ctx)
    # a mixed batch with the identity, against the oracle codec
    mix = orc.gen_points(5, 200)
    mix[1::2, 4:] = orc.field_op("fq", "neg", mix[1::2, 4:])          # both parities of y
    mix[7] = 0
    c04 = zk.g1_compress(mix, ctx, fmt="processed_v04")
    assert [bytes(c) for c in c04] == [op.compress_point_v04(p) for p in pm.limbs_to_affine(mix)]
    assert np.array_equal(zk.g1_decompress(c04, ctx, fmt="processed_v04"), mix)
    bad = c04.copy()
    bad[7, 0] = 1                                                    # identity flag with a non-zero body
    with pytest.raises(zk.ZkError):
        zk.g1_decompress(bad, ctx, fmt="processed_v04")
    raw = zk.g1_compress(mix, ctx, fmt="raw_bytes")
    assert raw.tobytes() == mix.tobytes()
    assert np.array_equal(zk.g1_decompress(raw, ctx, fmt="raw_bytes"), mix)
    off = mix.copy()
    off[3, 4] ^= np.uint64(1)                                        # y off the curve
    with pytest.raises(zk.ZkError):
        zk.g1_decompress(zk.g1_compress(off, ctx, fmt="raw_bytes"), ctx, fmt="raw_bytes")


@pytest.mark.parametrize("fmt", ["processed_v04", "raw_bytes"])
def test_params_file_formats(ctx, orc, fmt):
    from oracle import prover as op
    from oracle import pymodel as pm
    k = 5
    n = 1 << k
    g, gl = orc.gen_points(191, n), orc.gen_points(192, n)
    g[1::2, 4:] = orc.field_op("fq", "neg", g[1::2, 4:])              # both parities of y
    qb = 128 if fmt == "raw_bytes" else 64
    g2 = bytes(i & 0xFF for i in range(qb)); s_g2 = bytes((i * 7) & 0xFF for i in range(qb))
    if fmt == "raw_bytes":
        body = g.tobytes() + gl.tobytes()
    else:
        body = b"".join(op.compress_point_v04(p) for p in pm.limbs_to_affine(g)) + \
            b"".join(op.compress_point_v04(p) for p in pm.limbs_to_affine(gl))
    data = (k).to_bytes(4, "little") + body + g2 + s_g2
    params = zk.ParamsKZG.read(data, ctx, fmt=fmt)
    assert params.k == k and bytes(params.g2) == g2 and bytes(params.s_g2) == s_g2
    poly = orc.gen_scalars(193, n)
    assert np.array_equal(orc.g1_to_affine(params.commit(poly)), orc.g1_to_affine(orc.best_multiexp(poly, g)))
    assert np.array_equal(orc.g1_to_affine(params.commit_lagrange(poly)), orc.g1_to_affine(orc.best_multiexp(poly, gl)))
    assert params.write(fmt=fmt) == data
    if fmt == "processed_v04":                                       # bit 6 set: x >= p for the 0.3.1 reader
        with pytest.raises(zk.ZkError):
            zk.ParamsKZG.read(data, ctx, fmt="processed")
