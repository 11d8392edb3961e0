"""GPU parity at the PRODUCTION configuration (the sizes and window widths bench.py runs), bit-exact against the CPU
oracle's best_multiexp / create_proof:

  * resident-table MSM at k = 18, 19, 20 (window c = 16: 128 KB shared-memory digit sort, 128 x 256 row / column bucket
    reduction) on the three scalar classes of SURVEY 8d config 3, singly and as one batch of 4
    (BASELINE configs[2] "bit-exact vs halo2 best_multiexp"; reference call site circuits/src/sgx_dcap_verifier.rs:814-822);
  * the same c = 16 code paths forced on small inputs (k = 13..16) through ZK_MSM_C;
  * complete proofs, BYTES equal to the CPU prover's, on the QE3 shape at k = 13 and 15;
  * group mode (zk_group) with a group of one GPU: bytes equal to the ungrouped proof.
"""
import contextlib
import os
import random

import numpy as np
import pytest

import zkdcap_b200 as zk
from oracle import pymodel as pm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return zk.Context(0)


def affine(orc, g1):
    return orc.g1_to_affine(np.asarray(g1).reshape(-1, 12))


@contextlib.contextmanager
def env(**kv):
    old = {k: os.environ.get(k) for k in kv}
    os.environ.update({k: str(v) for k, v in kv.items()})
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def scalar_classes(orc, n, seed):
    """(i) uniform, (ii) 88-bit bounded, (iii) 90 % zeros / small (< 2^8) + 10 % full width"""
    rng = np.random.default_rng(seed)
    uniform = orc.gen_scalars(seed, n)
    raw = np.zeros((n, 4), dtype=np.uint64)
    raw[:, 0] = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    raw[:, 1] = rng.integers(0, 2**24, size=n, dtype=np.uint64)
    bits88 = orc.field_op("fr", "from_raw", raw)
    kind = rng.integers(0, 10, size=n)
    raw = np.zeros((n, 4), dtype=np.uint64)
    raw[:, 0] = np.where(kind < 5, 0, rng.integers(0, 256, size=n)).astype(np.uint64)
    small = orc.field_op("fr", "from_raw", raw)
    big = kind == 9
    small[big] = orc.gen_scalars(seed + 1, int(big.sum()))
    return {"uniform": uniform, "bits88": bits88, "small90": small, "uniform2": orc.gen_scalars(seed + 2, n)}


def commit_batch(ctx, n, cols, which=2):
    import ctypes as C
    lib = ctx.lib
    cnt = len(cols)
    flat = np.ascontiguousarray(np.stack(cols)).reshape(cnt * n, 4)
    d_s, d_o = C.c_void_p(), C.c_void_p()
    ctx.check(lib.zk_dev_alloc(ctx.h, flat.nbytes, C.byref(d_s)))
    ctx.check(lib.zk_dev_alloc(ctx.h, cnt * 96, C.byref(d_o)))
    ctx.check(lib.zk_h2d(ctx.h, d_s, flat.ctypes.data_as(C.c_void_p), flat.nbytes))
    ctx.check(lib.zk_msm_batch_dev(ctx.h, d_s, which, n, cnt, d_o))
    out = np.empty((cnt, 12), dtype=np.uint64)
    ctx.check(lib.zk_d2h(ctx.h, out.ctypes.data_as(C.c_void_p), d_o, out.nbytes))
    ctx.check(lib.zk_dev_free(ctx.h, d_s)); ctx.check(lib.zk_dev_free(ctx.h, d_o))
    return out


def window_bits(ctx, which=2):
    import ctypes as C
    c, w = C.c_uint32(), C.c_uint32()
    ctx.check(ctx.lib.zk_srs_window(ctx.h, which, C.byref(c), C.byref(w)))
    return c.value, w.value


@pytest.mark.parametrize("k", [18, 19, 20])
def test_resident_msm_production_window(ctx, orc, k):
    n = 1 << k
    g = orc.gen_points(0x5CA1A4 + k, n)
    params = zk.ParamsKZG(k, g, g, ctx, precompute=True)
    c, W = window_bits(ctx)
    assert c == 16 and W == 16, "bench sizes run the c = 16 window table"
    cls = scalar_classes(orc, n, 0x5CA1A4 + k)
    names = list(cls)
    want = {nm: affine(orc, orc.best_multiexp(cls[nm], g)) for nm in names}
    for nm in names[:3]:                                                   # singly (count = 1 plan, host-pointer call)
        assert np.array_equal(affine(orc, params.commit_lagrange(cls[nm])), want[nm]), nm
    got = commit_batch(ctx, n, [cls[nm] for nm in names])                  # one batch of 4 resident columns
    for i, nm in enumerate(names):
        assert np.array_equal(affine(orc, got[i]), want[nm]), "batch column %s" % nm
    m = n - 12345                                                          # a shorter polynomial on the same table
    assert np.array_equal(affine(orc, params.commit(cls["uniform"][:m])), affine(orc, orc.best_multiexp(cls["uniform"][:m], g[:m])))
    del params


@pytest.mark.parametrize("k,c", [(13, 16), (14, 16), (15, 16), (16, 16), (14, 20), (16, 20)])
def test_window16_forced_on_small_inputs(ctx, orc, k, c):
    """ZK_MSM_C=16 makes k = 13..16 take the code path of the bench sizes: 2^15 counters in shared memory, nR = 128,
    nC = 256, W = 16 table rows.  ZK_MSM_C=20 is the configuration of single commitments from 2^23 points on: 2^19 buckets,
    W = 13, digits through the global-atomic kernels, 512 x 1024 row / column reduction."""
    n = 1 << k
    g = orc.gen_points(0xC16 + k, n)
    g[5] = g[4]
    with env(ZK_MSM_C=c):
        params = zk.ParamsKZG(k, g, g, ctx, precompute=True)
    assert window_bits(ctx) == (c, 254 // c + 1)
    cls = scalar_classes(orc, n, 0xC16 + k)
    cls["ones"] = pm.to_mont_limbs([1] * n, pm.R_MOD)
    cls["top"] = pm.to_mont_limbs([pm.R_MOD - 1 - i for i in range(n)], pm.R_MOD)
    names = list(cls)
    want = {nm: affine(orc, orc.best_multiexp(cls[nm], g)) for nm in names}
    for nm in names:
        assert np.array_equal(affine(orc, params.commit_lagrange(cls[nm])), want[nm]), nm
    got = commit_batch(ctx, n, [cls[nm] for nm in names])
    for i, nm in enumerate(names):
        assert np.array_equal(affine(orc, got[i]), want[nm]), "batch column %s" % nm
    del params


@pytest.mark.parametrize("k", [13, 15])
def test_qe3_proof_bytes_at_larger_k(ctx, k):
    """k = 15 reaches every multi-pass code path of the bench size: two-pass NTTs (2^15, 2^17 extended), chunked digit
    sort, partial-slot levels, the 8-bit radix sort over > 2^13 rows."""
    from oracle import oracle as orc
    from oracle import prover as op
    from oracle import verifier as ov
    from tests import prover_common as pc
    case = pc.make_case("qe3", k, 5)
    p = pc.params_for(k)
    params = zk.ParamsKZG(k, p["g"], p["g_lagrange"], ctx, precompute=True)
    key = zk.ProverKey(params, case["circ"], case["fixed"], case["mapping"])
    opk = op.keygen(orc, p, case["circ"], case["fixed"], case["mapping"], case["circ"].transcript_repr)
    fc, pcm = key.vk_commitments()
    assert pm.limbs_to_affine(fc) == opk["vk"]["fixed_commitments"]
    assert pm.limbs_to_affine(pcm) == opk["vk"]["permutation_commitments"]
    want = op.create_proof(orc, p, opk, case["circ"], case["advice"], case["instances_int"], case["randoms"])
    got = key.create_proof(case["advice"], case["instances_limbs"], case["randoms"])
    assert got == want
    assert ov.verify_proof(pc.verifier_params(), opk["vk"], case["cs"], k, case["instances_int"], got)


@pytest.mark.parametrize("shape,k", [("toy", 6), ("qe3", 11)])
def test_group_of_one_gives_the_same_bytes(ctx, shape, k):
    """zk_group with world = 1: every sharding branch of csrc/prover.cu / group.cu runs (shard plans, NCCL all-gathers
    of one rank, row exchange) and the proof bytes must not change."""
    import torch  # noqa: F401  (loads the libnccl.so.2 the library binds at run time)
    from oracle import oracle as orc
    from oracle import prover as op
    from tests import prover_common as pc
    case = pc.make_case(shape, k, 7)
    p = pc.params_for(k)
    params = zk.ParamsKZG(k, p["g"], p["g_lagrange"], ctx, precompute=True)
    key = zk.ProverKey(params, case["circ"], case["fixed"], case["mapping"])
    plain = key.create_proof(case["advice"], case["instances_limbs"], case["randoms"])
    try:
        uid = zk.Group.unique_id()
    except zk.ZkError:
        pytest.skip("NCCL not loadable")
    grp = zk.Group(ctx, 0, 1, uid)
    key.set_group(grp)
    grouped = key.create_proof(case["advice"], case["instances_limbs"], case["randoms"])
    key.set_group(None)
    again = key.create_proof(case["advice"], case["instances_limbs"], case["randoms"])
    grp.close()
    assert grouped == plain and again == plain
    opk = op.keygen(orc, p, case["circ"], case["fixed"], case["mapping"], case["circ"].transcript_repr)
    assert plain == op.create_proof(orc, p, opk, case["circ"], case["advice"], case["instances_int"], case["randoms"])


@pytest.mark.parametrize("fmt", ["processed", "raw_bytes"])
def test_srs_file_to_verified_proof(ctx, fmt):
    """ParamsKZG::read of a kzg_bn254_k.srs image (G1 sets AND the G2 fields), keygen, create_proof, verify_proof with
    the file's own g2 / s_g2 (the file encodings must be decoded before they reach the verifier)."""
    from oracle import oracle as orc
    from oracle import prover as op
    from tests import prover_common as pc
    from tests.test_verifier_host import g2_bytes
    k = 6
    case = pc.make_case("toy", k, 9)
    p = pc.params_for(k)
    vp = pc.verifier_params()
    g2f, sg2f = zk.g2_encode(g2_bytes(vp["g2"]), fmt), zk.g2_encode(g2_bytes(vp["s_g2"]), fmt)
    if fmt == "raw_bytes":
        body = p["g"].tobytes() + p["g_lagrange"].tobytes()
    else:
        body = b"".join(op.compress_point(q) for q in pm.limbs_to_affine(p["g"])) + \
            b"".join(op.compress_point(q) for q in pm.limbs_to_affine(p["g_lagrange"]))
    data = (k).to_bytes(4, "little") + body + g2f + sg2f
    params = zk.ParamsKZG.read(data, ctx, fmt=fmt)
    assert params.g2_uncompressed == g2_bytes(vp["g2"]) and params.s_g2_uncompressed == g2_bytes(vp["s_g2"])
    key = zk.ProverKey(params, case["circ"], case["fixed"], case["mapping"])
    proof = key.create_proof(case["advice"], case["instances_limbs"], case["randoms"])
    fc, pcm = key.vk_commitments()
    g1 = pm.affine_to_limbs([(1, 2)])[0]
    assert zk.verify_proof(case["circ"], fc, pcm, g1, params.g2_uncompressed, params.s_g2_uncompressed, case["instances_limbs"], proof)
    with pytest.raises(zk.ZkError):
        zk.verify_proof(case["circ"], fc, pcm, g1, params.g2, params.s_g2, case["instances_limbs"], proof) if fmt != "raw_bytes" else \
            zk.verify_proof(case["circ"], fc, pcm, g1, params.g2[:64], params.s_g2[:64], case["instances_limbs"], proof)
    assert params.write(fmt=fmt) == data


def test_generated_evaluator_is_in_use_and_matches_the_interpreter(ctx, orc):
    """evaluate_h runs the circuit's own NVRTC-compiled kernel (csrc/evalh_codegen.hpp); the interpreter (ZK_EVALH=interp)
    and the CPU oracle give the same rows"""
    import ctypes as C
    from importlib import import_module
    from tests.test_evalh_host import make_case, oracle_h, L
    shapes = import_module("zk-dcap-verifier_b200.shapes")
    cs, k = shapes.qe3_shape(), 8
    case = make_case(cs, k, 33, orc)
    want, desc, keep = oracle_h(orc, cs, k, case)
    ch = pm.to_mont_limbs(case["challenges"], pm.R_MOD) if case["challenges"] else []
    got = {}
    for mode in ("generated", "interp"):
        with env(ZK_EVALH="interp" if mode == "interp" else "auto"):
            pk = zk.ProvingKey(cs, k, case["fixed_cosets"], case["sigma_cosets"], case["l0"], case["l_last"], case["l_active"], ctx=ctx)
        kind = ctx.lib.zk_pk_evaluator_kind(pk.h)
        secs = C.c_double()
        note = ctx.lib.zk_pk_evaluator_log(pk.h, C.byref(secs))
        assert kind == (1 if mode == "generated" else 0), note
        got[mode] = pk.evaluate_h(case["advice"], case["instance"], ch, L(case["y"]), L(case["beta"]), L(case["gamma"]),
                                  L(case["theta"]), case["lookup_polys"], case["z_polys"])
        assert np.array_equal(got[mode], want), mode


def test_work_counters(ctx, orc):
    """zk_ctx_counters: the device-side count of sorted bucket entries (= group additions of the accumulation) that bench.py's
    integer-pipe roofline is computed from"""
    k = 12
    n = 1 << k
    g = orc.gen_points(0xC7 + k, n)
    params = zk.ParamsKZG(k, g, g, ctx, precompute=True)
    c, W = window_bits(ctx)
    ctx.counters(reset=True)
    ones = pm.to_mont_limbs([1] * n, pm.R_MOD)
    params.commit_lagrange(ones)                                   # one non-zero window per scalar
    got = ctx.counters()
    assert got == {"msm_bucket_entries": n, "msms": 1}
    params.commit_lagrange(orc.gen_scalars(5, n))                  # uniform scalars: nearly every window is non-zero
    got = ctx.counters(reset=True)
    assert got["msms"] == 2 and 0.95 * n * W < got["msm_bucket_entries"] - n <= n * W
    assert ctx.counters() == {"msm_bucket_entries": 0, "msms": 0}
