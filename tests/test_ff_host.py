"""Host-emulated run of the device field arithmetic (csrc/ff.cuh) against big integers.

ff.cuh's carry-chain primitives have a 64-bit C emulation next to their PTX; this test
compiles the header for the CPU and checks the limb-level algorithms (lazy [0,2p) range,
staggered-accumulator CIOS multiplier).  The PTX itself is checked on the GPU
(tests/test_gpu_field.py).
"""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import pymodel as pm

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ffh():
    so = os.path.join(HERE, "host", "libffhost.so")
    src = os.path.join(HERE, "host", "ff_host_check.cpp")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, src])
    return C.CDLL(so)


def run(ffh, field, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    bb = np.ascontiguousarray(b, dtype=np.uint64) if b is not None else None
    ffh.ffh_op(C.c_int(field), C.c_int(op), a.ctypes.data_as(C.c_void_p),
               bb.ctypes.data_as(C.c_void_p) if bb is not None else None,
               out.ctypes.data_as(C.c_void_p), C.c_size_t(a.size // 4))
    return out


def raw(vals):
    return np.array([[(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for v in vals], dtype=np.uint64)


def ints(arr):
    return [sum(int(r[i]) << (64 * i) for i in range(4)) for r in arr.reshape(-1, 4)]


@pytest.mark.parametrize("field,mod", [(0, pm.R_MOD), (1, pm.P_MOD)])
def test_device_field_algorithms_on_host(ffh, field, mod):
    rnd = random.Random(42 + field)
    edge = [0, 1, 2, mod - 1, mod - 2, (1 << 253) + 12345, (1 << 32) - 1, 1 << 32, (1 << 224), mod >> 1,
            0xFFFFFFFF_FFFFFFFF_FFFFFFFF_FFFFFFFF % mod]
    xs = edge + [rnd.randrange(mod) for _ in range(3000)]
    ys = list(reversed(edge)) + [rnd.randrange(mod) for _ in range(3000)]
    a, b = raw(xs), raw(ys)
    rinv = pow(1 << 256, -1, mod)
    assert ints(run(ffh, field, 0, a, b)) == [(x + y) % mod for x, y in zip(xs, ys)]
    assert ints(run(ffh, field, 1, a, b)) == [(x - y) % mod for x, y in zip(xs, ys)]
    assert ints(run(ffh, field, 2, a, b)) == [x * y * rinv % mod for x, y in zip(xs, ys)]
    assert ints(run(ffh, field, 5, a)) == [x * x * rinv % mod for x in xs]
    # the dedicated squaring on lazy inputs ([p, 2p)) and chained, and against operator* on the same operands
    assert ints(run(ffh, field, 12, a)) == [x * x * rinv % mod for x in xs]
    assert ints(run(ffh, field, 13, a)) == [pow(x * x * rinv % mod, 2, mod) * rinv % mod for x in xs]
    assert np.array_equal(run(ffh, field, 5, a), run(ffh, field, 2, a, a))
    assert ints(run(ffh, field, 3, a)) == [(-x) % mod for x in xs]
    assert ints(run(ffh, field, 8, a)) == [2 * x % mod for x in xs]
    assert ints(run(ffh, field, 6, a)) == [x * (1 << 256) % mod for x in xs]
    assert ints(run(ffh, field, 7, a)) == [x * rinv % mod for x in xs]
    # operands in the upper half of the lazy range [p, 2p)
    assert ints(run(ffh, field, 9, a, b)) == [x * y * rinv % mod for x, y in zip(xs, ys)]
    assert ints(run(ffh, field, 10, a, b)) == [(x - y) % mod for x, y in zip(xs, ys)]
    assert ints(run(ffh, field, 11, a, b)) == [(x + y) % mod for x, y in zip(xs, ys)]
    small = xs[:40]
    got = ints(run(ffh, field, 4, raw(small)))
    R = 1 << 256
    # Montgomery inverse: inv(aR) = a^-1 R  ->  raw value v maps to v^-1 * R^2
    assert got == [(pow(v, -1, mod) * R * R % mod) if v else 0 for v in small]


def test_device_curve_formulas_on_host(ffh, orc):
    n = 64
    pts = orc.gen_points(4242, n)
    pts[5] = pts[4]                       # P + P  -> doubling branch of madd
    pts[9] = pts[8]; pts[9, 4:] = orc.field_op("fq", "neg", pts[8, 4:])   # P + (-P)
    pts[12] = 0                           # identity operand
    rnd = random.Random(5)
    signs = np.array([rnd.randrange(2) for _ in range(n)], dtype=np.uint8)
    signs[4] = signs[5]; signs[8] = signs[9]
    model_pts = pm.limbs_to_affine(pts)
    want = None
    for p, s in zip(model_pts, signs):
        if p is not None and s:
            p = (p[0], (-p[1]) % pm.P_MOD)
        want = pm.ec_add(want, p)
    for mode, expect in [(0, want), (1, want), (2, pm.ec_add(want, want)), (3, None)]:
        out = np.zeros(12, dtype=np.uint64)
        ffh.ech_sum(pts.ctypes.data_as(C.c_void_p), signs.ctypes.data_as(C.c_void_p), C.c_size_t(n),
                    out.ctypes.data_as(C.c_void_p), C.c_int(mode))
        assert pm.limbs_to_affine(orc.g1_to_affine(out))[0] == expect
        one = pm.to_mont_limbs([1], pm.P_MOD)[0]
        if expect is None:
            assert not out[:4].any() and np.array_equal(out[4:8], one) and not out[8:].any()
        else:
            assert np.array_equal(out[8:], one)
    # Jacobian (z != 1) -> XYZZ -> normalised
    j = orc.g1_add(orc.g1_mul(pts[0], orc.gen_scalars(1, 1)[0]), orc.g1_mul(pts[1], orc.gen_scalars(2, 1)[0]))
    out = np.zeros(12, dtype=np.uint64)
    ffh.ech_from_g1(j.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(orc.g1_to_affine(out), orc.g1_to_affine(j))


@pytest.mark.parametrize("c", [2, 3, 7, 8, 11, 13, 15, 16, 17, 20, 23])
def test_signed_digit_recoding(ffh, c):
    rnd = random.Random(c)
    W = 254 // c + 1
    vals = [0, 1, pm.R_MOD - 1, (1 << 254) - 1, (1 << 253), (1 << (c - 1)), (1 << c) - 1, (1 << 254) - (1 << (c - 1))]
    vals += [rnd.randrange(pm.R_MOD) for _ in range(300)]
    vals += [rnd.getrandbits(254) for _ in range(100)]
    for v in vals:
        canon = raw([v])
        d = np.zeros(W + 1, dtype=np.int32)
        ffh.ech_digits(canon.ctypes.data_as(C.c_void_p), C.c_int(c), C.c_int(W), d.ctypes.data_as(C.c_void_p))
        assert d[W] == 0, "carry out of the top window"
        assert all(-(1 << (c - 1)) < int(x) <= (1 << (c - 1)) for x in d[:W])
        assert sum(int(x) << (c * w) for w, x in enumerate(d[:W])) == v
        d2 = np.zeros(W + 1, dtype=np.int32)                              # the streaming (register-only) form
        ffh.ech_digits_stream(canon.ctypes.data_as(C.c_void_p), C.c_int(c), C.c_int(W), d2.ctypes.data_as(C.c_void_p))
        assert np.array_equal(d, d2)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_transcripts_host_build_match_python(ffh, kind):
    """csrc/transcript.hpp -- Blake2b-512 with personalisation / Challenge255 / point compression (kind 0), the
    Keccak-256 EVM transcript (kind 1) and the Poseidon transcript (kind 2: sponge, points absorbed mod r, bit-6 point
    compression) -- replayed on the host against oracle/prover.py's transcripts, across block / RATE boundaries."""
    import ctypes as C
    import random
    from oracle import prover as op
    from oracle import pymodel as pm
    rnd = random.Random(7 + kind)
    t = [op.Blake2bWrite, op.EvmWrite, op.PoseidonWrite][kind]()
    ops = bytearray()
    want_ch = []
    pts = [pm.ec_mul(rnd.randrange(1, pm.R_MOD), (1, 2)) for _ in range(6)]
    for step in range(60):
        what = rnd.choice([0, 1, 2, 2, 3])
        if what == 0:
            ops.append(0); want_ch.append(t.squeeze_challenge())
        elif what == 1:
            p = rnd.choice(pts)
            ops.append(1); ops += p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little"); t.write_point(p)
        elif what == 2:
            s = rnd.randrange(pm.R_MOD)
            ops.append(2); ops += s.to_bytes(32, "little"); t.write_scalar(s)
        else:
            s = rnd.randrange(pm.R_MOD)
            ops.append(3); ops += s.to_bytes(32, "little"); t.common_scalar(s)
    out = (C.c_ubyte * (32 * (len(want_ch) + 1)))()
    proof = (C.c_ubyte * (len(t.proof) + 64))()
    plen = C.c_size_t()
    ffh.transcript_replay.restype = C.c_size_t
    w = ffh.transcript_replay(bytes(ops), C.c_size_t(len(ops)), out, proof, C.byref(plen), C.c_int(kind))
    assert w == 32 * len(want_ch)
    got = [int.from_bytes(bytes(out[32 * i:32 * i + 32]), "little") for i in range(len(want_ch))]
    assert got == want_ch
    assert bytes(proof[:plen.value]) == t.finalize()


def test_keccak256_known_answers():
    from oracle.keccak import keccak256
    assert keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
