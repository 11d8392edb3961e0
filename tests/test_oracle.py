"""Pins the CPU oracle (oracle/zk_oracle.c).

The reference tree has no golden vector for MSM / FFT / evaluate_h, so the
oracle is pinned by: published constants, the EIP-196 point 2*(1,2), an independent Python
big-integer model (oracle/pymodel.py) and algebraic invariants.
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import pymodel as pm

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def limbs(x):
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def as_int(row):
    return sum(int(row[i]) << (64 * i) for i in range(4))


def test_constants_match_published(orc):
    # SURVEY.md A.1 -- constants published in halo2curves bn256/{fr,fq}.rs
    c = orc.fr_constants()
    assert as_int(c["R"]) == 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb
    assert as_int(c["R2"]) == 0x0216d0b17f4e44a58c49833d53bb808553fe3ab1e35c59e31bb8e645ae216da7
    assert orc.lib().orc_fr_inv64() == 0xc2e1f593efffffff
    assert orc.lib().orc_fq_inv64() == 0x87d20782e4866389
    q = orc.fq_constants()
    assert as_int(q["R"]) == 0x0e0a77c19a07df2f666ea36f7879462c0a78eb28f5c70b3dd35d438dc58f0d9d
    assert as_int(q["R2"]) == 0x06d89f71cab8351f47ab1eff0a417ff6b5e71911d44501fbf32cfc5b538afa89
    can = lambda v: pm.from_mont_limbs(v, pm.R_MOD)[0]
    assert can(c["ROOT_OF_UNITY"]) == 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c
    assert can(c["ROOT_OF_UNITY_INV"]) == 0x048127174daabc261bbe587180f34361b22625f59115aba70ed3e50a414e6dba
    assert can(c["DELTA"]) == 0x09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2
    assert can(c["ZETA"]) == pm.ZETA and pow(pm.ZETA, 3, pm.R_MOD) == 1 and pm.ZETA != 1
    assert pow(can(c["ROOT_OF_UNITY"]), 1 << 28, pm.R_MOD) == 1
    assert pow(can(c["ROOT_OF_UNITY"]), 1 << 27, pm.R_MOD) != 1


@pytest.mark.parametrize("field,mod", [("fr", pm.R_MOD), ("fq", pm.P_MOD)])
def test_field_ops_vs_bigint(orc, field, mod):
    rnd = random.Random(1234)
    edge = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, 1 << 253, (1 << 64) - 1, 1 << 64, mod >> 1]
    xs = edge + [rnd.randrange(mod) for _ in range(200)]
    ys = list(reversed(edge)) + [rnd.randrange(mod) for _ in range(200)]
    a, b = pm.to_mont_limbs(xs, mod), pm.to_mont_limbs(ys, mod)
    got = lambda op, *args: pm.from_mont_limbs(orc.field_op(field, op, *args), mod)
    assert got("add", a, b) == [(x + y) % mod for x, y in zip(xs, ys)]
    assert got("sub", a, b) == [(x - y) % mod for x, y in zip(xs, ys)]
    assert got("mul", a, b) == [(x * y) % mod for x, y in zip(xs, ys)]
    assert got("sqr", a) == [(x * x) % mod for x in xs]
    assert got("neg", a) == [(-x) % mod for x in xs]
    assert got("double", a) == [(2 * x) % mod for x in xs]
    assert got("inv", a) == [pow(x, -1, mod) if x else 0 for x in xs]
    # to_raw gives the canonical little-endian integer (`to_repr`)
    raw = orc.field_op(field, "to_raw", a)
    assert [as_int(r) for r in raw] == xs
    assert np.array_equal(orc.field_op(field, "from_raw", raw), a)


def test_from_u512(orc):
    rnd = random.Random(7)
    wide = [rnd.getrandbits(512) for _ in range(50)] + [0, (1 << 512) - 1]
    arr = np.array([[(w >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(8)] for w in wide], dtype=np.uint64)
    assert pm.from_mont_limbs(orc.fr_from_u512(arr), pm.R_MOD) == [w % pm.R_MOD for w in wide]


def test_eip196_double_generator(orc):
    g = pm.affine_to_limbs([(1, 2)])[0]
    gj = np.concatenate([g, pm.to_mont_limbs([1], pm.P_MOD)[0]])
    d = orc.g1_to_affine(orc.g1_double(gj))
    (x, y), = pm.limbs_to_affine(d)
    assert x == 0x030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3
    assert y == 0x15ed738c0e0a7c92e7845f96b2ae9c0a68a6a449e3538fc7ff3ebf7a5a18a2c4
    assert orc.g1_on_curve(d)


def test_curve_ops_vs_model(orc):
    rnd = random.Random(99)
    G = (1, 2)
    pts = [pm.ec_mul(rnd.randrange(1, pm.R_MOD), G) for _ in range(6)]
    one = pm.to_mont_limbs([1], pm.P_MOD)[0]
    J = lambda p: np.concatenate([pm.affine_to_limbs([p])[0], one]) if p else np.concatenate(
        [np.zeros(4, np.uint64), one, np.zeros(4, np.uint64)])
    A = lambda j: pm.limbs_to_affine(orc.g1_to_affine(j))[0]
    for p in pts:
        for q in pts:
            assert A(orc.g1_add(J(p), J(q))) == pm.ec_add(p, q)
            assert A(orc.g1_add_mixed(J(p), pm.affine_to_limbs([q])[0])) == pm.ec_add(p, q)
        neg = (p[0], (-p[1]) % pm.P_MOD)
        assert A(orc.g1_add(J(p), J(neg))) is None
        assert A(orc.g1_add_mixed(J(p), pm.affine_to_limbs([neg])[0])) is None
        assert A(orc.g1_add(J(p), J(None))) == p and A(orc.g1_add(J(None), J(p))) == p
        assert A(orc.g1_add_mixed(J(None), pm.affine_to_limbs([p])[0])) == p
        assert A(orc.g1_add_mixed(J(p), pm.affine_to_limbs([None])[0])) == p
        k = rnd.randrange(pm.R_MOD)
        assert A(orc.g1_mul(pm.affine_to_limbs([p])[0], pm.to_mont_limbs([k], pm.R_MOD)[0])) == pm.ec_mul(k, p)
    # non-normalised inputs (z != 1)
    a = orc.g1_add(J(pts[0]), J(pts[1])); b = orc.g1_double(J(pts[2]))
    assert A(orc.g1_add(a, b)) == pm.ec_add(pm.ec_add(pts[0], pts[1]), pm.ec_add(pts[2], pts[2]))
    assert A(orc.g1_add(a, a)) == A(orc.g1_double(a))


def test_gen_points_on_curve(orc):
    pts = orc.gen_points(0x5CA1A4, 300)
    assert orc.g1_on_curve(pts)
    assert np.array_equal(pts[:100], orc.gen_points(0x5CA1A4, 100, threads=1))
    for (x, y) in pm.limbs_to_affine(pts[:20]):
        assert (y * y - x * x * x - 3) % pm.P_MOD == 0 and y % 2 == 0


@pytest.mark.parametrize("n", [0, 1, 3, 4, 31, 32, 33, 100, 257])
def test_best_multiexp_vs_python_model(orc, n):
    rnd = random.Random(n)
    bases = orc.gen_points(1000 + n, max(n, 1))[:n]
    ks = [rnd.randrange(pm.R_MOD) for _ in range(n)]
    if n >= 4:
        ks[0] = 0; ks[1] = 1; ks[2] = pm.R_MOD - 1
    sc = pm.to_mont_limbs(ks, pm.R_MOD).reshape(n, 4)
    want = pm.msm(ks, pm.limbs_to_affine(bases))
    for threads in (1, 3, 8):
        got = pm.limbs_to_affine(orc.g1_to_affine(orc.best_multiexp(sc, bases, threads=threads)))[0]
        assert got == want
    assert pm.limbs_to_affine(orc.g1_to_affine(orc.msm_naive(sc, bases)))[0] == want


def test_best_multiexp_edge_scalars_and_duplicate_bases(orc):
    n = 200
    bases = orc.gen_points(5, n)
    bases[10:20] = bases[0]                       # repeated base -> P + P inside a bucket
    bases[20] = bases[0]
    bases[20, 4:] = orc.field_op("fq", "neg", bases[0, 4:])   # -P
    bases[30] = 0                                 # identity base
    rnd = random.Random(3)
    ks = [rnd.choice([0, 1, 2, 255, 1 << 88, pm.R_MOD - 1, rnd.randrange(pm.R_MOD)]) for _ in range(n)]
    for i in range(10, 21):
        ks[i] = 5
    sc = pm.to_mont_limbs(ks, pm.R_MOD)
    want = pm.msm(ks, pm.limbs_to_affine(bases))
    got = pm.limbs_to_affine(orc.g1_to_affine(orc.best_multiexp(sc, bases, threads=4)))[0]
    assert got == want


def test_best_multiexp_linearity_large(orc):
    n = 1 << 12
    bases = orc.gen_points(77, n)
    a, b = orc.gen_scalars(1, n), orc.gen_scalars(2, n)
    s = orc.field_op("fr", "add", a, b)
    lhs = orc.g1_to_affine(orc.best_multiexp(s, bases))
    rhs = orc.g1_to_affine(orc.g1_add(orc.best_multiexp(a, bases), orc.best_multiexp(b, bases, threads=2)))
    assert np.array_equal(lhs, rhs)


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8])
def test_best_fft_vs_dft(orc, k):
    n = 1 << k
    rnd = random.Random(k)
    xs = [rnd.randrange(pm.R_MOD) for _ in range(n)]
    w = pm.omega_for(k)
    a = pm.to_mont_limbs(xs, pm.R_MOD)
    wl = pm.to_mont_limbs([w], pm.R_MOD)[0]
    want = pm.dft(xs, w)
    for threads in (1, 2, 8, 64):
        assert pm.from_mont_limbs(orc.best_fft(a, wl, k, threads=threads), pm.R_MOD) == want
    assert pm.from_mont_limbs(orc.dft_naive(a, wl, k), pm.R_MOD) == want


def test_best_fft_large_roundtrip_and_threads(orc):
    k = 14
    a = orc.gen_scalars(0x177 + k, 1 << k)
    d = orc.Domain(3, k)
    f1 = orc.best_fft(a, d.omega, k, threads=1)
    f8 = orc.best_fft(a, d.omega, k, threads=8)
    assert np.array_equal(f1, f8)
    back = orc.best_fft(f1, d.omega_inv, k, threads=4)
    back = orc.field_op("fr", "mul", back, np.tile(d.ifft_divisor, (1 << k, 1)))
    assert np.array_equal(back, a)
    # a random evaluation: f1[i] == poly(omega^i)
    i = 12345 % (1 << k)
    wi = pm.to_mont_limbs([pow(pm.omega_for(k), i, pm.R_MOD)], pm.R_MOD)[0]
    assert np.array_equal(orc.eval_polynomial(a, wi), f1[i])


@pytest.mark.parametrize("j,k", [(3, 4), (5, 4), (4, 5), (9, 3), (2, 5)])
def test_domain_vs_model(orc, j, k):
    d = orc.Domain(j, k)
    assert d.extended_k == pm.extended_k_for(j, k)
    can = lambda v: pm.from_mont_limbs(v, pm.R_MOD)[0]
    assert can(d.omega) == pm.omega_for(k)
    assert can(d.extended_omega) == pm.omega_for(d.extended_k)
    assert can(d.omega_inv) == pow(pm.omega_for(k), -1, pm.R_MOD)
    assert can(d.g_coset) == pm.ZETA and can(d.g_coset_inv) == pow(pm.ZETA, -1, pm.R_MOD)
    assert can(d.ifft_divisor) == pow(1 << k, -1, pm.R_MOD)
    rnd = random.Random(j * 100 + k)
    n = 1 << k
    coeffs = [rnd.randrange(pm.R_MOD) for _ in range(n)]
    cl = pm.to_mont_limbs(coeffs, pm.R_MOD)
    ext = d.coeff_to_extended(cl)
    assert pm.from_mont_limbs(ext, pm.R_MOD) == pm.coeff_to_extended(coeffs, j, k)
    # lagrange_to_coeff inverts evaluation on <omega>
    evals = [pm.poly_eval(coeffs, pow(pm.omega_for(k), i, pm.R_MOD)) for i in range(n)]
    assert pm.from_mont_limbs(d.lagrange_to_coeff(pm.to_mont_limbs(evals, pm.R_MOD)), pm.R_MOD) == coeffs
    assert pm.from_mont_limbs(d.coeff_to_lagrange(cl), pm.R_MOD) == evals
    # extended_to_coeff(coeff_to_extended(p)) = p padded to n*(j-1)
    back = pm.from_mont_limbs(d.extended_to_coeff(ext), pm.R_MOD)
    assert back[:n] == coeffs and all(v == 0 for v in back[n:]) and len(back) == n * (j - 1)
    # t_evaluations = 1 / (x^n - 1) on the coset, period 2^(ext_k - k)
    te = pm.from_mont_limbs(d.t_evaluations(), pm.R_MOD)
    wx = pm.omega_for(d.extended_k)
    for i, t in enumerate(te):
        x = pm.ZETA * pow(wx, i, pm.R_MOD) % pm.R_MOD
        assert t * (pow(x, n, pm.R_MOD) - 1) % pm.R_MOD == 1
    ones = pm.to_mont_limbs([1] * d.ext_n, pm.R_MOD)
    div = pm.from_mont_limbs(d.divide_by_vanishing_poly(ones), pm.R_MOD)
    assert div == [te[i % len(te)] for i in range(d.ext_n)]
