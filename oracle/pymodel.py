"""
pymodel.py -- TEST INFRASTRUCTURE: independent Python big-integer model.

A second, deliberately naive implementation of the mathematics behind the hot
path (no Montgomery form, no windows, no butterflies) used to pin the C oracle
(oracle/zk_oracle.c), because the reference tree holds no golden vectors for
MSM / FFT / evaluate_h and its Rust dependencies cannot be built here
("parity unpinned").  Everything works on canonical
integers; conversion helpers translate to/from the 4 x u64 Montgomery layout
that halo2curves keeps in memory.

Only tests/ may import this module.
"""
import numpy as np

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001  # Fr
P_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47  # Fq
MONT_R = 1 << 256
S = 28
GENERATOR = 7
ROOT_OF_UNITY = pow(GENERATOR, (R_MOD - 1) >> S, R_MOD)
DELTA = pow(GENERATOR, 1 << S, R_MOD)
ZETA = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23


# ---- layout conversion ---------------------------------------------------------------
def to_mont_limbs(vals, mod):
    """canonical ints -> (n, 4) uint64 array in Montgomery form"""
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        m = (v % mod) * MONT_R % mod
        for j in range(4):
            out[i, j] = (m >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def from_mont_limbs(arr, mod):
    """(n, 4) uint64 Montgomery -> list of canonical ints"""
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    rinv = pow(MONT_R, -1, mod)
    out = []
    for row in arr:
        m = sum(int(row[j]) << (64 * j) for j in range(4))
        out.append(m * rinv % mod)
    return out


# ---- curve y^2 = x^3 + 3 over Fq, affine with None = identity --------------------------
def ec_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % P_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P_MOD) % P_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P_MOD) % P_MOD
    x3 = (lam * lam - x1 - x2) % P_MOD
    return (x3, (lam * (x1 - x3) - y1) % P_MOD)


def ec_mul(k, p):
    acc = None
    k %= R_MOD
    while k:
        if k & 1:
            acc = ec_add(acc, p)
        p = ec_add(p, p)
        k >>= 1
    return acc


def msm(scalars, points):
    acc = None
    for k, p in zip(scalars, points):
        acc = ec_add(acc, ec_mul(k, p))
    return acc


def affine_to_limbs(points):
    """list of (x, y) / None -> (n, 8) uint64 (G1Affine memory layout, identity = (0,0))"""
    out = np.zeros((len(points), 8), dtype=np.uint64)
    for i, p in enumerate(points):
        if p is None:
            continue
        out[i, :4] = to_mont_limbs([p[0]], P_MOD)[0]
        out[i, 4:] = to_mont_limbs([p[1]], P_MOD)[0]
    return out


def limbs_to_affine(arr):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 8)
    out = []
    for row in arr:
        x = from_mont_limbs(row[:4], P_MOD)[0]
        y = from_mont_limbs(row[4:], P_MOD)[0]
        out.append(None if (x == 0 and y == 0) else (x, y))
    return out


# ---- polynomials / domain ------------------------------------------------------------
def dft(a, omega):
    n = len(a)
    return [sum(a[j] * pow(omega, i * j, R_MOD) for j in range(n)) % R_MOD for i in range(n)]


def omega_for(k):
    return pow(ROOT_OF_UNITY, 1 << (S - k), R_MOD)


def poly_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R_MOD
    return acc


def extended_k_for(j, k):
    ek = k
    while (1 << ek) < (1 << k) * (j - 1):
        ek += 1
    return ek


def coeff_to_extended(coeffs, j, k):
    """evaluations of the polynomial on the coset zeta * <omega_ext>"""
    ek = extended_k_for(j, k)
    w = omega_for(ek)
    return [poly_eval(coeffs, ZETA * pow(w, i, R_MOD) % R_MOD) for i in range(1 << ek)]
