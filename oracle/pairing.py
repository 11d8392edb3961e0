"""
pairing.py -- TEST INFRASTRUCTURE: BN254 optimal-ate pairing on Python integers.

Used only by oracle/verifier.py for the final check of halo2's KZG verifier
(poly/kzg/strategy.rs DualMSM::check: e(left, [s]G2) * e(-right, G2) == 1), the check the reference
performs at circuits/src/sgx_dcap_verifier.rs:830-843 and bin/src/main.rs:244-251.
Fq12 = Fq[w] / (w^12 - 18 w^6 + 82), Fq2 = Fq[i] / (i^2 + 1), xi = 9 + i, w^6 = xi.
Pinned by tests/test_prover_oracle.py: G2 generator on the twist and of order r, bilinearity, non-degeneracy.

Only tests/ (through oracle/verifier.py) may import this module.
"""
P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
ATE_LOOP_COUNT = 29793968203157093288          # 6u + 2, u = 4965661367192848881
LOG_ATE = 63

# G2 generator (EIP-197), (c0, c1) = c0 + c1 * i
G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))


# ---- Fq2 ------------------------------------------------------------------------------------------
def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_scalar(a, k): return (a[0] * k % P, a[1] * k % P)
def f2_conj(a): return (a[0], (-a[1]) % P)


def f2_inv(a):
    d = pow((a[0] * a[0] + a[1] * a[1]) % P, -1, P)
    return (a[0] * d % P, (-a[1]) * d % P)


def f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_mul(a, a)
        e >>= 1
    return r


XI = (9, 1)
B2 = f2_mul((3, 0), f2_inv(XI))                 # twist curve: y^2 = x^3 + 3 / xi


def g2_on_curve(q):
    if q is None:
        return True
    x, y = q
    return f2_sub(f2_mul(y, y), f2_add(f2_mul(f2_mul(x, x), x), B2)) == (0, 0)


def g2_slope(p, q):
    """slope of the line through p and q (tangent when equal); None when the line is vertical"""
    if p[0] != q[0]:
        return f2_mul(f2_sub(q[1], p[1]), f2_inv(f2_sub(q[0], p[0])))
    if p[1] == q[1] and p[1] != (0, 0):
        return f2_mul(f2_scalar(f2_mul(p[0], p[0]), 3), f2_inv(f2_scalar(p[1], 2)))
    return None


def g2_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    m = g2_slope(p, q)
    if m is None:
        return None
    x3 = f2_sub(f2_sub(f2_mul(m, m), p[0]), q[0])
    return (x3, f2_sub(f2_mul(m, f2_sub(p[0], x3)), p[1]))


def g2_mul(k, q):
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, q)
        q = g2_add(q, q)
        k >>= 1
    return acc


def g2_neg(q):
    return None if q is None else (q[0], ((-q[1][0]) % P, (-q[1][1]) % P))


# ---- Fq12 as 12 coefficients of w ---------------------------------------------------------------------
def f12_mul(a, b):
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                t[i + j] += ai * bj
    for k in range(22, 11, -1):                  # w^12 = 18 w^6 - 82
        c = t[k]
        if c:
            t[k - 6] += 18 * c
            t[k - 12] -= 82 * c
    return [x % P for x in t[:12]]


F12_ONE = [1] + [0] * 11


def f12_pow(a, e):
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_mul(a, a)
        e >>= 1
    return r


def embed(c):
    """Fq2 -> Fq12: i = w^6 - 9"""
    out = [0] * 12
    out[0], out[6] = (c[0] - 9 * c[1]) % P, c[1] % P
    return out


def shift(a, k):
    """a * w^k for a sparse-friendly product"""
    w = [0] * 12
    w[k] = 1
    return f12_mul(a, w)


def line(m, t, p):
    """value at the G1 point p of the line with (untwisted) slope m through the twist point t.
    Twisting maps (x, y) -> (x w^2, y w^3) and the slope m -> m w."""
    if m is None:                                           # vertical line: xp - X_t
        o = shift(embed(t[0]), 2)
        o = [(-c) % P for c in o]
        o[0] = (o[0] + p[0]) % P
        return o
    M = shift(embed(m), 1)
    X1, Y1 = shift(embed(t[0]), 2), shift(embed(t[1]), 3)
    dx = [(-c) % P for c in X1]; dx[0] = (dx[0] + p[0]) % P      # xp - X1
    dy = [(-c) % P for c in Y1]; dy[0] = (dy[0] + p[1]) % P      # yp - Y1
    mm = f12_mul(M, dx)
    return [(a - b) % P for a, b in zip(mm, dy)]


# Frobenius on the twist: (x, y) -> (conj(x) * xi^((p-1)/3), conj(y) * xi^((p-1)/2))
GAMMA2 = f2_pow(XI, (P - 1) // 3)
GAMMA3 = f2_pow(XI, (P - 1) // 2)


def g2_frobenius(q):
    return (f2_mul(f2_conj(q[0]), GAMMA2), f2_mul(f2_conj(q[1]), GAMMA3))


def miller_loop(q, p):
    """q: G2 point on the twist (Fq2 coordinates), p: G1 point (ints).  Not final-exponentiated."""
    if q is None or p is None:
        return F12_ONE
    f, r = F12_ONE, q
    for i in range(LOG_ATE, -1, -1):
        m = g2_slope(r, r)
        f = f12_mul(f12_mul(f, f), line(m, r, p))
        r = g2_add(r, r)
        if ATE_LOOP_COUNT & (1 << i):
            m = g2_slope(r, q)
            f = f12_mul(f, line(m, r, p))
            r = g2_add(r, q)
    q1 = g2_frobenius(q)
    nq2 = g2_neg(g2_frobenius(q1))
    f = f12_mul(f, line(g2_slope(r, q1), r, p))
    r = g2_add(r, q1)
    f = f12_mul(f, line(g2_slope(r, nq2), r, p))
    return f


def final_exponentiation(f):
    return f12_pow(f, (P ** 12 - 1) // R)


def pairing(q, p):
    return final_exponentiation(miller_loop(q, p))


def pairing_check(pairs):
    """prod e(p_i, q_i) == 1 for pairs (G1 point, G2 point)"""
    f = F12_ONE
    for p, q in pairs:
        f = f12_mul(f, miller_loop(q, p))
    return final_exponentiation(f) == F12_ONE
