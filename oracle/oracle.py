"""
oracle.py -- TEST INFRASTRUCTURE: ctypes binding of oracle/libzkoracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import
this.  Arrays are numpy uint64: Fr/Fq = (..., 4), G1Affine = (..., 8),
G1 (Jacobian) = (..., 12), all in halo2curves' in-memory (Montgomery) layout.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libzkoracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("zk_oracle.c", "field_tmpl.h", "prover_tail.inc.h")]
    stale = not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libzkoracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_init()
        _LIB.orc_domain_new.restype = C.c_void_p
        _LIB.orc_domain_extended_k.restype = C.c_uint32
        _LIB.orc_fr_inv64.restype = C.c_uint64
        _LIB.orc_fq_inv64.restype = C.c_uint64
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a


def cpu_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# ---- field ---------------------------------------------------------------------------
_OPS = {"add": 0, "sub": 1, "mul": 2, "neg": 3, "inv": 4, "sqr": 5, "from_raw": 6, "to_raw": 7, "double": 8}


def field_op(field, op, a, b=None):
    a = _u64(a)
    out = np.empty_like(a)
    n = a.size // 4
    bb = _u64(b) if b is not None else None
    lib().orc_field_op(C.c_int(0 if field == "fr" else 1), C.c_int(_OPS[op]), _p(a),
                       _p(bb) if bb is not None else None, _p(out), C.c_size_t(n))
    return out


def fr_constants():
    out = np.zeros((6, 4), dtype=np.uint64)
    lib().orc_fr_constants(_p(out))
    return dict(zip(["ROOT_OF_UNITY", "ROOT_OF_UNITY_INV", "ZETA", "DELTA", "R", "R2"], out))


def fq_constants():
    out = np.zeros((3, 4), dtype=np.uint64)
    lib().orc_fq_constants(_p(out))
    return dict(zip(["R", "R2", "R3"], out))


def fr_from_u512(wide):
    wide = _u64(wide)
    n = wide.size // 8
    out = np.empty((n, 4), dtype=np.uint64)
    lib().orc_fr_from_u512(_p(wide), _p(out), C.c_size_t(n))
    return out


# ---- curve ---------------------------------------------------------------------------
def g1_add(a, b):
    out = np.empty(12, dtype=np.uint64)
    lib().orc_g1_add(_p(_u64(a)), _p(_u64(b)), _p(out))
    return out


def g1_double(a):
    out = np.empty(12, dtype=np.uint64)
    lib().orc_g1_double(_p(_u64(a)), _p(out))
    return out


def g1_add_mixed(a, b):
    out = np.empty(12, dtype=np.uint64)
    lib().orc_g1_add_mixed(_p(_u64(a)), _p(_u64(b)), _p(out))
    return out


def g1_to_affine(a):
    a = _u64(a).reshape(-1, 12)
    out = np.empty((a.shape[0], 8), dtype=np.uint64)
    lib().orc_g1_to_affine(_p(a), _p(out), C.c_size_t(a.shape[0]))
    return out


def g1_on_curve(a):
    a = _u64(a).reshape(-1, 8)
    return bool(lib().orc_g1_on_curve(_p(a), C.c_size_t(a.shape[0])))


def g1_mul(base, scalar):
    out = np.empty(12, dtype=np.uint64)
    lib().orc_g1_mul(_p(_u64(base)), _p(_u64(scalar)), _p(out))
    return out


def msm_naive(scalars, bases):
    scalars = _u64(scalars).reshape(-1, 4)
    bases = _u64(bases).reshape(-1, 8)
    out = np.empty(12, dtype=np.uint64)
    lib().orc_msm_naive(_p(scalars), _p(bases), C.c_size_t(scalars.shape[0]), _p(out))
    return out


def best_multiexp(scalars, bases, threads=None):
    scalars = _u64(scalars).reshape(-1, 4)
    bases = _u64(bases).reshape(-1, 8)
    assert scalars.shape[0] == bases.shape[0]
    out = np.empty(12, dtype=np.uint64)
    lib().orc_best_multiexp(_p(scalars), _p(bases), C.c_size_t(scalars.shape[0]),
                            C.c_int(threads or cpu_threads()), _p(out))
    return out


# ---- fft / domain ----------------------------------------------------------------------
def best_fft(a, omega, log_n, threads=None):
    a = _u64(a).copy().reshape(-1, 4)
    assert a.shape[0] == 1 << log_n
    lib().orc_best_fft(_p(a), _p(_u64(omega)), C.c_uint32(log_n), C.c_int(threads or cpu_threads()))
    return a


def dft_naive(a, omega, log_n):
    a = _u64(a).reshape(-1, 4)
    out = np.empty_like(a)
    lib().orc_dft_naive(_p(a), _p(_u64(omega)), C.c_uint32(log_n), _p(out))
    return out


def eval_polynomial(poly, point):
    poly = _u64(poly).reshape(-1, 4)
    out = np.empty(4, dtype=np.uint64)
    lib().orc_eval_polynomial(_p(poly), C.c_size_t(poly.shape[0]), _p(_u64(point)), _p(out))
    return out


class Domain:
    """EvaluationDomain::new(j, k)"""

    def __init__(self, j, k, threads=None):
        self.j, self.k = j, k
        self.threads = threads or cpu_threads()
        self.h = C.c_void_p(lib().orc_domain_new(C.c_uint32(j), C.c_uint32(k), C.c_int(self.threads)))
        self.extended_k = lib().orc_domain_extended_k(self.h)
        self.n = 1 << k
        self.ext_n = 1 << self.extended_k
        c = np.zeros((8, 4), dtype=np.uint64)
        lib().orc_domain_constants(self.h, _p(c))
        (self.omega, self.omega_inv, self.extended_omega, self.extended_omega_inv,
         self.g_coset, self.g_coset_inv, self.ifft_divisor, self.extended_ifft_divisor) = c

    def __del__(self):
        try:
            lib().orc_domain_free(self.h)
        except Exception:
            pass

    def t_evaluations(self):
        out = np.empty((1 << (self.extended_k - self.k), 4), dtype=np.uint64)
        lib().orc_domain_t_evaluations(self.h, _p(out))
        return out

    def lagrange_to_coeff(self, a):
        a = _u64(a).copy().reshape(self.n, 4)
        lib().orc_lagrange_to_coeff(self.h, _p(a))
        return a

    def coeff_to_lagrange(self, a):
        a = _u64(a).copy().reshape(self.n, 4)
        lib().orc_coeff_to_lagrange(self.h, _p(a))
        return a

    def coeff_to_extended(self, a):
        a = _u64(a).reshape(self.n, 4)
        out = np.empty((self.ext_n, 4), dtype=np.uint64)
        lib().orc_coeff_to_extended(self.h, _p(a), _p(out))
        return out

    def extended_to_coeff(self, a):
        a = _u64(a).copy().reshape(self.ext_n, 4)
        out = np.empty((self.n * (self.j - 1), 4), dtype=np.uint64)
        lib().orc_extended_to_coeff(self.h, _p(a), _p(out))
        return out

    def divide_by_vanishing_poly(self, a):
        a = _u64(a).copy().reshape(self.ext_n, 4)
        lib().orc_divide_by_vanishing_poly(self.h, _p(a))
        return a


# ---- generators ------------------------------------------------------------------------
def gen_scalars(seed, n):
    out = np.empty((n, 4), dtype=np.uint64)
    lib().orc_gen_scalars(C.c_uint64(seed), C.c_size_t(n), _p(out))
    return out


def gen_points(seed, n, threads=None):
    out = np.empty((n, 8), dtype=np.uint64)
    lib().orc_gen_points(C.c_uint64(seed), C.c_size_t(n), C.c_int(threads or cpu_threads()), _p(out))
    return out


def setup_params(k, s, threads=None):
    """ParamsKZG::setup with the known trapdoor `s` (canonical int) -> (g, g_lagrange), (n, 8) affine arrays"""
    n = 1 << k
    g, gl = np.empty((n, 8), dtype=np.uint64), np.empty((n, 8), dtype=np.uint64)
    raw = np.array([(s >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
    lib().orc_setup_params(C.c_uint32(k), _p(raw), C.c_int(threads or cpu_threads()), _p(g), _p(gl))
    return g, gl


def evaluate_h(desc_struct, fixed_cosets, perm_cosets, l0, l_last, l_active,
               advice, instance, challenges, y, beta, gamma, theta, lookup_polys, perm_products,
               ext_n, threads=None):
    """desc_struct: a ctypes zk_evaluator_desc built by the product's Python host layer
    (plain data; the oracle only reads it)."""
    def table(arrs):
        arrs = [_u64(a) for a in arrs]
        t = (C.c_void_p * max(1, len(arrs)))(*[a.ctypes.data for a in arrs])
        return t, arrs
    tf, kf = table(fixed_cosets)
    tp, kp = table(perm_cosets)
    ta, ka = table(advice)
    ti, ki = table(instance)
    tl, kl = table(lookup_polys)
    tz, kz = table(perm_products)
    ch = _u64(challenges) if len(challenges) else np.zeros((1, 4), dtype=np.uint64)
    out = np.empty((ext_n, 4), dtype=np.uint64)
    lib().orc_evaluate_h(C.byref(desc_struct), C.c_int(threads or cpu_threads()),
                         tf, tp, _p(_u64(l0)), _p(_u64(l_last)), _p(_u64(l_active)),
                         ta, ti, _p(ch), _p(_u64(y)), _p(_u64(beta)), _p(_u64(gamma)), _p(_u64(theta)),
                         tl, tz, _p(out))
    return out


# ---- prover tail (oracle/prover_tail.inc.h) ---------------------------------------------------------
def _table(arrs):
    arrs = [_u64(a) for a in arrs]
    t = (C.c_void_p * max(1, len(arrs)))(*[a.ctypes.data for a in arrs])
    return t, arrs


def eval_graph_rows(graph_struct, k, fixed, advice, instance, challenges, theta, threads=None):
    """graph (ctypes zk_graph, plain data) evaluated on the n Lagrange rows, rot_scale = 1."""
    tf, kf = _table(fixed)
    ta, ka = _table(advice)
    ti, ki = _table(instance)
    ch = _u64(challenges) if len(challenges) else np.zeros((1, 4), dtype=np.uint64)
    out = np.empty((1 << k, 4), dtype=np.uint64)
    lib().orc_eval_graph_rows(C.byref(graph_struct), C.c_int(threads or cpu_threads()), C.c_uint32(k), tf, ta, ti,
                              _p(ch), _p(_u64(theta)), _p(out))
    return out


def permute_expression_pair(inp, tab, usable):
    inp, tab = _u64(inp), _u64(tab)
    a = np.empty((usable, 4), dtype=np.uint64)
    s = np.empty((usable, 4), dtype=np.uint64)
    rc = lib().orc_permute_expression_pair(_p(inp), _p(tab), C.c_size_t(usable), _p(a), _p(s))
    if rc != 0:
        raise ValueError("ConstraintSystemFailure: lookup input not in table")
    return a, s


def permutation_product(k, values, sigma, beta, gamma, omega, deltaomega0, last_z, threads=None):
    tv, kv = _table(values)
    ts, ks = _table(sigma)
    out = np.empty((1 << k, 4), dtype=np.uint64)
    lib().orc_permutation_product(C.c_uint32(k), C.c_int(threads or cpu_threads()), C.c_uint32(len(values)), tv, ts,
                                  _p(_u64(beta)), _p(_u64(gamma)), _p(_u64(omega)), _p(_u64(deltaomega0)),
                                  _p(_u64(last_z)), _p(out))
    return out


def lookup_product(k, cin, ctab, pin, ptab, beta, gamma):
    out = np.empty((1 << k, 4), dtype=np.uint64)
    lib().orc_lookup_product(C.c_uint32(k), _p(_u64(cin)), _p(_u64(ctab)), _p(_u64(pin)), _p(_u64(ptab)),
                             _p(_u64(beta)), _p(_u64(gamma)), _p(out))
    return out


def kate_division(a, z):
    a = _u64(a).reshape(-1, 4)
    out = np.empty((a.shape[0] - 1, 4), dtype=np.uint64)
    lib().orc_kate_division(_p(a), C.c_size_t(a.shape[0]), _p(_u64(z)), _p(out))
    return out


def poly_lincomb(polys, scalars, threads=None):
    tp, kp = _table(polys)
    n = kp[0].size // 4
    sc = _u64(np.stack([_u64(s) for s in scalars]))
    out = np.empty((n, 4), dtype=np.uint64)
    lib().orc_poly_lincomb(C.c_size_t(n), C.c_int(threads or cpu_threads()), C.c_size_t(len(polys)), tp, _p(sc), _p(out))
    return out


def permutation_sigma(k, mapping, omega):
    """mapping: (ncols, n, 2) uint32 of (column', row') -> (ncols, n, 4) sigma values in Lagrange form"""
    mapping = np.ascontiguousarray(mapping, dtype=np.uint32)
    ncols = mapping.shape[0]
    out = np.empty((ncols, 1 << k, 4), dtype=np.uint64)
    lib().orc_permutation_sigma(C.c_uint32(k), C.c_uint32(ncols), _p(mapping), _p(_u64(omega)), _p(out))
    return out
