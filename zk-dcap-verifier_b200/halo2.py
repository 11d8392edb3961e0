"""
halo2.py -- host-side mirror of the halo2_proofs interface for the GPU hot path.

Same names, argument meaning and error behaviour as the functions the reference reaches
through `create_proof` / `keygen_pk` (circuits/src/sgx_dcap_verifier.rs:803-822):

    best_multiexp(coeffs, bases)                       arithmetic.rs
    best_fft(a, omega, log_n)                          arithmetic.rs
    EvaluationDomain(j, k).{lagrange_to_coeff, coeff_to_extended, extended_to_coeff, ...}   poly/domain.rs
    ParamsKZG.{commit, commit_lagrange}                poly/kzg/commitment.rs
    Expression / ConstraintSystem.{degree, blinding_factors}                 plonk/circuit.rs
    GraphEvaluator.add_expression / Evaluator.new(cs) / Evaluator.evaluate_h plonk/evaluation.rs

Every compute call goes through the C ABI of libzkdcap_b200.so (include/zkdcap_b200.h);
nothing here computes on the CPU and nothing here touches oracle/.  Field elements and points
are numpy uint64 arrays in halo2curves' in-memory layout (Montgomery limbs): Fr (..., 4),
G1Affine (..., 8), G1 (..., 12).
"""
import ctypes as C

import numpy as np

from . import _abi
from . import keyfile
from ._abi import ZkError

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
P_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
_MONT_R = 1 << 256


def fr_to_limbs(v):
    """canonical int -> Montgomery limbs (host-side constant preparation only)"""
    m = (v % R_MOD) * _MONT_R % R_MOD
    return np.array([(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def fq_to_limbs(v):
    m = (v % P_MOD) * _MONT_R % P_MOD
    return np.array([(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class Context:
    """One device + one stream (zk_ctx)."""

    def __init__(self, device=0, stream=None):
        self.lib = _abi.load()
        h = C.c_void_p()
        rc = self.lib.zk_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise ZkError("zk_ctx_create(device=%d) failed with %d: no CUDA device? (this backend has no CPU fallback)"
                          % (device, rc))
        self.h = h
        self.device = device

    @classmethod
    def borrowed(cls, handle, device=0):
        """Wrap a zk_ctx owned by another object (e.g. a key's second stream) for timing / launch statistics."""
        self = cls.__new__(cls)
        self.lib, self.h, self.device, self._borrowed = _abi.load(), C.c_void_p(handle), device, True
        return self

    def check(self, rc):
        if rc < 0:
            raise ZkError("%s (status %d)" % (self.lib.zk_last_error(self.h).decode(), rc))
        return rc

    def close(self):
        if getattr(self, "h", None):
            if not getattr(self, "_borrowed", False):
                self.lib.zk_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self.check(self.lib.zk_ctx_sync(self.h))

    def timing(self, enable=True):
        self.check(self.lib.zk_ctx_timing(self.h, 1 if enable else 0))

    def timing_reset(self):
        self.check(self.lib.zk_ctx_timing_reset(self.h))

    def timing_read(self):
        rows = (_abi.zk_kernel_stat * 64)()
        n = self.check(self.lib.zk_ctx_timing_read(self.h, rows, 64))
        return {rows[i].name.decode(): (int(rows[i].launches), float(rows[i].total_ms)) for i in range(n)}

    def launch_count(self):
        return int(self.lib.zk_ctx_launch_count(self.h))

    def counters(self, reset=False):
        """{"msm_bucket_entries": ..., "msms": ...} since creation / the last reset (device-side counters)"""
        out = (C.c_uint64 * 4)()
        self.check(self.lib.zk_ctx_counters(self.h, out, 1 if reset else 0))
        return {"msm_bucket_entries": int(out[0]), "msms": int(out[1])}

    def bench_modmul(self, field=0, iters=2000):
        out = C.c_double()
        self.check(self.lib.zk_bench_modmul(self.h, field, iters, C.byref(out)))
        return out.value

    def bench_imad(self, kind=0, iters=2000):
        out = C.c_double()
        self.check(self.lib.zk_bench_imad(self.h, kind, iters, C.byref(out)))
        return out.value

    def field_op(self, field, op, a, b=None):
        ops = {"add": 0, "sub": 1, "mul": 2, "neg": 3, "inv": 4, "sqr": 5, "to_mont": 6, "from_mont": 7, "double": 8}
        a = _u64(a)
        out = np.empty_like(a)
        bb = _u64(b) if b is not None else None
        self.check(self.lib.zk_field_op(self.h, 0 if field == "fr" else 1, ops[op], _ptr(a),
                                        _ptr(bb) if bb is not None else None, _ptr(out), a.size // 4))
        return out


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ---- arithmetic.rs ---------------------------------------------------------------------------------
def best_multiexp(coeffs, bases, ctx=None):
    """sum_i coeffs[i] * bases[i] -> G1 (12 limbs, normalised).  Panics (raises) on length mismatch like the
    Rust `assert_eq!(coeffs.len(), bases.len())`."""
    ctx = ctx or default_context()
    coeffs = _u64(coeffs, (-1, 4))
    bases = _u64(bases, (-1, 8))
    if coeffs.shape[0] != bases.shape[0]:
        raise AssertionError("best_multiexp: coeffs.len() != bases.len()")
    out = np.empty(12, dtype=np.uint64)
    ctx.check(ctx.lib.zk_msm(ctx.h, _ptr(coeffs), _ptr(bases), _abi.BASES_HOST, coeffs.shape[0], _ptr(out)))
    return out


def best_fft(a, omega, log_n, ctx=None):
    """In-place FFT of `a` (numpy (2^log_n, 4) uint64)."""
    ctx = ctx or default_context()
    if not (isinstance(a, np.ndarray) and a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]):
        raise TypeError("best_fft works in place on a C-contiguous uint64 array")
    if a.size != 4 << log_n:
        raise AssertionError("best_fft: a.len() != 1 << log_n")
    ctx.check(ctx.lib.zk_ntt(ctx.h, _ptr(a), _ptr(_u64(omega)), log_n))
    return a


# SerdeFormat: "processed" = halo2curves 0.3.1 compressed; "processed_v04" = halo2curves >= 0.4 / halo2curves-axiom
# compressed (parity in bit 6, identity flag in bit 7); "raw_bytes" = x | y Montgomery limbs
SERDE_FORMATS = {"processed": 0, "processed_v04": 1, "raw_bytes": 2}


def _g1_bytes(fmt):
    return 64 if SERDE_FORMATS[fmt] == 2 else 32


def g1_compress(points, ctx=None, fmt="processed"):
    """G1Affine::to_bytes (or SerdeObject::write_raw for "raw_bytes") for (n, 8) points -> (n, 32 | 64) uint8"""
    ctx = ctx or default_context()
    points = _u64(points, (-1, 8))
    out = np.empty((points.shape[0], _g1_bytes(fmt)), dtype=np.uint8)
    ctx.check(ctx.lib.zk_g1_encode(ctx.h, SERDE_FORMATS[fmt], _ptr(points), points.shape[0], _ptr(out)))
    return out


def g1_decompress(data, ctx=None, fmt="processed"):
    """G1Affine::from_bytes for (n, 32 | 64) uint8 -> (n, 8) points; raises on an invalid encoding"""
    ctx = ctx or default_context()
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, _g1_bytes(fmt))
    out = np.empty((data.shape[0], 8), dtype=np.uint64)
    ctx.check(ctx.lib.zk_g1_decode(ctx.h, SERDE_FORMATS[fmt], _ptr(data), data.shape[0], _ptr(out)))
    return out


def g1_sum(points, ctx=None):
    ctx = ctx or default_context()
    points = _u64(points, (-1, 12))
    out = np.empty(12, dtype=np.uint64)
    ctx.check(ctx.lib.zk_g1_sum(ctx.h, _ptr(points), points.shape[0], _ptr(out)))
    return out


def g2_decode(data, fmt="processed"):
    """ParamsKZG.g2 / s_g2 as an SRS file holds them (64 compressed bytes, or 128 bytes of Montgomery limbs for
    "raw_bytes") -> the 128 uncompressed canonical bytes verify_proof takes"""
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    if buf.size != 2 * _g1_bytes(fmt):
        raise ZkError("g2_decode: %d bytes, format %s takes %d" % (buf.size, fmt, 2 * _g1_bytes(fmt)))
    out = np.empty(128, dtype=np.uint8)
    if _abi.load().zk_g2_decode(SERDE_FORMATS[fmt], _ptr(buf), _ptr(out)) != 0:
        raise ZkError("g2_decode: not a G2 point encoding")
    return out.tobytes()


def g2_encode(uncompressed, fmt="processed"):
    buf = np.frombuffer(bytes(uncompressed), dtype=np.uint8)
    if buf.size != 128:
        raise ZkError("g2_encode: 128 uncompressed bytes expected")
    out = np.empty(2 * _g1_bytes(fmt), dtype=np.uint8)
    if _abi.load().zk_g2_encode(SERDE_FORMATS[fmt], _ptr(buf), _ptr(out)) != 0:
        raise ZkError("g2_encode: not a G2 point")
    return out.tobytes()


def encode_calldata(instances, proof):
    """snark-verifier `encode_calldata(&[instances], &proof)` (bin/src/main.rs:236; the input of the generated Solidity verifier):
    every instance value as 32 big-endian bytes of its canonical form, column after column, then the proof bytes (an
    EvmTranscript proof, `set_transcript("evm")`).  instances: one (len, 4) Montgomery-limb array per instance column."""
    rinv = pow(1 << 256, -1, keyfile.R_MOD)
    out = bytearray()
    for col in instances:
        for row in _u64(col, (-1, 4)):
            v = sum(int(row[i]) << (64 * i) for i in range(4)) * rinv % keyfile.R_MOD
            out += v.to_bytes(32, "big")
    return bytes(out) + bytes(proof)


# ---- poly/kzg/commitment.rs -----------------------------------------------------------------------
class ParamsKZG:
    """ParamsKZG { k, n, g, g_lagrange } with both base sets resident in HBM."""

    def __init__(self, k, g, g_lagrange, ctx=None, precompute=True):
        self.ctx = ctx or default_context()
        self.k, self.n = k, 1 << k
        g = _u64(g, (self.n, 8))
        gl = _u64(g_lagrange, (self.n, 8))
        self.ctx.check(self.ctx.lib.zk_srs_upload(self.ctx.h, k, _ptr(g), _ptr(gl), 1 if precompute else 0))

    @classmethod
    def read(cls, data, ctx=None, precompute=True, fmt="processed"):
        """ParamsKZG::read / read_custom: `data` = bytes of a kzg_bn254_{k}.srs file; compressed points are decompressed on
        the device."""
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        k = C.c_uint32()
        qb = 2 * _g1_bytes(fmt)
        self.g2, self.s_g2 = np.zeros(qb, np.uint8), np.zeros(qb, np.uint8)
        self.ctx.check(self.ctx.lib.zk_params_read_format(self.ctx.h, SERDE_FORMATS[fmt], _ptr(buf), buf.size,
                                                          1 if precompute else 0, C.byref(k), _ptr(self.g2), _ptr(self.s_g2)))
        self.k, self.n = k.value, 1 << k.value
        # the verifier's form (128 uncompressed canonical bytes); None when the file's G2 fields are not curve points
        try:
            self.g2_uncompressed, self.s_g2_uncompressed = g2_decode(self.g2, fmt), g2_decode(self.s_g2, fmt)
        except ZkError:
            self.g2_uncompressed = self.s_g2_uncompressed = None
        return self

    def write(self, g2=None, s_g2=None, fmt="processed"):
        """ParamsKZG::write / write_custom of the resident SRS -> bytes"""
        pb = _g1_bytes(fmt)
        out = np.empty(4 + 2 * self.n * pb + 4 * pb, dtype=np.uint8)
        ln = C.c_size_t()

        def g2_bytes(v, name):
            v = v if v is not None else getattr(self, name, None)
            v = np.zeros(2 * pb, np.uint8) if v is None else np.ascontiguousarray(v, dtype=np.uint8)
            if v.size != 2 * pb:
                raise ValueError("%s: %d bytes, the format takes %d" % (name, v.size, 2 * pb))
            return v
        g2, s_g2 = g2_bytes(g2, "g2"), g2_bytes(s_g2, "s_g2")
        self.ctx.check(self.ctx.lib.zk_params_write_format(self.ctx.h, SERDE_FORMATS[fmt], _ptr(g2), _ptr(s_g2), _ptr(out),
                                                           out.size, C.byref(ln)))
        return out[:ln.value].tobytes()

    def _commit(self, poly, which):
        poly = _u64(poly, (-1, 4))
        if poly.shape[0] > self.n:
            raise AssertionError("commit: polynomial longer than the SRS")
        out = np.empty(12, dtype=np.uint64)
        self.ctx.check(self.ctx.lib.zk_msm(self.ctx.h, _ptr(poly), None, which, poly.shape[0], _ptr(out)))
        return out

    def commit(self, poly, blind=None):
        """commit(&Polynomial<Fr, Coeff>, Blind) -- the blind is ignored by KZG, as upstream."""
        return self._commit(poly, _abi.BASES_G)

    def commit_lagrange(self, poly, blind=None):
        return self._commit(poly, _abi.BASES_G_LAGRANGE)


# ---- poly/domain.rs ---------------------------------------------------------------------------------
class EvaluationDomain:
    """EvaluationDomain::new(j, k)"""

    def __init__(self, j, k, ctx=None, _handle=None):
        self.ctx = ctx or default_context()
        self.j, self.k = j, k
        self._owned = _handle is None
        if _handle is None:
            h = C.c_void_p()
            self.ctx.check(self.ctx.lib.zk_domain_create(self.ctx.h, j, k, C.byref(h)))
            self.h = h
        else:
            self.h = _handle
        self.extended_k = self.ctx.lib.zk_domain_extended_k(self.h)
        self.n, self.ext_n = 1 << k, 1 << self.extended_k
        c = np.zeros((8, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.zk_domain_constants(self.h, _ptr(c)))
        (self.omega, self.omega_inv, self.extended_omega, self.extended_omega_inv,
         self.g_coset, self.g_coset_inv, self.ifft_divisor, self.extended_ifft_divisor) = c

    def __del__(self):
        try:
            if self._owned and self.h and self.ctx.h:
                self.ctx.lib.zk_domain_destroy(self.h)
        except Exception:
            pass

    def extended_len(self):
        return self.ext_n

    def get_quotient_poly_degree(self):
        return self.j - 1

    def lagrange_to_coeff(self, a):
        a = _u64(a, (self.n, 4)).copy()
        self.ctx.check(self.ctx.lib.zk_lagrange_to_coeff(self.h, _ptr(a)))
        return a

    def coeff_to_lagrange(self, a):
        a = _u64(a, (self.n, 4)).copy()
        self.ctx.check(self.ctx.lib.zk_coeff_to_lagrange(self.h, _ptr(a)))
        return a

    def coeff_to_extended(self, a):
        a = _u64(a, (self.n, 4))
        out = np.empty((self.ext_n, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.zk_coeff_to_extended(self.h, _ptr(a), _ptr(out)))
        return out

    def extended_to_coeff(self, a):
        a = _u64(a, (self.ext_n, 4))
        out = np.empty((self.n * (self.j - 1), 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.zk_extended_to_coeff(self.h, _ptr(a), _ptr(out)))
        return out

    def divide_by_vanishing_poly(self, a):
        a = _u64(a, (self.ext_n, 4)).copy()
        self.ctx.check(self.ctx.lib.zk_divide_by_vanishing_poly(self.h, _ptr(a)))
        return a


# ---- plonk/circuit.rs: Expression ---------------------------------------------------------------------
class Expression:
    """Expression<F> tree; field constants are canonical Python ints."""

    def __init__(self, tag, *args):
        self.tag, self.args = tag, args

    @staticmethod
    def Constant(v): return Expression("Constant", v % R_MOD)
    @staticmethod
    def Fixed(column_index, rotation=0): return Expression("Fixed", column_index, rotation)
    @staticmethod
    def Advice(column_index, rotation=0): return Expression("Advice", column_index, rotation)
    @staticmethod
    def Instance(column_index, rotation=0): return Expression("Instance", column_index, rotation)
    @staticmethod
    def Challenge(index): return Expression("Challenge", index)

    def __neg__(self): return Expression("Negated", self)
    def __add__(self, o): return Expression("Sum", self, o)
    def __sub__(self, o): return Expression("Sum", self, Expression("Negated", o))    # a + (-b), as upstream
    def __mul__(self, o):
        if isinstance(o, int):
            return Expression("Scaled", self, o % R_MOD)
        return Expression("Product", self, o)

    def degree(self):
        t, a = self.tag, self.args
        if t in ("Constant", "Challenge"):
            return 0
        if t in ("Fixed", "Advice", "Instance"):
            return 1
        if t in ("Negated", "Scaled"):
            return a[0].degree()
        if t == "Sum":
            return max(a[0].degree(), a[1].degree())
        return a[0].degree() + a[1].degree()


class ConstraintSystem:
    """The part of ConstraintSystem<F> that keygen and evaluate_h depend on."""

    def __init__(self, num_fixed=0, num_advice=0, num_instance=0, num_challenges=0):
        self.num_fixed_columns, self.num_advice_columns = num_fixed, num_advice
        self.num_instance_columns, self.num_challenges = num_instance, num_challenges
        self.gates = []            # list of gate polynomials (Expression), cs.gates order flattened
        self.lookups = []          # list of (input_expressions, table_expressions)
        self.permutation_columns = []   # list of (column_type, index); _abi.COL_*
        self.minimum_degree = None
        self.num_advice_queries = [0] * num_advice
        # cs.{advice,fixed,instance}_queries: (column index, rotation) in first-query order.  They fix the order of
        # the evaluations in the proof (plonk/prover.rs) and, through num_advice_queries, blinding_factors().
        self.advice_queries, self.fixed_queries, self.instance_queries = [], [], []

    def create_gate(self, polys):
        for p in polys:
            self.gates.append(p)
            self._count_queries(p)

    def lookup(self, inputs, tables):
        self.lookups.append((list(inputs), list(tables)))
        for e in list(inputs) + list(tables):
            self._count_queries(e)

    def enable_equality(self, column_type, index):
        # ConstraintSystem::enable_equality: query_any_index(column, Rotation::cur()) + permutation.add_column
        self._query({_abi.COL_ADVICE: "Advice", _abi.COL_FIXED: "Fixed", _abi.COL_INSTANCE: "Instance"}[column_type], index, 0)
        if (column_type, index) not in self.permutation_columns:
            self.permutation_columns.append((column_type, index))

    def _query(self, tag, column, rotation):
        lst = {"Advice": self.advice_queries, "Fixed": self.fixed_queries, "Instance": self.instance_queries}[tag]
        if (column, rotation) not in lst:
            lst.append((column, rotation))
            if tag == "Advice":
                self.num_advice_queries[column] += 1

    def _count_queries(self, e):
        # left-to-right walk of the expression: the order in which the closure would have called meta.query_*
        if e.tag in ("Advice", "Fixed", "Instance"):
            self._query(e.tag, e.args[0], e.args[1])
        for a in e.args:
            if isinstance(a, Expression):
                self._count_queries(a)

    def degree(self):
        degree = 3                                          # permutation::Argument::required_degree()
        for inputs, tables in self.lookups:
            idg = max([1] + [e.degree() for e in inputs])
            tdg = max([1] + [e.degree() for e in tables])
            degree = max(degree, max(4, 2 + idg + tdg))     # lookup::Argument::required_degree()
        for g in self.gates:
            degree = max(degree, g.degree())
        if self.minimum_degree is not None:
            degree = max(degree, self.minimum_degree)
        return degree

    def blinding_factors(self):
        factors = max([1] + self.num_advice_queries)
        return max(3, factors) + 2


# ---- plonk/evaluation.rs ------------------------------------------------------------------------------
class GraphEvaluator:
    """GraphEvaluator<C>: constants seeded with [0, 1, 2]; identical calculations share an intermediate."""

    def __init__(self):
        self.constants = [0, 1, 2]
        self.rotations = []
        self.calculations = []      # (calculation tuple, target)
        self.num_intermediates = 0

    def add_rotation(self, rot):
        if rot not in self.rotations:
            self.rotations.append(rot)
        return self.rotations.index(rot)

    def add_constant(self, c):
        c %= R_MOD
        if c not in self.constants:
            self.constants.append(c)
        return (_abi.VS_CONSTANT, self.constants.index(c), 0)

    def add_calculation(self, calc):
        for c, target in self.calculations:
            if c == calc:
                return (_abi.VS_INTERMEDIATE, target, 0)
        target = self.num_intermediates
        self.calculations.append((calc, target))
        self.num_intermediates += 1
        return (_abi.VS_INTERMEDIATE, target, 0)

    def add_expression(self, e):
        ZERO, ONE, TWO = (_abi.VS_CONSTANT, 0, 0), (_abi.VS_CONSTANT, 1, 0), (_abi.VS_CONSTANT, 2, 0)
        t, a = e.tag, e.args
        if t == "Constant":
            return self.add_constant(a[0])
        if t in ("Fixed", "Advice", "Instance"):
            kind = {"Fixed": _abi.VS_FIXED, "Advice": _abi.VS_ADVICE, "Instance": _abi.VS_INSTANCE}[t]
            return self.add_calculation(("Store", (kind, a[0], self.add_rotation(a[1]))))
        if t == "Challenge":
            return self.add_calculation(("Store", (_abi.VS_CHALLENGE, a[0], 0)))
        if t == "Negated":
            if a[0].tag == "Constant":
                return self.add_constant(-a[0].args[0])
            ra = self.add_expression(a[0])
            return ra if ra == ZERO else self.add_calculation(("Negate", ra))
        if t == "Sum":
            if a[1].tag == "Negated":
                ra, rb = self.add_expression(a[0]), self.add_expression(a[1].args[0])
                if ra == ZERO:
                    return self.add_calculation(("Negate", rb))
                if rb == ZERO:
                    return ra
                return self.add_calculation(("Sub", ra, rb))
            ra, rb = self.add_expression(a[0]), self.add_expression(a[1])
            if ra == ZERO:
                return rb
            if rb == ZERO:
                return ra
            return self.add_calculation(("Add", ra, rb) if ra <= rb else ("Add", rb, ra))
        if t == "Product":
            ra, rb = self.add_expression(a[0]), self.add_expression(a[1])
            if ra == ZERO or rb == ZERO:
                return ZERO
            if ra == ONE:
                return rb
            if rb == ONE:
                return ra
            if ra == TWO:
                return self.add_calculation(("Double", rb))
            if rb == TWO:
                return self.add_calculation(("Double", ra))
            if ra == rb:
                return self.add_calculation(("Square", ra))
            return self.add_calculation(("Mul", ra, rb) if ra <= rb else ("Mul", rb, ra))
        if t == "Scaled":
            if a[1] == 0:
                return ZERO
            if a[1] == 1:
                return self.add_expression(a[0])
            cst = self.add_constant(a[1])
            ra = self.add_expression(a[0])
            return self.add_calculation(("Mul", ra, cst))
        raise ValueError("Selector expressions never reach the evaluator (compressed into fixed columns at keygen)")

    def to_c(self):
        """-> (zk_graph, keepalive list)"""
        consts = np.stack([fr_to_limbs(c) for c in self.constants]) if self.constants else np.zeros((1, 4), np.uint64)
        rots = (C.c_int32 * max(1, len(self.rotations)))(*self.rotations)
        calcs = (_abi.zk_calculation * max(1, len(self.calculations)))()
        parts = []
        kinds = {"Add": _abi.CALC_ADD, "Sub": _abi.CALC_SUB, "Mul": _abi.CALC_MUL, "Square": _abi.CALC_SQUARE,
                 "Double": _abi.CALC_DOUBLE, "Negate": _abi.CALC_NEGATE, "Horner": _abi.CALC_HORNER,
                 "Store": _abi.CALC_STORE}
        for i, (calc, target) in enumerate(self.calculations):
            c = calcs[i]
            c.kind, c.target = kinds[calc[0]], target
            c.a = _abi.zk_value_source(*calc[1])
            if calc[0] in ("Add", "Sub", "Mul"):
                c.b = _abi.zk_value_source(*calc[2])
            if calc[0] == "Horner":           # ("Horner", start, parts, factor)
                c.b = _abi.zk_value_source(*calc[3])
                c.parts_off, c.parts_len = len(parts), len(calc[2])
                parts.extend(calc[2])
        hp = (_abi.zk_value_source * max(1, len(parts)))(*[_abi.zk_value_source(*p) for p in parts])
        g = _abi.zk_graph()
        g.constants, g.n_constants = consts.ctypes.data, len(self.constants)
        g.rotations, g.n_rotations = C.cast(rots, C.POINTER(C.c_int32)), len(self.rotations)
        g.calcs, g.n_calcs = C.cast(calcs, C.POINTER(_abi.zk_calculation)), len(self.calculations)
        g.horner_parts, g.n_horner_parts = C.cast(hp, C.POINTER(_abi.zk_value_source)), len(parts)
        g.num_intermediates = self.num_intermediates
        return g, [consts, rots, calcs, hp]


class Evaluator:
    """Evaluator<C> { custom_gates, lookups }"""

    def __init__(self):
        self.custom_gates = GraphEvaluator()
        self.lookups = []

    @staticmethod
    def new(cs):
        ev = Evaluator()
        parts = [ev.custom_gates.add_expression(p) for p in cs.gates]
        ev.custom_gates.add_calculation(("Horner", (_abi.VS_PREVIOUS_VALUE, 0, 0), tuple(parts), (_abi.VS_Y, 0, 0)))
        for inputs, tables in cs.lookups:
            g = GraphEvaluator()

            def evaluate_lc(exprs):
                ps = tuple(g.add_expression(e) for e in exprs)
                return g.add_calculation(("Horner", (_abi.VS_CONSTANT, 0, 0), ps, (_abi.VS_THETA, 0, 0)))
            compressed_input_coset = evaluate_lc(inputs)
            compressed_table_coset = evaluate_lc(tables)
            right_gamma = g.add_calculation(("Add", compressed_table_coset, (_abi.VS_GAMMA, 0, 0)))
            lc = g.add_calculation(("Add", compressed_input_coset, (_abi.VS_BETA, 0, 0)))
            g.add_calculation(("Mul", lc, right_gamma))
            ev.lookups.append(g)
        return ev

    def describe(self, cs, k):
        """-> (zk_evaluator_desc, keepalive) : the C image evaluate_h consumes"""
        keep = []
        d = _abi.zk_evaluator_desc()
        d.k, d.degree, d.blinding_factors = k, cs.degree(), cs.blinding_factors()
        d.n_fixed, d.n_advice = cs.num_fixed_columns, cs.num_advice_columns
        d.n_instance, d.n_challenges = cs.num_instance_columns, cs.num_challenges
        g, ka = self.custom_gates.to_c()
        d.custom_gates = g
        keep.append(ka)
        lk = (_abi.zk_graph * max(1, len(self.lookups)))()
        for i, lg in enumerate(self.lookups):
            lk[i], ka = lg.to_c()
            keep.append(ka)
        d.lookups, d.n_lookups = C.cast(lk, C.POINTER(_abi.zk_graph)), len(self.lookups)
        pc = (_abi.zk_column * max(1, len(cs.permutation_columns)))(
            *[_abi.zk_column(t, i) for t, i in cs.permutation_columns])
        d.perm_columns, d.n_perm_columns = C.cast(pc, C.POINTER(_abi.zk_column)), len(cs.permutation_columns)
        keep += [lk, pc]
        return d, keep


class Circuit:
    """What keygen / create_proof need to know about one circuit: the ConstraintSystem, k, the Evaluator image
    (zk_evaluator_desc) and, per lookup argument, the graphs of the theta-compressed input and table expressions
    that lookup::Argument::commit_permuted evaluates over the n Lagrange rows (plonk/lookup/prover.rs)."""

    def __init__(self, cs, k, transcript_repr=0):
        self.cs, self.k = cs, k
        self.transcript_repr = transcript_repr % R_MOD
        self.ev = Evaluator.new(cs)
        self.desc, self._keep = self.ev.describe(cs, k)
        self.lookup_input_graphs, self.lookup_table_graphs = [], []
        for inputs, tables in cs.lookups:
            for exprs, dst in ((inputs, self.lookup_input_graphs), (tables, self.lookup_table_graphs)):
                g = GraphEvaluator()
                parts = tuple(g.add_expression(e) for e in exprs)
                g.add_calculation(("Horner", (_abi.VS_CONSTANT, 0, 0), parts, (_abi.VS_THETA, 0, 0)))
                cg, keep = g.to_c()
                dst.append(cg)
                self._keep.append(keep)

    def to_c(self):
        """-> zk_circuit_desc (ctypes) for zk_keygen"""
        nl = len(self.cs.lookups)
        d = _abi.zk_circuit_desc()
        d.ev = self.desc
        li = (_abi.zk_graph * max(1, nl))(*self.lookup_input_graphs)
        lt = (_abi.zk_graph * max(1, nl))(*self.lookup_table_graphs)
        aq = (_abi.zk_query * max(1, len(self.cs.advice_queries)))(*[_abi.zk_query(c, r) for c, r in self.cs.advice_queries])
        fq = (_abi.zk_query * max(1, len(self.cs.fixed_queries)))(*[_abi.zk_query(c, r) for c, r in self.cs.fixed_queries])
        d.lookup_inputs, d.lookup_tables = C.cast(li, C.POINTER(_abi.zk_graph)), C.cast(lt, C.POINTER(_abi.zk_graph))
        d.advice_queries, d.n_advice_queries = C.cast(aq, C.POINTER(_abi.zk_query)), len(self.cs.advice_queries)
        d.fixed_queries, d.n_fixed_queries = C.cast(fq, C.POINTER(_abi.zk_query)), len(self.cs.fixed_queries)
        iq = (_abi.zk_query * max(1, len(self.cs.instance_queries)))(*[_abi.zk_query(c, r) for c, r in self.cs.instance_queries])
        d.instance_queries, d.n_instance_queries = C.cast(iq, C.POINTER(_abi.zk_query)), len(self.cs.instance_queries)
        self._keep.append(iq)
        tr = self.transcript_repr * _MONT_R % R_MOD
        for i in range(4):
            d.transcript_repr.l[i] = (tr >> (64 * i)) & 0xFFFFFFFFFFFFFFFF
        self._keep += [li, lt, aq, fq]
        return d


def _ptr_table(arrs):
    arrs = [_u64(a) for a in arrs]
    tab = (C.c_void_p * max(1, len(arrs)))(*[a.ctypes.data for a in arrs])
    return tab, arrs


class ProvingKey:
    """The slice of ProvingKey<C> evaluate_h reads (fixed_cosets, permutation.cosets, l0, l_last,
    l_active_row, ev), resident in HBM."""

    def __init__(self, cs, k, fixed, perm, l0, l_last, l_active_row, coeff_form=False, ctx=None):
        self.ctx = ctx or default_context()
        self.cs, self.k = cs, k
        self.ev = Evaluator.new(cs)
        self.desc, self._keep = self.ev.describe(cs, k)
        tf, kf = _ptr_table(fixed)
        tp, kp = _ptr_table(perm)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.zk_pk_upload(self.ctx.h, C.byref(self.desc), tf, tp, _ptr(_u64(l0)),
                                                 _ptr(_u64(l_last)), _ptr(_u64(l_active_row)),
                                                 1 if coeff_form else 0, C.byref(h)))
        self.h = h
        self.domain = EvaluationDomain(cs.degree(), k, self.ctx, _handle=C.c_void_p(self.ctx.lib.zk_pk_domain(h)))

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.zk_pk_destroy(self.h)
        except Exception:
            pass

    @classmethod
    def read(cls, data, cs, ctx=None, fmt="raw_bytes_unchecked", num_selectors=0):
        """ProvingKey::read / from_bytes (bin/src/main.rs:205-213 reads ./params/pk.bin with SerdeFormat::RawBytesUnchecked): the
        extended-domain parts of the file go to HBM, the parsed key (vk commitments, Lagrange / coefficient forms) stays on the
        instance as `file`.  Byte layout: keyfile.py (restated, unpinned)."""
        k = int.from_bytes(bytes(data[:4]), "big")
        ext_k = k
        while (1 << ext_k) < (1 << k) * (cs.degree() - 1):
            ext_k += 1
        pk = keyfile.read_pk(data, len(cs.permutation_columns), num_selectors, fmt, extended_k=ext_k)
        if len(pk["fixed_cosets"]) != cs.num_fixed_columns:
            raise keyfile.KeyFileError("the key holds %d fixed columns, the circuit %d" % (len(pk["fixed_cosets"]), cs.num_fixed_columns))
        self = cls(cs, k, pk["fixed_cosets"], pk["permutation_cosets"], pk["l0"], pk["l_last"], pk["l_active_row"], ctx=ctx)
        self.file = pk
        return self

    def program_stats(self):
        m, c = C.c_uint64(), C.c_uint64()
        self.ctx.check(self.ctx.lib.zk_pk_program_stats(self.h, C.byref(m), C.byref(c)))
        return {"muls_per_row": m.value, "columns_read": c.value}

    def evaluate_h(self, advice_polys, instance_polys, challenges, y, beta, gamma, theta, lookups, permutations):
        """Evaluator::evaluate_h(pk, advice_polys, instance_polys, challenges, y, beta, gamma, theta, lookups,
        permutations) for one circuit instance.  lookups: list of (product_poly, permuted_input_poly,
        permuted_table_poly); permutations: list of permutation_product_poly; all in coefficient form."""
        ta, ka = _ptr_table(advice_polys)
        ti, ki = _ptr_table(instance_polys)
        flat = [p for triple in lookups for p in triple]
        tl, kl = _ptr_table(flat)
        tz, kz = _ptr_table(permutations)
        ch = _u64(challenges, (-1, 4)) if len(challenges) else np.zeros((1, 4), np.uint64)
        out = np.empty((self.domain.ext_n, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.zk_evaluate_h(self.h, ta, ti, _ptr(ch), _ptr(_u64(y)), _ptr(_u64(beta)),
                                                  _ptr(_u64(gamma)), _ptr(_u64(theta)), tl, tz, _ptr(out)))
        return out


class Group:
    """The GPUs (one process each) that compute one proof together (zk_group): `uid` comes from Group.unique_id() on
    rank 0 and reaches the other ranks out of band (e.g. torch.distributed.broadcast_object_list)."""

    @staticmethod
    def unique_id():
        lib = _abi.load()
        buf = (C.c_uint8 * 256)()
        if lib.zk_group_unique_id(buf) != 0:
            raise ZkError("zk_group_unique_id failed (NCCL not available)")
        return bytes(buf)

    def __init__(self, ctx, rank, world, uid):
        self.ctx, self.rank, self.world = ctx, rank, world
        h = C.c_void_p()
        buf = (C.c_uint8 * 256)(*uid)
        ctx.check(ctx.lib.zk_group_create(ctx.h, rank, world, buf, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.zk_group_destroy(self.h)
            self.h = None

    @property
    def p2p(self):
        """True when the small exchanges go through peer-mapped windows (stores over NVLink + flags) instead of NCCL"""
        return bool(self.ctx.lib.zk_group_p2p_enabled(self.h))

    def msm_sharded_dev(self, d_scalars, n_local, d_out, lagrange=True):
        """ONE MSM sharded by base range: this rank's slice of the scalars (device) against its resident slice of the
        bases; d_out (device) receives the normalised G1 of the whole sum on every rank"""
        which = _abi.BASES_G_LAGRANGE if lagrange else _abi.BASES_G
        self.ctx.check(self.ctx.lib.zk_msm_sharded_dev(self.h, _vp(d_scalars), which, n_local, _vp(d_out)))

    @staticmethod
    def shard_plan(rank, world, count, n):
        """(mode, per, lo, hi): mode 0 = columns [lo, hi) per rank, mode 1 = base range [lo, hi) per rank"""
        lib = _abi.load()
        mode, per, lo, hi = C.c_int(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        lib.zk_group_shard_plan(rank, world, count, n, C.byref(mode), C.byref(per), C.byref(lo), C.byref(hi))
        return mode.value, per.value, lo.value, hi.value


# ---- plonk/keygen.rs + plonk/prover.rs ------------------------------------------------------------------
class ProverKey:
    """keygen_vk + keygen_pk (circuits/src/sgx_dcap_verifier.rs:803, :807): the whole ProvingKey resident in HBM,
    together with the per-proof work area.  `params` must hold the SRS of the circuit's k on the same context."""

    def __init__(self, params, circuit, fixed_values, perm_mapping):
        self.ctx, self.circuit, self.params = params.ctx, circuit, params
        cs, k = circuit.cs, circuit.k
        n = 1 << k
        if len(fixed_values) != cs.num_fixed_columns:
            raise AssertionError("keygen: fixed column count")
        self._cd = circuit.to_c()
        tf, kf = _ptr_table([_u64(f, (n, 4)) for f in fixed_values])
        mp = np.ascontiguousarray(perm_mapping, dtype=np.uint32).reshape(len(cs.permutation_columns), n, 2)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.zk_keygen(self.ctx.h, C.byref(self._cd), tf, _ptr(mp), C.byref(h)))
        self.h = h
        aux = self.ctx.lib.zk_pkey_aux_ctx(h)
        self.aux_ctx = Context.borrowed(aux, self.ctx.device) if aux else None
        self.randoms_needed = int(self.ctx.lib.zk_proof_randoms(h))
        self.max_proof_bytes = int(self.ctx.lib.zk_proof_max_bytes(h))

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.zk_pkey_destroy(self.h)
        except Exception:
            pass

    def set_transcript(self, kind):
        """"blake2b" = Blake2bWrite / Challenge255 (default), "evm" = snark-verifier's EvmTranscript (Keccak-256), "poseidon" =
        snark-verifier's PoseidonTranscript over the native loader (gen_proof, crates/p256-ecdsa/src/base.rs:200-211)"""
        self.ctx.check(self.ctx.lib.zk_pkey_set_transcript(self.h, _abi.TRANSCRIPTS[kind]))
        self.max_proof_bytes = int(self.ctx.lib.zk_proof_max_bytes(self.h))
        if hasattr(self, "_proof_buf"):
            del self._proof_buf

    def set_group(self, group):
        """Proofs of this key are computed jointly by the group's GPUs (every rank must make the same calls)."""
        self.ctx.check(self.ctx.lib.zk_pkey_set_group(self.h, group.h if group is not None else None))
        self._group = group

    def vk_commitments(self):
        """(vk.fixed_commitments, vk.permutation.commitments) as (count, 8) G1Affine arrays"""
        cs = self.circuit.cs
        fc = np.zeros((max(1, cs.num_fixed_columns), 8), dtype=np.uint64)
        pc = np.zeros((max(1, len(cs.permutation_columns)), 8), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.zk_pkey_vk_commitments(self.h, _ptr(fc), _ptr(pc)))
        return fc[:cs.num_fixed_columns], pc[:len(cs.permutation_columns)]

    def create_proof(self, advice_values, instances, randoms):
        """create_proof(params, pk, &[circuit], &[instances], rng, Blake2bWrite) -> transcript.finalize()
        (sgx_dcap_verifier.rs:814-823).  advice_values: Lagrange columns left by synthesize; instances: one (len, 4)
        array per instance column; randoms: (count, 4) -- the RNG's Fr::random outputs in draw order."""
        cs = self.circuit.cs
        n = 1 << self.circuit.k
        if len(advice_values) != cs.num_advice_columns or len(instances) != cs.num_instance_columns:
            raise AssertionError("create_proof: column counts")
        ta, ka = _ptr_table([_u64(a, (n, 4)) for a in advice_values])
        inst = [_u64(i, (-1, 4)) for i in instances]
        ti, ki = _ptr_table(inst)
        lens = (C.c_uint32 * max(1, len(inst)))(*[i.shape[0] for i in inst])
        rnd = _u64(randoms, (-1, 4))
        out = np.empty(self.max_proof_bytes, dtype=np.uint8)
        ln = C.c_size_t()
        self.ctx.check(self.ctx.lib.zk_create_proof(self.h, ta, ti, lens, _ptr(rnd), rnd.shape[0], _ptr(out), out.size,
                                                    C.byref(ln)))
        return out[:ln.value].tobytes()


    def create_proof_raw(self, advice_ptrs, randoms_ptr, n_randoms):
        """Same with raw addresses (pinned host memory or device memory) for the advice columns and the randoms;
        circuits without instance columns."""
        if len(advice_ptrs) != self.circuit.cs.num_advice_columns or self.circuit.cs.num_instance_columns:
            raise AssertionError("create_proof_raw: column counts")
        ta = (C.c_void_p * max(1, len(advice_ptrs)))(*[int(p) for p in advice_ptrs])
        if not hasattr(self, "_proof_buf"):
            self._proof_buf = np.empty(self.max_proof_bytes, dtype=np.uint8)
        ln = C.c_size_t()
        self.ctx.check(self.ctx.lib.zk_create_proof(self.h, ta, None, None, C.c_void_p(int(randoms_ptr)), n_randoms,
                                                    _ptr(self._proof_buf), self._proof_buf.size, C.byref(ln)))
        return self._proof_buf[:ln.value].tobytes()


def verify_proof(circuit, fixed_commitments, permutation_commitments, g1, g2, s_g2, instances, proof, transcript="blake2b"):
    """verify_proof::<KZGCommitmentScheme<Bn256>, VerifierSHPLONK<_>, Challenge255<_>, Blake2bRead<..>, SingleStrategy<_>>
    (circuits/src/sgx_dcap_verifier.rs:827-843) on the host.  fixed / permutation commitments: (count, 8) G1Affine;
    g1: (8,) G1Affine = params.g[0]; g2, s_g2: 128 uncompressed bytes each; instances: one (len, 4) array per column."""
    lib = _abi.load()
    cd = circuit.to_c()
    fc = _u64(fixed_commitments, (-1, 8)) if len(fixed_commitments) else np.zeros((1, 8), np.uint64)
    pc = _u64(permutation_commitments, (-1, 8)) if len(permutation_commitments) else np.zeros((1, 8), np.uint64)
    vp = _abi.zk_verifier_params()
    vp.fixed_commitments, vp.permutation_commitments = fc.ctypes.data, pc.ctypes.data
    for i, v in enumerate(_u64(g1, (8,))):
        vp.g1[i] = int(v)
    b2, bs = np.frombuffer(bytes(g2), dtype=np.uint8), np.frombuffer(bytes(s_g2), dtype=np.uint8)
    if b2.size != 128 or bs.size != 128:
        raise ZkError("verify_proof: g2 / s_g2 must be 128 uncompressed bytes each (got %d / %d); decode the SRS file's "
                      "fields with g2_decode() or use ParamsKZG.read(...).g2_uncompressed" % (b2.size, bs.size))
    vp.g2, vp.s_g2 = b2.ctypes.data, bs.ctypes.data
    vp.transcript_kind = _abi.TRANSCRIPTS[transcript]
    inst = [_u64(i, (-1, 4)) for i in instances]
    ti, ki = _ptr_table(inst)
    lens = (C.c_uint32 * max(1, len(inst)))(*[i.shape[0] for i in inst])
    pb = np.frombuffer(bytes(proof), dtype=np.uint8)
    ok = C.c_int(0)
    rc = lib.zk_verify_proof(C.byref(cd), C.byref(vp), ti, lens, _ptr(pb) if pb.size else None, pb.size, C.byref(ok))
    if rc != 0:
        raise ZkError("zk_verify_proof: bad arguments (status %d)" % rc)
    return bool(ok.value)


# ---- device-pointer variants (operands already resident in HBM; asynchronous on the context stream) -----
def _vp(x):
    return C.c_void_p(int(x)) if x else None


def commit_dev(params, d_scalars, count, d_out, lagrange=True, n=None):
    """`count` commitments over the resident SRS: d_scalars = count x n Fr (device), d_out = count G1 (device)."""
    ctx = params.ctx
    which = _abi.BASES_G_LAGRANGE if lagrange else _abi.BASES_G
    ctx.check(ctx.lib.zk_msm_batch_dev(ctx.h, _vp(d_scalars), which, n or params.n, count, _vp(d_out)))


def msm_dev(ctx, d_scalars, d_bases, n, d_out):
    """best_multiexp with scalars and bases in device memory."""
    ctx.check(ctx.lib.zk_msm_dev(ctx.h, _vp(d_scalars), _vp(d_bases), _abi.BASES_HOST, n, _vp(d_out)))


def lagrange_to_coeff_dev(domain, d_a, count=1):
    domain.ctx.check(domain.ctx.lib.zk_lagrange_to_coeff_dev(domain.h, _vp(d_a), count))


def coeff_to_extended_dev(domain, d_a, d_out, count=1):
    domain.ctx.check(domain.ctx.lib.zk_coeff_to_extended_dev(domain.h, _vp(d_a), _vp(d_out), count))


def extended_to_coeff_dev(domain, d_a_ext, d_out):
    domain.ctx.check(domain.ctx.lib.zk_extended_to_coeff_dev(domain.h, _vp(d_a_ext), _vp(d_out)))


def divide_by_vanishing_poly_dev(domain, d_a_ext):
    domain.ctx.check(domain.ctx.lib.zk_divide_by_vanishing_poly_dev(domain.h, _vp(d_a_ext)))


def ntt_dev(ctx, d_a, omega, log_n):
    ctx.check(ctx.lib.zk_ntt_dev(ctx.h, _vp(d_a), _ptr(_u64(omega)), log_n))


def evaluate_h_dev(pk, d_advice, d_instance, challenges, y, beta, gamma, theta, d_lookups, d_permutations, d_out):
    """Pointer lists hold DEVICE addresses of coefficient-form polynomials; d_out = device buffer of ext_n Fr."""
    def tab(ptrs):
        return (C.c_void_p * max(1, len(ptrs)))(*[int(p) for p in ptrs])
    ch = _u64(challenges, (-1, 4)) if len(challenges) else np.zeros((1, 4), np.uint64)
    pk.ctx.check(pk.ctx.lib.zk_evaluate_h_dev(pk.h, tab(d_advice), tab(d_instance), _ptr(ch), _ptr(_u64(y)),
                                              _ptr(_u64(beta)), _ptr(_u64(gamma)), _ptr(_u64(theta)),
                                              tab(d_lookups), tab(d_permutations), _vp(d_out)))


def srs_upload_dev(ctx, k, d_g, d_g_lagrange, precompute=True):
    ctx.check(ctx.lib.zk_srs_upload_dev(ctx.h, k, _vp(d_g), _vp(d_g_lagrange), 1 if precompute else 0))


def fixed_base_mul_dev(ctx, base_affine, d_scalars, n, d_out):
    ctx.check(ctx.lib.zk_g1_fixed_base_mul_dev(ctx.h, _ptr(_u64(base_affine)), _vp(d_scalars), n, _vp(d_out)))


class ResidentParams:
    """ParamsKZG whose bases were produced on the device (synthetic SRS for benchmarks)."""

    def __init__(self, ctx, k):
        self.ctx, self.k, self.n = ctx, k, 1 << k
