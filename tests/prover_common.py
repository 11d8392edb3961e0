"""Shared fixtures of the prover tests: synthetic circuits, SRS with a known trapdoor, oracle keys."""
from importlib import import_module

import numpy as np

import zkdcap_b200 as zk
from oracle import oracle as orc
from oracle import pairing
from oracle import prover as op
from oracle import pymodel as pm

shapes = import_module("zk-dcap-verifier_b200.shapes")

# s = Fr::random of ChaCha20Rng::from_seed([0; 32]); any value would do for the tests
TRAPDOOR = 0x1c59a59b6cff4308740943526ade1d8c09f71b337a67269cc89586bcdd6dfcba


class OrcOps:
    """field ops for the synthetic witness generator, backed by the CPU oracle"""
    mul = staticmethod(lambda a, b: orc.field_op("fr", "mul", a, b))
    add = staticmethod(lambda a, b: orc.field_op("fr", "add", a, b))
    to_mont = staticmethod(lambda a: orc.field_op("fr", "from_raw", a))


_cache = {}


def params_for(k):
    """ParamsKZG::setup with the tests' trapdoor: the C oracle's fixed-base comb (checked against the plain
    double-and-add loop op.setup_params in tests/test_prover_oracle.py)."""
    if ("params", k) not in _cache:
        g, gl = orc.setup_params(k, TRAPDOOR)
        _cache[("params", k)] = {"k": k, "n": 1 << k, "g": g, "g_lagrange": gl, "s": TRAPDOOR}
    return _cache[("params", k)]


def verifier_params():
    if "vp" not in _cache:
        _cache["vp"] = {"g0": (1, 2), "g2": pairing.G2_GEN, "s_g2": pairing.g2_mul(TRAPDOOR, pairing.G2_GEN)}
    return _cache["vp"]


def make_case(shape, k, seed):
    """-> dict(cs, circ, fixed, advice, mapping, instances_limbs, instances_int, randoms)"""
    if shape == "qe3":
        cs = shapes.qe3_shape()
        fixed, advice, mapping = shapes.synth_qe3_witness(cs, k, seed, OrcOps)
        inst = []
    else:
        cs = shapes.toy_prover_shape()
        fixed, advice, mapping, inst = shapes.synth_toy_witness(cs, k, seed, OrcOps)
    circ = zk.Circuit(cs, k, transcript_repr=0xDCA9000 + seed)
    randoms = orc.gen_scalars(1000 + seed, op.randoms_needed(cs, k))
    return {"cs": cs, "k": k, "circ": circ, "fixed": fixed, "advice": advice, "mapping": mapping,
            "instances_limbs": inst, "instances_int": [pm.from_mont_limbs(i, pm.R_MOD) for i in inst], "randoms": randoms}


def oracle_key(case):
    return op.keygen(orc, params_for(case["k"]), case["circ"], case["fixed"], case["mapping"], case["circ"].transcript_repr)


# ---- random SATISFIED circuits -------------------------------------------------------------------------------------------
def _random_input_expression(rng, E, n_in, ni, const_col, depth, max_degree):
    """random expression over the input advice columns, the instance columns and one fixed column (rotations -2..2)"""
    def leaf():
        kind = ["advice", "advice", "advice", "fixed", "const"][rng.integers(0, 5)] if not ni else \
            ["advice", "advice", "advice", "fixed", "const", "instance"][rng.integers(0, 6)]
        if kind == "advice":
            return E.Advice(int(rng.integers(0, n_in)), int(rng.integers(-2, 3))), 1
        if kind == "fixed":
            return E.Fixed(const_col, int(rng.integers(-2, 3))), 1
        if kind == "instance":
            return E.Instance(int(rng.integers(0, ni)), int(rng.integers(-1, 2))), 1
        return E.Constant(int(rng.integers(0, 1 << 62))), 0

    def build(d):
        if d == 0 or rng.integers(0, 4) == 0:
            return leaf()
        kind = rng.integers(0, 5)
        a, da = build(d - 1)
        if kind == 0:
            return -a, da
        if kind == 1:
            return a * int(rng.integers(1, 1 << 30)), da
        b, db = build(d - 1)
        if kind == 2 and da + db <= max_degree:
            return a * b, da + db
        return (a - b, max(da, db)) if kind == 3 else (a + b, max(da, db))
    return build(depth)


def random_satisfied_case(seed):
    """A random circuit WITH a satisfying witness (same dict as make_case).  Fixed columns: 0 lookup selector, 1..G gate
    selectors, then two table columns and one column of constants.  Advice: input columns, then one OUTPUT column per gate
    polynomial `q_g * (out - e)` with `e` a random expression over the inputs (rotations -2..2), the instance column and the
    constants; the outputs are computed row by row from the definition.  0-2 lookups of `q * input` into the tables, copy
    constraints between input cells, instance cells and constant cells."""
    from tests.test_evalh_host import eval_expr
    rng = np.random.default_rng(seed)
    E = zk.Expression
    R = pm.R_MOD
    n_in, n_gates, ni = int(rng.integers(2, 5)), int(rng.integers(1, 4)), int(rng.integers(0, 2))
    polys_per_gate = [int(rng.integers(1, 3)) for _ in range(n_gates)]
    n_out = sum(polys_per_gate)
    T0, T1, CONST = 1 + n_gates, 2 + n_gates, 3 + n_gates
    cs = zk.ConstraintSystem(num_fixed=4 + n_gates, num_advice=n_in + n_out, num_instance=ni)
    max_degree = int(rng.integers(1, 7))
    exprs, o = [], n_in
    for g in range(n_gates):
        polys = []
        for _ in range(polys_per_gate[g]):
            e, _d = _random_input_expression(rng, E, n_in, ni, CONST, 3, max_degree)
            exprs.append((g, o, e))
            polys.append(E.Fixed(1 + g) * (E.Advice(o) - e))
            o += 1
        cs.create_gate(polys)
    lookups = []
    n_lk = int(rng.integers(0, 3))
    if n_lk >= 1:
        a_cols = [0, 1] if rng.integers(0, 2) else [0]
        lookups.append((a_cols, [T0, T1][:len(a_cols)]))
    if n_lk >= 2 and n_in >= 3:
        lookups.append(([n_in - 1], [T1]))
    for a_cols, t_cols in lookups:
        cs.lookup([E.Fixed(0) * E.Advice(a) for a in a_cols], [E.Fixed(t) for t in t_cols])
    eq_cols = [(0, c) for c in range(n_in) if rng.integers(0, 3)]
    if ni and rng.integers(0, 2):
        eq_cols.append((2, 0))
    if rng.integers(0, 2):
        eq_cols.append((1, CONST))
    for c in eq_cols:
        cs.enable_equality(*c)

    k = int(rng.integers(5, 8))
    n = 1 << k
    usable = n - (cs.blinding_factors() + 1)
    assert usable >= 12
    rand_col = lambda: [int.from_bytes(rng.bytes(40), "little") % R for _ in range(n)]
    fixed = [[0] * n for _ in range(cs.num_fixed_columns)]
    advice = [rand_col() for _ in range(n_in)] + [[0] * n for _ in range(n_out)]
    inst = [[int(v) for v in rng.integers(1, 1 << 62, size=int(rng.integers(1, min(usable, 9))))] for _ in range(ni)]
    inst_cols = [v + [0] * (n - len(v)) for v in inst]
    fixed[CONST] = rand_col()
    tab_rows = int(rng.integers(2, usable // 2))
    for t in (T0, T1):
        col = rand_col()
        fixed[t] = [0] + col[1:tab_rows] + [0] * (n - tab_rows)
    sel = [int(rng.integers(0, 2)) if r < usable else 0 for r in range(n)]
    fixed[0] = sel
    for a_cols, t_cols in lookups:
        for r in range(usable):
            if sel[r]:
                row = int(rng.integers(0, tab_rows))
                for a, t in zip(a_cols, t_cols):
                    advice[a][r] = fixed[t][row]
    # copy constraints: 2-cycles between cells whose values are made equal (only rows the lookups do not read)
    pcol = {col: i for i, col in enumerate(cs.permutation_columns)}
    mapping = np.empty((len(cs.permutation_columns), n, 2), dtype=np.uint32)
    mapping[:, :, 0] = np.arange(len(cs.permutation_columns), dtype=np.uint32)[:, None]
    mapping[:, :, 1] = np.arange(n, dtype=np.uint32)[None, :]
    adv_eq = [c for (t, c) in cs.permutation_columns if t == 0]
    free = [(c, r) for c in adv_eq for r in range(usable) if not sel[r]]
    free = [free[i] for i in rng.permutation(len(free))]

    def tie(a, b):
        mapping[pcol[a[0]], a[1]] = (pcol[b[0]], b[1])
        mapping[pcol[b[0]], b[1]] = (pcol[a[0]], a[1])
    for _ in range(int(rng.integers(0, max(1, len(free) // 3)))):
        if len(free) < 2:
            break
        a, b = free.pop(), free.pop()
        advice[b[0]][b[1]] = advice[a[0]][a[1]]
        tie(((0, a[0]), a[1]), ((0, b[0]), b[1]))
    if (2, 0) in pcol and free:
        for j in range(len(inst[0])):
            if not free:
                break
            c = free.pop()
            advice[c[0]][c[1]] = inst[0][j]
            tie(((2, 0), j), ((0, c[0]), c[1]))
    if (1, CONST) in pcol:
        for j in rng.permutation(usable)[:4]:
            if not free:
                break
            c = free.pop()
            advice[c[0]][c[1]] = fixed[CONST][int(j)]
            tie(((1, CONST), int(j)), ((0, c[0]), c[1]))
    # gate selectors on rows whose rotated reads stay inside the usable rows; outputs from the definition
    env = {"Fixed": fixed, "Advice": advice, "Instance": inst_cols, "rot_scale": 1}
    for g in range(n_gates):
        fixed[1 + g] = [int(rng.integers(0, 2)) if 2 <= r < usable - 2 else 0 for r in range(n)]
    for g, o, e in exprs:
        for r in range(2, usable - 2):
            if fixed[1 + g][r]:
                advice[o][r] = eval_expr(e, r, env)
            else:
                advice[o][r] = int.from_bytes(rng.bytes(40), "little") % R
    to_limbs = lambda col: pm.to_mont_limbs(col, R)
    circ = zk.Circuit(cs, k, transcript_repr=0xF00D000 + seed)
    return {"cs": cs, "k": k, "circ": circ, "fixed": [to_limbs(c) for c in fixed], "advice": [to_limbs(c) for c in advice],
            "mapping": mapping, "instances_limbs": [to_limbs(v) for v in inst], "instances_int": inst,
            "randoms": orc.gen_scalars(2000 + seed, op.randoms_needed(cs, k)),
            "first_output": n_in, "first_gate_rows": [r for r in range(n) if fixed[1][r]], "lookups": lookups,
            "lookup_rows": [r for r in range(usable) if sel[r]]}
