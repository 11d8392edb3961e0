"""evaluate_h: the Python GraphEvaluator (host mirror of Evaluator::new), the word-stream compiler and
the per-row interpreter (csrc/evalh.cuh, run here through its host build) against
  (1) the oracle's interpreter over the same zk_graph (oracle/zk_oracle.c orc_evaluate_h), and
  (2) a direct big-integer evaluation of the Expression trees (the mathematical definition)."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

import zkdcap_b200 as zk
from oracle import pymodel as pm
from importlib import import_module

shapes = import_module("zk-dcap-verifier_b200.shapes")
_abi = zk._abi
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ffh():
    so = os.path.join(HERE, "host", "libffhost.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "host", "ff_host_check.cpp")])
    return C.CDLL(so)


def eval_expr(e, row, env):
    t, a = e.tag, e.args
    R = pm.R_MOD
    if t == "Constant":
        return a[0]
    if t in ("Fixed", "Advice", "Instance"):
        col = env[t][a[0]]
        return col[(row + a[1] * env["rot_scale"]) % len(col)]
    if t == "Challenge":
        return env["challenges"][a[0]]
    if t == "Negated":
        return (-eval_expr(a[0], row, env)) % R
    if t == "Sum":
        return (eval_expr(a[0], row, env) + eval_expr(a[1], row, env)) % R
    if t == "Product":
        return eval_expr(a[0], row, env) * eval_expr(a[1], row, env) % R
    return eval_expr(a[0], row, env) * a[1] % R


def direct_h(cs, k, env, y, beta, gamma, theta, z_sets, lookups, l0, l_last, l_active, sigma):
    """evaluate_h from the definition on canonical integers"""
    R = pm.R_MOD
    d = cs.degree()
    ek = pm.extended_k_for(d, k)
    E = 1 << ek
    rs = 1 << (ek - k)
    env["rot_scale"] = rs
    wx = pm.omega_for(ek)
    bf = cs.blinding_factors()
    chunk = d - 2
    out = []
    for idx in range(E):
        v = 0
        for g in cs.gates:
            v = (v * y + eval_expr(g, idx, env)) % R
        if cs.permutation_columns:
            ns = len(z_sets)
            r_next, r_last = (idx + rs) % E, (idx - (bf + 1) * rs) % E
            v = (v * y + (1 - z_sets[0][idx]) * l0[idx]) % R
            zl = z_sets[-1][idx]
            v = (v * y + (zl * zl - zl) * l_last[idx]) % R
            for s in range(1, ns):
                v = (v * y + (z_sets[s][idx] - z_sets[s - 1][r_last]) * l0[idx]) % R
            cur = beta * pm.ZETA % R * pow(wx, idx, R) % R
            for s in range(ns):
                cols = cs.permutation_columns[s * chunk:(s + 1) * chunk]
                left, right = z_sets[s][r_next], z_sets[s][idx]
                for ci, (ct, ix) in enumerate(cols):
                    val = env[{0: "Advice", 1: "Fixed", 2: "Instance"}[ct]][ix][idx]
                    left = left * (val + beta * sigma[s * chunk + ci][idx] + gamma) % R
                    right = right * (val + cur + gamma) % R
                    cur = cur * pm.DELTA % R
                v = (v * y + (left - right) * l_active[idx]) % R
        for (inputs, tables), (Z, A, S) in zip(cs.lookups, lookups):
            ci = 0
            for e in inputs:
                ci = (ci * theta + eval_expr(e, idx, env)) % R
            ct = 0
            for e in tables:
                ct = (ct * theta + eval_expr(e, idx, env)) % R
            tv = (ci + beta) * (ct + gamma) % R
            r_next, r_prev = (idx + rs) % E, (idx - rs) % E
            ams = (A[idx] - S[idx]) % R
            v = (v * y + (1 - Z[idx]) * l0[idx]) % R
            v = (v * y + (Z[idx] * Z[idx] - Z[idx]) * l_last[idx]) % R
            v = (v * y + (Z[r_next] * (A[idx] + beta) * (S[idx] + gamma) - Z[idx] * tv) * l_active[idx]) % R
            v = (v * y + ams * l0[idx]) % R
            v = (v * y + ams * (A[idx] - A[r_prev]) * l_active[idx]) % R
        out.append(v)
    return out


def make_case(cs, k, seed, orc):
    """random polynomials everywhere (evaluate_h is a polynomial identity evaluator: no satisfying witness needed)"""
    rnd = random.Random(seed)
    n = 1 << k
    dom = orc.Domain(cs.degree(), k, threads=2)
    E = dom.ext_n
    rp = lambda: pm.to_mont_limbs([rnd.randrange(pm.R_MOD) for _ in range(n)], pm.R_MOD)
    sc = shapes.counts(cs)
    case = {
        "advice": [rp() for _ in range(sc["advice"])],
        "instance": [rp() for _ in range(sc["instance"])],
        "fixed_polys": [rp() for _ in range(sc["fixed"])],
        "sigma_polys": [rp() for _ in range(sc["perm_columns"])],
        "z_polys": [rp() for _ in range(sc["perm_sets"])],
        "lookup_polys": [(rp(), rp(), rp()) for _ in range(sc["lookups"])],
        "l0_poly": rp(), "l_last_poly": rp(), "l_blind_poly": rp(),
        "challenges": [rnd.randrange(pm.R_MOD) for _ in range(cs.num_challenges)],
        "y": rnd.randrange(pm.R_MOD), "beta": rnd.randrange(pm.R_MOD), "gamma": rnd.randrange(pm.R_MOD),
        "theta": rnd.randrange(pm.R_MOD),
    }
    ext = dom.coeff_to_extended
    case["fixed_cosets"] = [ext(p) for p in case["fixed_polys"]]
    case["sigma_cosets"] = [ext(p) for p in case["sigma_polys"]]
    case["l0"], case["l_last"] = ext(case["l0_poly"]), ext(case["l_last_poly"])
    one = np.tile(pm.to_mont_limbs([1], pm.R_MOD), (E, 1))
    lb = ext(case["l_blind_poly"])
    case["l_active"] = orc.field_op("fr", "sub", one, orc.field_op("fr", "add", case["l_last"], lb))
    case["dom"] = dom
    return case


def L(v):
    return pm.to_mont_limbs([v], pm.R_MOD)[0]


def oracle_h(orc, cs, k, case):
    ev = zk.Evaluator.new(cs)
    desc, keep = ev.describe(cs, k)
    ch = pm.to_mont_limbs(case["challenges"], pm.R_MOD) if case["challenges"] else []
    flat = [p for t in case["lookup_polys"] for p in t]
    return orc.evaluate_h(desc, case["fixed_cosets"], case["sigma_cosets"], case["l0"], case["l_last"], case["l_active"],
                          case["advice"], case["instance"], ch, L(case["y"]), L(case["beta"]), L(case["gamma"]),
                          L(case["theta"]), flat, case["z_polys"], case["dom"].ext_n, threads=4), desc, keep


def host_interp_h(ffh, cs, k, case, desc):
    dom = case["dom"]
    ext = dom.coeff_to_extended
    cols = list(case["fixed_cosets"]) + [ext(p) for p in case["advice"]] + [ext(p) for p in case["instance"]]
    cols += list(case["sigma_cosets"]) + [ext(p) for p in case["z_polys"]]
    cols += [ext(p) for t in case["lookup_polys"] for p in t]
    cols += [case["l0"], case["l_last"], case["l_active"]]
    cols = [np.ascontiguousarray(c) for c in cols]
    tab = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    ch = pm.to_mont_limbs(case["challenges"], pm.R_MOD) if case["challenges"] else np.zeros((1, 4), np.uint64)
    out = np.zeros((dom.ext_n, 4), dtype=np.uint64)
    stats = (C.c_ulonglong * 4)()
    err = C.create_string_buffer(256)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = ffh.evh_run(C.byref(desc), tab, p(ch), p(L(case["y"])), p(L(case["beta"])), p(L(case["gamma"])),
                     p(L(case["theta"])), p(np.ascontiguousarray(dom.extended_omega)), p(out), stats, err)
    assert rc == 0, err.value
    assert stats[2] == len(cols)
    return out, list(stats)


def test_toy_shape_three_way(ffh, orc):
    cs, k = shapes.toy_shape(), 3
    case = make_case(cs, k, 11, orc)
    got_orc, desc, keep = oracle_h(orc, cs, k, case)
    got_dev, stats = host_interp_h(ffh, cs, k, case, desc)
    assert np.array_equal(got_orc, got_dev)
    # independent: straight from the expression trees on integers
    dom = case["dom"]
    ints = lambda a: pm.from_mont_limbs(a, pm.R_MOD)
    env = {"Fixed": [ints(c) for c in case["fixed_cosets"]],
           "Advice": [ints(dom.coeff_to_extended(p)) for p in case["advice"]],
           "Instance": [ints(dom.coeff_to_extended(p)) for p in case["instance"]],
           "challenges": case["challenges"]}
    want = direct_h(cs, k, env, case["y"], case["beta"], case["gamma"], case["theta"],
                    [ints(dom.coeff_to_extended(p)) for p in case["z_polys"]],
                    [tuple(ints(dom.coeff_to_extended(p)) for p in t) for t in case["lookup_polys"]],
                    ints(case["l0"]), ints(case["l_last"]), ints(case["l_active"]),
                    [ints(c) for c in case["sigma_cosets"]])
    assert ints(got_orc) == want


def test_qe3_shape_counts_and_interpreters_agree(ffh, orc):
    cs = shapes.qe3_shape()
    c = shapes.counts(cs)
    # A ~ 28, L ~ 11, fixed ~ 21, 18 equality columns, degree 5 -> 6 permutation sets, bf = 6
    assert (c["advice"], c["fixed"], c["lookups"], c["perm_columns"]) == (28, 21, 11, 18)
    assert c["degree"] == 5 and c["perm_sets"] == 6 and c["blinding_factors"] == 6
    calls = shapes.hot_path_calls(cs)
    assert calls["msm_lagrange"] + calls["msm_coeff"] == 74 and calls["intt_n"] == 67
    k = 4
    case = make_case(cs, k, 5, orc)
    got_orc, desc, keep = oracle_h(orc, cs, k, case)
    got_dev, stats = host_interp_h(ffh, cs, k, case, desc)
    assert np.array_equal(got_orc, got_dev)
    assert stats[1] >= 100     # ~109 extended columns read per row
    assert stats[0] > 200      # field multiplications per row
    # the program optimiser (store forwarding, Horner splitting, slot allocation) is value preserving and
    # shrinks the per-row state from ~70 intermediates to a handful
    ffh.evh_set_optimise(0)
    got_raw, stats_raw = host_interp_h(ffh, cs, k, case, desc)
    ffh.evh_set_optimise(1)
    assert np.array_equal(got_raw, got_orc)
    assert stats_raw[0] == stats[0] and stats[3] <= 16 < stats_raw[3]


def test_graph_evaluator_dedup_rules():
    E = zk.Expression
    g = zk.GraphEvaluator()
    a, b = E.Advice(0), E.Advice(1, 1)
    r1 = g.add_expression(a * b)
    r2 = g.add_expression(b * a)            # operands are ordered -> same calculation
    assert r1 == r2
    assert g.add_expression(a * a) == g.add_expression(a * a)
    assert g.calculations[-1][0][0] == "Square"
    assert g.add_expression(a * E.Constant(1)) == g.add_expression(a)
    assert g.add_expression(a * E.Constant(0)) == (_abi.VS_CONSTANT, 0, 0)
    assert g.add_expression(a + E.Constant(0)) == g.add_expression(a)
    n = len(g.calculations)
    g.add_expression(a * E.Constant(2))
    assert g.calculations[-1][0][0] == "Double" and len(g.calculations) == n + 1
    g.add_expression(a - b)
    assert g.calculations[-1][0][0] == "Sub"
    assert g.add_expression(-E.Constant(5)) == (_abi.VS_CONSTANT, g.constants.index(pm.R_MOD - 5), 0)
    assert g.constants[:3] == [0, 1, 2] and g.rotations == [0, 1]


# ---- evaluate_h as GENERATED code (csrc/evalh_codegen.hpp) ----------------------------------------------------------------------
def generated_source(desc, host):
    lib = zk._abi.load()
    ln = C.c_size_t()
    assert lib.zk_evalh_generate_source(C.byref(desc), 1 if host else 0, None, 0, C.byref(ln)) == 0
    buf = C.create_string_buffer(ln.value + 1)
    assert lib.zk_evalh_generate_source(C.byref(desc), 1 if host else 0, buf, ln.value + 1, C.byref(ln)) == 0
    return buf.value.decode()


def host_generated_h(cs, k, case, desc, tmp_path):
    """compile the host flavour of the generated text with g++ and run every extended row"""
    import subprocess
    src = generated_source(desc, host=True)
    cpath, so = tmp_path / "evalh_gen.cpp", tmp_path / "evalh_gen.so"
    cpath.write_text(src)
    inc = os.path.join(os.path.dirname(HERE), "zk-dcap-verifier_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-I", inc, "-o", str(so), str(cpath)])
    gen = C.CDLL(str(so))
    dom = case["dom"]
    ext = dom.coeff_to_extended
    cols = list(case["fixed_cosets"]) + [ext(p) for p in case["advice"]] + [ext(p) for p in case["instance"]]
    cols += list(case["sigma_cosets"]) + [ext(p) for p in case["z_polys"]]
    cols += [ext(p) for t in case["lookup_polys"] for p in t]
    cols += [case["l0"], case["l_last"], case["l_active"]]
    cols = [np.ascontiguousarray(c) for c in cols]
    tab = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    ch = pm.to_mont_limbs(case["challenges"], pm.R_MOD) if case["challenges"] else np.zeros((1, 4), np.uint64)
    beta = case["beta"]
    sc = pm.to_mont_limbs([beta, case["gamma"], case["theta"], case["y"], beta * pm.ZETA % pm.R_MOD, pm.DELTA, 1], pm.R_MOD)
    out = np.zeros((dom.ext_n, 4), dtype=np.uint64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    gen.zk_evalh_gen_host(tab, p(ch), p(np.ascontiguousarray(sc)), p(np.ascontiguousarray(dom.extended_omega)), p(out),
                          C.c_size_t(0), C.c_size_t(dom.ext_n))
    return out, src


@pytest.mark.parametrize("shape,k", [("toy_shape", 3), ("qe3_shape", 4)])
def test_generated_evaluator_equals_oracle_on_host(orc, tmp_path, shape, k):
    """the text zk_pk_upload hands to NVRTC, compiled for the CPU instead: every row equals the oracle's evaluate_h"""
    cs = getattr(shapes, shape)()
    case = make_case(cs, k, 21, orc)
    want, desc, keep = oracle_h(orc, cs, k, case)
    got, src = host_generated_h(cs, k, case, desc, tmp_path)
    assert np.array_equal(got, want)
    assert "ev_fma" in src and "zk_evalh_gen_host" in src


def test_generated_evaluator_compiles_for_sm_100a():
    """NVRTC (libnvrtc.so.12 of the toolkit; no GPU needed) accepts the device flavour for sm_100a"""
    try:
        nv = C.CDLL("libnvrtc.so.12")
    except OSError:
        pytest.skip("libnvrtc.so.12 not loadable")
    cs = shapes.qe3_shape()
    desc, keep = zk.Evaluator.new(cs).describe(cs, 10)
    src = generated_source(desc, host=False).encode()
    prog = C.c_void_p()
    assert nv.nvrtcCreateProgram(C.byref(prog), src, b"zk_evalh_gen.cu", 0, None, None) == 0
    opts = (C.c_char_p * 3)(b"--gpu-architecture=sm_100a", b"-std=c++17", b"-lineinfo")
    rc = nv.nvrtcCompileProgram(prog, 3, opts)
    if rc != 0:
        ls = C.c_size_t()
        nv.nvrtcGetProgramLogSize(prog, C.byref(ls))
        log = C.create_string_buffer(ls.value + 1)
        nv.nvrtcGetProgramLog(prog, log)
        raise AssertionError(log.value.decode()[-3000:])
    cs_ = C.c_size_t()
    assert nv.nvrtcGetCUBINSize(prog, C.byref(cs_)) == 0 and cs_.value > 10000
    cubin = C.create_string_buffer(cs_.value)
    assert nv.nvrtcGetCUBIN(prog, cubin) == 0
    assert b"zk_evalh_gen" in cubin.raw
