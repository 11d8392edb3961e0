"""
verifier.py -- TEST INFRASTRUCTURE: CPU restatement of halo2's KZG/SHPLONK verifier.

PARITY UNPINNED (see oracle/zk_oracle.c): restated from halo2_proofs 0.2.0 @ zkwebauthn/halo2 c254c75f
(Cargo.lock:1314-1316) -- plonk/verifier.rs `verify_proof`, plonk/{vanishing,permutation,lookup}/verifier.rs,
poly/kzg/multiopen/shplonk/verifier.rs, poly/kzg/strategy.rs (SingleStrategy -> DualMSM::check), whose sources
are not vendored.  Reference call sites: circuits/src/sgx_dcap_verifier.rs:827-843 (verify_proof with
VerifierSHPLONK / Blake2bRead / SingleStrategy), bin/src/main.rs:244-251.

It is the acceptance test of the provers in this repository: the device prover's bytes and the oracle prover's
bytes must be accepted here.  All arithmetic is on Python integers (naive, independent of the C oracle); the final
check is a real pairing (oracle/pairing.py) against [s]G2.

Only tests/, __graft_entry__.smoke() and bench.py's checking leg may import this module.
"""
from . import pairing
from . import pymodel as pm
from .prover import (Blake2bRead, EvmRead, PoseidonRead, construct_intermediate_sets, eval_small, evaluate_vanishing_polynomial,
                     lagrange_interpolate, rotate_omega)

R = pm.R_MOD


def eval_expression(e, fixed_evals, advice_evals, instance_evals, challenges, cs):
    """Expression::evaluate with the query-index lookups of plonk/verifier.rs"""
    t, a = e.tag, e.args
    if t == "Constant":
        return a[0] % R
    if t == "Fixed":
        return fixed_evals[cs.fixed_queries.index((a[0], a[1]))]
    if t == "Advice":
        return advice_evals[cs.advice_queries.index((a[0], a[1]))]
    if t == "Instance":
        return instance_evals[cs.instance_queries.index((a[0], a[1]))]
    if t == "Challenge":
        return challenges[a[0]]
    if t == "Negated":
        return (-eval_expression(a[0], fixed_evals, advice_evals, instance_evals, challenges, cs)) % R
    if t == "Sum":
        return (eval_expression(a[0], fixed_evals, advice_evals, instance_evals, challenges, cs) +
                eval_expression(a[1], fixed_evals, advice_evals, instance_evals, challenges, cs)) % R
    if t == "Product":
        return (eval_expression(a[0], fixed_evals, advice_evals, instance_evals, challenges, cs) *
                eval_expression(a[1], fixed_evals, advice_evals, instance_evals, challenges, cs)) % R
    if t == "Scaled":
        return eval_expression(a[0], fixed_evals, advice_evals, instance_evals, challenges, cs) * a[1] % R
    raise ValueError(t)


def l_i_range(k, x, xn, rotations):
    """EvaluationDomain::l_i_range: l_i(x) = omega^i (x^n - 1) / (n (x - omega^i))"""
    n = 1 << k
    common = (xn - 1) * pow(n, -1, R) % R
    out = []
    for r in rotations:
        wi = rotate_omega(k, 1, r)
        out.append(pow((x - wi) % R, -1, R) * common % R * wi % R)
    return out


def msm_eval(terms):
    """sum scalar * point over a list of (scalar, affine point)"""
    acc = None
    for s, p in terms:
        acc = pm.ec_add(acc, pm.ec_mul(s % R, p))
    return acc


def verify_proof(params, vk, cs, k, instances, proof, transcript="blake2b"):
    """params: {"g0": G1 generator used by the SRS (affine ints), "g2": G2, "s_g2": [s]G2}.
    vk: {"fixed_commitments", "permutation_commitments", "transcript_repr"}.  transcript: "blake2b" (Blake2bRead,
    Challenge255), "evm" (snark-verifier EvmTranscript, Keccak-256) or "poseidon" (snark-verifier PoseidonTranscript, the
    CLI's verifier at bin/src/main.rs:239-252).  Returns True / False."""
    try:
        return _verify(params, vk, cs, k, instances, proof, transcript)
    except (ValueError, KeyError, ZeroDivisionError):
        return False


def _verify(params, vk, cs, k, instances, proof, transcript="blake2b"):
    n, bf, d = 1 << k, cs.blinding_factors(), cs.degree()
    chunk = d - 2
    tr = {"blake2b": Blake2bRead, "evm": EvmRead, "poseidon": PoseidonRead}[transcript](proof)
    if len(instances) != cs.num_instance_columns:
        return False
    tr.common_scalar(vk["transcript_repr"])
    for col in instances:
        if len(col) > n - (bf + 1):
            return False
        for v in col:
            tr.common_scalar(v % R)
    advice_c = [tr.read_point() for _ in range(cs.num_advice_columns)]
    challenges = [tr.squeeze_challenge() for _ in range(cs.num_challenges)]
    theta = tr.squeeze_challenge()
    lookups = [{"pin_c": tr.read_point(), "ptab_c": tr.read_point()} for _ in cs.lookups]
    beta = tr.squeeze_challenge()
    gamma = tr.squeeze_challenge()
    npc = len(cs.permutation_columns)
    n_sets = (npc + chunk - 1) // chunk if npc else 0
    perm_c = [tr.read_point() for _ in range(n_sets)]
    for lk in lookups:
        lk["z_c"] = tr.read_point()
    random_c = tr.read_point()
    y = tr.squeeze_challenge()
    h_c = [tr.read_point() for _ in range(d - 1)]
    x = tr.squeeze_challenge()
    xn = pow(x, n, R)
    # instance evaluations are computed by the verifier (KZG: QUERY_INSTANCE == false)
    instance_evals = []
    if cs.instance_queries:
        rots = [r for _, r in cs.instance_queries]
        min_rot, max_rot = min(min(rots), 0), max(max(rots), 0)
        max_len = max([len(c) for c in instances] + [0])
        l_i_s = l_i_range(k, x, xn, range(-max_rot, max_len + abs(min_rot)))
        for col, r in cs.instance_queries:
            vals = instances[col]
            off = max_rot - r
            instance_evals.append(sum(v * l for v, l in zip(vals, l_i_s[off:off + len(vals)])) % R)
    advice_evals = [tr.read_scalar() for _ in cs.advice_queries]
    fixed_evals = [tr.read_scalar() for _ in cs.fixed_queries]
    random_eval = tr.read_scalar()
    sigma_evals = [tr.read_scalar() for _ in range(npc)]
    perm_evals = []
    for si in range(n_sets):
        pe = {"x": tr.read_scalar(), "next": tr.read_scalar()}
        if si + 1 < n_sets:
            pe["last"] = tr.read_scalar()
        perm_evals.append(pe)
    for lk in lookups:
        lk["z"], lk["z_next"], lk["pin"], lk["pin_inv"], lk["ptab"] = (tr.read_scalar() for _ in range(5))

    # ---- vanishing argument: expected h(x) ---------------------------------------------------------
    l_evals = l_i_range(k, x, xn, range(-(bf + 1), 1))
    assert len(l_evals) == 2 + bf
    l_last, l_blind, l_0 = l_evals[0], sum(l_evals[1:1 + bf]) % R, l_evals[1 + bf]
    active = (1 - (l_last + l_blind)) % R
    ee = lambda e: eval_expression(e, fixed_evals, advice_evals, instance_evals, challenges, cs)
    exprs = [ee(g) for g in cs.gates]

    def any_eval(ct, ix):
        return (advice_evals[cs.advice_queries.index((ix, 0))] if ct == 0 else
                fixed_evals[cs.fixed_queries.index((ix, 0))] if ct == 1 else
                instance_evals[cs.instance_queries.index((ix, 0))])
    if n_sets:
        exprs.append(l_0 * (1 - perm_evals[0]["x"]) % R)
        zl = perm_evals[-1]["x"]
        exprs.append((zl * zl - zl) * l_last % R)
        for si in range(1, n_sets):
            exprs.append((perm_evals[si]["x"] - perm_evals[si - 1]["last"]) * l_0 % R)
        for si in range(n_sets):
            cols = cs.permutation_columns[si * chunk:(si + 1) * chunk]
            sig = sigma_evals[si * chunk:(si + 1) * chunk]
            left = perm_evals[si]["next"]
            for (ct, ix), se in zip(cols, sig):
                left = left * (any_eval(ct, ix) + beta * se + gamma) % R
            right = perm_evals[si]["x"]
            cur = beta * x % R * pow(pm.DELTA, si * chunk, R) % R
            for ct, ix in cols:
                right = right * (any_eval(ct, ix) + cur + gamma) % R
                cur = cur * pm.DELTA % R
            exprs.append((left - right) * active % R)
    for (inputs, tables), lk in zip(cs.lookups, lookups):
        ci = 0
        for e in inputs:
            ci = (ci * theta + ee(e)) % R
        ct_ = 0
        for e in tables:
            ct_ = (ct_ * theta + ee(e)) % R
        left = lk["z_next"] * (lk["pin"] + beta) % R * (lk["ptab"] + gamma) % R
        right = lk["z"] * (ci + beta) % R * (ct_ + gamma) % R
        exprs.append(l_0 * (1 - lk["z"]) % R)
        exprs.append(l_last * (lk["z"] * lk["z"] - lk["z"]) % R)
        exprs.append((left - right) * active % R)
        exprs.append(l_0 * (lk["pin"] - lk["ptab"]) % R)
        exprs.append((lk["pin"] - lk["ptab"]) * (lk["pin"] - lk["pin_inv"]) % R * active % R)
    expected_h = 0
    for v in exprs:
        expected_h = (expected_h * y + v) % R
    expected_h = expected_h * pow((xn - 1) % R, -1, R) % R
    h_commitment = [(pow(xn, i, R), c) for i, c in enumerate(h_c)]        # an MSM, kept symbolic

    # ---- queries (commitments are identified by their position, like the pointer equality upstream) ---
    rot = lambda r: rotate_omega(k, x, r)
    x_next, x_last, x_inv = rot(1), rot(-(bf + 1)), rot(-1)
    comm = {}
    queries = []

    def q(key, terms, pt, ev):
        comm[key] = terms
        queries.append((key, pt, ev))
    for qi, (c, r) in enumerate(cs.advice_queries):
        q(("advice", c), [(1, advice_c[c])], rot(r), advice_evals[qi])
    for si in range(n_sets):
        q(("perm", si), [(1, perm_c[si])], x, perm_evals[si]["x"])
        q(("perm", si), [(1, perm_c[si])], x_next, perm_evals[si]["next"])
    for si in range(n_sets - 2, -1, -1):
        q(("perm", si), [(1, perm_c[si])], x_last, perm_evals[si]["last"])
    for li, lk in enumerate(lookups):
        q(("lz", li), [(1, lk["z_c"])], x, lk["z"])
        q(("lpin", li), [(1, lk["pin_c"])], x, lk["pin"])
        q(("lptab", li), [(1, lk["ptab_c"])], x, lk["ptab"])
        q(("lpin", li), [(1, lk["pin_c"])], x_inv, lk["pin_inv"])
        q(("lz", li), [(1, lk["z_c"])], x_next, lk["z_next"])
    for qi, (c, r) in enumerate(cs.fixed_queries):
        q(("fixed", c), [(1, vk["fixed_commitments"][c])], rot(r), fixed_evals[qi])
    for i in range(npc):
        q(("sigma", i), [(1, vk["permutation_commitments"][i])], x, sigma_evals[i])
    q(("h",), h_commitment, x, expected_h)
    q(("random",), [(1, random_c)], x, random_eval)

    # ---- VerifierSHPLONK::verify_proof ------------------------------------------------------------------
    sets, super_points = construct_intermediate_sets(queries)
    yy = tr.squeeze_challenge()
    v = tr.squeeze_challenge()
    h1 = tr.read_point()
    u = tr.squeeze_challenge()
    h2 = tr.read_point()
    if tr.pos != (len(tr.buf) if transcript == "blake2b" else tr.buf_len):
        return False
    outer, r_outer = [], 0
    z_0 = z_0_diff_inv = 0
    for i, (pts, comms) in enumerate(sets):
        z_diff = evaluate_vanishing_polynomial([p for p in super_points if p not in pts], u)
        if i == 0:
            z_0 = evaluate_vanishing_polynomial(pts, u)
            z_0_diff_inv = pow(z_diff, -1, R)
            z_diff = 1
        else:
            z_diff = z_diff * z_0_diff_inv % R
        pv = pow(v, i, R)
        r_inner = 0
        for j, (key, evs) in enumerate(comms):
            py = pow(yy, j, R)
            r_x = lagrange_interpolate(pts, evs)
            r_inner = (r_inner + py * eval_small(r_x, u)) % R
            outer += [(s * py % R * pv % R * z_diff % R, c) for s, c in comm[key]]
        r_outer = (r_outer + pv * r_inner % R * z_diff) % R
    outer.append(((-r_outer) % R, params["g0"]))
    outer.append(((-z_0) % R, h1))
    outer.append((u, h2))
    right = msm_eval(outer)
    # DualMSM::check: e(left, [s]G2) * e(-right, G2) == 1
    neg_right = None if right is None else (right[0], (-right[1]) % pm.P_MOD)
    return pairing.pairing_check([(h2, params["s_g2"]), (neg_right, params["g2"])])
