"""CPU: the oracle's create_proof restatement (oracle/prover.py) is accepted by the oracle's verifier
(oracle/verifier.py, real pairing) -- the acceptance test the reference itself uses
(circuits/src/sgx_dcap_verifier.rs:827-843: verify_proof(..).is_ok()) -- and its building blocks are pinned."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import pairing as pr
from oracle import prover as op
from oracle import pymodel as pm
from oracle import verifier as ov
from tests import prover_common as pc
from tests.test_evalh_host import eval_expr


def test_pairing_is_bilinear_and_nondegenerate():
    assert pr.g2_on_curve(pr.G2_GEN)
    assert pr.g2_mul(pr.R, pr.G2_GEN) is None                      # order r
    e1 = pr.pairing(pr.G2_GEN, (1, 2))
    assert e1 != pr.F12_ONE and pr.f12_pow(e1, pr.R) == pr.F12_ONE
    a, b = 0x1234567, 0x7654321
    assert pr.pairing(pr.g2_mul(b, pr.G2_GEN), pm.ec_mul(a, (1, 2))) == pr.f12_pow(e1, a * b)
    s = pc.TRAPDOOR
    neg = pm.ec_mul(pr.R - s, (1, 2))
    assert pr.pairing_check([((1, 2), pr.g2_mul(s, pr.G2_GEN)), (neg, pr.G2_GEN)])
    assert not pr.pairing_check([((1, 2), pr.g2_mul(s + 1, pr.G2_GEN)), (neg, pr.G2_GEN)])


def test_blake2b_transcript_known_answers():
    import hashlib
    t = op.Blake2bWrite()
    t.common_scalar(5)
    h = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
    h.update(b"\x02" + (5).to_bytes(32, "little") + b"\x00")
    assert t.squeeze_challenge() == int.from_bytes(h.digest(), "little") % pm.R_MOD
    t.write_point((1, 2))
    assert t.finalize() == (1).to_bytes(32, "little")              # y = 2 is even: no sign bit
    assert op.decompress_point(op.compress_point((1, pm.P_MOD - 2))) == (1, pm.P_MOD - 2)


def test_srs_closed_form_matches_the_definition():
    k = 4
    p = pc.params_for(k)
    n, s = 1 << k, pc.TRAPDOOR
    g = pm.limbs_to_affine(p["g"])
    assert g[0] == (1, 2) and g[3] == pm.ec_mul(pow(s, 3, pm.R_MOD), (1, 2))
    # sum_i g_lagrange[i] * f(omega^i) == sum_j g[j] * coeff_j  for a random polynomial
    coeffs = [(7 * j + 3) % pm.R_MOD for j in range(n)]
    w = pm.omega_for(k)
    evals = [pm.poly_eval(coeffs, pow(w, i, pm.R_MOD)) for i in range(n)]
    assert pm.msm(evals, pm.limbs_to_affine(p["g_lagrange"])) == pm.msm(coeffs, g)


def test_prover_tail_loops_against_bigints():
    rnd = np.random.default_rng(5)
    n = 64
    vals = [int(x) for x in rnd.integers(0, 1 << 62, size=n)]
    z = 0x1234567890abcdef
    q = pm.from_mont_limbs(orc.kate_division(pm.to_mont_limbs(vals, pm.R_MOD), op.L(z)), pm.R_MOD)
    # (X - z) * q(X) + a(z) == a(X)
    prod = [0] * n
    for i, c in enumerate(q):
        prod[i + 1] = (prod[i + 1] + c) % pm.R_MOD
        prod[i] = (prod[i] - z * c) % pm.R_MOD
    prod[0] = (prod[0] + pm.poly_eval(vals, z)) % pm.R_MOD
    assert prod == [v % pm.R_MOD for v in vals]
    # permute_expression_pair: sorted input, table permuted so that first occurrences match
    inp = [3, 1, 1, 2, 3, 3, 0, 0]
    tab = [0, 1, 2, 3, 9, 8, 7, 7]
    a, s = orc.permute_expression_pair(pm.to_mont_limbs(inp, pm.R_MOD), pm.to_mont_limbs(tab, pm.R_MOD), 8)
    a, s = pm.from_mont_limbs(a, pm.R_MOD), pm.from_mont_limbs(s, pm.R_MOD)
    assert a == sorted(inp) and sorted(s) == sorted(tab)
    assert all(s[i] == a[i] for i in range(8) if i == 0 or a[i] != a[i - 1])
    # leftovers ascending are handed to the repeated rows from the LAST one backwards (Vec::pop)
    assert [s[i] for i in (1, 3, 6, 7)] == [9, 8, 7, 7]
    with pytest.raises(ValueError):
        orc.permute_expression_pair(pm.to_mont_limbs([4], pm.R_MOD), pm.to_mont_limbs([5], pm.R_MOD), 1)


def test_compress_graph_rows_match_expression_trees():
    case = pc.make_case("toy", 5, 3)
    cs, k = case["cs"], 5
    theta = 0xabcdef1234
    got = orc.eval_graph_rows(case["circ"].lookup_input_graphs[0], k, case["fixed"], case["advice"],
                              [np.zeros((32, 4), np.uint64)], [], op.L(theta), threads=2)
    env = {"Fixed": [pm.from_mont_limbs(f, pm.R_MOD) for f in case["fixed"]],
           "Advice": [pm.from_mont_limbs(a, pm.R_MOD) for a in case["advice"]], "Instance": [[0] * 32],
           "challenges": [], "rot_scale": 1}
    want = []
    for row in range(32):
        acc = 0
        for e in cs.lookups[0][0]:
            acc = (acc * theta + eval_expr(e, row, env)) % pm.R_MOD
        want.append(acc)
    assert pm.from_mont_limbs(got, pm.R_MOD) == want


@pytest.mark.parametrize("shape,k", [("toy", 5), ("toy", 7), ("qe3", 9)])
def test_oracle_proof_verifies(shape, k):
    case = pc.make_case(shape, k, 1)
    pk = pc.oracle_key(case)
    proof = op.create_proof(orc, pc.params_for(k), pk, case["circ"], case["advice"], case["instances_int"], case["randoms"])
    vp = pc.verifier_params()
    assert ov.verify_proof(vp, pk["vk"], case["cs"], k, case["instances_int"], proof)
    # soundness smoke: a flipped bit anywhere must be rejected; so must a wrong public input
    for pos in (0, 40, len(proof) // 2, len(proof) - 1):
        bad = bytearray(proof); bad[pos] ^= 1
        assert not ov.verify_proof(vp, pk["vk"], case["cs"], k, case["instances_int"], bytes(bad))
    if case["instances_int"]:
        wrong = [list(case["instances_int"][0])]
        wrong[0][0] = (wrong[0][0] + 1) % pm.R_MOD
        assert not ov.verify_proof(vp, pk["vk"], case["cs"], k, wrong, proof)


def test_unsatisfied_witness_is_rejected():
    case = pc.make_case("toy", 5, 2)
    pk = pc.oracle_key(case)
    adv = [a.copy() for a in case["advice"]]
    adv[0][3] = adv[0][7]                                          # break the first gate's output cell
    proof = op.create_proof(orc, pc.params_for(5), pk, case["circ"], adv, case["instances_int"], case["randoms"])
    assert not ov.verify_proof(pc.verifier_params(), pk["vk"], case["cs"], 5, case["instances_int"], proof)


# ---- the reference's golden proof pins the halo2curves >= 0.4 point encoding ---------------------------------------
def _golden_points():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_proof_points.json")) as f:
        return json.load(f)


def test_golden_proof_point_encoding_v04():
    from oracle import prover as op
    g = _golden_points()
    chunks = [bytes.fromhex(c) for c in g["chunks"]]
    assert len(chunks) == 47 and g["flagged"] == [1, 2, 3, 4, 45, 46]
    for i in g["flagged"]:
        pt = op.decompress_point_v04(chunks[i])
        assert pt == tuple(int(v, 16) for v in g["points"][str(i)])
        assert (pt[1] * pt[1] - pt[0] ** 3 - 3) % op.P == 0 and pt[1] & 1 == 1
        assert op.compress_point_v04(pt) == chunks[i]
        # the 0.3.1 reader takes bit 6 for part of x and must refuse (x >= p)
        with pytest.raises(ValueError):
            op.decompress_point(chunks[i])
    for i, xy in g["points"].items():
        pt = tuple(int(v, 16) for v in xy)
        assert op.decompress_point_v04(chunks[int(i)]) == pt and op.compress_point_v04(pt) == chunks[int(i)]
    assert op.decompress_point_v04(op.compress_point_v04(None)) is None


def test_fast_setup_matches_the_plain_loop():
    """orc.setup_params (fixed-base comb in C) against op.setup_params (one double-and-add product per point)"""
    for k in (3, 6):
        slow = op.setup_params(orc, k, pc.TRAPDOOR)
        g, gl = orc.setup_params(k, pc.TRAPDOOR)
        assert np.array_equal(g, slow["g"]) and np.array_equal(gl, slow["g_lagrange"])


def test_srs_secret_is_the_zero_seed_chacha20_draw():
    """halo2-base `gen_srs(k)` = `ParamsKZG::setup(k, ChaCha20Rng::from_seed([0; 32]))` ([UPSTREAM?]): the secret is
    `Fr::random` = the first 64 keystream bytes (RFC 7539 block function, zero key / counter / nonce) as a little-endian integer
    mod r.  The tests' and the bench's TRAPDOOR is that value, so their SRS is the one the reference would generate."""
    import struct

    def rotl(v, c):
        return ((v << c) & 0xFFFFFFFF) | (v >> (32 - c))

    def quarter(s, a, b, c, d):
        s[a] = (s[a] + s[b]) & 0xFFFFFFFF; s[d] = rotl(s[d] ^ s[a], 16)
        s[c] = (s[c] + s[d]) & 0xFFFFFFFF; s[b] = rotl(s[b] ^ s[c], 12)
        s[a] = (s[a] + s[b]) & 0xFFFFFFFF; s[d] = rotl(s[d] ^ s[a], 8)
        s[c] = (s[c] + s[d]) & 0xFFFFFFFF; s[b] = rotl(s[b] ^ s[c], 7)

    init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + [0] * 12
    w = list(init)
    for _ in range(10):
        for q in ((0, 4, 8, 12), (1, 5, 9, 13), (2, 6, 10, 14), (3, 7, 11, 15), (0, 5, 10, 15), (1, 6, 11, 12), (2, 7, 8, 13), (3, 4, 9, 14)):
            quarter(w, *q)
    stream = struct.pack("<16I", *[(a + b) & 0xFFFFFFFF for a, b in zip(w, init)])
    assert stream[:16].hex() == "76b8e0ada0f13d90405d6ae55386bd28"          # the published zero-key test vector
    s = int.from_bytes(stream, "little") % pm.R_MOD
    assert s == pc.TRAPDOOR
    import bench
    assert bench.TRAPDOOR == s
    g = pm.limbs_to_affine(pc.params_for(4)["g"][:2])
    assert g[0] == (1, 2) and g[1] == pm.ec_mul(s, (1, 2))
    assert g[1] == (0x0ec379d622d571768ccf88ad66c436878c1f3023a56b611f076f39283681af49,
                    0x1ec57fbfc73af5f417110eaea857cb7854476803c5ba11248d090398017fb020)
