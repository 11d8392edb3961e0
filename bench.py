#!/usr/bin/env python
"""
bench.py -- hot path of one DCAP attestation proof on B200 (see DESIGN.md "Measurement").

A "step" is every MSM / NTT / h(X) call that one halo2 `create_proof` makes for the full
certificate-chain circuit shape (BASELINE.json configs[1]; SURVEY.md 3.2 / 8d: the QE3 column
set at k = 19, extended domain 2^21), on synthetic witness columns:

    74 MSM(2^19)  =  A + 3L + P commit_lagrange  +  1 + (d-1) + 2 commit
    67 iNTT(2^19) =  lagrange_to_coeff of the advice / lookup / permutation columns
    67 coset-NTT(2^19 -> 2^21) + the fused per-row evaluate_h over 2^21 rows
     1 divide_by_vanishing_poly + extended_to_coeff (iNTT 2^21)

  value : steps/s with the Lagrange-form witness columns already resident in HBM
  e2e   : the same calls through the host-pointer C ABI (the drop-in boundary behind
          best_multiexp / best_fft / EvaluationDomain / evaluate_h), host<->device copies timed
  --impl reference : the CPU oracle (a restatement of halo2's own algorithms; the Rust prover
          cannot be built here) on all host cores, on a bounded sample of the same workload.

N > 1 (torchrun): independent proofs per GPU (BASELINE configs[4]), no data-path collective;
`--mode msm-sharded` instead shards ONE MSM by base range with a partial-point exchange (NCCL
all_gather of 96-byte points + zk_g1_sum).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "DCAP full-chain proof (halo2 create_proof, KZG/SHPLONK, Blake2b transcript), proofs/s"
METRIC_HOTPATH = "DCAP full-chain proof hot path (MSM + NTT + h(X) of one create_proof), proofs/s"
DTYPE = "u32 limbs (254-bit modular)"
DATA = ("synthetic: a SATISFYING witness of a hand-restated circuit SHAPE (QE3 column / gate / lookup counts of SURVEY 8d at k rows), "
        "NOT the reference circuit's real witness -- the reference's circuits cannot be synthesised here (no Rust toolchain); "
        "SRS from a known trapdoor")


def shape_counts():
    import importlib
    shapes = importlib.import_module("zk-dcap-verifier_b200.shapes")
    return shapes.counts(shapes.full_chain_shape())


class OrcOps:
    """field ops of the synthetic witness generator on the CPU oracle (reference arm)"""
    @staticmethod
    def _o():
        from oracle import oracle as orc
        return orc
    mul = staticmethod(lambda a, b: OrcOps._o().field_op("fr", "mul", np.ascontiguousarray(a), np.ascontiguousarray(b)))
    add = staticmethod(lambda a, b: OrcOps._o().field_op("fr", "add", np.ascontiguousarray(a), np.ascontiguousarray(b)))
    to_mont = staticmethod(lambda a: OrcOps._o().field_op("fr", "from_raw", np.ascontiguousarray(a)))


def cpu_reference_proof(k, seed, verify=True):
    """The reference arm: ParamsKZG::setup, keygen and ONE complete create_proof on the host cores through the CPU
    restatement of halo2 (oracle/prover.py + oracle/zk_oracle.c), on exactly the inputs ProofWorkload gives the device."""
    import hashlib
    import importlib
    from oracle import oracle as orc
    from oracle import prover as op
    shapes = importlib.import_module("zk-dcap-verifier_b200.shapes")
    zk = importlib.import_module("zk-dcap-verifier_b200")
    threads = orc.cpu_threads()
    cs = shapes.full_chain_shape()
    t0 = time.perf_counter()
    g, gl = orc.setup_params(k, TRAPDOOR, threads)
    params = {"k": k, "n": 1 << k, "g": g, "g_lagrange": gl, "s": TRAPDOOR}
    t_setup = time.perf_counter() - t0
    fixed, advice, mapping = shapes.synth_qe3_witness(cs, k, seed, OrcOps)
    circ = zk.Circuit(cs, k, transcript_repr=0xDCA90002)
    t0 = time.perf_counter()
    pk = op.keygen(orc, params, circ, fixed, mapping, circ.transcript_repr, threads)
    t_keygen = time.perf_counter() - t0
    rnd = rand_fr_host(np.random.default_rng(seed), op.randoms_needed(cs, k))
    trace = {"keep_h_ext": False}
    t0 = time.perf_counter()
    proof = op.create_proof(orc, params, pk, circ, advice, [], rnd, threads=threads, trace=trace)
    t_prove = time.perf_counter() - t0
    ok = None
    if verify:
        from oracle import pairing, verifier as ov
        vp = {"g0": (1, 2), "g2": pairing.G2_GEN, "s_g2": pairing.g2_mul(TRAPDOOR, pairing.G2_GEN)}
        ok = bool(ov.verify_proof(vp, pk["vk"], cs, k, [], proof))
    return {"prove_s": t_prove, "keygen_s": t_keygen, "setup_s": t_setup, "cores": threads, "proof": proof,
            "proof_sha256": hashlib.sha256(proof).hexdigest(), "proof_bytes": len(proof), "verified": ok,
            "sections": {kk: vv for kk, vv in trace.items() if kk.startswith("t_")}}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--k", type=int, default=19, help="log2 rows (19 = full-chain shape, 17 = QE3 shape)")
    ap.add_argument("--mode", default="proof", choices=["proof", "proof-sharded", "hotpath", "msm-sharded", "sweep"],
                    help="proof = complete create_proof on the device, independent proofs per GPU (default); proof-sharded = ONE "
                         "proof computed jointly by all GPUs (strong scaling); hotpath = only the MSM / NTT / h(X) calls")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--streams", type=int, default=3, help="1 = one stream, 3 = commitments on two streams + NTT/h(X) on a third")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep-max", type=int, default=24)
    ap.add_argument("--sweep-min", type=int, default=16)
    ap.add_argument("--sweep-oracle-max", type=int, default=22, help="largest k whose MSM is checked against the oracle's best_multiexp")
    ap.add_argument("--sweep-coset-max", type=int, default=24)
    ap.add_argument("--sweep-classes", action="store_true",
                    help="sweep mode: also time the resident MSM on 88-bit scalars and on 90 %% zero / small scalars (SURVEY 8d config 3)")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the one-proof / one-MSM-over-all-GPUs measurements")
    ap.add_argument("--no-dropin", action="store_true", help="skip the one-call-per-halo2-call-site e2e measurement")
    ap.add_argument("--check-bytes", action="store_true",
                    help="cpu_baseline leg: also run ONE complete CPU create_proof of the same inputs and compare the proof bytes")
    ap.add_argument("--inflight", type=int, default=2, help="proof mode: also measure throughput with this many proofs in flight per GPU")
    return ap.parse_args()


# ---- clocks ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile(prefix="clocks", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm)); out["sm_max_mhz"] = float(max(mx)); out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


_STREAMS = {}


def work_stream(torch, dev):
    """A non-default torch stream made current, so that the library's kernels, torch's copies and the
    torch.cuda.Event timers all live on ONE stream (the legacy default stream has handle 0, which the ABI reads
    as "create a private stream")."""
    key = str(dev)
    if key not in _STREAMS:
        _STREAMS[key] = torch.cuda.Stream(dev)
        torch.cuda.set_stream(_STREAMS[key])
    return _STREAMS[key].cuda_stream


# ---- synthetic inputs -----------------------------------------------------------------------------------
def rand_fr(torch, n, device, gen):
    """n x 4 int64 limbs, every element < 2^252 < r: a valid Montgomery-form Fr (uniform over 2^252 values)"""
    t = torch.randint(-2**63, 2**63 - 1, (n, 4), dtype=torch.int64, device=device, generator=gen)
    t[:, 3] &= 0x0FFFFFFFFFFFFFFF
    return t


def rand_fr_host(rng, n):
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    return a


class Workload:
    """Everything one step needs, resident on one GPU."""

    def __init__(self, zk, shapes, torch, device, k, seed, with_host=False, streams=3):
        self.zk, self.torch, self.k, self.n = zk, torch, k, 1 << k
        self.h2 = __import__("importlib").import_module("zk-dcap-verifier_b200.halo2")
        dev = torch.device("cuda", device)
        self.dev = dev
        self.ctx = zk.Context(device, work_stream(torch, dev))
        # Independent parts of a proof run on separate streams (one zk_ctx each): the latency- and atomic-bound
        # stages of one MSM batch (digit sort, partial levels, bucket reduction) overlap the integer-pipe-bound
        # stages of another, and the NTT / h(X) chain overlaps the commitments.  streams=1 serialises everything.
        self.multi = streams > 1
        self.sA = torch.cuda.current_stream(dev)
        if self.multi:
            self.sB, self.sC = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            self.ctxB = zk.Context(device, self.sB.cuda_stream)
            self.ctxC = zk.Context(device, self.sC.cuda_stream)
        else:
            self.ctxB = self.ctxC = self.ctx
        self.ctxs = [self.ctx, self.ctxB, self.ctxC] if self.multi else [self.ctx]
        self.cs = shapes.full_chain_shape()
        self.counts = shapes.counts(self.cs)
        self.calls = shapes.hot_path_calls(self.cs)
        gen = torch.Generator(device=dev); gen.manual_seed(seed)
        n, c = self.n, self.counts
        # synthetic SRS: g[i] = r_i * G, g_lagrange[i] = r'_i * G for random r (valid curve points; cofactor 1)
        G = np.concatenate([zk.fq_to_limbs(1), zk.fq_to_limbs(2)])
        sc = rand_fr(torch, 2 * n, dev, gen)
        pts = torch.empty((2 * n, 8), dtype=torch.int64, device=dev)
        self.h2.fixed_base_mul_dev(self.ctx, G, sc.data_ptr(), 2 * n, pts.data_ptr())
        self.h2.srs_upload_dev(self.ctx, k, pts[:n].data_ptr(), pts[n:].data_ptr(), precompute=True)
        self.params = self.h2.ResidentParams(self.ctx, k)
        torch.cuda.synchronize()
        if self.multi:
            self.ctxB.check(self.ctxB.lib.zk_ctx_share_srs(self.ctxB.h, self.ctx.h))
        self.paramsB = self.h2.ResidentParams(self.ctxB, k)
        del sc, pts
        # proving key side of evaluate_h: random fixed / sigma / l0 / l_last / l_blind polynomials (coefficient form)
        rng = np.random.default_rng(seed)
        fixed = [rand_fr_host(rng, n) for _ in range(c["fixed"])]
        sigma = [rand_fr_host(rng, n) for _ in range(c["perm_columns"])]
        l0, l_last, l_blind = (rand_fr_host(rng, n) for _ in range(3))
        self.pk = zk.ProvingKey(self.cs, k, fixed, sigma, l0, l_last, l_blind, coeff_form=True, ctx=self.ctxC)
        torch.cuda.synchronize()
        del fixed, sigma
        self.dom = self.pk.domain
        self.E = self.dom.ext_n
        self.stats = self.pk.program_stats()
        # witness columns in Lagrange form: advice | lookup (A', S' per lookup, then Z per lookup) | permutation z
        self.n_lag = self.calls["msm_lagrange"]           # A + 3L + P
        self.lag = rand_fr(torch, self.n_lag * n, dev, gen).view(self.n_lag, n, 4)
        self.work = torch.empty_like(self.lag)
        self.commits = torch.empty((self.n_lag + self.calls["msm_coeff"], 12), dtype=torch.int64, device=dev)
        self.h_ext = torch.empty((self.E, 4), dtype=torch.int64, device=dev)
        self.h_coeff = torch.empty((n * (c["degree"] - 1), 4), dtype=torch.int64, device=dev)
        self.instance = rand_fr(torch, max(1, c["instance"]) * n, dev, gen).view(-1, n, 4)
        self.chal = rand_fr_host(rng, 8)
        self.host = None
        if with_host:
            self.host = torch.empty((self.n_lag, n, 4), dtype=torch.int64).pin_memory()
            self.host.copy_(self.lag)
            self.host_work = torch.empty((self.n_lag, n, 4), dtype=torch.int64).pin_memory()
            self.host_h = torch.empty((self.E, 4), dtype=torch.int64).pin_memory()
            self.host_hc = torch.empty((n * (c["degree"] - 1), 4), dtype=torch.int64).pin_memory()
            self.host_commit = torch.empty((self.n_lag + self.calls["msm_coeff"], 12), dtype=torch.int64).pin_memory()

    # ---- one step, operands resident --------------------------------------------------------------------
    def step(self):
        if self.multi:
            return self.step_multi()
        h2, c, n = self.h2, self.counts, self.n
        A, Lk, P = c["advice"], c["lookups"], c["perm_sets"]
        w = self.work
        w.copy_(self.lag)                                   # create_proof keeps the Lagrange columns: work on a copy
        # commit_lagrange in the groups create_proof forms: advice | permuted (2L) | perm z (P) | lookup z (L)
        off = 0
        for cnt in (A, 2 * Lk, P, Lk):
            if cnt:
                h2.commit_dev(self.params, w[off].data_ptr(), cnt, self.commits[off].data_ptr(), lagrange=True)
            off += cnt
        # lagrange_to_coeff of every committed column
        h2.lagrange_to_coeff_dev(self.dom, w.data_ptr(), self.n_lag)
        # vanishing argument's random polynomial commitment
        ci = self.n_lag
        h2.commit_dev(self.params, w[0].data_ptr(), 1, self.commits[ci].data_ptr(), lagrange=False); ci += 1
        # evaluate_h: (A + I + P + 3L) coset NTTs + the fused per-row kernel
        adv = [w[i].data_ptr() for i in range(A)]
        ins = [self.instance[i].data_ptr() for i in range(c["instance"])]
        look = []
        for l in range(Lk):      # (product, permuted_input, permuted_table)
            look += [w[A + 2 * Lk + P + l].data_ptr(), w[A + 2 * l].data_ptr(), w[A + 2 * l + 1].data_ptr()]
        zs = [w[A + 2 * Lk + i].data_ptr() for i in range(P)]
        ch = self.chal
        h2.evaluate_h_dev(self.pk, adv, ins, ch[4:4 + self.cs.num_challenges], ch[0], ch[1], ch[2], ch[3], look, zs,
                          self.h_ext.data_ptr())
        # vanishing::construct
        h2.divide_by_vanishing_poly_dev(self.dom, self.h_ext.data_ptr())
        h2.extended_to_coeff_dev(self.dom, self.h_ext.data_ptr(), self.h_coeff.data_ptr())
        pieces = c["degree"] - 1
        h2.commit_dev(self.params, self.h_coeff.data_ptr(), pieces, self.commits[ci].data_ptr(), lagrange=False); ci += pieces
        # SHPLONK: two more commitments of degree < n polynomials
        h2.commit_dev(self.params, w[1].data_ptr(), 2, self.commits[ci].data_ptr(), lagrange=False)

    serialize = False

    def _j(self):
        """breakdown pass: force the three streams to run one call at a time so per-kernel timings do not overlap"""
        if self.serialize:
            for s_ in (self.sA, self.sB, self.sC):
                for t_ in (self.sA, self.sB, self.sC):
                    if s_ is not t_:
                        t_.wait_stream(s_)

    def step_multi(self):
        """Same calls as step(), issued on three streams with events for the true data dependencies."""
        torch, h2, c, n = self.torch, self.h2, self.counts, self.n
        A, Lk, P = c["advice"], c["lookups"], c["perm_sets"]
        sA, sB, sC = self.sA, self.sB, self.sC
        lag, w = self.lag, self.work
        sB.wait_stream(sA); sC.wait_stream(sA)
        # stream C: coefficient forms, evaluate_h, quotient
        with torch.cuda.stream(sC):
            w.copy_(lag)
        h2.lagrange_to_coeff_dev(self.dom, w.data_ptr(), self.n_lag)
        self._j()
        ev_coeff = torch.cuda.Event(); ev_coeff.record(sC)
        # streams A / B: commit_lagrange of the Lagrange columns (read-only), split by group
        offs, off = [], 0
        for cnt in (A, 2 * Lk, P, Lk):
            offs.append((off, cnt)); off += cnt
        for i, (o, cnt) in enumerate(offs):
            if cnt:
                h2.commit_dev(self.params if i % 2 == 0 else self.paramsB, lag[o].data_ptr(), cnt,
                              self.commits[o].data_ptr(), lagrange=True)
                self._j()
        adv = [w[i].data_ptr() for i in range(A)]
        ins = [self.instance[i].data_ptr() for i in range(c["instance"])]
        look = []
        for l in range(Lk):
            look += [w[A + 2 * Lk + P + l].data_ptr(), w[A + 2 * l].data_ptr(), w[A + 2 * l + 1].data_ptr()]
        zs = [w[A + 2 * Lk + i].data_ptr() for i in range(P)]
        ch = self.chal
        h2.evaluate_h_dev(self.pk, adv, ins, ch[4:4 + self.cs.num_challenges], ch[0], ch[1], ch[2], ch[3], look, zs,
                          self.h_ext.data_ptr())
        self._j()
        h2.divide_by_vanishing_poly_dev(self.dom, self.h_ext.data_ptr())
        self._j()
        h2.extended_to_coeff_dev(self.dom, self.h_ext.data_ptr(), self.h_coeff.data_ptr())
        self._j()
        ev_h = torch.cuda.Event(); ev_h.record(sC)
        ci = self.n_lag
        sA.wait_event(ev_coeff); sB.wait_event(ev_coeff)
        h2.commit_dev(self.params, w[0].data_ptr(), 1, self.commits[ci].data_ptr(), lagrange=False); ci += 1
        self._j()
        pieces = c["degree"] - 1
        h2.commit_dev(self.paramsB, w[1].data_ptr(), 2, self.commits[ci + pieces].data_ptr(), lagrange=False)
        self._j()
        sA.wait_event(ev_h)
        h2.commit_dev(self.params, self.h_coeff.data_ptr(), pieces, self.commits[ci].data_ptr(), lagrange=False)
        self._j()
        sA.wait_stream(sB); sA.wait_stream(sC)

    # ---- one step through the host-pointer ABI (the drop-in boundary) ---------------------------------
    def step_e2e(self):
        zk, c, n = self.zk, self.counts, self.n
        lib, ctx = self.ctx.lib, self.ctx
        A, Lk, P = c["advice"], c["lookups"], c["perm_sets"]
        hp = lambda t: C.c_void_p(t.data_ptr())
        hw = self.host_work
        hw.copy_(self.host)
        h2d = d2h = 0
        for i in range(self.n_lag):                          # ParamsKZG::commit_lagrange(column)
            ctx.check(lib.zk_msm(ctx.h, hp(hw[i]), None, 2, n, hp(self.host_commit[i])))
            h2d += n * 32; d2h += 96
        for i in range(self.n_lag):                          # EvaluationDomain::lagrange_to_coeff(column)
            ctx.check(lib.zk_lagrange_to_coeff(self.dom.h, hp(hw[i])))
            h2d += n * 32; d2h += n * 32
        ci = self.n_lag
        ctx.check(lib.zk_msm(ctx.h, hp(hw[0]), None, 1, n, hp(self.host_commit[ci]))); ci += 1
        h2d += n * 32; d2h += 96
        tab = lambda ts: (C.c_void_p * max(1, len(ts)))(*[t.data_ptr() for t in ts])
        adv = [hw[i] for i in range(A)]
        look = []
        for l in range(Lk):
            look += [hw[A + 2 * Lk + P + l], hw[A + 2 * l], hw[A + 2 * l + 1]]
        zs = [hw[A + 2 * Lk + i] for i in range(P)]
        ch = self.chal
        np_p = lambda a: a.ctypes.data_as(C.c_void_p)
        chal = np.ascontiguousarray(ch[4:5])
        ctx.check(lib.zk_evaluate_h(self.pk.h, tab(adv), tab([]), np_p(chal), np_p(ch[0]), np_p(ch[1]), np_p(ch[2]),
                                    np_p(ch[3]), tab(look), tab(zs), hp(self.host_h)))
        h2d += (A + 3 * Lk + P) * n * 32; d2h += self.E * 32
        ctx.check(lib.zk_divide_by_vanishing_poly(self.dom.h, hp(self.host_h)))
        h2d += self.E * 32; d2h += self.E * 32
        ctx.check(lib.zk_extended_to_coeff(self.dom.h, hp(self.host_h), hp(self.host_hc)))
        h2d += self.E * 32; d2h += n * (c["degree"] - 1) * 32
        for p in range(c["degree"] - 1):
            ctx.check(lib.zk_msm(ctx.h, C.c_void_p(self.host_hc.data_ptr() + p * n * 32), None, 1, n, hp(self.host_commit[ci]))); ci += 1
            h2d += n * 32; d2h += 96
        for p in range(2):
            ctx.check(lib.zk_msm(ctx.h, hp(hw[1 + p]), None, 1, n, hp(self.host_commit[ci]))); ci += 1
            h2d += n * 32; d2h += 96
        return h2d, d2h

    def work_units(self):
        """algorithmic work of one step: field multiplications and bytes per kernel family"""
        n, E, c = self.n, self.E, self.counts
        k, ek = self.k, self.dom.extended_k
        n_msm = self.calls["msm_lagrange"] + self.calls["msm_coeff"]
        n_coset = self.calls["coset_ntt_ext"]
        return {
            "msm_points": n_msm * n, "msm_bytes": 96 * n_msm * n,
            "ntt_bytes": 64 * (self.calls["intt_n"] * n + (n_coset + 1) * E),
            "ntt_muls": self.calls["intt_n"] * (n // 2) * k + (n_coset + 1) * (E // 2) * ek,
            "evalh_bytes": 32 * E * (self.stats["columns_read"] + 1), "evalh_muls": E * self.stats["muls_per_row"],
        }


# ---- complete proofs -------------------------------------------------------------------------------------
# SRS secret: Fr::random of ChaCha20Rng::from_seed([0; 32]) -- what halo2-base's gen_srs(k) would draw ([UPSTREAM?];
# derivation checked in tests/test_prover_oracle.py)
TRAPDOOR = 0x1c59a59b6cff4308740943526ade1d8c09f71b337a67269cc89586bcdd6dfcba


class DevOps:
    """field ops for the synthetic witness generator, run by the device library on host arrays"""

    def __init__(self, ctx):
        self.ctx = ctx

    def mul(self, a, b): return self.ctx.field_op("fr", "mul", np.ascontiguousarray(a), np.ascontiguousarray(b))
    def add(self, a, b): return self.ctx.field_op("fr", "add", np.ascontiguousarray(a), np.ascontiguousarray(b))
    def sub(self, a, b): return self.ctx.field_op("fr", "sub", np.ascontiguousarray(a), np.ascontiguousarray(b))
    def inv(self, a): return self.ctx.field_op("fr", "inv", np.ascontiguousarray(a))
    def to_mont(self, a): return self.ctx.field_op("fr", "to_mont", np.ascontiguousarray(a))

    def powers(self, zk, base, n):
        """[base^0 .. base^(n-1)] in Montgomery form, by doubling"""
        R = (1 << 256)
        arr = zk.fr_to_limbs(1).reshape(1, 4)
        m = 1
        while m < n:
            step = np.tile(zk.fr_to_limbs(pow(base, m, R_MOD)), (m, 1))
            arr = np.concatenate([arr, self.mul(arr, step)])
            m *= 2
        return arr[:n]


R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


class ProofWorkload:
    """A complete create_proof per step: ParamsKZG::setup with a known trapdoor, a satisfying synthetic witness of
    the full-chain shape, keygen on the device, then zk_create_proof."""

    def __init__(self, zk, shapes, torch, device, k, seed, with_host=True):
        self.zk, self.torch, self.k, self.n = zk, torch, k, 1 << k
        self.h2 = __import__("importlib").import_module("zk-dcap-verifier_b200.halo2")
        dev = torch.device("cuda", device)
        self.dev = dev
        self.ctx = zk.Context(device, work_stream(torch, dev))
        self.ctxs = [self.ctx]
        ops = DevOps(self.ctx)
        n = self.n
        self.cs = shapes.full_chain_shape()
        self.counts = shapes.counts(self.cs)
        self.calls = shapes.hot_path_calls(self.cs)
        # ParamsKZG::setup: g[i] = s^i G, g_lagrange[i] = L_i(s) G with L_i(s) = (s^n - 1)/n * w^i / (s - w^i)
        s = TRAPDOOR
        w = pow(pow(7, (R_MOD - 1) >> 28, R_MOD), 1 << (28 - k), R_MOD)
        sp = ops.powers(zk, s, n)
        wp = ops.powers(zk, w, n)
        c = (pow(s, n, R_MOD) - 1) * pow(n, -1, R_MOD) % R_MOD
        li = ops.mul(ops.mul(wp, np.tile(zk.fr_to_limbs(c), (n, 1))), ops.inv(ops.sub(np.tile(zk.fr_to_limbs(s), (n, 1)), wp)))
        G = np.concatenate([zk.fq_to_limbs(1), zk.fq_to_limbs(2)])
        sc = torch.from_numpy(np.concatenate([sp, li]).view(np.int64)).to(dev)
        pts = torch.empty((2 * n, 8), dtype=torch.int64, device=dev)
        self.h2.fixed_base_mul_dev(self.ctx, G, sc.data_ptr(), 2 * n, pts.data_ptr())
        self.h2.srs_upload_dev(self.ctx, k, pts[:n].data_ptr(), pts[n:].data_ptr(), precompute=True)
        self.params = self.h2.ResidentParams(self.ctx, k)
        torch.cuda.synchronize()
        self.host_gl = pts[n:].cpu().numpy().view(np.uint64).reshape(n, 8)      # for the cpu_baseline checker only
        del sc, pts, sp, wp, li
        fixed, advice, mapping = shapes.synth_qe3_witness(self.cs, k, seed, ops)
        self.circ = zk.Circuit(self.cs, k, transcript_repr=0xDCA90002)
        t0 = time.perf_counter()
        self.key = zk.ProverKey(self.params, self.circ, fixed, mapping)
        torch.cuda.synchronize()
        self.keygen_s = time.perf_counter() - t0
        if self.key.aux_ctx is not None:
            self.ctxs.append(self.key.aux_ctx)
        self._keygen_inputs = (fixed, mapping)
        self.pk_stats = self._program_stats()
        rng = np.random.default_rng(seed)
        self.n_rand = self.key.randoms_needed
        rnd = rand_fr_host(rng, self.n_rand)
        A = self.counts["advice"]
        self.host_adv = torch.empty((A, n, 4), dtype=torch.int64).pin_memory()
        for i in range(A):
            self.host_adv[i].copy_(torch.from_numpy(advice[i].view(np.int64)))
        self.host_rnd = torch.from_numpy(rnd.view(np.int64)).pin_memory()
        self.dev_adv = self.host_adv.to(dev)
        self.dev_rnd = self.host_rnd.to(dev)
        del advice
        self.E = 1 << self.ext_k()
        self.last_proof = None

    def ext_k(self):
        ek = self.k
        while (1 << ek) < (1 << self.k) * (self.counts["degree"] - 1):
            ek += 1
        return ek

    def _program_stats(self):
        m, c = C.c_uint64(), C.c_uint64()
        lib = self.ctx.lib
        pk = C.c_void_p(lib.zk_pkey_pk(self.key.h))
        self.ctx.check(lib.zk_pk_program_stats(pk, C.byref(m), C.byref(c)))
        secs = C.c_double()
        log = lib.zk_pk_evaluator_log(pk, C.byref(secs))
        return {"muls_per_row": m.value, "columns_read": c.value,
                "evaluator": "generated (NVRTC, sm_100a)" if lib.zk_pk_evaluator_kind(pk) else "interpreter",
                "evaluator_note": (log or b"").decode()[:200], "nvrtc_compile_s": secs.value}

    def step(self):
        A = self.counts["advice"]
        self.last_proof = self.key.create_proof_raw([self.dev_adv[i].data_ptr() for i in range(A)], self.dev_rnd.data_ptr(), self.n_rand)

    def step_e2e(self):
        A, n = self.counts["advice"], self.n
        self.last_proof = self.key.create_proof_raw([self.host_adv[i].data_ptr() for i in range(A)], self.host_rnd.data_ptr(), self.n_rand)
        return A * n * 32 + self.n_rand * 32, len(self.last_proof)

    # ---- several proofs in flight on one GPU (BASELINE configs[4]: batches of independent proofs per GPU) ----------
    def enable_inflight(self, p, fixed_mapping):
        """p - 1 more provers (own context = own streams, own key and work area, the SRS shared) driven by host threads."""
        zk, torch = self.zk, self.torch
        fixed, mapping = fixed_mapping
        self.extra = []
        for _ in range(p - 1):
            cx = zk.Context(self.dev.index)
            cx.check(cx.lib.zk_ctx_share_srs(cx.h, self.ctx.h))
            params = self.h2.ResidentParams(cx, self.k)
            key = zk.ProverKey(params, self.circ, fixed, mapping)
            self.extra.append((cx, key))
        torch.cuda.synchronize()

    def step_inflight(self):
        """one proof per prover, concurrently; every call returns only when its proof is complete"""
        A = self.counts["advice"]
        ptrs = [self.dev_adv[i].data_ptr() for i in range(A)]
        out = [None] * (1 + len(self.extra))

        def work(i, key):
            out[i] = key.create_proof_raw(ptrs, self.dev_rnd.data_ptr(), self.n_rand)
        ths = [threading.Thread(target=work, args=(i + 1, key)) for i, (_, key) in enumerate(self.extra)]
        for t in ths:
            t.start()
        work(0, self.key)
        for t in ths:
            t.join()
        self.last_inflight = out[0]
        ref = getattr(self, "inflight_ref", None) or out[0]
        for i, o in enumerate(out):
            if o != ref:                                  # diagnostic, never fatal: which transcript items differ
                self.inflight_mismatches = getattr(self, "inflight_mismatches", 0) + 1
                items = [j // 32 for j in range(0, max(len(o), len(ref)), 32) if o[j:j + 32] != ref[j:j + 32]]
                print("[bench] device %s: in-flight prover %d returned a different proof: %d of %d 32-byte items differ, first %d"
                      % (self.dev, i, len(items), len(ref) // 32, items[0] if items else -1), file=sys.stderr, flush=True)
        return len(out)

    serialize = False

    def work_units(self):
        n, E, c = self.n, self.E, self.counts
        k, ek = self.k, self.ext_k()
        n_msm = self.calls["msm_lagrange"] + self.calls["msm_coeff"]
        n_coset = self.calls["coset_ntt_ext"]
        return {
            "msm_points": n_msm * n, "msm_bytes": 96 * n_msm * n,
            "ntt_bytes": 64 * (self.calls["intt_n"] * n + (n_coset + 1) * E),
            "ntt_muls": self.calls["intt_n"] * (n // 2) * k + (n_coset + 1) * (E // 2) * ek,
            "evalh_bytes": 32 * E * (self.pk_stats["columns_read"] + 1), "evalh_muls": E * self.pk_stats["muls_per_row"],
        }


# ---- CPU baseline (oracle port) -----------------------------------------------------------------------
def verify_gpu_proof(wl):
    """CPU verifier (oracle/verifier.py: transcript replay, expected h(x), SHPLONK, pairing) on the device's proof."""
    from oracle import pairing, pymodel as pm, verifier as ov
    fc, pc = wl.key.vk_commitments()
    vk = {"fixed_commitments": pm.limbs_to_affine(fc), "permutation_commitments": pm.limbs_to_affine(pc),
          "transcript_repr": wl.circ.transcript_repr}
    vp = {"g0": (1, 2), "g2": pairing.G2_GEN, "s_g2": pairing.g2_mul(TRAPDOOR, pairing.G2_GEN)}
    return bool(ov.verify_proof(vp, vk, wl.cs, wl.k, [], wl.last_proof))


def verify_native(wl):
    """The library's own verifier (zk_verify_proof: host code with its own pairing) on the device's proof."""
    from zkdcap_b200 import verify_proof
    fc, pc = wl.key.vk_commitments()
    # [s]G2 of the synthetic setup, computed with the product's scalar field only: G2 arithmetic lives in the verifier,
    # so the uncompressed encodings are produced once by the small affine ladder below
    g2, s_g2 = g2_encodings(TRAPDOOR)
    g1 = np.concatenate([wl.zk.fq_to_limbs(1), wl.zk.fq_to_limbs(2)])
    return bool(verify_proof(wl.circ, fc, pc, g1, g2, s_g2, [], wl.last_proof))


P_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))


def g2_encodings(s):
    """(G2, [s]G2) as 128 uncompressed bytes each (x.c0, x.c1, y.c0, y.c1 little-endian): plain affine double-and-add
    over Fq2 on Python integers -- setup of the synthetic SRS, not part of any measured path."""
    p = P_MOD
    mul = lambda a, b: ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)
    sub = lambda a, b: ((a[0] - b[0]) % p, (a[1] - b[1]) % p)

    def inv(a):
        d = pow((a[0] * a[0] + a[1] * a[1]) % p, -1, p)
        return (a[0] * d % p, (-a[1]) * d % p)

    def add(P, Q):
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            if P[1] != Q[1] or P[1] == (0, 0):
                return None
            xx = mul(P[0], P[0])
            m = mul(((3 * xx[0]) % p, (3 * xx[1]) % p), inv(((2 * P[1][0]) % p, (2 * P[1][1]) % p)))
        else:
            m = mul(sub(Q[1], P[1]), inv(sub(Q[0], P[0])))
        x3 = sub(sub(mul(m, m), P[0]), Q[0])
        return (x3, sub(mul(m, sub(P[0], x3)), P[1]))
    acc, base, e = None, G2_GEN, s
    while e:
        if e & 1:
            acc = add(acc, base)
        base = add(base, base)
        e >>= 1
    enc = lambda q: b"".join(v.to_bytes(32, "little") for v in (q[0][0], q[0][1], q[1][0], q[1][1]))
    return enc(G2_GEN), enc(acc)


def cpu_tail_sample(orc, zk, shapes, cs, k, threads):
    """Times the prover's loops BETWEEN the hot-path calls on the CPU oracle (one lookup argument, one permutation
    set, a few evaluations, one kate division, one linear combination) and scales by their counts."""
    n = 1 << k
    c = shapes.counts(cs)
    bf = cs.blinding_factors()
    usable = n - (bf + 1)
    circ = zk.Circuit(cs, k)
    dom_omega = orc.Domain(cs.degree(), k, threads).omega
    fixed = [orc.gen_scalars(300 + i, n) for i in range(c["fixed"])]
    advice = [orc.gen_scalars(400 + i, n) for i in range(c["advice"])]
    sc = orc.gen_scalars(9, 8)
    t0 = time.time()
    cin = orc.eval_graph_rows(circ.lookup_input_graphs[0], k, fixed, advice, [], [], sc[0], threads)
    ctab = orc.eval_graph_rows(circ.lookup_table_graphs[0], k, fixed, advice, [], [], sc[0], threads)
    t_compress = time.time() - t0
    # inputs drawn from the table so that the permutation succeeds
    idx = np.random.default_rng(1).integers(0, usable, size=usable)
    inp = np.ascontiguousarray(ctab[:usable][idx])
    t0 = time.time(); a_s, s_s = orc.permute_expression_pair(inp, ctab[:usable], usable); t_permute = time.time() - t0
    pin = np.concatenate([a_s, ctab[usable:]]); ptab = np.concatenate([s_s, ctab[usable:]])
    t0 = time.time(); orc.lookup_product(k, cin, ctab, pin, ptab, sc[1], sc[2]); t_lprod = time.time() - t0
    chunk = c["degree"] - 2
    t0 = time.time()
    orc.permutation_product(k, advice[:chunk], fixed[:chunk], sc[1], sc[2], dom_omega, sc[3], sc[4], threads)
    t_pprod = time.time() - t0
    t0 = time.time()
    for i in range(4):
        orc.eval_polynomial(advice[i], sc[5])
    t_eval = (time.time() - t0) / 4
    t0 = time.time(); orc.kate_division(advice[0], sc[6]); t_kate = time.time() - t0
    t0 = time.time(); orc.poly_lincomb(advice[:10], [sc[7]] * 10, threads); t_lin = (time.time() - t0) / 10
    L, P = c["lookups"], c["perm_sets"]
    n_evals = len(cs.advice_queries) + len(cs.fixed_queries) + 1 + c["perm_columns"] + 3 * P - 1 + 5 * L + 1
    n_open_polys = c["advice"] + c["fixed"] + c["perm_columns"] + P + 3 * L + 2
    n_kate = 4 + 3 + 2 + 2 + 2 + 1 + 1        # points of the six rotation sets of this shape, + the final division by (X - u)
    total = L * (t_compress + t_permute + t_lprod) + P * t_pprod + n_evals * t_eval + n_kate * t_kate + 2 * n_open_polys * t_lin
    return total, {"lookup_compress_s": t_compress, "lookup_permute_s": t_permute, "lookup_product_s": t_lprod,
                   "permutation_product_s": t_pprod, "eval_polynomial_s": t_eval, "kate_division_s": t_kate,
                   "lincomb_per_poly_s": t_lin, "tail_total_scaled_s": total}


def cpu_baseline_proof(k, threads=None, full=False):
    """Times the CPU oracle (halo2-algorithm restatement) on a bounded sample and scales to one step."""
    from oracle import oracle as orc
    import importlib
    shapes = importlib.import_module("zk-dcap-verifier_b200.shapes")
    zk = importlib.import_module("zk-dcap-verifier_b200")
    from tests.test_evalh_host import make_case, oracle_h
    threads = threads or orc.cpu_threads()
    cs = shapes.full_chain_shape()
    calls = shapes.hot_path_calls(cs)
    n = 1 << k
    t0 = time.time()
    bases = orc.gen_points(1, n, threads)
    sc = orc.gen_scalars(2, n)
    t_gen = time.time() - t0
    t0 = time.time(); orc.best_multiexp(sc, bases, threads); t_msm = time.time() - t0
    dom = orc.Domain(cs.degree(), k, threads)
    t0 = time.time(); coeff = dom.lagrange_to_coeff(sc); t_intt = time.time() - t0
    t0 = time.time(); ext = dom.coeff_to_extended(coeff); t_coset = time.time() - t0
    t0 = time.time(); dom.extended_to_coeff(ext); t_iext = time.time() - t0
    # evaluate_h (includes its own coset NTTs) on 1/2^(k-ks) of the rows
    ks = min(k, 15)
    import random
    case = make_case_fast(cs, ks, orc)
    t0 = time.time(); oracle_h_fast(orc, zk, cs, ks, case, threads); t_h = (time.time() - t0) * (1 << (k - ks))
    n_msm = calls["msm_lagrange"] + calls["msm_coeff"]
    total = n_msm * t_msm + calls["intt_n"] * t_intt + t_h + t_iext
    parts = {"msm_s": t_msm, "intt_s": t_intt, "coset_ntt_s": t_coset, "intt_ext_s": t_iext,
             "evaluate_h_scaled_s": t_h, "gen_inputs_s": t_gen}
    sample = ("1 of %d MSM(2^%d), 1 of %d iNTT(2^%d), 1 iNTT(ext), evaluate_h (with its %d coset NTTs) on 2^%d "
              "of 2^%d rows; scaled by call counts / rows" % (n_msm, k, calls["intt_n"], k, calls["coset_ntt_ext"], ks, k))
    if full:
        t_tail, tail_parts = cpu_tail_sample(orc, zk, shapes, cs, k, threads)
        total += t_tail
        parts.update(tail_parts)
        sample += "; prover loops between them: 1 of %d lookup arguments, 1 of %d permutation sets, 4 evaluations, 1 kate division, " \
                  "a 10-polynomial linear combination, scaled by counts" % (len(cs.lookups), shapes.counts(cs)["perm_sets"])
    return {"seconds_per_proof": total, "cores": threads, "parts": parts, "sample": sample}


def make_case_fast(cs, k, orc):
    import importlib
    shapes = importlib.import_module("zk-dcap-verifier_b200.shapes")
    n = 1 << k
    sc = shapes.counts(cs)
    dom = orc.Domain(cs.degree(), k)
    seed = [100]

    def rp():
        seed[0] += 1
        return orc.gen_scalars(seed[0], n)
    case = {"advice": [rp() for _ in range(sc["advice"])], "instance": [rp() for _ in range(sc["instance"])],
            "z_polys": [rp() for _ in range(sc["perm_sets"])],
            "lookup_polys": [(rp(), rp(), rp()) for _ in range(sc["lookups"])]}
    E = dom.ext_n
    re_ = lambda: orc.gen_scalars(seed[0] + 7777, E)
    case["fixed_cosets"] = [re_() for _ in range(sc["fixed"])]
    case["sigma_cosets"] = [re_() for _ in range(sc["perm_columns"])]
    case["l0"] = case["l_last"] = case["l_active"] = re_()
    case["dom"] = dom
    return case


def oracle_h_fast(orc, zk, cs, k, case, threads):
    ev = zk.Evaluator.new(cs)
    desc, keep = ev.describe(cs, k)
    one = orc.gen_scalars(5, 4)
    flat = [p for t in case["lookup_polys"] for p in t]
    return orc.evaluate_h(desc, case["fixed_cosets"], case["sigma_cosets"], case["l0"], case["l_last"], case["l_active"],
                          case["advice"], case["instance"], [], one[0], one[1], one[2], one[3], flat, case["z_polys"],
                          case["dom"].ext_n, threads=threads)


# ---- main ------------------------------------------------------------------------------------------------
_JSON_FD = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else written to fd 1 meanwhile (NCCL's version banner,
    library chatter) has been diverted to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse()
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    full = args.mode in ("proof", "proof-sharded")
    sharded = args.mode == "proof-sharded"
    metric = METRIC if full else METRIC_HOTPATH
    config = {"workload": "full-chain DCAP shape: QE3 column set at k=%d (ext_k=%d), %s per step"
                          % (args.k, args.k + 2, "one complete create_proof (satisfying synthetic witness, proof verified by the "
                             "CPU verifier)" if full else "one create_proof hot path"),
              "shape": shape_counts(), "l2": "inputs (>= 1 GiB of witness columns per step) exceed the 126 MB L2",
              "parallelism": ("one proof sharded over the GPUs: commitments by column / base range, h(X) by rows; NCCL all-gather"
                              if sharded else "independent proofs per GPU (replicas)") if args.gpus > 1 else "1 GPU",
              "streams": (1 if os.environ.get("ZK_PROVER_STREAMS") == "1" else 2) if full else args.streams}

    if args.impl == "reference":
        if rank != 0:
            return 0
        if full:
            # ONE real, complete CPU create_proof (oracle/prover.py over the C loops of oracle/zk_oracle.c) of the same
            # circuit, witness, SRS and RNG stream as the device arm: steps = 1, warmup = 0, measured, not extrapolated
            ref = cpu_reference_proof(args.k, seed=0xDCA90002)
            sec = ref["prove_s"]
            sample = ("one complete create_proof of the k=%d witness (same seed, SRS trapdoor and RNG stream as the device arm), "
                      "all host threads; keygen (%.1f s) and ParamsKZG::setup (%.1f s) outside the timed step" % (args.k, ref["keygen_s"], ref["setup_s"]))
            extra = {"keygen_s": ref["keygen_s"], "proof_sha256": ref["proof_sha256"], "proof_bytes": ref["proof_bytes"],
                     "proof_verified": ref["verified"]}
            cores, parts, steps, warm = ref["cores"], ref["sections"], 1, 0
        else:
            best = min((cpu_baseline_proof(args.k, full=False) for _ in range(max(1, min(args.steps, 2)))), key=lambda r: r["seconds_per_proof"])
            sec, sample, cores, parts, extra, steps, warm = best["seconds_per_proof"], best["sample"], best["cores"], best["parts"], {}, args.steps, args.warmup
        v = 1.0 / sec
        config["shape"] = shape_counts()
        line = {"impl": "reference", "metric": metric, "value": v, "unit": "proofs/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
                "data": DATA, "config": config,
                "cpu_baseline": {"value": v, "unit": "proofs/s", "cores": cores, "kind": "port",
                                 "sample": sample, "parts": parts},
                "e2e": {"value": v, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        line.update(extra)
        emit(line)
        return 0

    import importlib
    import torch
    zk = importlib.import_module("zk-dcap-verifier_b200")
    shapes = importlib.import_module("zk-dcap-verifier_b200.shapes")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this backend has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    if args.mode == "sweep":
        return sweep(args, zk, torch, dev, local_rank)
    if args.mode == "msm-sharded":
        return msm_sharded(args, zk, torch, dist, dev, rank, local_rank, world)

    if full:
        # every rank proves the SAME witness (replicas are still independent work): the sharded measurement of the default
        # N > 1 run and the proof hash need identical inputs everywhere
        wl = ProofWorkload(zk, shapes, torch, local_rank, args.k, seed=0xDCA90002)
        if sharded and world > 1:
            uid = [zk.Group.unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            wl.group = zk.Group(wl.ctx, rank, world, uid[0])
            wl.key.set_group(wl.group)
    else:
        wl = Workload(zk, shapes, torch, local_rank, args.k, seed=0xDCA90002 + rank, with_host=(rank == 0 and not args.no_e2e), streams=args.streams)
    assert config["shape"] == wl.counts
    ctx = wl.ctx
    imad_peak = ctx.bench_imad(1, 2000) if rank == 0 else None
    modmul_peak = ctx.bench_modmul(1, 2000) if rank == 0 else None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        wl.step()
    barrier()
    launches0 = sum(cx.launch_count() for cx in wl.ctxs)
    for cx in wl.ctxs:
        cx.counters(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        wl.step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = sum(cx.launch_count() for cx in wl.ctxs) - launches0
    work = [cx.counters() for cx in wl.ctxs]
    bucket_adds_step = sum(w["msm_bucket_entries"] for w in work) / args.steps
    msms_step = sum(w["msms"] for w in work) / args.steps
    # per-kernel breakdown: the same steps again with the streams forced to take turns, timed per launch with
    # CUDA events by the library (overlapping kernels would inflate each other's durations)
    for cx in wl.ctxs:
        cx.timing(True); cx.timing_reset()
    wl.serialize = True
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); s0.record()
    for _ in range(args.steps):
        wl.step()
    s1.record(); barrier()
    ms_serial = s0.elapsed_time(s1) / args.steps
    wl.serialize = False
    kstats = {}
    for cx in wl.ctxs:
        for name, (cnt, tot) in cx.timing_read().items():
            c0, t0 = kstats.get(name, (0, 0.0))
            kstats[name] = (c0 + cnt, t0 + tot)
        cx.timing(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = (1 if sharded else world) * args.steps / (ms_max * 1e-3)

    e2e = None
    if sharded and not args.no_e2e:           # every rank takes part: the proof is a joint computation
        wl.step_e2e()
        barrier()
        t0 = time.perf_counter()
        reps = max(1, min(args.steps, 3))
        for _ in range(reps):
            h2d, d2h = wl.step_e2e()
        barrier()
        dt = (time.perf_counter() - t0) / reps
        e2e = {"value": 1.0 / dt, "unit": "proofs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": dt * 1e3, "note": "zk_create_proof on every rank with the advice columns and the RNG stream in "
                                                "pinned HOST memory; one proof computed jointly"}
    elif not args.no_e2e and (full or rank == 0):
        # every rank uploads its own witness at the same time (PCIe / host-memory contention included); max over ranks
        wl.step_e2e()
        barrier() if full else torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = max(1, min(args.steps, 3))
        for _ in range(reps):
            h2d, d2h = wl.step_e2e()
        torch.cuda.synchronize()
        dt_loc = (time.perf_counter() - t0) / reps
        te = torch.tensor([dt_loc], dtype=torch.float64, device=dev)
        if dist is not None and full:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        e2e = {"value": world / dt, "unit": "proofs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": dt * 1e3,
               "note": ("zk_create_proof with the advice columns and the RNG stream in pinned HOST memory, proof bytes returned to "
                        "the host; all ranks run it concurrently (one proof each), wall clock, max over ranks") if full else
                       ("host-pointer C ABI, one call per halo2 call site (74 zk_msm, 67 zk_lagrange_to_coeff, "
                        "zk_evaluate_h, zk_divide_by_vanishing_poly, zk_extended_to_coeff); pinned host buffers; "
                        "measured on rank 0 and scaled by the GPU count (ranks are independent)")}
    batch = None
    proof_ref = wl.last_proof if full else None
    if full and args.inflight > 1 and not sharded:
        # throughput with several proofs in flight per GPU: the latency- and atomic-bound stages of one proof overlap
        # the integer-pipe-bound stages of another (every rank does this; rank 0 reports max over ranks)
        wl.enable_inflight(args.inflight, wl._keygen_inputs)
        wl.inflight_ref = proof_ref
        for _ in range(2):
            wl.step_inflight()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); b0.record()
        done = 0
        for _ in range(args.steps):
            done += wl.step_inflight()
        b1.record(); barrier()
        tb = torch.tensor([b0.elapsed_time(b1)], dtype=torch.float64, device=dev)
        bad = torch.tensor([getattr(wl, "inflight_mismatches", 0)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
            dist.all_reduce(bad, op=dist.ReduceOp.SUM)
        batch = {"inflight_per_gpu": args.inflight, "proofs": done * world, "ms": float(tb.item()),
                 "value": world * done / (float(tb.item()) * 1e-3), "unit": "proofs/s",
                 "proofs_identical": int(bad.item()) == 0, "mismatching_proofs": int(bad.item()),
                 "proofs_compared": (done + 2 * args.inflight) * world,
                 "note": "independent provers on one GPU (own streams, key and work area; shared SRS), one host thread each; "
                         "`proofs_identical`: every proof of every prover on every rank (warm-up included) equals the proof "
                         "that rank computed alone"}
    # ---- one proof / one MSM over all GPUs (strong scaling), measured in the SAME run so that the driver's --gpus N
    # lines carry them ----------------------------------------------------------------------------
    sharded_proof = sharded_msm = None
    proof_plain = wl.last_proof if full else None
    if full and not sharded and world > 1 and not args.no_sharded:
        sharded_proof = measure_sharded_proof(args, zk, torch, dist, dev, wl, rank, world, barrier, proof_plain)
        sharded_msm = measure_sharded_msm(args, zk, torch, dist, dev, rank, local_rank, world, barrier)
    e2e_dropin = None
    if full and world == 1 and not args.no_e2e and not args.no_dropin:
        e2e_dropin = measure_dropin(args, zk, shapes, torch, local_rank, wl)
    if dist is not None:
        dist.barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    wu = wl.work_units()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    fam = {}
    for name, (cnt, tot) in kstats.items():
        fam[name] = {"launches": cnt, "ms_per_step": tot / args.steps}
    hot = [kname for kname in fam if kname.startswith(("msm", "ntt")) or kname == "evaluate_h"]
    dom_name = max(hot, key=lambda kname: fam[kname]["ms_per_step"]) if hot else None
    # what the committed ncu captures say about each kernel (profiles/ncu_summary.json, written by tools/ncu_digest.py from
    # the .ncu-rep of THIS round's kernels); never a literal in this file
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
    except Exception:
        pass
    # roofline of the dominant kernel family
    roof = None
    if dom_name:
        d = fam[dom_name]
        per_launch_ms = d["ms_per_step"] * args.steps / max(1, d["launches"])
        if dom_name.startswith("msm"):
            bytes_step, muls_step = wu["msm_bytes"], None
        elif dom_name.startswith("ntt"):
            bytes_step, muls_step = wu["ntt_bytes"], wu["ntt_muls"]
        else:
            bytes_step, muls_step = wu["evalh_bytes"], wu["evalh_muls"]
        launches_per_step = d["launches"] / args.steps
        ach = bytes_step / launches_per_step / (per_launch_ms * 1e-3) / 1e9
        prof = ncu.get(dom_name) or {}
        traffic = None
        if prof.get("dram_bytes_per_launch") and prof.get("k") == args.k:
            traffic = prof["dram_bytes_per_launch"]      # mean over the launches of this kernel in ONE captured proof
        roof = {"kernel": dom_name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": traffic,
                "traffic_source": (prof.get("source") if traffic else None), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_step / launches_per_step,
                "avg_launch_ms": per_launch_ms, "share_of_step": d["ms_per_step"] / ms_serial,
                "from_profile": {kk: vv for kk, vv in prof.items() if kk not in ("source",)} or None,
                "note": "integer-pipe bound kernel (254-bit modular arithmetic): the binding roofline is `int_pipe`"}
    # integer pipe: field multiplications per second against the measured register-resident peak
    msm_ms = sum(v["ms_per_step"] for kname, v in fam.items() if kname.startswith("msm"))
    ntt_ms = sum(v["ms_per_step"] for kname, v in fam.items() if kname.startswith("ntt"))
    evh_ms = fam.get("evaluate_h", {"ms_per_step": 0})["ms_per_step"]
    acc_ms = sum(fam.get(nm, {"ms_per_step": 0})["ms_per_step"] for nm in ("msm_accumulate", "msm_pair_add"))
    int_pipe = {"modmul_peak_G_per_s": modmul_peak, "imad_wide_peak_T_per_s": imad_peak,
                "ntt_modmul_G_per_s": wu["ntt_muls"] / (ntt_ms * 1e-3) / 1e9 if ntt_ms else None,
                "evalh_modmul_G_per_s": wu["evalh_muls"] / (evh_ms * 1e-3) / 1e9 if evh_ms else None,
                "msm_Gpts_per_s": wu["msm_points"] / (msm_ms * 1e-3) / 1e9 if msm_ms else None,
                "ntt_GB_per_s": wu["ntt_bytes"] / (ntt_ms * 1e-3) / 1e9 if ntt_ms else None,
                "evalh_GB_per_s": wu["evalh_bytes"] / (evh_ms * 1e-3) / 1e9 if evh_ms else None,
                "msm_bucket_additions_per_step": bucket_adds_step, "msms_per_step": msms_step,
                "msm_additions_per_point": bucket_adds_step / wu["msm_points"] if wu["msm_points"] else None}
    if int_pipe["ntt_modmul_G_per_s"] and modmul_peak:
        int_pipe["ntt_frac_of_modmul_peak"] = int_pipe["ntt_modmul_G_per_s"] / modmul_peak
    if int_pipe["evalh_modmul_G_per_s"] and modmul_peak:
        int_pipe["evalh_frac_of_modmul_peak"] = int_pipe["evalh_modmul_G_per_s"] / modmul_peak
    if acc_ms and bucket_adds_step:
        # bucket accumulation: every sorted entry is one mixed XYZZ addition (8M + 2S = 10 modular multiplications;
        # MUL_PER_ADD is the library's formula count, csrc/ec.cuh); SURVEY 8d: frac(IMAD) = modmuls * 136 / (t * imad_wide_peak)
        mm = bucket_adds_step * MUL_PER_ADD
        g = mm / (acc_ms * 1e-3) / 1e9
        acc = {"kernel": "msm_accumulate", "modmuls_per_step": mm, "mul_per_addition": MUL_PER_ADD, "ms_per_step": acc_ms,
               "achieved_G_per_s": g, "frac_of_modmul_peak": g / modmul_peak if modmul_peak else None,
               "frac_of_imad_wide_peak": g * 136 / (imad_peak * 1e3) if imad_peak else None}
        int_pipe["msm_accumulate"] = acc
        if roof and roof["kernel"] == "msm_accumulate":
            roof["int_pipe"] = acc
            if roof.get("traffic"):
                roof["dram_bytes_per_bucket_addition_from_profile"] = roof["traffic"] * fam["msm_accumulate"]["launches"] / args.steps / bucket_adds_step

    import hashlib
    cpu = None
    if sharded and world > 1 and not args.no_cpu_baseline:      # checker only: the jointly computed proof must verify
        cpu = {"value": None, "unit": "proofs/s", "cores": 0, "kind": "port", "sample": "verification only",
               "gpu_proof_verified": verify_gpu_proof(wl), "proof_bytes": len(wl.last_proof)}
    if not args.no_cpu_baseline and world == 1:
        try:
            b = cpu_baseline_proof(args.k, full=full)
            cpu = {"value": 1.0 / b["seconds_per_proof"], "unit": "proofs/s", "cores": b["cores"], "kind": "port",
                   "sample": b["sample"], "parts": b["parts"]}
            if full:      # the CPU leg is also the checker: the device's proof must pass the CPU verifier (real pairing)
                cpu["gpu_proof_verified"] = verify_gpu_proof(wl)
                cpu["proof_bytes"] = len(proof_plain)
                cpu["commitments_vs_oracle"] = check_commitments(wl, proof_plain)
                if args.check_bytes:      # one complete CPU proof of the same inputs: the BYTES must be equal
                    ref = cpu_reference_proof(args.k, seed=0xDCA90002, verify=False)
                    cpu["proof_bytes_equal_cpu_prover"] = ref["proof"] == proof_plain
                    cpu["cpu_prover_s"] = ref["prove_s"]
        except Exception as ex:          # the oracle is a checker; its absence must not hide the GPU number
            cpu = {"value": None, "unit": "proofs/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}
    verified = None
    if full:
        try:
            wl.last_proof = proof_plain
            verified = verify_native(wl)
        except Exception as ex:
            verified = "failed: %r" % (ex,)
    line = {"metric": metric, "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "ms_per_step_one_stream": ms_serial,
            "higher_is_better": True, "scaling": "strong" if sharded else "weak",
            "vs_baseline": None, "dtype": DTYPE, "data": DATA, "config": config,
            "clocks": clocks, "e2e": e2e, "e2e_dropin": e2e_dropin, "gpu_launches": launches, "roofline": roof, "int_pipe": int_pipe,
            "kernels": fam, "cpu_baseline": cpu, "batch": batch, "proof_verified": verified,
            "sharded_proof": sharded_proof, "sharded_msm": sharded_msm}
    if full:
        line["evaluate_h"] = wl.pk_stats
        line["keygen_s"] = wl.keygen_s
        line["proof_sha256"] = hashlib.sha256(proof_plain).hexdigest()
        line["proof_bytes"] = len(proof_plain)
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


MUL_PER_ADD = 10      # XYZZ mixed addition, madd-2008-s: 8M + 2S (csrc/ec.cuh)


def check_commitments(wl, proof, columns=(0, 15, 27)):
    """cpu_baseline leg, checker: the first A 32-byte items of the proof are the advice commitments.  Recompute a few of
    them with the CPU oracle's best_multiexp over the same SRS (downloaded from the device) and compare the bytes."""
    from oracle import oracle as orc
    from oracle import prover as op
    from oracle import pymodel as pm
    n, bf = wl.n, wl.cs.blinding_factors()
    gl = wl.host_gl
    rnd = wl.host_rnd.numpy().view(np.uint64).reshape(-1, 4)
    out = {}
    for c in columns:
        if c >= wl.counts["advice"]:
            continue
        col = wl.host_adv[c].numpy().view(np.uint64).reshape(n, 4).copy()
        col[n - (bf + 1):] = rnd[c * (bf + 1):(c + 1) * (bf + 1)]          # create_proof's blinding rows, in draw order
        pt = pm.limbs_to_affine(orc.g1_to_affine(orc.best_multiexp(col, gl)))[0]
        out["advice[%d]" % c] = op.compress_point(pt) == proof[32 * c:32 * c + 32]
    out["all_equal"] = all(out.values())
    return out


def measure_sharded_proof(args, zk, torch, dist, dev, wl, rank, world, barrier, proof_plain):
    """ONE proof computed jointly by all ranks (zk_group): commitments by column / base range, lookup arguments by owner,
    transforms by owner with a row-slice exchange, h(X) by rows.  Same key, witness and RNG stream on every rank."""
    uid = [zk.Group.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    group = zk.Group(wl.ctx, rank, world, uid[0])
    wl.key.set_group(group)
    for _ in range(2):
        wl.step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); e0.record()
    for _ in range(args.steps):
        wl.step()
    e1.record(); barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    same = torch.tensor([1 if wl.last_proof == proof_plain else 0], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    wl.key.set_group(None)
    barrier()
    group.close()
    return {"ms_per_proof": ms, "value": 1e3 / ms, "unit": "proofs/s", "n_gpus": world, "scaling": "strong",
            "bytes_equal_single_gpu_proof_on_all_ranks": bool(same.item()),
            "note": "zk_group: one create_proof over all GPUs, device-timed, max over ranks"}


def measure_sharded_msm(args, zk, torch, dist, dev, rank, local_rank, world, barrier, k=24):
    """ONE MSM of 2^k points sharded by contiguous base range: every rank keeps ITS slice of the bases resident with its own
    window table (N/G x 64 B x W), runs the resident-table Pippenger on its slice of the scalars, and the partial points
    are exchanged and summed.  Checked against the single-GPU result of the same MSM computed slice by slice on rank 0."""
    h2 = __import__("importlib").import_module("zk-dcap-verifier_b200.halo2")
    ctx = zk.Context(local_rank, work_stream(torch, dev))
    n = 1 << k
    per = n // world
    ks = k - (world.bit_length() - 1)
    G = np.concatenate([zk.fq_to_limbs(1), zk.fq_to_limbs(2)])

    def slice_inputs(r):
        gen = torch.Generator(device=dev); gen.manual_seed(0x5CA1A4 + 1000 * k + r)
        return rand_fr(torch, per, dev, gen), rand_fr(torch, per, dev, gen)
    sc, bs = slice_inputs(rank)
    pts = torch.empty((per, 8), dtype=torch.int64, device=dev)
    h2.fixed_base_mul_dev(ctx, G, bs.data_ptr(), per, pts.data_ptr())
    h2.srs_upload_dev(ctx, ks, pts.data_ptr(), pts.data_ptr(), precompute=True)
    params = h2.ResidentParams(ctx, ks)
    out = torch.empty((1, 12), dtype=torch.int64, device=dev)
    uid = [zk.Group.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    group = zk.Group(ctx, rank, world, uid[0])

    def step():
        # zk_msm_sharded_dev: local Pippenger on the resident slice; its last kernel stores the partial point into every
        # peer's window and raises a flag, the wait kernel gathers, one more kernel sums and normalises
        group.msm_sharded_dev(sc.data_ptr(), per, out.data_ptr(), lagrange=True)
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); e0.record()
    reps = max(3, args.steps)
    for _ in range(reps):
        step()
    e1.record(); barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / reps
    res = {"k": k, "ms": ms, "Gpts_s": n / (ms * 1e-3) / 1e9, "n_gpus": world, "scaling": "strong",
           "bases": "resident per rank: slice of 2^%d bases with its own window table" % ks,
           "exchange": ("peer-to-peer: the MSM's last kernel stores its 128-byte partial point into every peer's window over NVLink and raises a "
                        "flag; no collective call") if group.p2p else "NCCL all-gather of one 128-byte point per rank"}
    # single-GPU reference of the SAME MSM on rank 0: the slices one after the other on its own context
    if rank == 0:
        parts = torch.empty((world, 12), dtype=torch.int64, device=dev)
        t1 = 0.0
        for r in range(world):
            sc_r, bs_r = slice_inputs(r)
            h2.fixed_base_mul_dev(ctx, G, bs_r.data_ptr(), per, pts.data_ptr())
            h2.srs_upload_dev(ctx, ks, pts.data_ptr(), pts.data_ptr(), precompute=True)
            prm = h2.ResidentParams(ctx, ks)
            h2.commit_dev(prm, sc_r.data_ptr(), 1, parts[r].data_ptr(), lagrange=True)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            h2.commit_dev(prm, sc_r.data_ptr(), 1, parts[r].data_ptr(), lagrange=True)
            a1.record(); torch.cuda.synchronize()
            t1 += a0.elapsed_time(a1)
        single = torch.empty((1, 12), dtype=torch.int64, device=dev)
        ctx.check(ctx.lib.zk_g1_sum_dev(ctx.h, C.c_void_p(parts.data_ptr()), world, C.c_void_p(single.data_ptr())))
        torch.cuda.synchronize()
        # `single` comes from zk_g1_sum over Jacobian records, `out` from the XYZZ sum: both normalised (z = 1)
        res["matches_single_gpu"] = bool(torch.equal(single, out))
        res["single_gpu_ms_slice_by_slice"] = t1
        res["result_point"] = [int(x) & 0xFFFFFFFFFFFFFFFF for x in out.cpu().view(-1).tolist()]
    barrier()
    group.close()
    return res


def measure_dropin(args, zk, shapes, torch, local_rank, wl_proof):
    """The boundary north_star names: one host-pointer call per halo2 call site (best_multiexp behind commit /
    commit_lagrange, lagrange_to_coeff, evaluate_h, divide_by_vanishing_poly, extended_to_coeff), every call copying its
    operands in and its result out.  Random columns of the same shape (the calls do not depend on satisfiability)."""
    try:
        w2 = Workload(zk, shapes, torch, local_rank, args.k, seed=0xDCA90002, with_host=True, streams=1)
        w2.step_e2e()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            h2d, d2h = w2.step_e2e()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        res = {"value": 1.0 / dt, "unit": "proofs/s", "ms_per_step": dt * 1e3, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "calls_per_step": w2.n_lag * 2 + 1 + 3 + (w2.counts["degree"] - 1) + 2,
               "note": "hot path only (74 zk_msm, 67 zk_lagrange_to_coeff, zk_evaluate_h, zk_divide_by_vanishing_poly, "
                       "zk_extended_to_coeff) through the host-pointer ABI behind best_multiexp / best_fft / evaluate_h; the "
                       "prover's other loops stay on the CPU in that integration and are not in this number"}
        del w2
        torch.cuda.empty_cache()
        return res
    except Exception as ex:
        return {"value": None, "note": "failed: %r" % (ex,)}


def sweep(args, zk, torch, dev, local_rank):
    """MSM Gpts/s and NTT / coset-NTT GB/s sweeps (BASELINE configs[2], [3]); one JSON line per size, EVERY timed result
    checked (the checker is the CPU oracle -- never inside a timed region):
      MSM   k <= --sweep-oracle-max: bit-exact against the oracle's best_multiexp on the downloaded bases and scalars;
            above: the bases are r_i * G with known r_i, so sum s_i * (r_i * G) must equal (sum s_i * r_i) * G -- the dot
            product and the single scalar multiplication are done by the oracle (exact, size independent)
      NTT   k <= 22: bit-exact against the oracle's best_fft; above: spot evaluations f[i] = poly(omega^i) by the oracle
      coset k <= 20: bit-exact against the oracle's coeff_to_extended; above: spot evaluations at zeta * omega_ext^i"""
    h2 = __import__("importlib").import_module("zk-dcap-verifier_b200.halo2")
    from oracle import oracle as orc
    from oracle import pymodel as pm
    ctx = zk.Context(local_rank, work_stream(torch, dev))
    gen = torch.Generator(device=dev); gen.manual_seed(0x5CA1A4)
    G = np.concatenate([zk.fq_to_limbs(1), zk.fq_to_limbs(2)])
    R_MOD = h2.R_MOD
    root = pow(7, (R_MOD - 1) >> 28, R_MOD)
    one = zk.fr_to_limbs(1)
    host = lambda t: t.cpu().numpy().view(np.uint64)
    aff = lambda g1: orc.g1_to_affine(np.asarray(g1, dtype=np.uint64).reshape(-1, 12))

    def timed(fn, reps=5):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def msm_expected(sc_h, pts_h, r_h, k):
        if k <= args.sweep_oracle_max:
            return aff(orc.best_multiexp(sc_h, pts_h)), "exact: oracle best_multiexp"
        dot = orc.eval_polynomial(orc.field_op("fr", "mul", sc_h, r_h), one)       # sum s_i * r_i  (Horner at 1)
        return aff(orc.g1_mul(G, dot)), "exact: sum s_i*(r_i*G) == (sum s_i*r_i)*G, dot product and scalar multiplication by the oracle"

    for k in range(args.sweep_min, args.sweep_max + 1):
        n = 1 << k
        sc = rand_fr(torch, n, dev, gen)
        rs = rand_fr(torch, n, dev, gen)                   # discrete logs of the bases
        pts = torch.empty((n, 8), dtype=torch.int64, device=dev)
        h2.fixed_base_mul_dev(ctx, G, rs.data_ptr(), n, pts.data_ptr())
        h2.srs_upload_dev(ctx, k, pts.data_ptr(), pts.data_ptr(), precompute=True)
        params = h2.ResidentParams(ctx, k)
        cw = (C.c_uint32(), C.c_uint32())
        ctx.check(ctx.lib.zk_srs_window(ctx.h, 2, C.byref(cw[0]), C.byref(cw[1])))
        out = torch.empty((1, 12), dtype=torch.int64, device=dev)
        res = {"k": k, "window_bits": cw[0].value, "windows": cw[1].value}
        sc_h, r_h = host(sc).reshape(n, 4), host(rs).reshape(n, 4)
        pts_h = host(pts).reshape(n, 8) if k <= args.sweep_oracle_max else None
        want, how = msm_expected(sc_h, pts_h, r_h, k)
        ok = True
        for name, fn in (("msm_resident", lambda: h2.commit_dev(params, sc.data_ptr(), 1, out.data_ptr(), lagrange=True)),
                         ("msm_plain", lambda: h2.msm_dev(ctx, sc.data_ptr(), pts.data_ptr(), n, out.data_ptr()))):
            ms = timed(fn)
            res[name + "_ms"] = ms; res[name + "_Gpts_s"] = n / (ms * 1e-3) / 1e9
            good = bool(np.array_equal(aff(host(out)), want))
            res[name + "_parity"] = how if good else "MISMATCH"
            ok = ok and good
        if args.sweep_classes and k <= 24:
            # canonical values built on the host, converted to Montgomery form by the library
            rng = np.random.default_rng(0x5CA1A4 + k)
            for cname in ("bits88", "small90"):
                raw = np.zeros((n, 4), dtype=np.uint64)
                if cname == "bits88":
                    raw[:, 0] = rng.integers(0, 2**64, size=n, dtype=np.uint64)
                    raw[:, 1] = rng.integers(0, 2**24, size=n, dtype=np.uint64)
                else:
                    kind = rng.integers(0, 10, size=n)
                    raw[:, 0] = np.where(kind < 5, 0, rng.integers(0, 256, size=n)).astype(np.uint64)   # 50 % zeros, 40 % bytes
                    big = kind == 9                                                                      # 10 % full width
                    raw[big] = rand_fr_host(rng, int(big.sum()))
                mont = ctx.field_op("fr", "to_mont", np.ascontiguousarray(raw))
                sc_c = torch.from_numpy(mont.view(np.int64)).to(dev)
                ms = timed(lambda: h2.commit_dev(params, sc_c.data_ptr(), 1, out.data_ptr(), lagrange=True))
                res["msm_resident_%s_ms" % cname] = ms
                res["msm_resident_%s_Gpts_s" % cname] = n / (ms * 1e-3) / 1e9
                want_c, how_c = msm_expected(mont, pts_h, r_h, k)
                good = bool(np.array_equal(aff(host(out)), want_c))
                res["msm_resident_%s_parity" % cname] = how_c if good else "MISMATCH"
                ok = ok and good
                del sc_c
        del pts, rs, pts_h, r_h
        # ---- best_fft ------------------------------------------------------------------------------------------------
        w = pow(root, 1 << (28 - k), R_MOD)
        omega = zk.fr_to_limbs(w)
        a = rand_fr(torch, n, dev, gen)
        a_h = host(a).reshape(n, 4).copy()
        work = a.clone()
        ms = timed(lambda: h2.ntt_dev(ctx, work.data_ptr(), omega, k))
        res["ntt_ms"] = ms; res["ntt_GB_s"] = 64 * n / (ms * 1e-3) / 1e9
        res["ntt_Gmul_s"] = (n // 2) * k / (ms * 1e-3) / 1e9
        work.copy_(a)
        h2.ntt_dev(ctx, work.data_ptr(), omega, k)
        f = host(work).reshape(n, 4)
        if k <= 22:
            good = bool(np.array_equal(f, orc.best_fft(a_h, omega, k))); how = "exact: oracle best_fft"
        else:
            good = all(np.array_equal(orc.eval_polynomial(a_h, zk.fr_to_limbs(pow(w, i, R_MOD))), f[i]) for i in (0, 1, n // 2 + 7, n - 1))
            how = "spot: 4 outputs equal the oracle's evaluation of the input polynomial at omega^i"
        res["ntt_parity"] = how if good else "MISMATCH"
        ok = ok and good
        del work
        # ---- coeff_to_extended (coset NTT k -> k + 2, the transform evaluate_h's inputs take) ------------------------
        if k + 2 <= 28 and k <= args.sweep_coset_max:
            dom = zk.EvaluationDomain(5, k, ctx)
            ext = torch.empty((4 * n, 4), dtype=torch.int64, device=dev)
            ms = timed(lambda: h2.coeff_to_extended_dev(dom, a.data_ptr(), ext.data_ptr(), 1))
            res["coset_ntt_ms"] = ms; res["coset_ntt_GB_s"] = (32 * n + 32 * 4 * n) / (ms * 1e-3) / 1e9
            e_h = host(ext).reshape(4 * n, 4)
            if k <= 20:
                good = bool(np.array_equal(e_h, orc.Domain(5, k).coeff_to_extended(a_h))); how = "exact: oracle coeff_to_extended"
            else:
                we = pow(root, 1 << (28 - (k + 2)), R_MOD)
                good = all(np.array_equal(orc.eval_polynomial(a_h, zk.fr_to_limbs(pm.ZETA * pow(we, i, R_MOD) % R_MOD)), e_h[i])
                           for i in (0, 1, 2 * n + 5, 4 * n - 1))
                how = "spot: 4 outputs equal the oracle's evaluation at zeta * omega_ext^i"
            res["coset_ntt_parity"] = how if good else "MISMATCH"
            ok = ok and good
            del ext, dom
        res["parity"] = "exact" if ok else "MISMATCH"
        emit(res)
        del sc, a
        torch.cuda.empty_cache()
    return 0


def msm_sharded(args, zk, torch, dist, dev, rank, local_rank, world):
    """`--mode msm-sharded`: ONE MSM of 2^k points sharded by base range over the ranks (zk_msm_sharded_dev, the same measurement
    the default N > 1 line carries as `sharded_msm`)."""
    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
    if world == 1:
        raise SystemExit("bench.py --mode msm-sharded needs torchrun with N > 1 (the single-GPU figures are in --mode sweep)")
    res = measure_sharded_msm(args, zk, torch, dist, dev, rank, local_rank, world, barrier, k=args.k)
    if rank == 0:
        emit({"metric": "BN254 G1 MSM sharded by base range, Gpts/s", "value": res["Gpts_s"], "unit": "Gpts/s", "n_gpus": world,
              "steps": max(3, args.steps), "warmup": 3, "ms_per_step": res["ms"], "higher_is_better": True, "scaling": "strong",
              "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
              "config": {"workload": "one MSM of 2^%d points, bases sharded by contiguous range, resident per rank" % args.k}, "sharded_msm": res})
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
