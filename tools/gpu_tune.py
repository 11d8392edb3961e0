"""Parameter sweeps on the GPU box.  Each configuration runs in a fresh process (tunables are read once).
usage: python tools/gpu_tune.py            -> runs the sweep, one JSON line per configuration
       python tools/gpu_tune.py --one      -> measures the current environment's configuration"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import numpy as np
    import torch
    import zkdcap_b200 as zk
    from importlib import import_module
    h2 = import_module("zk-dcap-verifier_b200.halo2")
    import bench
    dev = torch.device("cuda", 0)
    ctx = zk.Context(0, bench.work_stream(torch, dev))
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    k = int(os.environ.get("TUNE_K", "19"))
    n = 1 << k
    res = {"env": {e: os.environ[e] for e in os.environ if e.startswith("ZK_")}, "k": k}

    def timeit(fn, reps=5):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    what = os.environ.get("TUNE_WHAT", "ntt,msm")
    if "ntt" in what:
        dom = zk.EvaluationDomain(5, k, ctx)
        cnt = 16
        a = bench.rand_fr(torch, cnt * n, dev, gen)
        ext = torch.empty((cnt * dom.ext_n, 4), dtype=torch.int64, device=dev)
        res["intt_n_ms_per_poly"] = timeit(lambda: h2.lagrange_to_coeff_dev(dom, a.data_ptr(), cnt)) / cnt
        res["coset_ext_ms_per_poly"] = timeit(lambda: h2.coeff_to_extended_dev(dom, a.data_ptr(), ext.data_ptr(), cnt)) / cnt
        muls = (n // 2) * k, (dom.ext_n // 2) * dom.extended_k
        res["intt_Gmul_s"] = muls[0] / res["intt_n_ms_per_poly"] / 1e6
        res["coset_Gmul_s"] = muls[1] / res["coset_ext_ms_per_poly"] / 1e6
        del a, ext
    if "msm" in what:
        G = np.concatenate([zk.fq_to_limbs(1), zk.fq_to_limbs(2)])
        sc = bench.rand_fr(torch, n, dev, gen)
        pts = torch.empty((n, 8), dtype=torch.int64, device=dev)
        h2.fixed_base_mul_dev(ctx, G, sc.data_ptr(), n, pts.data_ptr())
        h2.srs_upload_dev(ctx, k, pts.data_ptr(), pts.data_ptr(), precompute=True)
        params = h2.ResidentParams(ctx, k)
        # skewed scalar classes of real witness columns: all ones, random bits, bytes, 88-bit limbs
        one = torch.tensor(zk.fr_to_limbs(1).astype(np.int64), device=dev)
        for name, make in (("ones", lambda: one.repeat(n, 1)),
                           ("bits", lambda: one.repeat(n, 1) * torch.randint(0, 2, (n, 1), device=dev, generator=gen)),
                           ("sparse", lambda: bench.rand_fr(torch, n, dev, gen) * (torch.rand((n, 1), device=dev, generator=gen) < 0.1))):
            s = make().contiguous()
            out = torch.empty((1, 12), dtype=torch.int64, device=dev)
            res["msm_%s_ms" % name] = timeit(lambda: h2.commit_dev(params, s.data_ptr(), 1, out.data_ptr(), lagrange=True), reps=3)
        for cnt in (1, 28):
            s = bench.rand_fr(torch, cnt * n, dev, gen)
            out = torch.empty((cnt, 12), dtype=torch.int64, device=dev)
            ctx.timing(True); ctx.timing_reset()
            ms = timeit(lambda: h2.commit_dev(params, s.data_ptr(), cnt, out.data_ptr(), lagrange=True), reps=3)
            st = ctx.timing_read(); ctx.timing(False)
            res["msm_x%d_ms_per_msm" % cnt] = ms / cnt
            res["msm_x%d_parts" % cnt] = {kk: round(v[1] / 5 / cnt, 4) for kk, v in st.items()}
            del s, out
    print(json.dumps(res), flush=True)


def sweep():
    cfgs = []
    for tl, mt in ((10, 8), (11, 8), (9, 8), (10, 7), (11, 9)):
        cfgs.append({"ZK_NTT_TILE_LOG": str(tl), "ZK_NTT_MAX_T": str(mt), "TUNE_WHAT": "ntt"})
    pass
    for cfg in cfgs:
        env = dict(os.environ); env.update(cfg)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True, text=True,
                           timeout=300)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + r.stderr[-400:]
        print(line, flush=True)


if __name__ == "__main__":
    one() if "--one" in sys.argv else sweep()
