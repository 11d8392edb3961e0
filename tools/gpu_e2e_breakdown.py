"""Wall-clock breakdown of one e2e step (host-pointer ABI) by call family."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import zkdcap_b200 as zk
from importlib import import_module
import bench
shapes = import_module("zk-dcap-verifier_b200.shapes")
wl = bench.Workload(zk, shapes, torch, 0, 19, seed=1, with_host=True)
lib, ctx, n = wl.ctx.lib, wl.ctx, wl.n
hp = lambda t: C.c_void_p(t.data_ptr())
hw = wl.host_work; hw.copy_(wl.host)
def T(f, reps=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
res = {}
res["msm_lagrange_ms_each"] = T(lambda: [ctx.check(lib.zk_msm(ctx.h, hp(hw[i]), None, 2, n, hp(wl.host_commit[i]))) for i in range(20)]) / 20
res["lagrange_to_coeff_ms_each"] = T(lambda: [ctx.check(lib.zk_lagrange_to_coeff(wl.dom.h, hp(hw[i]))) for i in range(20)]) / 20
# raw copies
buf = torch.empty((n, 4), dtype=torch.int64, device="cuda")
res["h2d_16MB_ms"] = T(lambda: buf.copy_(hw[0], non_blocking=True), 10)
res["d2h_16MB_ms"] = T(lambda: hw[1].copy_(buf, non_blocking=True), 10)
big = torch.empty((wl.E, 4), dtype=torch.int64, device="cuda")
res["h2d_64MB_ms"] = T(lambda: big.copy_(wl.host_h, non_blocking=True), 5)
t0 = time.perf_counter(); wl.step_e2e(); torch.cuda.synchronize(); res["step_e2e_ms"] = (time.perf_counter() - t0) * 1e3
ctx.timing(True); ctx.timing_reset()
ctx.check(lib.zk_msm(ctx.h, hp(hw[0]), None, 2, n, hp(wl.host_commit[0])))
res["single_msm_kernels"] = {k: round(v[1], 4) for k, v in ctx.timing_read().items()}
print(json.dumps(res, indent=1))
