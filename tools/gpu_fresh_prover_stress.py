"""Stress of the FIRST proofs of freshly created provers running next to an established one (the pattern of bench.py's
in-flight warm-up): every cycle creates a context + key on the shared SRS, runs `steps` concurrent rounds, compares every
proof with the proof computed alone, and destroys the prover again.
usage: gpu_fresh_prover_stress.py [seconds] [steps per cycle] [k] [device]"""
import gc
import importlib
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
k = int(sys.argv[3]) if len(sys.argv) > 3 else 19
dev = int(sys.argv[4]) if len(sys.argv) > 4 else 0
zk = importlib.import_module("zk-dcap-verifier_b200")
shapes = importlib.import_module("zk-dcap-verifier_b200.shapes")
wl = bench.ProofWorkload(zk, shapes, torch, dev, k, seed=0xDCA90002, with_host=False)
wl.step()
good = wl.last_proof
A = wl.counts["advice"]
ptrs = [wl.dev_adv[i].data_ptr() for i in range(A)]
bad = cycles = proofs = 0
t0 = time.time()
while time.time() - t0 < seconds:
    wl.enable_inflight(2, wl._keygen_inputs)
    keys = [wl.key] + [key for _, key in wl.extra]
    for s in range(steps):
        out = [None] * len(keys)

        def work(i):
            out[i] = keys[i].create_proof_raw(ptrs, wl.dev_rnd.data_ptr(), wl.n_rand)
        ths = [threading.Thread(target=work, args=(i,)) for i in range(1, len(keys))]
        for t in ths:
            t.start()
        work(0)
        for t in ths:
            t.join()
        for i, o in enumerate(out):
            proofs += 1
            if o != good:
                bad += 1
                items = [j // 32 for j in range(0, len(good), 32) if o[j:j + 32] != good[j:j + 32]]
                print("dev %d cycle %d step %d prover %d (%s) differs: first item %d, %d of %d items differ"
                      % (dev, cycles, s, i, "established" if i == 0 else "fresh", items[0] if items else -1, len(items), len(good) // 32), flush=True)
    cycles += 1
    keys = None
    for cx, key in wl.extra:
        del key
    extra, wl.extra = wl.extra, []
    ctxs = [cx for cx, _ in extra]
    del extra
    gc.collect()
    torch.cuda.synchronize()
    for cx in ctxs:
        cx.close()
print("dev %d: cycles %d proofs %d mismatching %d" % (dev, cycles, proofs, bad))
