"""Stress of several provers in flight on one GPU: every proof must equal the proof computed alone.
usage: gpu_inflight_stress.py [steps] [provers] [k]"""
import importlib
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
provers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
k = int(sys.argv[3]) if len(sys.argv) > 3 else 19
zk = importlib.import_module("zk-dcap-verifier_b200")
shapes = importlib.import_module("zk-dcap-verifier_b200.shapes")
wl = bench.ProofWorkload(zk, shapes, torch, 0, k, seed=0xDCA90002, with_host=False)
wl.step()
good = wl.last_proof
wl.step()
assert wl.last_proof == good, "a lone prover is not deterministic"
wl.enable_inflight(provers, wl._keygen_inputs)
A = wl.counts["advice"]
ptrs = [wl.dev_adv[i].data_ptr() for i in range(A)]
keys = [wl.key] + [key for _, key in wl.extra]
bad = 0
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
        if o != good:
            bad += 1
            first = next((j // 32 for j in range(0, len(good), 32) if o[j:j + 32] != good[j:j + 32]), -1)
            ndiff = sum(1 for j in range(0, len(good), 32) if o[j:j + 32] != good[j:j + 32])
            print("step %d prover %d differs: first item %d, %d of %d items differ" % (s, i, first, ndiff, len(good) // 32), flush=True)
print("steps %d provers %d mismatching proofs %d" % (steps, provers, bad))
