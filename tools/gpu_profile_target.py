"""ncu target: N complete k=19 proofs (bench.py's ProofWorkload), nothing else.  usage: gpu_profile_target.py [steps] [k]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
k = int(sys.argv[2]) if len(sys.argv) > 2 else 19
zk = importlib.import_module("zk-dcap-verifier_b200")
shapes = importlib.import_module("zk-dcap-verifier_b200.shapes")
wl = bench.ProofWorkload(zk, shapes, torch, 0, k, seed=0xDCA90002, with_host=False)
torch.cuda.synchronize()
print("setup done", flush=True)
for _ in range(steps):
    wl.step()
torch.cuda.synchronize()
print("proof bytes", len(wl.last_proof))
