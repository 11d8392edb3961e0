"""Builds libzkdcap_b200.so (hand-written sm_100a kernels + the C ABI) in-tree with nvcc."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libzkdcap_b200.so")
SOURCES = ["api.cu", "ntt.cu", "msm.cu", "evalh.cu", "prover.cu", "group.cu", "params.cu", "verifier.cu"]
HEADERS = ["ff.cuh", "ec.cuh", "common.cuh", "evalh.cuh", "evalh_compile.hpp", "evalh_codegen.hpp", "prover_kernels.cuh", "transcript.hpp", "poseidon.hpp", os.path.join("..", "..", "include", "zkdcap_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + os.environ.get("ZK_NVCC_EXTRA", "").split()


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def embed_ff():
    """csrc/ff.cuh as a C++ raw string literal (build/ff_embed.inc): the prelude of the evaluate_h kernels that evalh.cu
    generates and compiles with NVRTC at key generation"""
    src = open(os.path.join(CSRC, "ff.cuh")).read().replace("#include <stdint.h>", "")
    out = os.path.join(HERE, "build", "ff_embed.inc")
    text = 'R"ZKFF(' + src + ')ZKFF"\n'
    if not os.path.exists(out) or open(out).read() != text:
        with open(out, "w") as f:
            f.write(text)
    return out


def build(force=False, verbose=False):
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    embed_ff()
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    if jobs or force or _stale(OUT, objs):
        run([NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-ldl"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
