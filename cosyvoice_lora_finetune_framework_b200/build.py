"""Build libcvflow.so (hand-written sm_100a CUDA behind the C ABI of include/cvflow.h) in-tree.

nvcc cross-compiles without a GPU. Objects are rebuilt only when a source or header is newer.
Usage: python -m cosyvoice_lora_finetune_framework_b200.build [--force] [--profiling]

--profiling additionally builds libcvflow_prof.so (-DCVFLOW_PROFILING_BUILD: honours CVFLOW_SKIP, which drops kernel
classes from the step so that profiles/prof_marginals.py can time their in-graph share; never loaded by the product,
select it explicitly with CVFLOW_LIB_PATH).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libcvflow.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(HERE, "..", "include", "cvflow.h"))
    return max(os.path.getmtime(h) for h in hs)


def build(force: bool = False, verbose: bool = True, profiling: bool = False) -> str:
    OBJ = os.path.join(CSRC, "build_prof" if profiling else "build")
    LIB = os.path.join(HERE, "libcvflow_prof.so" if profiling else "libcvflow.so")
    FLAGS = globals()["FLAGS"] + (["-DCVFLOW_PROFILING_BUILD"] if profiling else [])
    os.makedirs(OBJ, exist_ok=True)
    hm = _headers_mtime()
    jobs = []
    objs = []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hm):
            jobs.append((s, o))

    def run(job):
        s, o = job
        cmd = [NVCC] + FLAGS + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (s, r.stdout, r.stderr))
        if verbose:
            print("[cvflow build] compiled", os.path.basename(s))

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        if verbose:
            print("[cvflow build] linked", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    if "--profiling" in sys.argv:
        build(force="--force" in sys.argv, profiling=True)
