"""Build libstpyb.so (the C-ABI CUDA library) in-tree for sm_100a.

nvcc cross-compiles without a GPU; objects are compiled in parallel and linked
into stpy_b200/libstpyb.so with a static cudart, so the library has no runtime
dependency beyond the driver (collectives are issued by the host side through
torch.distributed; the library itself never calls NCCL).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libstpyb.so")
BUILD = os.path.join(HERE, "build")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def headers_mtime():
    m = 0.0
    for root in (HERE, os.path.join(os.path.dirname(PKG), "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(BUILD, exist_ok=True)
    hm = headers_mtime()
    jobs = []
    objs = []
    for src in sources():
        s = os.path.join(HERE, src)
        o = os.path.join(BUILD, src[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hm):
            cmd = [nvcc] + ARCH + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), r.stderr))
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(run, jobs):
                if verbose and log:
                    print(log)
    if jobs or not os.path.exists(OUT):
        link = [nvcc] + ARCH + ["-shared", "-cudart", "static", "-o", OUT] + objs + ["-ldl"]
        run(link)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
