"""Build libfitsnap_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m fitsnap_b200.csrc.build [--force] [--verbose]

The shared object lands in fitsnap_b200/_lib/ (git-ignored, travels to the GPU box).
cudart is linked statically so the library has no dependency on torch's CUDA runtime;
torch tensors only supply raw device pointers and the stream handle.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
LIB_DIR = os.path.join(PKG, "_lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libfitsnap_b200.so")

SOURCES = ["fsb_api.cu", "gram.cu", "solve.cu", "stream_ops.cu", "lasso.cu", "gram_small.cu", "gram_tma.cu", "pinv.cu", "gram_i8.cu", "comm.cu"]
HEADERS = ["fsb_common.cuh", os.path.join(ROOT, "include", "fitsnap_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--cudart", "static",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = [os.path.join(HERE, s) for s in SOURCES]
    hdrs = [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    if not force and _newer(LIB_PATH, srcs + hdrs + [os.path.abspath(__file__)]):
        return LIB_PATH

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src).replace(".cu", ".o"))
        if not force and _newer(obj, [src] + hdrs + [os.path.abspath(__file__)]):
            return obj, ""
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    log = "\n".join(l for _, l in results if l)
    with open(os.path.join(LIB_DIR, "ptxas.log"), "w") as f:
        f.write(log)
    if verbose:
        print(log)
    cmd = [nvcc, "-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", LIB_PATH] + objs + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
