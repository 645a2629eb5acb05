"""In-tree build of libdcb200.so (hand-written sm_100a kernels + C ABI) with plain nvcc.

    python -m dcb200.build            # or: python diffusion-classifier_b200/dcb200/build.py

nvcc cross-compiles for sm_100a without a GPU; the .so lands next to this file so it travels with the tree.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libdcb200.so")
SOURCES = ["api.cu", "gemm_tc.cu", "gemm_tc2.cu", "gemm_tc2x.cu", "gemm_tc3.cu", "gemm_simt.cu", "norm.cu", "elementwise.cu", "pack.cu", "attention.cu", "attention_tc.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]
FLAGS += os.environ.get("DCB_EXTRA_NVCC_FLAGS", "").split()       # experiments only (e.g. -DDCB_ATTN_POLY=4)
if os.environ.get("DCB_BUILD_SUFFIX"):                               # ... built next to the product library, never replacing it
    OBJ = os.path.join(HERE, "build" + os.environ["DCB_BUILD_SUFFIX"])
    LIB = os.path.join(HERE, "build", "libdcb200" + os.environ["DCB_BUILD_SUFFIX"] + ".so")
if os.environ.get("DCB_PROBES"):     # issue-loop experiments of tools/gemm_micro.py / epi_micro.py (never in the product build)
    FLAGS.append("-DDCB_PROBES")


def _stale(src, obj):
    if not os.path.exists(obj):
        return True
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(HERE, "..", "..", "include", "dcb200.h"))
    return any(os.path.getmtime(d) > os.path.getmtime(obj) for d in deps)


def _compile(name, verbose):
    src, obj = os.path.join(CSRC, name), os.path.join(OBJ, name.replace(".cu", ".o"))
    if not _stale(src, obj):
        return name, "up to date"
    r = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
    with open(obj + ".ptxas.log", "w") as f:
        f.write(r.stderr)
    return name, (r.stderr if verbose else "compiled")


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(lambda n: _compile(n, verbose), SOURCES))
    objs = [os.path.join(OBJ, n.replace(".cu", ".o")) for n in SOURCES]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB, results


if __name__ == "__main__":
    lib, res = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    for n, msg in res:
        print(f"[{n}] {msg}")
    print(lib)
