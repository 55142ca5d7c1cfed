"""Builds the C-ABI shared library (soccdpt_b200/_lib/libsoccdpt_b200.so) with nvcc for sm_100a.

    python -m soccdpt_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is built IN-TREE so that it travels to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB_DIR = os.path.join(PKG, "_lib")
LIB = os.path.join(LIB_DIR, "libsoccdpt_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
          "-I", CSRC, "--expt-relaxed-constexpr"]
# per-file extras: the bit-exact post-processing spells its roundings with explicit _rn intrinsics (never contracted);
# IEEE division / sqrt and no flush-to-zero are required there
EXTRA = {"postprocess.cu": ["-prec-div=true", "-prec-sqrt=true", "-ftz=false"]}


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    m = os.path.getmtime(os.path.join(ROOT, "include", "soccdpt_b200.h"))
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".h")):
            m = max(m, os.path.getmtime(os.path.join(CSRC, f)))
    return m


def _compile(src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    cmd = [NVCC, *ARCH, *COMMON, *EXTRA.get(src, []), *os.environ.get("SOCCDPT_NVCC_FLAGS", "").split(),
           "-c", os.path.join(CSRC, src), "-o", obj]     # SOCCDPT_NVCC_FLAGS: experiment switches (-DSOCCDPT_...), with --force
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = _sources()
    hdr_m = _headers_mtime()
    todo, objs = [], []
    for s in srcs:
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        objs.append(obj)
        if (force or not os.path.exists(obj)
                or os.path.getmtime(obj) < max(os.path.getmtime(os.path.join(CSRC, s)), hdr_m)):
            todo.append(s)
    logs = []
    if todo:
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for obj, log in ex.map(lambda s: _compile(s, verbose), todo):
                logs.append(log)
    if todo or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
