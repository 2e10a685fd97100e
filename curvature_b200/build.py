"""Build libcurvature_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

    python -m curvature_b200.build [--force]

The library is a plain C-ABI shared object (include/curvature_b200.h); Python binds it with
ctypes (curvature_b200/_native.py).  Built artefacts are git-ignored but travel to the GPU box.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcurvature_b200.so")
STAMP = os.path.join(HERE, ".build_stamp")
SOURCES = ["api.cu", "syrk_simt.cu", "syrk_tc.cu", "gemm_simt.cu", "gemm_tc.cu", "gemm_chain.cu", "elementwise.cu", "chol.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
         "-Xcompiler", "-fPIC", "-Xptxas=-v"]


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode() + b"\0" + f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        failed |= p.returncode != 0
    text = "\n".join(log)
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(text)
    if failed:
        sys.stderr.write(text)
        raise RuntimeError("nvcc failed (see curvature_b200/build.log)")
    if verbose:
        print(text)
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"]
    subprocess.check_call(cmd)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
