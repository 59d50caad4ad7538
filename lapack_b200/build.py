"""Build liblapack_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "liblapack_b200.so")
OBJ = os.path.join(HERE, "build")

# never --use_fast_math: FP64 parity with the reference
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=default"]


def sources(minimal: bool = False):
    if minimal:
        return ["gemm_f64.cu", "gemm_tma.cu", "runtime.cu", "capi.cu"]
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(minimal: bool = False, force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs += [os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))]
    hdr_time = max(os.path.getmtime(h) for h in hdrs)
    objs, procs = [], []
    for src in sources(minimal):
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ, src.replace(".cu", ".min.o" if minimal else ".o"))
        objs.append(op)
        if force or not os.path.exists(op) or os.path.getmtime(op) < max(os.path.getmtime(sp), hdr_time):
            cmd = ["nvcc", *NVCC_FLAGS, "-c", sp, "-o", op] + (["-DLB_MINIMAL"] if minimal else [])
            if verbose:
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src}\n{out.decode()}\n")
        elif verbose and out:
            sys.stderr.write(out.decode())
    if failed:
        raise RuntimeError("nvcc build failed")
    if procs or not os.path.exists(OUT) or force:
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *objs,
               "-cudart", "static"]
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(minimal="--minimal" in sys.argv, force="--force" in sys.argv, verbose=True))
