"""Build csrc/*.cu into csrc/librgbd_b200.so for sm_100a with nvcc (cross-compiles without a GPU).

Bit-exact integer/IEEE files (decompose.cu, gradfeat.cu) are compiled with -fmad=false so no
multiply-add is contracted; everything else uses the default.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(CSRC, "librgbd_b200.so")

SOURCES = ["api.cu", "dggm.cu", "gradfeat.cu", "decompose.cu", "pack.cu", "conv_gemm.cu", "ratio_chain.cu", "ratio_front.cu", "ratio_tail.cu", "ratio_feat.cu", "dsam_bwd.cu", "postproc.cu", "resize.cu", "msda.cu", "maskattn.cu", "winattn.cu", "layernorm.cu"]
NO_FMAD = {"gradfeat.cu", "decompose.cu", "postproc.cu", "resize.cu", "maskattn.cu"}
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(CSRC, name), "rb").read())
    h.update(open(os.path.join(INCLUDE, "rgbd_b200.h"), "rb").read())
    # flags and toolkit are part of the identity: the bit-exact kernels depend on -fmad=false and on the compiler
    h.update(repr((ARCH, COMMON, sorted(NO_FMAD), SOURCES)).encode())
    try:
        h.update(subprocess.run([_nvcc(), "--version"], capture_output=True, text=True).stdout.encode())
    except OSError:
        pass
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = os.path.join(CSRC, ".build_stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = _nvcc()
    objs = []

    def compile_one(src):
        obj = os.path.join(CSRC, src[:-3] + ".o")
        cmd = [nvcc, *ARCH, *COMMON, "-c", os.path.join(CSRC, src), "-o", obj]
        if src in NO_FMAD:
            cmd.insert(1, "-fmad=false")
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
