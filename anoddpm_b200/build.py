"""In-tree nvcc build of libanoddpm_b200.so for sm_100a (no JIT cache: the .so ships with the repo snapshot).

    python -m anoddpm_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libanoddpm_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
# exact-arithmetic translation units: no fma contraction so fp64/fp32 results match the CPU reference bit for bit
EXACT = {"simplex.cu", "diffusion.cu", "optim.cu"}
SOURCES = ["abi.cu", "simplex.cu", "diffusion.cu", "unet_ops.cu", "conv_tc.cu", "conv_tc2.cu", "conv_tc3.cu", "attention_tc.cu",
           "optim.cu", "conv_wgrad.cu", "unet_bwd.cu", "edge_convs.cu"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    hdrs = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "diffusion_step.cuh"), os.path.join(ROOT, "include", "anoddpm_b200.h"), __file__]
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(bdir, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC, *ARCH, *COMMON, "-c", src, "-o", obj]
            if s in EXACT:
                cmd += ["-fmad=false"]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            print("[build]", " ".join(cmd[-6:]), flush=True)
            subprocess.check_call(cmd)
    if force or _stale(OUT, objs):
        subprocess.check_call([NVCC, *ARCH, "-shared", "-o", OUT, *objs, "-lcudart"])
        print("[build] linked", OUT, flush=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
