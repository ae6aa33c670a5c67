"""Build libedadm.so in-tree with nvcc for sm_100a (no torch extension ABI: plain C symbols)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libedadm.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--use_fast_math=false"]
# per-file extras: the quantizer files must reproduce torch's separate mul/add roundings
EXTRA = {"elementwise.cu": ["-fmad=false"], "optim.cu": ["-fmad=false"], "pack.cu": ["-fmad=false"], "qgemm2_sm100.cu": ["-fmad=false"], "gemm_bf16x3_sm100.cu": ["-fmad=false"]}
SOURCES = ["common.cu", "elementwise.cu", "optim.cu", "pack.cu", "qgemm_sm100.cu", "qgemm2_sm100.cu", "gemm_bf16x3_sm100.cu", "qattn.cu", "conv_small.cu"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(deps):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    for src in srcs:
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < _newest([src] + [d for d in deps if d.endswith((".cuh", ".h"))]):
            cmd = [nvcc] + ARCH + [c for c in COMMON if not c.startswith("--use_fast_math")] + EXTRA.get(os.path.basename(src), []) + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
        objs.append(obj)
    cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
