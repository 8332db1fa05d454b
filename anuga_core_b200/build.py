"""Build libswk.so (the CUDA kernels + C ABI) in-tree for sm_100a."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "swk_api.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "swk_kernels.cuh"), os.path.join(HERE, "csrc", "swk_math.cuh"),
        os.path.join(HERE, "..", "include", "swk.h")]
OUT = os.path.join(HERE, "libswk.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    # parity build: no FMA contraction on the device or in the host-side geometry set-up
    "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-shared", "-ldl",
]


def is_fresh():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force=False, verbose=False):
    if is_fresh() and not force:
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libswk.so")
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
