"""Builds libtsdf_b200.so (sm_100a only) in-tree with nvcc.

    python -m union_thesis_slam_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtsdf_b200.so")
SOURCES = ["capi_grid.cu", "dense_kernels.cu", "capi_hash.cu", "hash_kernels.cu", "mesh_kernels.cu", "bucket_kernels.cu", "capi_fanout.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    # the arithmetic contract is spelled out with explicit intrinsics; never let nvcc fuse on its own
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared", "-Xptxas", "-v",
]


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "tsdf_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale() and "TSDF_LIB_OUT" not in os.environ:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("TSDF_NVCC_EXTRA", "").split()  # e.g. -DINTEGRATE_STAGED=0 for A/B runs
    out = os.environ.get("TSDF_LIB_OUT", LIB)  # variants for A/B runs are built beside the product library
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", out] + sources()
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(os.path.join(HERE, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + res.stdout)
    if res.returncode != 0 or verbose:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed (see union_thesis_slam_b200/build.log)")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
