"""Builds the product's CUDA library in-tree: multi_orb_slam_b200/liborb_b200.so (sm_100a only).

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
repo snapshot.  Run as `python -m multi_orb_slam_b200.build` or via __graft_entry__.build()."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# development knobs (kernel A/B runs): ORB_B200_LIB = path of the library to build / load instead of the in-tree
# default, ORB_B200_NVCC_FLAGS = extra nvcc flags (e.g. -DBF_QPT=4)
LIB = os.environ.get("ORB_B200_LIB") or os.path.join(HERE, "liborb_b200.so")
SOURCES = ["extract_kernels.cu", "extract_api.cu", "matcher.cu", "dist_api.cu", "pipeline_api.cu"]
HEADERS = ["device_guard.h", "extract_kernels.h", "orb_geom.h", "octree_core.h", "orb_pattern.h", "../../include/orb_b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              # host-side float projections must not be contracted into FMAs either (--fmad covers device code only;
              # every device float op on the parity path is an explicit __f*_rn intrinsic)
              "-Xcompiler", "-ffp-contract=off", "-shared", "-cudart", "shared", "-ldl"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get("ORB_B200_NVCC_FLAGS", "").split() + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building liborb_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
