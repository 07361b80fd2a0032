"""Builds libgs3d_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

No torch headers are involved: the library's only dependency is the CUDA runtime.  The .so is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libgs3d_b200.so"
SOURCES = ["project.cu", "binning.cu", "composite.cu", "exchange.cu", "bands.cu", "train.cu", "loss.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    return "nvcc"


def is_fresh():
    if not LIB.exists():
        return False
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [CSRC / "common.cuh", PKG.parent / "include" / "gs3d_b200.h"]
    return all(d.stat().st_mtime <= t for d in deps)


def build(force=False, verbose=False, extra_flags=()):
    if not force and is_fresh():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-o", str(LIB)] + [str(CSRC / s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    env = dict(os.environ)
    # the image's default CC (/opt/gcc) is fine for nvcc; nothing else to set
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="-f" in sys.argv, verbose=True,
                extra_flags=("-Xptxas", "-v") if "-v" in sys.argv else ()))
