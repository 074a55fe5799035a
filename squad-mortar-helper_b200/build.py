"""Builds the native libraries of the package in-tree (the .so files travel to the GPU box).

  libsmh_b200.so       CUDA kernels + C ABI (include/smh_b200.h), sm_100a only, static cudart
  libsmh_hostcheck.so  CPU-only build of the shared per-pixel arithmetic (test helper)
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
ROOT = PKG.parent

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # results feed truncating casts: no FMA contraction anywhere
    "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and Path(c).exists():
            return c
    raise RuntimeError("nvcc not found")


def _gxx() -> str:
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper); prefer the system compiler
    return "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def cuda_sources():
    return [CSRC / "smh_b200.cu", CSRC / "kernels.cuh", CSRC / "kernels_lsd.cuh", CSRC / "kernels_px.cuh", CSRC / "kernels_next.cuh", CSRC / "kernels_icons.cuh", CSRC / "smh_host_widen.cpp", CSRC / "smh_pixel.h", CSRC / "smh_synth.h", CSRC / "smh_lsd_filter.h", CSRC / "smh_px_plan.h",
            ROOT / "include" / "smh_b200.h", ROOT / "include" / "smh_consts.h"]


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    out = PKG / "libsmh_b200.so"
    if force or _stale(out, cuda_sources()):
        cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", _gxx(), "-o", str(out), str(CSRC / "smh_b200.cu"), str(CSRC / "smh_host_widen.cpp")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True, cwd=str(CSRC))
    return out


def build_hostcheck(force: bool = False) -> Path:
    out = PKG / "libsmh_hostcheck.so"
    src = [CSRC / "hostcheck.cpp", CSRC / "smh_pixel.h", CSRC / "smh_lsd_filter.h", CSRC / "smh_px_plan.h", ROOT / "include" / "smh_consts.h"]
    if force or _stale(out, src):
        subprocess.run([_gxx(), "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                        str(CSRC / "hostcheck.cpp"), "-o", str(out)], check=True)
    return out


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_hostcheck(force)


if __name__ == "__main__":
    import sys
    build_all(force="--force" in sys.argv)
    print("built", PKG / "libsmh_b200.so")
