"""In-tree build recipe: nvcc for the CUDA library (sm_100a only), g++ for the CPython host module.

    python loco_hd_b200/build.py            # build what is stale (run as a script: importing the package needs
    python loco_hd_b200/build.py --force    # the built extension, so `-m loco_hd_b200.build` only works afterwards)

Outputs (git-ignored, shipped to the GPU box with the tree):
    loco_hd_b200/liblocohd_b200.so                 C ABI + kernels   (include/locohd_b200.h)
    loco_hd/loco_hd.cpython-312-x86_64-linux-gnu.so  CPython host module at the reference's module path
                                                   `loco_hd.loco_hd` (replaces src/lib.rs of the reference)
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
CUDA_LIB = PKG / "liblocohd_b200.so"
HOST_MOD = ROOT / "loco_hd" / ("loco_hd" + sysconfig.get_config_var("EXT_SUFFIX"))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build_cuda_lib(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / "locohd_kernels.cu", CSRC / "locohd_capi.cu"]
    deps = srcs + [CSRC / "locohd_kernels.cuh", CSRC / "locohd_math.cuh", ROOT / "include" / "locohd_b200.h"]
    if force or _stale(CUDA_LIB, deps):
        cmd = [NVCC, *NVCC_FLAGS, "-o", str(CUDA_LIB), *map(str, srcs)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return CUDA_LIB


def build_host_module(force: bool = False, verbose: bool = False) -> Path:
    src = CSRC / "host_module.cpp"
    deps = [src, ROOT / "include" / "locohd_b200.h"]
    if force or _stale(HOST_MOD, deps) or _stale(HOST_MOD, [CUDA_LIB]):
        import pybind11

        cmd = [
            "g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
            f"-I{pybind11.get_include()}", f"-I{sysconfig.get_paths()['include']}", f"-I{ROOT / 'include'}",
            str(src), "-o", str(HOST_MOD),
            f"-L{PKG}", "-llocohd_b200", "-Wl,-rpath,$ORIGIN/../loco_hd_b200",
        ]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return HOST_MOD


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda_lib(force, verbose)
    build_host_module(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
