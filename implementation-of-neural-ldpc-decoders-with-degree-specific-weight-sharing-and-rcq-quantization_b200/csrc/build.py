"""In-tree build of the CUDA library: nvcc -> ../libldpc_b200.so (sm_100a only, no torch headers)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SOURCES = [os.path.join(HERE, "ldpc_kernels.cu"), os.path.join(HERE, "ldpc_api.cu")]
HEADERS = [os.path.join(HERE, "ldpc_device.cuh"), os.path.join(HERE, "ldpc_internal.h"),
           os.path.join(os.path.dirname(PKG), "include", "ldpc_b200.h")]
OUT = os.path.join(PKG, "libldpc_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",               # parity: llr + alpha*s and beta*raw round the product first
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found")
    return cand


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(p) <= t for p in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    subprocess.run(cmd, check=True)
    return OUT


def build_variant(name: str, defines: list) -> str:
    """Tuning builds (kernel-geometry A/B on the GPU box): <repo>/tuning/libldpc_b200_<name>.so, selected
    at run time with LDPC_B200_LIB=<path>.  Not used by the product path."""
    outdir = os.path.join(os.path.dirname(PKG), "tuning")
    os.makedirs(outdir, exist_ok=True)
    out = os.path.join(outdir, f"libldpc_b200_{name}.so")
    cmd = [nvcc_path()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", out] + SOURCES
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 2 and sys.argv[1] == "variant":      # build.py variant NAME DEF1=V1 DEF2=V2 ...
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        print(build(force=True, verbose="-v" in sys.argv))
