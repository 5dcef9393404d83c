"""In-tree build of the CUDA library: nvcc -> <repo>/ldpc_b200/libldpc_b200.so (sm_100a only, no torch headers).

The library lives next to the short import alias (``ldpc_b200/``) rather than inside the package directory:
its path then stays short enough for every tool that lists the shared objects a process has mapped."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.realpath(__file__))
PKG = os.path.dirname(HERE)
SOURCES = [os.path.join(HERE, f) for f in ("ldpc_cn.cu", "ldpc_vn.cu", "ldpc_misc.cu", "ldpc_small.cu", "ldpc_resident.cu", "ldpc_train.cu", "ldpc_api.cu")]
HEADERS = [os.path.join(HERE, "ldpc_device.cuh"), os.path.join(HERE, "ldpc_kernel_common.cuh"), os.path.join(HERE, "ldpc_cn_common.cuh"),
           os.path.join(HERE, "ldpc_internal.h"), os.path.join(os.path.dirname(PKG), "include", "ldpc_b200.h")]
OBJDIR = os.path.join(HERE, "build")
OUT = os.path.join(os.path.dirname(PKG), "ldpc_b200", "libldpc_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",               # parity: llr + alpha*s and beta*raw round the product first
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
]
# the CUDA runtime is linked dynamically (the image's libcudart.so.12; a process that imported torch first shares
# torch's copy): the shipped artefact then carries no runtime symbol table of its own
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared",
              "-Xlinker", "-rpath,/usr/local/cuda/lib64"]


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


def _compile_all(out: str, defines, verbose: bool, tag: str) -> str:
    """One nvcc process per translation unit (they run in parallel), then one link step."""
    objdir = os.path.join(OBJDIR, tag)
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        cmd = ([nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) +
               ["-c", "-o", obj, src])
        procs.append((cmd, obj, subprocess.Popen(cmd)))
    objs = []
    for cmd, obj, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
        objs.append(obj)
    subprocess.run([nvcc] + LINK_FLAGS + ["-o", out] + objs, check=True)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    return _compile_all(OUT, [], verbose, "main")


def build_variant(name: str, defines: list) -> str:
    """Tuning builds (kernel-geometry A/B on the GPU box): <repo>/tuning/libldpc_b200_<name>.so, selected
    at run time with LDPC_B200_LIB=<path>.  Not used by the product path."""
    outdir = os.path.join(os.path.dirname(PKG), "tuning")
    os.makedirs(outdir, exist_ok=True)
    return _compile_all(os.path.join(outdir, f"libldpc_b200_{name}.so"), list(defines), False, "variant_" + name)


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 2 and sys.argv[1] == "variant":      # build.py variant NAME DEF1=V1 DEF2=V2 ...
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        print(build(force=True, verbose="-v" in sys.argv))
