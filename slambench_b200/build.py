"""In-tree build of the CUDA product (libkfb200.so) and, where the reference tree is
mounted, of the drop-in `kfusion-benchmark-b200` binary.

Everything is compiled for sm_100a only.  Parity-critical flags:
  --fmad=false        a*b+c stays two IEEE roundings (the reference targets baseline x86-64: no FMA)
  -prec-div/-prec-sqrt=true, -ftz=false   correctly rounded / and sqrtf, denormals kept
  -Xcompiler -ffp-contract=off            same for the host-side pose algebra
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libkfb200.so")
REFERENCE_ROOT = "/root/reference"
BUILD_DIR = os.path.join(ROOT, "build")
BENCH_BIN = os.path.join(BUILD_DIR, "kfusion-benchmark-b200")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off",
]


def nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def sources() -> list[str]:
    out = [os.path.join(ROOT, "include", "kfb200.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if force or _stale(LIB, sources()):
        cmd = [nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB, os.path.join(CSRC, "kfb_api.cu"), "-ldl"]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        subprocess.check_call(cmd, cwd=CSRC)
    return LIB


def build_variant(name: str, defines: list[str]) -> str:
    """A tuning variant of the same library (extra -D macros) next to the product: libkfb200_<name>.so.  Selected at run
    time with KFB_LIB=<path> (kfusion.load_library); never loaded by default."""
    out = os.path.join(PKG, f"libkfb200_{name}.so")
    cmd = [nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-shared", "-o", out, os.path.join(CSRC, "kfb_api.cu"), "-ldl"]
    subprocess.check_call(cmd, cwd=CSRC)
    return out


def build_benchmark(force: bool = False) -> str | None:
    """`kfusion-benchmark-b200`: the reference's UNMODIFIED benchmark.cpp + PowerMonitor.cpp linked
    against our Kfusion backend glue (csrc/kfusion_b200.cpp) and libkfb200.so.  Needs the reference
    headers, so it is only (re)built where /root/reference is mounted; the binary travels to the
    GPU box inside build/."""
    glue = os.path.join(CSRC, "kfusion_b200.cpp")
    if not os.path.isdir(REFERENCE_ROOT) or not os.path.exists(glue):
        return BENCH_BIN if os.path.exists(BENCH_BIN) else None
    build_lib()
    if not (force or _stale(BENCH_BIN, [glue, LIB])):
        return BENCH_BIN
    os.makedirs(BUILD_DIR, exist_ok=True)
    ref = os.path.join(REFERENCE_ROOT, "kfusion")
    cmd = [
        "g++", "-g", "-O3", "-std=gnu++11", "-w", "-ffp-contract=off",
        "-I", os.path.join(ROOT, "third_party_shims"), "-I", os.path.join(ref, "include"),
        "-I", os.path.join(ref, "thirdparty"), "-I", os.path.join(ROOT, "include"),
        os.path.join(ref, "src", "benchmark.cpp"), os.path.join(ref, "src", "PowerMonitor.cpp"), glue,
        "-o", BENCH_BIN, "-L", PKG, "-lkfb200", "-Wl,-rpath,$ORIGIN/../slambench_b200", "-lrt", "-lpthread",
    ]
    subprocess.check_call(cmd)
    return BENCH_BIN


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
    print(build_benchmark(force=True))
