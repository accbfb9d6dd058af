"""The drop-in boundary: libkfb200.so loads, exports every symbol include/kfb200.h declares, and
fails loudly (no CPU fallback) when there is no CUDA device.  No compute calls here."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import HAS_GPU, ROOT, T0
from slambench_b200 import build as b
from slambench_b200 import kfusion as kf

HEADER = os.path.join(ROOT, "include", "kfb200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(kfb_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_builds_and_exports_every_declared_symbol():
    lib_path = b.build_lib()
    assert os.path.exists(lib_path)
    lib = C.CDLL(lib_path)
    names = declared_functions()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.kfb_abi_version.restype = C.c_int
    assert lib.kfb_abi_version() == 1
    # every exported kfb_ symbol is declared (no undocumented entry points)
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = sorted({l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("kfb_")})
    assert exported == names, set(exported) ^ set(names)


def test_library_is_sm_100a_only():
    out = subprocess.check_output(["cuobjdump", "--list-elf", b.build_lib()], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_config_struct_layout_matches_header():
    """ctypes mirror vs the C struct (compiled with gcc from the header itself)."""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "kfb200.h"
int main(void) {
	printf("%zu %zu %zu %zu %zu %zu\n", sizeof(kfb_config), offsetof(kfb_config, init_pose), offsetof(kfb_config, iterations),
		offsetof(kfb_config, flags), sizeof(kfb_stats), offsetof(kfb_stats, ms_preprocess));
	return 0;
}'''
    exe = "/tmp/kfb_layout_check"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=prog, text=True, check=True)
    got = [int(v) for v in subprocess.check_output([exe], text=True).split()]
    want = [C.sizeof(kf.KfbConfig), kf.KfbConfig.init_pose.offset, kf.KfbConfig.iterations.offset, kf.KfbConfig.flags.offset,
            C.sizeof(kf.KfbStats), kf.KfbStats.ms_preprocess.offset]
    assert got == want


@pytest.mark.skipif(HAS_GPU, reason="this box has a GPU")
def test_no_cpu_fallback_without_a_device():
    with pytest.raises(kf.KfbError, match="no CUDA device|CUDA"):
        kf.Kfusion((640, 480), 32, 4.8, T0, (10, 5, 4))


def test_bad_arguments_are_reported_not_crashed():
    lib = kf.load_library()
    h = C.c_void_p()
    assert lib.kfb_create(None, C.byref(h)) != 0
    cfg = kf.KfbConfig()
    cfg.compute_w, cfg.compute_h = 640, 480
    cfg.n_levels = 9
    assert lib.kfb_create(C.byref(cfg), C.byref(h)) != 0
    assert b"pyramid" in lib.kfb_last_error()
    cfg.n_levels = 3
    assert lib.kfb_create(C.byref(cfg), C.byref(h)) != 0      # empty volume
    assert b"empty" in lib.kfb_last_error()


def test_product_never_imports_the_oracle():
    """The product path (slambench_b200/) must not import, link or dlopen anything under oracle/."""
    pkg = os.path.join(ROOT, "slambench_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "libkfusion_oracle" not in src \
                    and "libkfusion_ref" not in src, f
    out = subprocess.check_output(["ldd", b.build_lib()], text=True)
    assert "oracle" not in out
    # ... nor may any of its build command lines reach into oracle/ (include paths, sources, libraries): record what the
    # build functions hand to the compilers
    calls = []
    real = b.subprocess.check_call
    b.subprocess.check_call = lambda cmd, **kw: calls.append(list(cmd)) or 0
    try:
        b.build_lib(force=True)
        b.build_variant("probe", ["KFB_PROBE=1"])
        if os.path.isdir(b.REFERENCE_ROOT):
            b.build_benchmark(force=True)
    finally:
        b.subprocess.check_call = real
    assert calls
    for cmd in calls:
        assert not any("oracle" in str(a) for a in cmd), cmd
