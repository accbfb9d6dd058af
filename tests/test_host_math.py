"""Host-side arithmetic of the product (slambench_b200/csrc/kfb_hostmath.h, kfb_expf.h) — pure host
functions of libkfb200.so, no device needed — against the oracle."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import K, ROOT
from slambench_b200 import kfusion as kf
from slambench_b200 import synth


@pytest.fixture(scope="module")
def lib():
    return kf.load_library()


def _call16(fn, *arrs):
    out = np.empty(16, np.float32)
    fn(out.ctypes.data_as(C.c_void_p), *[np.ascontiguousarray(a, np.float32).ctypes.data_as(C.c_void_p) for a in arrs])
    return out.reshape(4, 4)


def test_matrix_helpers_bit_exact_vs_oracle(lib, port):
    rng = np.random.default_rng(11)
    for _ in range(100):
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] = synth.rpy_to_R(*rng.uniform(-1, 1, 3)).astype(np.float32)
        m[:3, 3] = rng.uniform(0, 4.8, 3)
        r = rng.normal(size=(4, 4)).astype(np.float32)
        for a in (m, r):
            assert np.array_equal(_call16(lib.kfb_inverse4, a).view(np.uint32), port.inverse(a).view(np.uint32))
        assert np.array_equal(_call16(lib.kfb_matmul4, m, r).view(np.uint32), port.matmul(m, r).view(np.uint32))
    assert np.array_equal(_call16(lib.kfb_camera_matrix, K), port.camera_matrix(K))
    assert np.array_equal(_call16(lib.kfb_inverse_camera_matrix, K).view(np.uint32), port.inverse_camera_matrix(K).view(np.uint32))
    z = _call16(lib.kfb_inverse4, np.zeros((4, 4), np.float32))
    assert np.isnan(z).any() and np.array_equal(np.isnan(z), np.isnan(port.inverse(np.zeros((4, 4), np.float32))))


def test_update_and_check_pose_vs_oracle(lib, port):
    rng = np.random.default_rng(13)
    for trial in range(60):
        J = rng.normal(size=(500, 6)) * rng.uniform(0.2, 3, 6)
        e = rng.normal(size=500) * 10 ** rng.uniform(-5, -2)
        red = np.zeros((8, 32), np.float32)
        JTJ = J.T @ J
        red[0, 0] = (e * e).sum()
        red[0, 1:7] = J.T @ e
        red[0, 7:28] = JTJ[np.triu_indices(6)]
        red[0, 28] = 500 if trial % 3 else 30
        pose = np.eye(4, dtype=np.float32)
        pose[:3, 3] = rng.uniform(1, 3, 3)
        want, wc = port.update_pose(pose, red, 1e-5)
        got = pose.reshape(16).copy()
        conv = C.c_int(0)
        lib.kfb_k_update_pose(got.ctypes.data_as(C.c_void_p), red[0].ctypes.data_as(C.c_void_p), C.c_float(1e-5), C.byref(conv))
        assert bool(conv.value) == wc
        assert np.abs(got.reshape(4, 4) - want).max() <= 5e-7
        w2, ok2 = port.check_pose(want, pose, red, (64, 48))
        g2 = want.reshape(16).copy()
        ok = C.c_int(0)
        lib.kfb_k_check_pose(g2.ctypes.data_as(C.c_void_p), pose.ctypes.data_as(C.c_void_p), red[0].ctypes.data_as(C.c_void_p),
                             C.c_uint32(64), C.c_uint32(48), C.c_float(0.15), C.byref(ok))
        assert bool(ok.value) == ok2 and np.array_equal(g2.reshape(4, 4), w2)
    # nothing tracked: 0/0 = NaN compares false, the inlier-ratio test rejects (start-up frames)
    red = np.zeros(32, np.float32)
    p = np.eye(4, dtype=np.float32).reshape(16)
    ok = C.c_int(1)
    lib.kfb_k_check_pose(p.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), red.ctypes.data_as(C.c_void_p),
                         C.c_uint32(640), C.c_uint32(480), C.c_float(0.15), C.byref(ok))
    assert ok.value == 0


def _build_expf_check():
    exe = "/tmp/kfb_expf_exhaustive"
    src = os.path.join(ROOT, "tests", "native", "expf_exhaustive.c")
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-I", os.path.join(ROOT, "slambench_b200", "csrc"),
                           src, "-o", exe, "-lm"])
    return exe


def test_expf_strided_vs_libm():
    """kfb_expf_nonpos == glibc expf on every 61st float of [-104.5, 0] (~18 M values)."""
    out = subprocess.check_output([_build_expf_check(), "61"], text=True)
    assert out.startswith("mismatches 0 of"), out


@pytest.mark.slow
def test_expf_exhaustive():
    """All 1.1e9 floats of [-104.5, 0]."""
    out = subprocess.check_output([_build_expf_check(), "1"], text=True)
    assert out.startswith("mismatches 0 of"), out
