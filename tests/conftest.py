"""Shared fixtures.  `-m "not gpu"` runs here on the CPU box; `-m gpu` runs on a B200.

The oracle (oracle/) is loaded ONLY from this test tree (and smoke()/bench.py's CPU legs):
it is the checker, never the product.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cpu_backend as cb  # noqa: E402
from slambench_b200 import synth  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s on the CPU box")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


K = np.array(synth.K_DEFAULT, np.float32)
T0 = (np.array(synth.INIT_POS_FACTOR, np.float32) * np.float32(synth.VOLUME_DIM)).astype(np.float32)


@pytest.fixture(scope="session")
def port() -> cb.CpuKfusion:
    """Our plain-C restatement of the reference (oracle/kfusion_oracle.c)."""
    cb.build_port()
    return cb.CpuKfusion(cb.PORT_LIB)


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference C++ backend compiled in place (only where it was built)."""
    if os.path.isdir(cb.REFERENCE_ROOT):
        cb.build_ref()
    if not cb.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return cb.CpuKfusion(cb.REF_LIB)


@pytest.fixture(scope="session")
def seq16():
    """First 16 frames of the synthetic 640x480 sequence: (uint16[n,480,640], gt poses)."""
    return synth.make_sequence(16)


def run_cpu_pipeline(backend: cb.CpuKfusion, depth, n_frames, vres, mu=0.1, csize=(640, 480), k=K, vdim=4.8,
                     pyramid=(10, 5, 4), on_frame=None):
    """Drive a CPU backend exactly like benchmark.cpp:125-150 does. Returns poses-after-frame, tracked, integrated."""
    backend.create(csize, vres, vdim, T0, pyramid)
    poses, tracked, integrated = [], [], []
    try:
        for f in range(n_frames):
            backend.preprocessing(depth[f])
            tr = backend.tracking(k, 1e-5, 1, f)
            it = backend.integration(k, 1, mu, f)
            backend.raycasting(k, mu, f)
            poses.append(backend.get_pose().copy())
            tracked.append(tr)
            integrated.append(it)
            if on_frame is not None:
                on_frame(f, backend)
    finally:
        backend.destroy()
    return np.stack(poses), tracked, integrated
