"""Pin the oracle restatement against the UNMODIFIED reference C++ backend compiled in place
(oracle/_ref/libkfusion_ref.so) — bit for bit, every function, on larger and more varied inputs
than the golden fixtures.  Skipped where oracle/_ref was not built (no /root/reference)."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import K, T0, run_cpu_pipeline
from oracle import cpu_backend as cb
from slambench_b200 import synth


def b32(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def frames():
    return synth.make_sequence(8)[0]


def test_image_kernels_bit_exact(port, ref, frames):
    rng = np.random.default_rng(3)
    d = frames[5].copy()
    d[rng.random(d.shape) < 0.1] = 0
    d[:, :3] = 0
    for size in [(640, 480), (320, 240), (160, 120)]:
        a, b = port.mm2meters(d, size), ref.mm2meters(d, size)
        assert np.array_equal(b32(a), b32(b))
    raw = ref.mm2meters(d, (320, 240))
    g = ref.gaussian()
    assert np.array_equal(b32(port.gaussian()), b32(g))
    fa, fb = port.bilateral(raw, g), ref.bilateral(raw, g)
    assert np.array_equal(b32(fa), b32(fb))
    ha, hb = port.halfsample(fb), ref.halfsample(fb)
    assert np.array_equal(b32(ha), b32(hb))
    invK = ref.inverse_camera_matrix(K / 2)
    assert np.array_equal(b32(port.inverse_camera_matrix(K / 2)), b32(invK))
    va, vb = port.depth2vertex(fb, invK), ref.depth2vertex(fb, invK)
    assert np.array_equal(b32(va), b32(vb))
    init = rng.random((240, 320, 3)).astype(np.float32)     # invalid pixels keep .y/.z of the old buffer
    na, nb = port.vertex2normal(vb, init), ref.vertex2normal(vb, init)
    assert np.array_equal(b32(na), b32(nb))


def test_matrix_helpers_bit_exact(port, ref):
    rng = np.random.default_rng(5)
    for _ in range(50):
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] = synth.rpy_to_R(*rng.uniform(-1, 1, 3)).astype(np.float32)
        m[:3, 3] = rng.uniform(0, 4.8, 3)
        assert np.array_equal(b32(port.inverse(m)), b32(ref.inverse(m)))
        r = rng.normal(size=(4, 4)).astype(np.float32)
        assert np.array_equal(b32(port.inverse(r)), b32(ref.inverse(r)))
        assert np.array_equal(b32(port.matmul(m, r)), b32(ref.matmul(m, r)))
    z = np.zeros((4, 4), np.float32)
    assert np.array_equal(np.isnan(port.inverse(z)), np.isnan(ref.inverse(z)))


def test_solve_and_se3_against_the_toon_stand_in(port, ref):
    """Both sides restate TooN (absent): Jacobi pseudo-inverse (port) vs the stand-in's GR_SVD."""
    rng = np.random.default_rng(7)
    for trial in range(40):
        J = rng.normal(size=(200, 6)) * rng.uniform(0.1, 10, 6)
        if trial % 5 == 0:
            J[:, 4] = 0  # rank-deficient: the pseudo-inverse cut-off (sigma * 1e6 <= sigma_max) must act
        e = rng.normal(size=200)
        JTJ, JTe = J.T @ J, J.T @ e
        vals = np.concatenate([JTe, JTJ[np.triu_indices(6)]]).astype(np.float32)
        xa, xb = port.solve(vals), ref.solve(vals)
        assert np.allclose(xa, xb, rtol=1e-7, atol=1e-9 * max(1.0, np.abs(xb).max())), trial
    for s in (1e-6, 1e-4, 2e-3, 0.05, 1.0, 3.0):
        x = rng.normal(size=6) * s
        assert np.abs(port.se3_exp(x) - ref.se3_exp(x)).max() <= 1.2e-7


def test_volume_kernels_bit_exact(port, ref, frames):
    dim = np.array([4.8] * 3, np.float32)
    N = 96
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = synth.rpy_to_R(0.01, 0.02, -0.015).astype(np.float32)
    pose[:3, 3] = T0 + np.array([0.02, -0.01, 0.015], np.float32)
    raw = ref.mm2meters(frames[6], (640, 480))
    Kmat = ref.camera_matrix(K)
    va, vb = port.init_volume((N,) * 3), ref.init_volume((N,) * 3)
    assert np.array_equal(va, vb)
    for _ in range(2):
        port.integrate(va, dim, raw, port.inverse(pose), Kmat, 0.1)
        ref.integrate(vb, dim, raw, ref.inverse(pose), Kmat, 0.1)
    assert np.array_equal(va, vb)
    assert (vb[..., 1] > 0).mean() > 0.03
    view = ref.matmul(pose, ref.inverse_camera_matrix(K))
    (pv, pn), (rv, rn) = port.raycast(vb, dim, (640, 480), view), ref.raycast(vb, dim, (640, 480), view)
    assert np.array_equal(b32(pv), b32(rv)) and np.array_equal(b32(pn), b32(rn))
    assert (rn[..., 0] != -2).mean() > 0.8   # 96^3 after two integrates: coarse, some rays slip through
    # track + reduce against these maps
    filt = ref.bilateral(raw, ref.gaussian())
    inV = ref.depth2vertex(filt, ref.inverse_camera_matrix(K))
    inN = ref.vertex2normal(inV)
    pose2 = pose.copy()
    pose2[:3, 3] += np.array([0.003, 0.002, -0.002], np.float32)
    proj = ref.matmul(Kmat, ref.inverse(pose))
    ta, tb = port.track(inV, inN, rv, rn, pose2, proj), ref.track(inV, inN, rv, rn, pose2, proj)
    assert ta.tobytes() == tb.tobytes()
    ra, rb = port.reduce(ta, (640, 480)), ref.reduce(tb, (640, 480))
    assert np.array_equal(b32(ra), b32(rb))
    (pa, ca), (pb, cb_) = port.update_pose(pose2, ra), ref.update_pose(pose2, rb)
    assert ca == cb_ and np.abs(pa - pb).max() <= 5e-7
    assert port.check_pose(pa, pose2, ra, (640, 480))[1] == ref.check_pose(pb, pose2, rb, (640, 480))[1]
    # renders
    assert np.array_equal(port.render_depth(raw), ref.render_depth(raw))
    assert np.array_equal(port.render_track(ta), ref.render_track(tb))
    assert np.array_equal(port.render_volume(vb, dim, (320, 240), view), ref.render_volume(vb, dim, (320, 240), view))


def test_whole_pipeline_matches_reference(port, ref, frames):
    """8 frames at -c 2 / 64^3 through both `Kfusion` drivers: same flags, poses within 1e-6."""
    small = frames[:, ::2, ::2].copy()
    k2 = (K / 2).astype(np.float32)
    pa, ta, ia = run_cpu_pipeline(port, small, 8, 64, csize=(320, 240), k=k2)
    pb, tb, ib = run_cpu_pipeline(ref, small, 8, 64, csize=(320, 240), k=k2)
    assert ta == tb and ia == ib
    assert ta == [False] * 4 + [True] * 4
    assert np.abs(pa - pb).max() <= 2e-6
