"""The synthetic `.raw` sequence generator."""
from __future__ import annotations

import numpy as np

from slambench_b200 import synth


def test_raw_container_round_trip(tmp_path):
    depth, _ = synth.make_sequence(3)
    p = str(tmp_path / "seq.raw")
    synth.write_raw(p, depth)
    import os

    # scene2raw.cpp:170-176 layout: (8 + w*h*2 + 8 + w*h*3) bytes per frame
    assert os.path.getsize(p) == 3 * (8 + 640 * 480 * 2 + 8 + 640 * 480 * 3) == 3 * 1536016
    assert np.array_equal(synth.read_raw(p), depth)
    hdr = np.fromfile(p, dtype=np.uint32, count=2)
    assert tuple(hdr) == (640, 480)


def test_sequence_properties():
    depth, gt = synth.make_sequence(8)
    assert depth.dtype == np.uint16 and depth.shape == (8, 480, 640)
    # frames 0-3 static (the reference cannot track before its first raycast)
    for f in range(1, 4):
        assert np.array_equal(depth[f], depth[0]) and np.array_equal(gt[f], gt[0])
    assert not np.array_equal(depth[4], depth[3])
    assert depth.min() > 400 and depth.max() < 4000          # inside the raycaster's [0.4, 4.0] m range, no invalid px
    step = np.linalg.norm(gt[5][:3, 3] - gt[4][:3, 3])
    assert 0.003 < step < 0.008
    # deterministic
    d2, _ = synth.make_sequence(8)
    assert np.array_equal(depth, d2)
    # eight distinct trajectories for the one-sequence-per-GPU mode
    ends = {tuple(np.round(synth.trajectory_pose(50, s)[:3, 3], 4)) for s in range(8)}
    assert len(ends) == 8


def test_long_run_stays_inside_the_room():
    for f in range(0, 1000, 37):
        t = synth.trajectory_pose(f, 0, long_run=True)[:3, 3]
        assert np.all(t > synth.ROOM_LO + 0.3) and np.all(t < synth.ROOM_HI - 0.3)


def test_expected_pose_is_ground_truth_in_the_trackers_frame():
    """A tracker starts at identity rotation (kernels.h:106-109); the 1000-frame trajectory does not, so its ground
    truth is re-expressed in the tracker's frame.  Short sequences start at the initial pose: nothing changes."""
    _, gt = synth.make_sequence(8, long_run=False)
    assert np.allclose(synth.expected_pose(gt, 7), gt[7])
    gl = np.stack([synth.trajectory_pose(f, 0, True) for f in (0, 120)])
    e = synth.expected_pose(gl, 1)
    assert np.allclose(synth.expected_pose(gl, 0), np.block([[np.eye(3), gl[0][:3, 3:4]], [np.zeros((1, 3)), np.ones((1, 1))]]))
    # relative motion is preserved: inverse(first) * frame is the same in both frames
    p0 = synth.expected_pose(gl, 0)
    assert np.allclose(np.linalg.inv(p0) @ e, np.linalg.inv(gl[0]) @ gl[1])
    assert not np.allclose(e[:3, 3], gl[1][:3, 3], atol=1e-3)
