"""Does the 1000-frame trajectory (synth.trajectory_pose(long_run=True), period 400 frames) track?  One full period at a
small volume on one GPU:  python tools/long_run_check.py [volume] [frames]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from multiprocessing import Pool
from slambench_b200 import synth


def render(f):
    return synth.render_depth_mm(synth.trajectory_pose(f, 0, True))


if __name__ == "__main__":
    vres = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 404
    t0 = time.time()
    with Pool(min(16, os.cpu_count() or 1)) as pool:
        depth = np.stack(pool.map(render, range(n)))
    gt = np.stack([synth.trajectory_pose(f, 0, True) for f in range(n)])
    print(f"rendered {n} frames in {time.time() - t0:.1f} s")
    from slambench_b200 import kfusion as kf
    K = np.array(synth.K_DEFAULT, np.float32)
    T0 = (np.array(synth.INIT_POS_FACTOR, np.float32) * np.float32(synth.VOLUME_DIM)).astype(np.float32)
    with kf.Kfusion((640, 480), vres, 4.8, T0, (10, 5, 4)) as g:
        tracked, worst = 0, 0.0
        t0 = time.time()
        for f in range(n):
            g.preprocessing(depth[f]); tr = g.tracking(K, 1e-5, 1, f); g.integration(K, 1, 0.1, f); g.raycasting(K, 0.1, f)
            tracked += int(tr)
            if f % 50 == 0 or f == n - 1:
                err = float(np.abs(g.getPose()[:3, 3] - synth.expected_pose(gt, f)[:3, 3]).max())
                worst = max(worst, err)
                print(f"frame {f}: tracked so far {tracked}, position error {err * 1e3:.2f} mm")
        g.synchroniseDevices()
        print(f"{vres}^3: {n} frames, tracked {tracked}/{n - 4}, worst sampled error {worst * 1e3:.2f} mm, {n / (time.time() - t0):.0f} fps wall")
