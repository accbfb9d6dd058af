"""A/B of the side-stream overlap in one process: python tools/overlap_ab.py [volume] [frames]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from multiprocessing import Pool
from slambench_b200 import synth


def render(f):
    return synth.render_depth_mm(synth.trajectory_pose(f, 0, False))


if __name__ == "__main__":
    vres = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    with Pool(min(16, os.cpu_count() or 1)) as pool:
        depth = np.stack(pool.map(render, range(n)))
    import torch
    host = torch.from_numpy(depth).pin_memory()
    depth = host.numpy()
    from slambench_b200 import kfusion as kf
    K = np.array(synth.K_DEFAULT, np.float32)
    T0 = (np.array(synth.INIT_POS_FACTOR, np.float32) * np.float32(synth.VOLUME_DIM)).astype(np.float32)
    out = {}
    for rep in range(2):
        for no in ("1", "0"):
            os.environ["KFB_NO_OVERLAP"] = no
            with kf.Kfusion((640, 480), vres, 4.8, T0, (10, 5, 4)) as g:
                for f in range(4):
                    g.preprocessing(depth[f]); g.tracking(K, 1e-5, 1, f); g.integration(K, 1, 0.1, f); g.raycasting(K, 0.1, f)
                g.synchroniseDevices()
                t0 = time.perf_counter()
                tracked = 0
                for f in range(4, n):
                    g.preprocessing(depth[f]); tracked += g.tracking(K, 1e-5, 1, f); g.integration(K, 1, 0.1, f); g.raycasting(K, 0.1, f)
                g.synchroniseDevices()
                dt = time.perf_counter() - t0
                out[no] = (g.getPose().copy(), g.read(kf.BUF_VOLUME).copy())
                print(f"rep {rep} no_overlap={no}: {(n - 4) / dt:.0f} fps, tracked {tracked}/{n - 4}")
    same = np.array_equal(out["0"][0], out["1"][0]) and np.array_equal(out["0"][1], out["1"][1])
    print("pose and volume bit-identical with and without the overlap:", same)
