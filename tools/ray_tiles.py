"""Distribution of the raycaster's per-tile cost (SM cycles per 8x4-pixel tile): python tools/ray_tiles.py [volume] [frames]   (RAY_TILES_OUT=<file.npy> keeps the array)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from slambench_b200 import kfusion as kf, synth

vres = int(sys.argv[1]) if len(sys.argv) > 1 else 512
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 30
K = np.array(synth.K_DEFAULT, np.float32)
T0 = (np.array(synth.INIT_POS_FACTOR, np.float32) * np.float32(synth.VOLUME_DIM)).astype(np.float32)
depth, _ = synth.make_sequence(frames)
with kf.Kfusion((640, 480), vres, 4.8, T0, (10, 5, 4)) as g:
    for f in range(frames):
        g.preprocessing(depth[f]); g.tracking(K, 1e-5, 1, f); g.integration(K, 1, 0.1, f); g.raycasting(K, 0.1, f)
    g.synchroniseDevices()
    c = g.read(kf.BUF_RAYTILECOST).astype(np.float64)
    if os.environ.get("RAY_TILES_OUT"): np.save(os.environ["RAY_TILES_OUT"], c)
    warps = 148 * 36
    print(f"{vres}^3 frame {frames - 1}: tiles {c.size}, sum {c.sum() / 1e6:.1f} Mcycles = {c.sum() / warps / 1.965e3:.1f} us per warp slot at 1.965 GHz")
    print("percentiles (kcycles): " + ", ".join(f"p{q}={np.percentile(c, q) / 1e3:.1f}" for q in (50, 90, 99, 99.9, 100)))
    rows = c.max(axis=1)
    worst = np.argsort(rows)[-5:]
    print("slowest tile rows (row*4 = pixel y):", [(int(r) * 4, round(rows[r] / 1e3, 1)) for r in worst])
    print("slowest tile: %.1f us" % (c.max() / 1.965e3))
