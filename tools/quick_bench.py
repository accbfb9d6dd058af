"""Quick stage-timing probe (development aid): python tools/quick_bench.py [vres] [frames]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from slambench_b200 import synth, kfusion as kf

vres = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 30
K = np.array(synth.K_DEFAULT, np.float32)
T0 = (np.array(synth.INIT_POS_FACTOR, np.float32) * np.float32(4.8)).astype(np.float32)
t = time.time(); depth, gt = synth.make_sequence(nf); print("synth %.1fs" % (time.time() - t))
for rep in range(2):
    with kf.Kfusion((640, 480), vres, 4.8, T0, (10, 5, 4)) as g:
        g.enable_timing(True); g.reset_stats()
        t0 = time.time(); tr_n = 0
        for f in range(nf):
            g.preprocessing(depth[f]); tr = g.tracking(K, 1e-5, 1, f); it = g.integration(K, 1, 0.1, f); g.raycasting(K, 0.1, f); tr_n += tr
        g.synchroniseDevices(); wall = time.time() - t0
        st = g.stats()
        err = np.abs(g.getPose()[:3, 3] - gt[nf - 1][:3, 3]).max()
        print(f"vres {vres} rep {rep}: wall {wall*1e3/nf:.3f} ms/frame ({nf/wall:.1f} fps) tracked {tr_n}/{nf} err {err:.4f} m")
        print(f"  per-frame GPU ms: pre {st['ms_preprocess']/nf:.4f} track {st['ms_track']/nf:.4f} integ {st['ms_integrate']/nf:.4f} ray {st['ms_raycast']/max(1,nf-3):.4f}")
        print(f"  icp iters/frame {st['icp_iterations_total']/nf:.1f} launches {st['kernel_launches']} N_upd last {st['voxels_updated_last']} ({st['voxels_updated_last']/vres**3:.3f} of volume)")
