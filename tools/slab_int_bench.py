"""Time integrate on ONE z-slab of a big volume (a single GPU stands in for one rank of a z-slab group):
    python tools/slab_int_bench.py [volume] [z0] [z1] [repeats]      (KFB_LIB selects another build of the library)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from slambench_b200 import kfusion as kf, synth

vres = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
z0, z1 = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (928, 944)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
K = np.array(synth.K_DEFAULT, np.float32)
T0 = (np.array(synth.INIT_POS_FACTOR, np.float32) * np.float32(synth.VOLUME_DIM)).astype(np.float32)
depth, poses = synth.make_sequence(10)
with kf.Kfusion((640, 480), vres, 4.8, T0, (10, 5, 4), slab=(z0, z1), flags=kf.FLAG_BRICKS_MERGED) as g:
    g.enable_timing(True)
    for f in range(8):
        g.preprocessing(depth[f])
        g.integrateKernel(g.inverse(poses[f].astype(np.float32)), g.cameraMatrix(K), 0.1)
    g.synchroniseDevices()
    g.reset_stats()
    pose = poses[9].astype(np.float32)
    g.preprocessing(depth[9])
    for _ in range(reps):
        g.integrateKernel(g.inverse(pose), g.cameraMatrix(K), 0.1)
    s = g.stats()
    print(f"{os.environ.get('KFB_LIB', 'default')}: {vres}^3 slab [{z0},{z1}) integrate {s['ms_integrate'] / reps * 1e3:.1f} us/launch, N_upd {s['voxels_updated_last']}")
