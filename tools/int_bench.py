"""Time the integrate stage alone (plan + run [+ pyramid]) on a warmed-up volume:
    python tools/int_bench.py [volume] [repeats]
The volume is first built with the default library's full pipeline for 30 frames, saved, then each repeat writes nothing
back: the same frame is integrated again and again at the ground-truth pose (timing only, not a parity run)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from slambench_b200 import kfusion as kf, synth

vres = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
K = np.array(synth.K_DEFAULT, np.float32)
T0 = (np.array(synth.INIT_POS_FACTOR, np.float32) * np.float32(synth.VOLUME_DIM)).astype(np.float32)
frames = 24
depth, poses = synth.make_sequence(frames + 1)
with kf.Kfusion((640, 480), vres, 4.8, T0, (10, 5, 4)) as g:
    g.enable_timing(True)
    for f in range(frames):
        pose = poses[f].astype(np.float32)
        g.preprocessing(depth[f])
        g.integrateKernel(g.inverse(pose), g.cameraMatrix(K), 0.1)
    g.synchroniseDevices()
    g.reset_stats()
    pose = poses[frames].astype(np.float32)
    g.preprocessing(depth[frames])
    for _ in range(reps):
        g.integrateKernel(g.inverse(pose), g.cameraMatrix(K), 0.1)
    s = g.stats()
    if not (int(os.environ.get("KFB_FLAGS", "0"), 0) & 4):
        cls = g.read(kf.BUF_BRICKCLASS)
        n = cls.size
        print(f"  brick classes of the last launch: skip {int((cls == 0).sum())} free {int((cls == 1).sum())} per-voxel {int((cls == 2).sum())} of {n}"
              f"  -> voxels free {int((cls == 1).sum()) * 512} per-voxel {int((cls == 2).sum()) * 512}")
    print(f"{os.environ.get('KFB_LIB', 'default')}: {vres}^3 integrate {s['ms_integrate'] / reps * 1e3:.1f} us/launch, N_upd {s['voxels_updated_last']}")
