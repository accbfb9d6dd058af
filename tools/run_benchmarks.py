"""Run kfusion-benchmark-{b200,cuda,openmp,cpp} on the same synthetic .raw with the same flags and print
frames/s from the `computation` column (benchmark.cpp:166) over frames >= 4, as SURVEY §8d asks.
    python tools/run_benchmarks.py [--volume 256] [--frames 100] [--cpp-frames 20]
"""
import argparse, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from slambench_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--volume", type=int, default=256)
ap.add_argument("--frames", type=int, default=100)
ap.add_argument("--cpp-frames", type=int, default=24, help="frames for the single-thread cpp binary (it is slow)")
a = ap.parse_args()
bins = {"b200": os.path.join(ROOT, "build", "kfusion-benchmark-b200"),
        "openmp": os.path.join(ROOT, "oracle", "_ref", "kfusion-benchmark-openmp"),
        "cuda": os.path.join(ROOT, "oracle", "_ref", "kfusion-benchmark-cuda"),   # the reference's own CUDA backend, sm_100a
        "cpp": os.path.join(ROOT, "oracle", "_ref", "kfusion-benchmark-cpp")}
depth, _ = synth.make_sequence(a.frames)
with tempfile.TemporaryDirectory() as tmp:
    for name, exe in bins.items():
        if not os.path.exists(exe):
            print(f"{name}: {exe} not built"); continue
        n = a.cpp_frames if name == "cpp" else a.frames
        raw = os.path.join(tmp, f"{name}.raw")
        synth.write_raw(raw, depth[:n])
        log = os.path.join(tmp, f"{name}.log")
        subprocess.run([exe, "-i", raw, "-s", "4.8", "-p", "0.5,0.5,0.25", "-z", "1000000", "-c", "1", "-r", "1", "-t", "1", "-m", "0.1",
                        "-y", "10,5,4", "-k", "481.2,480,320,240", "-v", str(a.volume), "-o", log], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        rows = np.array([[float(v) for v in l.split()] for l in open(log) if len(l.split()) == 14 and l.split()[0].isdigit()])
        r = rows[4:]
        print(f"{name:7s} {a.volume}^3 frames {len(rows)}: computation {1e3 * r[:, 7].mean():.3f} ms/frame = {1 / r[:, 7].mean():.1f} fps | "
              f"pre {1e3 * r[:, 2].mean():.3f} track {1e3 * r[:, 3].mean():.3f} integ {1e3 * r[:, 4].mean():.3f} ray {1e3 * r[:, 5].mean():.3f} "
              f"render {1e3 * r[:, 6].mean():.3f} ms | tracked {int(r[:, 12].sum())}/{len(r)} | final XYZ {rows[-1, 9:12]} | cores {os.cpu_count()}")
