#!/usr/bin/env python
"""bench.py — frames/s of the KinectFusion per-frame pipeline (preprocess -> track -> integrate ->
raycast) on B200, through the C ABI of libkfb200.so (include/kfb200.h).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--volume 512]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A STEP is one 640x480 depth frame through the whole pipeline: one `Kfusion::computeFrame` call
(kernels.h:158-166; `--staged`: the four stage calls exactly as kfusion/src/benchmark.cpp:125-150 issues them).  The workload is BASELINE.json configs[1]:
the synthetic analytic-room sequence, 512^3 volume, 4.8 m, mu 0.1, pyramid 10,5,4, -r 1 -t 1.
With W >= 4 the warm-up covers the reference's start-up frames 0-3 (untracked by construction,
SURVEY §8a a18), so every timed frame is tracked + integrated + raycast.

  value     whole-job frames/s with the uint16 sensor frames already resident in HBM
  e2e       the same frames from pinned HOST memory through kfb_preprocess (H2D inside the timed
            region) with the pose / ICP sums read back every frame (D2H inside)
  roofline  the integrate kernel: algorithmic bytes (8*N_upd + 4*P per launch; N_upd
            counted exactly by the kernel) / its CUDA-event time measured in the timed region
  cpu_baseline  the unmodified reference C++ backend (oracle/_ref, OpenMP) or, where that was not
            built, the plain-C restatement, timed on this box's host cores on a bounded sample

N > 1: one process per GPU, one independent sequence per GPU (BASELINE configs[2]), no data-path
collective; value = sum of frames / max-over-ranks time ("weak" scaling).  The same line then carries
a "sharded" key: ONE sequence on ONE volume cut into z-slabs over the ranks (configs[3] 1024^3, and
configs[4] 2048^3 from 8 GPUs on), moved over NVLink peer memory by the library itself, with the slab
cut picked at set-up by timing a few candidate partitions (`--no-slab-tuning`, `--no-sharded`,
`--sharded-steps`).  `--mode sharded --volume V` prints that run as a line of its own ("strong" scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "frames/sec end-to-end (640x480)"
UNIT = "frames/s"
W_IMG, H_IMG = 640, 480
P_PIX = W_IMG * H_IMG
MU = 0.1
ICP_THRESHOLD = 1e-5
PYRAMID = (10, 5, 4)
VOLUME_DIM = 4.8


def workload_name(vres: int) -> str:
    which = {256: "configs[0]", 512: "configs[1]"}.get(vres, "configs[1] family")
    return (f"{which}: synthetic analytic-room 640x480 .raw-format depth sequence, {vres}^3 short2 TSDF, 4.8 m, "
            f"mu 0.1, pyramid 10,5,4, integration/tracking rate 1")


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_bytes(args):
    """(DRAM bytes per integrate launch, note) from the committed ncu capture (profiles/integrate_traffic.json), or (None, why).
    It is a CAPTURE, not a live measurement: the note names its file and the N_upd it was taken at."""
    if args.traffic_bytes is not None:
        return args.traffic_bytes, "--traffic-bytes"
    try:
        e = json.load(open(os.path.join(ROOT, "profiles", "integrate_traffic.json"))).get(str(args.volume))
        if e is None:
            return None, "no ncu capture committed for this volume"
        return e["bytes"], f"ncu capture {e['source']} at N_upd = {e['n_upd']} (algorithmic bytes of that launch: {e['algorithmic_bytes']})"
    except Exception as ex:  # noqa: BLE001
        return None, f"profiles/integrate_traffic.json unreadable: {ex}"


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled from a
    background thread every ~2 ms (the timed region of the default run is only ~50 ms, too short for
    `nvidia-smi -lms`)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples, self.bits = [], 0
        self.max_mhz = None
        self._stop = False
        self._thread = None
        try:
            import pynvml  # noqa: PLC0415

            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self._stop:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        import threading  # noqa: PLC0415

        self._thread = threading.Thread(target=self._poll, daemon=True)
        self._thread.start()

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if self._thread is None:
            return out
        self._stop = True
        self._thread.join(timeout=2)
        if self.samples:
            out.update(sm_mhz=statistics.median(self.samples), samples=len(self.samples))
        out["reasons"] = sorted(n for b, n in self.REASONS.items() if self.bits & b)
        return out


# ----------------------------------------------------------------------------- CPU baseline
def set_omp_threads(n: int) -> int:
    """Make the OpenMP runtime of the CPU legs use `n` threads and return what it will really use.  torchrun
    pre-sets OMP_NUM_THREADS=1 in every rank, so the variable is overwritten (not defaulted) and, because a
    runtime that is already loaded no longer reads it, omp_set_num_threads is called on libgomp as well."""
    import ctypes  # noqa: PLC0415

    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(ctypes.c_int(n))
        return int(gomp.omp_get_max_threads())
    except OSError:
        return n


def run_cpu_frames(depth, vres: int, n_warm: int, n_timed: int, budget_s: float | None, single_thread: bool = False):
    """Drive the reference CPU backend over depth[0 : n_warm + n_timed]; returns
    (frames timed, seconds, kind, cores, impl name).  Test-infrastructure code path: this is the one
    place bench.py executes oracle/ as the thing measured (cpu_baseline / --impl reference).
    `single_thread`: the reference's `-cpp` build (kfusion-benchmark-cpp's backend) instead of `-openmp`."""
    from oracle import cpu_backend as cb
    from slambench_b200 import synth

    if single_thread and os.path.exists(cb.REF_LIB):
        lib, kind, cores = cb.REF_LIB, "reference", 1
    elif os.path.exists(cb.REF_OMP_LIB):
        lib, kind, cores = cb.REF_OMP_LIB, "reference", 1 if single_thread else host_cores()
    elif os.path.exists(cb.REF_LIB):
        lib, kind, cores = cb.REF_LIB, "reference", 1
    else:
        cb.build_port()
        lib, kind, cores = cb.PORT_LIB, "port", 1 if single_thread else host_cores()   # the C restatement is built with -fopenmp too
    be = cb.CpuKfusion(lib)
    used = set_omp_threads(cores)
    if lib != cb.REF_LIB:
        cores = used                                   # the thread count the OpenMP runtime reports, not the wish
    K = np.array(synth.K_DEFAULT, np.float32)
    T0 = (np.array(synth.INIT_POS_FACTOR, np.float32) * np.float32(VOLUME_DIM)).astype(np.float32)
    be.create((W_IMG, H_IMG), vres, VOLUME_DIM, T0, PYRAMID)
    done, t_sum = 0, 0.0
    try:
        for f in range(min(len(depth), n_warm + n_timed)):
            t0 = time.perf_counter()
            be.preprocessing(depth[f])
            be.tracking(K, ICP_THRESHOLD, 1, f)
            be.integration(K, 1, MU, f)
            be.raycasting(K, MU, f)
            dt = time.perf_counter() - t0
            if f >= n_warm:
                done += 1
                t_sum += dt
                if budget_s is not None and t_sum > budget_s and done >= 4:
                    break
    finally:
        be.destroy()
    return done, t_sum, kind, cores, be.name


def run_reference_cuda(depth, vres: int, n_frames: int):
    """GPU comparator: the reference's OWN CUDA backend (kfusion/src/cuda/kernels.cu, unmodified, compiled for sm_100a into
    oracle/_ref/kfusion-benchmark-cuda) on the same frames, on this GPU.  fps from the `computation` column over frames >= 4
    (benchmark.cpp:166), i.e. with its own per-stage synchronisation.  None when the binary was not built."""
    import tempfile

    from oracle import cpu_backend as cb
    from slambench_b200 import synth

    if not os.path.exists(cb.REF_CUDA_BIN):
        return None
    with tempfile.TemporaryDirectory() as tmp:
        raw, log = os.path.join(tmp, "seq.raw"), os.path.join(tmp, "cuda.log")
        synth.write_raw(raw, depth[:n_frames])
        r = subprocess.run([cb.REF_CUDA_BIN, "-i", raw, "-s", "4.8", "-p", "0.5,0.5,0.25", "-z", "1000000", "-c", "1", "-r", "1", "-t", "1",
                            "-m", "0.1", "-y", "10,5,4", "-k", "481.2,480,320,240", "-v", str(vres), "-o", log],
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=300)
        if r.returncode != 0 or not os.path.exists(log):
            return {"unavailable": f"kfusion-benchmark-cuda exited {r.returncode}: {r.stderr.decode(errors='replace')[-200:]}"}
        rows = np.array([[float(v) for v in l.split()] for l in open(log) if len(l.split()) == 14 and l.split()[0].isdigit()])
    t = rows[4:]
    return {"value": float(1.0 / t[:, 7].mean()), "unit": UNIT, "kind": "reference-cuda (kfusion/src/cuda/kernels.cu, unmodified, sm_100a)",
            "stage_ms_per_frame": {"preprocess": 1e3 * float(t[:, 2].mean()), "track": 1e3 * float(t[:, 3].mean()),
                                   "integrate": 1e3 * float(t[:, 4].mean()), "raycast": 1e3 * float(t[:, 5].mean())},
            "tracked_frames": int(t[:, 12].sum()), "sample": f"{len(t)} frames (from frame 4) of the same sequence and volume"}


def reference_arm(args, rank: int):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's cores."""
    if rank != 0:
        return
    from slambench_b200 import synth

    n = args.warmup + args.steps
    depth, _ = synth.make_sequence(n, long_run=False if n <= 400 else None)
    done, secs, kind, cores, name = run_cpu_frames(depth, args.volume, args.warmup, args.steps, None)
    fps = done / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": done, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / done, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(args.volume), "volume": args.volume, "backend": name},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"frames {args.warmup}..{args.warmup + done - 1} of the same sequence, whole pipeline per frame"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------- B200 arm
def b200_arm(args, rank: int, world: int, local_rank: int):
    import torch

    from slambench_b200 import kfusion as kf
    from slambench_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libkfb200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    K = np.array(synth.K_DEFAULT, np.float32)
    T0 = (np.array(synth.INIT_POS_FACTOR, np.float32) * np.float32(VOLUME_DIM)).astype(np.float32)
    n = args.warmup + args.steps
    # one independent sequence per GPU (seed = rank); deterministic, no RNG
    depth_np, gt = synth.make_sequence(n, seed=rank % 8, long_run=False if n <= 400 else None)
    host = torch.from_numpy(depth_np).pin_memory()                   # pinned host frames (e2e arm)
    dev = host.to(f"cuda:{local_rank}", non_blocking=False)          # HBM-resident frames (value arm)
    frame_bytes = W_IMG * H_IMG * 2

    def run(resident: bool, time_mask: int, collective: bool = True):
        """W warm-up frames, then exactly K timed frames.  Returns dict of measurements.  `collective=False`: a pass
        that only some ranks execute (the per-stage breakdown on rank 0) must not enter the cross-rank barrier."""
        sync = barrier if collective else torch.cuda.synchronize
        with kf.Kfusion((W_IMG, H_IMG), args.volume, VOLUME_DIM, T0, PYRAMID, device=local_rank) as g:
            stream = torch.cuda.ExternalStream(g.stream(), device=local_rank)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

            def frame(f):
                if not args.staged:
                    # the reference's one-call entry point (Kfusion::computeFrame, kernels.h:158-166): the library enqueues
                    # all four stages before it waits for the pose; tracked / integrated / pose are read back every frame
                    if resident:
                        g.computeFrame_device(dev[f].data_ptr(), (W_IMG, H_IMG), K, 1, 1, ICP_THRESHOLD, MU, f)
                    else:
                        g.computeFrame(depth_np[f], None, K, 1, 1, ICP_THRESHOLD, MU, f)   # numpy view of the pinned tensor's memory
                    return g.getTracked(), g.getIntegrated(), g.getPose()
                # the four stage calls, exactly as benchmark.cpp:125-150 issues them
                if resident:
                    g.preprocessing_device(dev[f].data_ptr(), (W_IMG, H_IMG))
                else:
                    g.preprocessing(depth_np[f])
                tr = g.tracking(K, ICP_THRESHOLD, 1, f)
                it = g.integration(K, 1, MU, f)
                g.raycasting(K, MU, f)
                return tr, it, g.getPose()                # pose is on the host when tracking returns

            tracked = integrated = 0
            for f in range(args.warmup):
                frame(f)
            g.synchroniseDevices()
            g.enable_timing(time_mask)
            g.reset_stats()
            sync()
            clocks = ClockSampler(local_rank)
            if time_mask and rank == 0:
                clocks.start()
            t0 = time.perf_counter()
            ev0.record(stream)
            for f in range(args.warmup, n):
                tr, it, pose = frame(f)
                tracked += tr
                integrated += it
            ev1.record(stream)
            g.synchroniseDevices()
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            ck = clocks.stop() if (time_mask and rank == 0) else None
            sync()
            ms = ev0.elapsed_time(ev1)
            st = g.stats()
            err = float(np.abs(pose[:3, 3] - synth.expected_pose(gt, n - 1)[:3, 3]).max())
            return dict(ms=ms, wall_ms=wall * 1e3, st=st, tracked=tracked, integrated=integrated, err=err, clocks=ck)

    # depth_np must alias the pinned buffer for the e2e arm
    depth_np = host.numpy()
    res = run(resident=True, time_mask=4)        # value: HBM-resident input; integrate timed with CUDA events
    e2e = run(resident=False, time_mask=0)       # e2e: pinned host frames, H2D + D2H inside the timed region
    # rank 0 only, and therefore without any collective inside (round 1 deadlocked here at N > 1)
    diag = run(resident=True, time_mask=15, collective=False) if rank == 0 and not args.no_breakdown else None

    def reduce_max(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ms = reduce_max(res["ms"])
    ms_e2e = reduce_max(max(e2e["ms"], e2e["wall_ms"]))   # host-visible completion: the slower of device and wall time
    launches = reduce_sum(float(res["st"]["kernel_launches"]))
    all_tracked = reduce_sum(float(res["tracked"]))
    K_steps = args.steps
    value = world * K_steps / (ms * 1e-3)
    e2e_value = world * K_steps / (ms_e2e * 1e-3)

    if rank == 0:
        st = res["st"]
        peak, peak_src = peaks()
        n_int = max(1, int(st["frames_integrated"]))
        alg_bytes = (8.0 * st["voxels_updated_total"] + 4.0 * P_PIX * n_int) / n_int
        t_int_s = (st["ms_integrate"] / n_int) * 1e-3
        achieved = alg_bytes / t_int_s / 1e9 if t_int_s > 0 else 0.0
        full_sweep_bytes = 8.0 * args.volume ** 3 + 4.0 * P_PIX
        roof = {
            "bound": "hbm", "kernel": "k_integrate_plan2 + k_integrate_free_runs + k_integrate_run2", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic_bytes(args)[0], "traffic_source": traffic_bytes(args)[1], "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": t_int_s * 1e6,
            "n_upd_per_launch": st["voxels_updated_total"] / n_int,
            "full_sweep_gbs": full_sweep_bytes / t_int_s / 1e9 if t_int_s > 0 else 0.0,
            "launches_timed": n_int,
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K_steps, "warmup": args.warmup,
            "ms_per_step": ms / K_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args.volume), "volume": args.volume, "image": [W_IMG, H_IMG],
                       "frames": n, "parallelism": "1 sequence per GPU, no collective" if world > 1 else "single GPU",
                       "l2": f"no flush: the {4 * args.volume ** 3 / 1e6:.0f} MB volume streamed every frame exceeds the 126 MB L2"
                             if args.volume >= 512 else "no flush: volume fits L2 (the reference's own case); see --volume 512",
                       "tracked_frames": int(all_tracked), "final_pose_err_m": res["err"]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": frame_bytes,
                    "d2h_bytes_per_step": e2e["st"]["d2h_bytes"] / K_steps, "ms_per_step": ms_e2e / K_steps},
            "gpu_launches": int(launches),
            "clocks": res["clocks"],
            "roofline": roof,
        }
        if diag is not None:
            d = diag["st"]
            line["stage_ms_per_frame"] = {
                "preprocess": d["ms_preprocess"] / K_steps, "track": d["ms_track"] / K_steps,
                "integrate": d["ms_integrate"] / max(1, d["frames_integrated"]), "raycast": d["ms_raycast"] / K_steps,
                "icp_iterations": d["icp_iterations_total"] / K_steps, "launches": d["kernel_launches"] / K_steps,
                "note": "separate pass with per-stage CUDA events, which keeps the stages serial; in the timed region frame k+1's "
                        "copy + preprocess + pyramid overlap frame k's raycast, so the stages sum to more than ms_per_step",
            }
        if world == 1 and not args.no_cpu_baseline:
            done, secs, kind, cores, name = run_cpu_frames(depth_np, args.volume, min(4, args.warmup), args.steps, args.cpu_budget)
            line["cpu_baseline"] = {"value": done / secs, "unit": UNIT, "cores": cores, "kind": kind, "backend": name,
                                    "sample": f"{done} frames (from frame {min(4, args.warmup)}) of the same sequence and volume, "
                                              f"whole pipeline per frame, bounded to ~{args.cpu_budget:.0f} s"}
            # north_star: next to kfusion-benchmark-openmp AND -cpp: the reference's single-thread backend, same frames
            d1, s1, kind1, _, name1 = run_cpu_frames(depth_np, args.volume, min(4, args.warmup), min(args.steps, 8),
                                                     args.cpu_budget / 2, single_thread=True)
            line["cpu_baseline"]["cpp_1thread"] = {"value": d1 / s1, "unit": UNIT, "cores": 1, "kind": kind1, "backend": name1,
                                                   "sample": f"{d1} frames (from frame {min(4, args.warmup)}), bounded to ~{args.cpu_budget / 2:.0f} s"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gb = run_reference_cuda(depth_np, args.volume, min(n, 40))
        if gb is not None:
            line["gpu_baseline"] = gb
    sh = None
    if dist is not None and not args.no_sharded:
        # N > 1: the driver only ever runs `bench.py --gpus N`, so the z-slab mode (BASELINE configs[3]: ONE sequence,
        # 1024^3 cut into N slabs; configs[4]: 2048^3, at 8 GPUs) is timed in the same run and reported next to `value`
        sh = {}
        vols = [1024] + ([2048] if world >= 8 else [])
        for vol in vols:
            r = sharded_run(args, rank, world, local_rank, torch, dist, vol, args.sharded_steps, max(4, min(args.warmup, 8)))
            if rank == 0:
                sh[str(vol)] = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "scaling": "strong",
                                "steps": r["steps"], "workload": r["config"]["workload"], "parallelism": r["config"]["parallelism"],
                                "slabs": r["config"]["slabs"], "slab_calibration": r["config"]["slab_calibration"], "tracked_frames": r["config"]["tracked_frames"],
                                "final_pose_err_m": r["config"]["final_pose_err_m"], "roofline": r["roofline"],
                                "stage_ms_per_frame": r["stage_ms_per_frame"]}
    if rank == 0:
        if sh:
            line["sharded"] = sh
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def render_sequence_distributed(total: int, rank: int, world: int, torch, dist, device):
    """Frames f == rank (mod world) rendered here, all frames gathered from the ranks (bytes: NCCL has no int16).
    Returns (uint16[total, H, W], gt[total, 4, 4]) identical on every rank."""
    from slambench_b200 import synth

    long_run = total > 400
    per = (total + world - 1) // world
    loc = torch.zeros((per, H_IMG, W_IMG * 2), dtype=torch.uint8, device=device)
    for i, f in enumerate(range(rank, total, world)):
        img = np.ascontiguousarray(synth.render_depth_mm(synth.trajectory_pose(f, 0, long_run)))
        loc[i] = torch.from_numpy(img.view(np.uint8).reshape(H_IMG, W_IMG * 2)).to(device)
    allf = torch.empty((world * per, H_IMG, W_IMG * 2), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(allf, loc)
    allf = allf.view(world, per, H_IMG, W_IMG * 2).permute(1, 0, 2, 3).reshape(world * per, H_IMG, W_IMG * 2)[:total]
    depth = np.ascontiguousarray(allf.contiguous().cpu().numpy()).view(np.uint16).reshape(total, H_IMG, W_IMG)
    gt = np.stack([synth.trajectory_pose(f, 0, long_run) for f in range(total)])
    return depth, gt


def sharded_arm(args, rank: int, world: int, local_rank: int):
    """`--mode sharded`: the z-slab run alone, printed as the bench line (strong scaling: the work is fixed)."""
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = sharded_run(args, rank, world, local_rank, torch, dist, args.volume, args.steps, args.warmup)
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def sharded_run(args, rank: int, world: int, local_rank: int, torch, dist, volume: int, steps: int, warmup: int):
    """ONE sequence, the volume cut into z-slabs over the ranks (BASELINE configs[3], [4]): integrate local, raycast
    over NVLink peer slabs, ICP replicated or all-reduced.  Returns the bench line (rank 0) or None."""
    from slambench_b200 import sharded, synth

    K = np.array(synth.K_DEFAULT, np.float32)
    T0 = (np.array(synth.INIT_POS_FACTOR, np.float32) * np.float32(VOLUME_DIM)).astype(np.float32)
    n = warmup + steps
    n_diag = 6                                    # extra frames for the per-stage breakdown (after the timed region)
    if n + n_diag <= 200:
        depth_np, gt = synth.make_sequence(n + n_diag, long_run=False)
    else:
        # long runs (configs[4]: 1000 frames): every rank renders the frames f == rank (mod world) and the ranks
        # all-gather them over NCCL — the sequence is identical on every rank and to synth.make_sequence()
        depth_np, gt = render_sequence_distributed(n + n_diag, rank, world, torch, dist, f"cuda:{local_rank}")
    host = torch.from_numpy(depth_np).pin_memory()
    depth_np = host.numpy()
    slabs, calib = None, []
    if not args.even_slabs and not args.no_slab_tuning:
        # Set-up (untimed): a few a-priori partitions (sharded.candidate_slabs) are each run on the first frames and the one
        # with the shortest frame time is kept.  A volume does not depend on how it is cut; what a cut costs does.
        from slambench_b200 import kfusion as kf_

        n_cal = min(n, 12)
        cands = sharded.candidate_slabs(volume, world, VOLUME_DIM, kf_.identity_pose(T0), K, (W_IMG, H_IMG), far=float(depth_np[0].max()) / 1000.0)
        best = None
        for name, cs in cands.items():
            if any(cs == c["slabs_t"] for c in calib):
                continue
            with sharded.ShardedKfusion((W_IMG, H_IMG), volume, VOLUME_DIM, T0, PYRAMID, rank=rank, world=world, device=local_rank,
                                        icp_mode=args.icp_mode, slabs=cs) as s:
                t0c = 0.0
                for f in range(n_cal):
                    if f == 4:
                        s.synchroniseDevices()
                        dist.barrier()
                        torch.cuda.synchronize()
                        t0c = time.perf_counter()
                    s.preprocessing(depth_np[f]); s.tracking(K, ICP_THRESHOLD, 1, f); s.integration(K, 1, MU, f); s.raycasting(K, MU, f)
                s.synchroniseDevices()
                torch.cuda.synchronize()
                tt = torch.tensor([(time.perf_counter() - t0c) * 1e3 / (n_cal - 4)], dtype=torch.float64, device=f"cuda:{local_rank}")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            calib.append({"candidate": name, "slabs_t": cs, "slabs": [list(z) for z in cs], "ms_per_frame": round(float(tt[0]), 4)})
            if best is None or float(tt[0]) < best[0]:
                best = (float(tt[0]), cs)
        slabs = best[1]
        for c in calib:
            del c["slabs_t"]
    with sharded.ShardedKfusion((W_IMG, H_IMG), volume, VOLUME_DIM, T0, PYRAMID, rank=rank, world=world, device=local_rank,
                                icp_mode=args.icp_mode, balance_k=None if args.even_slabs else K,
                                balance_far=float(depth_np[0].max()) / 1000.0, slabs=slabs) as s:
        g = s.local
        stream = g.torch_stream(torch)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def frame(f):
            s.preprocessing(depth_np[f])
            tr = s.tracking(K, ICP_THRESHOLD, 1, f)
            s.integration(K, 1, MU, f)
            s.raycasting(K, MU, f)
            return tr

        for f in range(warmup):
            frame(f)
        s.synchroniseDevices()
        g.enable_timing(4)
        g.reset_stats()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev0.record(stream)
        tracked = 0
        for f in range(warmup, n):
            tracked += frame(f)
        ev1.record(stream)
        s.synchroniseDevices()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        dist.barrier()
        ms = max(ev0.elapsed_time(ev1), wall)
        st = g.stats()
        t = torch.tensor([ms, float(st["voxels_updated_total"]), st["ms_integrate"]], dtype=torch.float64, device=f"cuda:{local_rank}")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        per_rank = torch.zeros(world, dtype=torch.float64, device=f"cuda:{local_rank}")
        per_rank[rank] = st["ms_integrate"] / max(1, int(st["frames_integrated"])) * 1e3
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
        err = float(np.abs(s.getPose()[:3, 3] - synth.expected_pose(gt, n - 1)[:3, 3]).max())
        # per-stage breakdown: the same calls with a full sync after each stage (host-visible time, max over ranks)
        stage = np.zeros(4)
        for f in range(n, n + n_diag):
            marks = [time.perf_counter()]
            for call in (lambda: s.preprocessing(depth_np[f]), lambda: s.tracking(K, ICP_THRESHOLD, 1, f),
                         lambda: s.integration(K, 1, MU, f), lambda: s.raycasting(K, MU, f)):
                call()
                s.synchroniseDevices()
                torch.cuda.synchronize()
                marks.append(time.perf_counter())
            stage += np.diff(marks) * 1e3 / n_diag
        tstage = torch.tensor(stage, dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(tstage, op=dist.ReduceOp.MAX)
        if rank == 0:
            peak, peak_src = peaks()
            ms_all = float(tmax[0])
            n_int = max(1, int(st["frames_integrated"]))
            alg = (8.0 * float(tsum[1]) + 4.0 * P_PIX * n_int * world) / n_int          # all slabs together
            t_int = float(tmax[2]) / n_int * 1e-3                                          # slowest slab
            line = {
                "metric": METRIC, "value": steps / (ms_all * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": ms_all / steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(volume), "volume": volume, "frames": n,
                           "parallelism": f"z-slab x{world}: integrate local, raycast via NVLink peer slabs, map bands + brick flags stored into the "
                                          f"peers, peer-memory barriers ({s.transport} transport), ICP {args.icp_mode}",
                           "slabs": [list(z) for z in s.slabs], "slab_calibration": calib,
                           "tracked_frames": int(tracked), "final_pose_err_m": err},
                "e2e": {"value": steps / (ms_all * 1e-3), "unit": UNIT, "h2d_bytes_per_step": W_IMG * H_IMG * 2,
                        "d2h_bytes_per_step": st["d2h_bytes"] / steps},
                "gpu_launches": int(st["kernel_launches"]) * world,
                "roofline": {"bound": "hbm", "kernel": "k_integrate_run", "achieved": alg / t_int / 1e9, "peak": peak * world, "unit": "GB/s",
                             "frac": alg / t_int / 1e9 / (peak * world), "traffic": None, "peak_source": peak_src + f" x{world}",
                             "us_per_launch": t_int * 1e6, "us_per_launch_by_rank": [round(float(v), 1) for v in per_rank]},
                "stage_ms_per_frame": {"preprocess": float(tstage[0]), "track": float(tstage[1]),
                                       "integrate+barrier": float(tstage[2]), "raycast+band exchange": float(tstage[3]),
                                       "note": "synchronised after every stage, max over ranks"},
            }
            return line
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=96)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--volume", type=int, default=512, help="volume resolution N (N^3 voxels); default = BASELINE configs[1]")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--staged", action="store_true",
                    help="drive the four stage calls of benchmark.cpp (preprocessing / tracking / integration / raycasting) instead of "
                         "one Kfusion::computeFrame call per frame")
    ap.add_argument("--no-breakdown", action="store_true", help="skip the extra per-stage timing pass")
    ap.add_argument("--mode", default="sequences", choices=["sequences", "sharded"],
                    help="N > 1: one independent sequence per GPU (weak scaling, default) or ONE sequence on a z-slab sharded volume (strong)")
    ap.add_argument("--icp-mode", default="replicated", choices=["replicated", "allreduce"])
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the extra z-slab (configs[3]/[4]) runs")
    ap.add_argument("--sharded-steps", type=int, default=20, help="timed frames of the z-slab runs added to the N > 1 line")
    ap.add_argument("--even-slabs", action="store_true", help="sharded mode: equal z-slabs instead of the load-aware boundaries")
    ap.add_argument("--no-slab-tuning", action="store_true",
                    help="sharded mode: take the default a-priori slab boundaries instead of timing a few candidate partitions at set-up")
    ap.add_argument("--traffic-bytes", type=float, default=None, help="dram bytes/launch of k_integrate from the ncu capture in profiles/")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one process per GPU)")
    if args.mode == "sharded" and world > 1:
        sharded_arm(args, rank, world, local_rank)
        return
    b200_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
