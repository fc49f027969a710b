#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native hot paths (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): depth frames/s integrated into a 512^3 TSDF (configs[1]: 1000 synthetic 640x480
frames, rigid trajectory, 512^3 = 8^3 blocks of 64^3 voxels @ 3/512 m, one B200).  A "step" is one pass over
that sequence: --frames-per-step (default: all 1000) consecutive frames handed to the library in one call.

  value       frames/s with the frames already resident in HBM (er_tsdf_integrate_device)
  e2e         frames/s through the C-ABI with HOST buffers: pinned host frames -> er_tsdf_integrate_batch
              (host->device copies inside the timed region) + a device->host read of the step's result
  roofline    integrate kernel: algorithmic bytes (16 B x updated voxels + 2 B x W x H per frame)
              / its event-bracketed device time, against the measured HBM peak
  cpu_baseline  the verbatim reference binary (oracle/_ref/Integrate_ref, its 8 OpenMP threads) on a bounded
              prefix of the same frames, timed per frame on this box's host cores
  icp         secondary metric of BASELINE.json: fragment-pair ICP iterations/s (same JSON line)

N > 1 (torchrun, one rank per GPU): every rank integrates its own frame stream into its own volume -- the
path shards by independent volumes/fragments with no data-path collective -- so scaling is "weak" and
value = total frames of all ranks / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from elasticreconstruction_b200 import synth  # noqa: E402

SEQ_FRAMES = 1000            # configs[1]
W, H = 640, 480
BYTES_PER_FRAME = W * H * 2


# --------------------------------------------------------------------------------------------- helpers
def load_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))[kernel]
        return t["dram_bytes_per_launch"] / t.get("units_per_launch", 1.0)
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled every 10 ms DURING the timed region through NVML (the same
    counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints; polling nvidia-smi itself is too
    slow for a timed region of a few hundred milliseconds)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return

        def loop():
            while not self._stop.is_set():
                try:
                    self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                self._stop.wait(0.01)
        self._thr = threading.Thread(target=loop, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def make_sequence(rank: int, world: int):
    """The configs[1] trajectory (camera->world) and a generator for its depth frames.  Rank r of a
    multi-GPU run starts r/world of the way round the loop (its own independent stream)."""
    poses = synth.quantize_log(synth.loop_trajectory(SEQ_FRAMES))
    off = (rank * SEQ_FRAMES) // max(world, 1)
    order = (np.arange(SEQ_FRAMES) + off) % SEQ_FRAMES
    return poses[order]


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


# --------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """--impl reference: the reference's own CPU Integrate (oracle/_ref/Integrate_ref, verbatim sources,
    8 OpenMP threads) on bounded samples of the same workload; rank 0 only."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    fps_step = args.ref_frames_per_step
    nsteps = args.warmup + args.steps
    n = fps_step * nsteps
    poses = make_sequence(0, 1)[: n + 1]
    kind = "reference"
    with tempfile.TemporaryDirectory() as d:
        frames = synth.room_depth(poses[:n])
        synth.write_depth_container(os.path.join(d, "depth.raw"), frames)
        synth.write_log(os.path.join(d, "traj.log"), poses)
        tfile = os.path.join(d, "stamps.txt")
        if os.path.exists(pyoracle.REF_INTEGRATE):
            env = dict(os.environ, ER_SHIM_TIMING_FILE=tfile)
            subprocess.run([pyoracle.REF_INTEGRATE, "-oni", os.path.join(d, "depth.raw"), "--ref_traj",
                            os.path.join(d, "traj.log"), "--save_to", os.path.join(d, "w.pcd")], cwd=d, env=env,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
            stamps = np.loadtxt(tfile)
            per_frame = np.diff(stamps)[:n]
        else:
            kind = "port"
            pyoracle.build()
            vol = pyoracle.OracleTSDF(threads=8)
            per_frame = np.empty(n)
            for f in range(n):
                t0 = time.perf_counter()
                vol.integrate(frames[f], poses[f])
                per_frame[f] = time.perf_counter() - t0
    step_s = per_frame.reshape(nsteps, fps_step).sum(axis=1)[args.warmup:]
    total = float(step_s.sum())
    value = fps_step * args.steps / total
    cores = min(8, os.cpu_count() or 1)
    sample = "%d steps x %d frames (prefix of the 1000-frame sequence), per-frame wall time of Execute()" % (
        args.steps, fps_step)
    line = {
        "impl": "reference", "metric": "depth frames/sec into 512^3 TSDF", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "TSDF integrate: 1000 synthetic 640x480 frames, rigid trajectory, 512^3 @ 3/512 m (BASELINE configs[1])",
                   "frames_per_step": fps_step, "threads": 8, "host_cores": os.cpu_count()},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_leg(poses, n_frames: int):
    """cpu_baseline of the main line: the verbatim reference binary on the first n_frames frames."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    with tempfile.TemporaryDirectory() as d:
        frames = synth.room_depth(poses[:n_frames])
        if os.path.exists(pyoracle.REF_INTEGRATE):
            synth.write_depth_container(os.path.join(d, "depth.raw"), frames)
            synth.write_log(os.path.join(d, "traj.log"), poses[: n_frames + 1])
            tfile = os.path.join(d, "stamps.txt")
            env = dict(os.environ, ER_SHIM_TIMING_FILE=tfile)
            subprocess.run([pyoracle.REF_INTEGRATE, "-oni", os.path.join(d, "depth.raw"), "--ref_traj",
                            os.path.join(d, "traj.log"), "--save_to", os.path.join(d, "w.pcd")], cwd=d, env=env,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
            per = np.diff(np.loadtxt(tfile))[:n_frames]
            kind = "reference"
        else:
            pyoracle.build()
            vol = pyoracle.OracleTSDF(threads=8)
            per = np.empty(n_frames)
            for f in range(n_frames):
                t0 = time.perf_counter()
                vol.integrate(frames[f], poses[f])
                per[f] = time.perf_counter() - t0
            kind = "port"
    skip = min(5, n_frames // 4)       # first frames pay the block allocations
    fps = (n_frames - skip) / float(per[skip:].sum())
    return {"value": fps, "unit": "frames/s", "cores": min(8, os.cpu_count() or 1), "kind": kind,
            "sample": "first %d frames of the sequence (%d warm-up frames dropped), per-frame wall time of "
                      "Execute(); host has %d cores, reference hard-codes 8 threads" % (n_frames, skip, os.cpu_count())}


# --------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=SEQ_FRAMES,
                    help="frames handed to the library per call; must divide the 1000-frame sequence")
    ap.add_argument("--ref-frames-per-step", type=int, default=8)
    ap.add_argument("--cpu-frames", type=int, default=120, help="frames timed for cpu_baseline (0 = skip)")
    ap.add_argument("--fusion", type=int, default=32, help="frames per fused launch triple (1..32)")
    ap.add_argument("--no-icp", action="store_true")
    ap.add_argument("--icp-fragments", type=int, default=50)
    ap.add_argument("--icp-points", type=int, default=200000)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout must carry exactly one JSON line, but NCCL prints its version banner to stdout when NCCL_DEBUG is set
    # (and libraries may print there too): park the real stdout and point fd 1 at stderr until the line is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from elasticreconstruction_b200 import capi

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    F = args.frames_per_step
    nsteps = args.warmup + args.steps
    poses = make_sequence(rank, world)
    # every step's frames, generated once on the host (synthetic, analytic room)
    need = min(SEQ_FRAMES, F * nsteps)
    frames_np = synth.room_depth(poses[:need])
    step_idx = [(np.arange(F) + s * F) % need for s in range(nsteps)]

    # ---------------- value: frames resident in HBM
    d_frames = torch.from_numpy(frames_np.view(np.int16)).cuda(local)
    vol = capi.TSDFVolume(device=local, max_blocks=1024)
    vol.set_fusion(args.fusion)
    torch.cuda.synchronize()

    assert need % F == 0, "--frames-per-step must divide the resident sequence so every step is one contiguous run"

    def run_step_device(s):
        idx = step_idx[s]
        vol.integrate_device(d_frames.data_ptr() + int(idx[0]) * BYTES_PER_FRAME, poses[idx], F)

    for s in range(args.warmup):
        run_step_device(s)
    vol.sync()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = vol.launch_count()
    vol.event_record(0)
    t0 = time.perf_counter()
    for s in range(args.warmup, nsteps):
        run_step_device(s)
    vol.event_record(1)
    vol.sync()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_ms = vol.event_elapsed_ms(0, 1)
    launches = vol.launch_count() - l0
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    total_frames = F * args.steps * world
    value = total_frames / (dev_ms_max * 1e-3)

    # ---------------- roofline: per-kernel time + exact updated-voxel count on the timed frames
    vol2 = capi.TSDFVolume(device=local, max_blocks=1024)
    vol2.set_fusion(args.fusion)
    for s in range(args.warmup):
        idx = step_idx[s]
        vol2.integrate_device(d_frames.data_ptr() + int(idx[0]) * BYTES_PER_FRAME, poses[idx], F)
    vol2.set_stats(True)
    vol2.sync()
    st0 = vol2.stats()
    t_touch = t_cls = t_int = 0.0
    for s in range(args.warmup, nsteps):
        idx = step_idx[s]
        a, b, c = vol2.profile_device(d_frames.data_ptr() + int(idx[0]) * BYTES_PER_FRAME, poses[idx], F)
        t_touch += a
        t_cls += b
        t_int += c
    st1 = vol2.stats()
    n_prof = F * args.steps
    fusion = args.fusion
    n_launch = args.steps * ((F + fusion - 1) // fusion)          # integrate_kernel launches in the profiled pass
    upd = st1["updated_voxels"] - st0["updated_voxels"]
    blocks = st1["touched_blocks"] - st0["touched_blocks"]
    alg_bytes_per_launch = (16.0 * upd + 2.0 * W * H * n_prof) / n_launch
    int_s_per_launch = t_int * 1e-3 / n_launch
    peak, peak_src = load_peaks()
    achieved = alg_bytes_per_launch / int_s_per_launch / 1e9
    roofline = {"bound": "hbm", "kernel": "integrate_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "traffic": load_traffic("integrate_kernel"),
                "traffic_source": "ncu --set full capture of one 32-frame launch (profiles/traffic.json)",
                "algorithmic_bytes_per_launch": alg_bytes_per_launch, "us_per_launch": int_s_per_launch * 1e6,
                "frames_per_launch": n_prof / n_launch,
                "note": "algorithmic bytes = 16 B x updated voxels + 2 B x W x H per frame (SURVEY 8d); one launch "
                        "applies a whole batch of frames to register-resident voxels, so achieved may exceed DRAM traffic",
                "touch_us_per_frame": t_touch * 1e3 / n_prof, "classify_us_per_frame": t_cls * 1e3 / n_prof,
                "integrate_us_per_frame": t_int * 1e3 / n_prof,
                "updated_voxels_per_frame": upd / n_prof, "active_blocks_per_frame": blocks / n_prof,
                "one_voxels_per_frame": 64.0 * (st1["one_pairs"] - st0["one_pairs"]) / n_prof,
                "full_voxels_per_frame": 64.0 * (st1["full_pairs"] - st0["full_pairs"]) / n_prof,
                "dense_bytes_per_frame": 16.0 * 262144 * blocks / n_prof + 2.0 * W * H}
    del vol2

    # ---------------- e2e: host buffers through the C ABI, H2D inside the timed region
    pinned = torch.from_numpy(frames_np.view(np.int16)).pin_memory()
    vol3 = capi.TSDFVolume(device=local, max_blocks=1024)
    vol3.set_fusion(args.fusion)
    for s in range(args.warmup):
        idx = step_idx[s]
        vol3.integrate_ptr(pinned.data_ptr() + int(idx[0]) * BYTES_PER_FRAME, poses[idx], F)
        vol3.num_blocks()
    barrier()
    t0 = time.perf_counter()
    for s in range(args.warmup, nsteps):
        idx = step_idx[s]
        vol3.integrate_ptr(pinned.data_ptr() + int(idx[0]) * BYTES_PER_FRAME, poses[idx], F)
        vol3.num_blocks()            # device->host read of the step's result (forces completion)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = {"value": total_frames / float(t.item()), "unit": "frames/s", "h2d_bytes_per_step": F * BYTES_PER_FRAME,
           "d2h_bytes_per_step": 4 + 32, "timing": "wall clock around the C-ABI calls, barrier+sync both sides"}
    del vol3

    # ---------------- ICP secondary metric, then the cpu baseline (rank 0, N = 1 only; last, so that no GPU leg starts
    # from the idle clocks the GPU drops to while the host is busy)
    icp = None
    if not args.no_icp:
        del vol, d_frames, pinned
        torch.cuda.empty_cache()
        from elasticreconstruction_b200 import icp_bench
        icp = icp_bench.run(local, rank, world, barrier, peak_gbs=peak, num_fragments=args.icp_fragments,
                            points=args.icp_points, clock_sampler=ClockSampler)
    cpu = None
    if rank == 0 and world == 1 and args.cpu_frames > 0:
        cpu = cpu_baseline_leg(poses, args.cpu_frames)

    if rank == 0:
        line = {
            "metric": "depth frames/sec into 512^3 TSDF", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "TSDF integrate: 1000 synthetic 640x480 frames, rigid trajectory, 512^3 @ 3/512 m "
                                   "(BASELINE configs[1])",
                       "frames_per_step": F, "frames_per_launch": args.fusion, "block_pool": 1024, "sharding": "independent volume per rank" if world > 1 else "single volume",
                       "l2": "inputs larger than L2 (resident depth sequence %.0f MB + 2 GiB block pool); consecutive "
                             "frames legitimately reuse their active blocks through L2" % (need * BYTES_PER_FRAME / 1e6)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "wall_s_timed_region": wall, "icp": icp,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
