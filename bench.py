#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native hot paths (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--metric tsdf|icp|fragopt]

Metric (BASELINE.json, first half): depth frames/s integrated into a 512^3 TSDF (configs[1]: 1000 synthetic 640x480
frames, rigid trajectory, 512^3 = 8^3 blocks of 64^3 voxels @ 3/512 m).  A "step" is one pass over that sequence:
--frames-per-step (default: all 1000) consecutive frames handed to the library in one call.

  value         frames/s with the frames already resident in HBM (er_tsdf_integrate_device)
  e2e           frames/s through the C ABI with HOST buffers: pinned host frames -> device (inside the timed region),
                integration, and the step's result -- the SaveWorld voxel list (er_tsdf_extract) -- back in pinned
                host memory
  roofline      integrate kernel (the production instantiation): algorithmic bytes (16 B x updated voxels + 2 B x W x H
                per frame) / its event-bracketed device time, against the measured HBM peak; the updated
                voxels are counted by a separate pass with the statistics build of the kernels
  cpu_baseline  the verbatim reference binary (oracle/_ref/Integrate_ref, its 8 OpenMP threads) on a bounded
                prefix of the same frames, timed per frame on this box's host cores
  ctrlgrid      BASELINE configs[3]: the ControlGrid (nonrigid re-projection) mode on a 6 m room, pool of 4096 blocks
  icp           second half of the metric: fragment-pair ICP iterations/s (icp_bench.py); `--metric icp` prints it as
                the line itself, and `--impl reference --metric icp` times the CPU port of that path
  fragopt       the stage between the two halves: SLAC iterations of FragmentOptimizer over all
                correspondences of the same 50 fragments (fragopt_bench.py); `--metric fragopt` / `--impl reference --metric fragopt`

N > 1 (torchrun, one rank per GPU): ONE volume, its 64^3 blocks partitioned over the ranks ((xi+yi+zi) % N == rank,
er_tsdf_set_partition); every rank sees every frame, no data-path collective during integration ("scaling":
"strong": the sequence is fixed, value = its frames / max-over-ranks time).  The e2e feed is sharded: rank r copies
1/N of each chunk of frames over its own PCIe link and the ranks all-gather the chunk over NVLink (NCCL).  After the
timed regions the union of the partitions is checked against an unpartitioned volume (order-independent digest of the
extracted voxels).  The independent-replica number of round 1 (every rank its own volume and frame stream, weak
scaling) is kept as `replica`.
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
TSDF_METRIC = "depth frames/sec into 512^3 TSDF"
ICP_METRIC = "fragment-pair ICP iterations/sec"
WORKLOAD = "TSDF integrate: 1000 synthetic 640x480 frames, rigid trajectory, 512^3 @ 3/512 m (BASELINE configs[1])"


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


def make_sequence(rank: int = 0, world: int = 1):
    """The configs[1] trajectory (camera->world).  Rank r of a replica run starts r/world of the way round the loop."""
    poses = synth.quantize_log(synth.loop_trajectory(SEQ_FRAMES))
    off = (rank * SEQ_FRAMES) // max(world, 1)
    order = (np.arange(SEQ_FRAMES) + off) % SEQ_FRAMES
    return poses[order]


def room_depth_device(poses, device, lo=0.25, hi=2.75):
    """synth.room_depth evaluated with torch on the GPU (same float64 formula; input generation is not timed)."""
    import torch
    T = torch.as_tensor(np.asarray(poses, dtype=np.float64).reshape(-1, 4, 4), device=device)
    u = (torch.arange(W, dtype=torch.float64, device=device) - synth.CX) / synth.FX
    v = (torch.arange(H, dtype=torch.float64, device=device) - synth.CY) / synth.FY
    dirs = torch.stack([u[None, :].expand(H, W), v[:, None].expand(H, W), torch.ones((H, W), dtype=torch.float64, device=device)], -1)
    out = torch.empty((len(T), H, W), dtype=torch.int16, device=device)
    for i in range(len(T)):
        d = dirs @ T[i, :3, :3].T
        o = T[i, :3, 3]
        t = torch.where(d > 0, (hi - o) / d, (lo - o) / d)
        t = torch.where(d == 0, torch.full_like(t, float("inf")), t)
        mm = torch.round(t.min(dim=-1).values * 1000.0)          # round half to even, like np.rint
        mm = torch.where((mm > 0) & (mm < 65535), mm, torch.zeros_like(mm))
        out[i] = mm.to(torch.int32).to(torch.int16)               # u16 bit pattern
    return out


def voxel_digest(xyzi: np.ndarray):
    """Order-independent digest of an extracted voxel list [n,4] float32: (count, sum of 64-bit row hashes mod 2^64 as
    two 32-bit halves).  Equal multisets give equal digests whatever the order or the partition."""
    if len(xyzi) == 0:
        return np.array([0, 0, 0], dtype=np.int64)
    r = np.ascontiguousarray(xyzi, dtype=np.float32).view(np.uint32).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = r[:, 0] * np.uint64(0x9E3779B97F4A7C15)
        h ^= (r[:, 1] + np.uint64(0x7F4A7C15)) * np.uint64(0xC2B2AE3D27D4EB4F)
        h ^= (r[:, 2] + np.uint64(0x165667B1)) * np.uint64(0xD6E8FEB86659FD93)
        h ^= (r[:, 3] + np.uint64(0x27D4EB4F)) * np.uint64(0xA0761D6478BD642F)
        h ^= h >> np.uint64(29)
        h *= np.uint64(0xBF58476D1CE4E5B9)
        h ^= h >> np.uint64(32)
        s = np.uint64(h.sum(dtype=np.uint64))
    return np.array([len(xyzi), int(s & np.uint64(0xFFFFFFFF)), int(s >> np.uint64(32))], dtype=np.int64)


def digest_equal(total, single):
    """`total` = element-wise sum over ranks of voxel_digest() triples; compare with one triple modulo 2^64."""
    if int(total[0]) != int(single[0]):
        return False
    a = (int(total[1]) + (int(total[2]) << 32)) % (1 << 64)
    b = (int(single[1]) + (int(single[2]) << 32)) % (1 << 64)
    return a == b


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
    if args.metric == "icp":
        import icp_bench
        return icp_bench.run_reference(args)
    if args.metric == "fragopt":
        import fragopt_bench
        return fragopt_bench.run_reference(args)
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
        "impl": "reference", "metric": TSDF_METRIC, "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": fps_step, "threads": 8, "host_cores": os.cpu_count()},
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


# --------------------------------------------------------------------------------------------- ControlGrid leg
def ctrlgrid_scene(n_frames: int, interval: int = 50, res: int = 8, length: float = 3.0):
    """BASELINE configs[3] inputs: a loop through a 6 m room (16^3 blocks of 64^3 voxels = "~1024^3 hashed"),
    fragments of `interval` frames, one control lattice per fragment = the fragment's pose applied to the regular
    lattice (Pose2Ctr form, FragmentOptimizer/OptApp.cpp:735-750) plus a smooth displacement <= 1 cm; seg_traj = the
    frame's pose relative to its fragment's volume (base pose t = (1.5, 1.5, -0.3), CorresApp.cpp:46-49)."""
    num = n_frames // interval
    world = synth.quantize_log(synth.loop_trajectory(n_frames, center=(3.0, 3.0, 3.0), radius=1.6, bob=0.4))
    base = synth.BASEPOSE
    base_inv = np.linalg.inv(base)
    pose, seg = [], []
    for c in range(num):
        pose.append(world[c * interval] @ base_inv)
        first_inv = np.linalg.inv(world[c * interval])
        for j in range(interval):
            seg.append(base @ first_inv @ world[c * interval + j])
    seg.append(seg[-1])
    pose = synth.quantize_log(np.array(pose))
    seg = synth.quantize_log(np.array(seg))
    frames0 = [np.linalg.inv(pose[0]) @ pose[c] for c in range(num)]
    grids = synth.quantize_ctr(synth.make_control_grids(num, res, length, amp=0.01, frames=frames0))
    return dict(world=world, pose=pose, seg=seg, grids=grids, num=num, interval=interval, res=res, length=length)


def ctrlgrid_leg(args, local, rank, world, barrier, capi, torch, dist):
    """Frames resident in HBM -> er_tsdf_reproject-on-device + ScaleDepth + Integrate per batch
    (IntegrateApp.cpp:217-225 per frame), device-timed; reproject_kernel timed on its own."""
    n = args.ctr_frames
    sc = ctrlgrid_scene(n)
    # traj[i*I+j] = pose[i] * seg[i*I+j] with the library's Eigen-order product (IntegrateApp.cpp:73)
    traj = np.array([capi.mat4_mul(sc["pose"][f // sc["interval"]], sc["seg"][f]) for f in range(n)])
    d_frames = room_depth_device(traj, torch.device("cuda", local), lo=0.25, hi=5.75)
    chunks = (np.arange(n) // sc["interval"]).astype(np.int32)
    vol = capi.TSDFVolume(device=local, max_blocks=args.ctr_pool)
    vol.set_fusion(args.fusion)
    vol.set_control_grids(sc["grids"], sc["res"], sc["length"])
    if world > 1:
        vol.set_partition(rank, world)
    T0, A0 = traj[0], sc["seg"][0]

    def run():
        vol.integrate_ctr_device(d_frames.data_ptr(), traj, T0, A0, sc["seg"][:n], chunks)

    for _ in range(3):
        run()
    vol.sync()
    barrier()
    l0 = vol.launch_count()
    vol.event_record(0)
    for _ in range(args.ctr_steps):
        run()
    vol.event_record(1)
    vol.sync()
    barrier()
    ms = vol.event_elapsed_ms(0, 1)
    launches = vol.launch_count() - l0
    rp_ms = vol.profile_reproject_device(d_frames.data_ptr(), traj, T0, A0, sc["seg"][:n], chunks)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nb = vol.num_blocks()
    nbt = torch.tensor([float(nb)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(nbt, op=dist.ReduceOp.SUM)
    out = {"metric": "depth frames/sec, ControlGrid mode (BASELINE configs[3])",
           "value": n * args.ctr_steps / (float(t.item()) * 1e-3), "unit": "frames/s",
           "config": {"workload": "TSDF integrate with ControlGrid nonrigid re-projection: %d synthetic 640x480 frames of a loop "
                                  "through a 6 m room, %d fragments x %d frames, lattice res %d, block pool %d per rank, blocks "
                                  "partitioned over %d rank(s)" % (n, sc["num"], sc["interval"], sc["res"], args.ctr_pool, world),
                      "frames": n, "passes_timed": args.ctr_steps},
           "ms_per_pass": float(t.item()) / args.ctr_steps, "allocated_blocks": int(nbt.item()), "gpu_launches": int(launches),
           "reproject_us_per_frame": 1e3 * rp_ms / n,
           "reproject_note": "reproject_kernel + z-buffer reset/convert of one frame, event-bracketed on the stream; algorithmic "
                             "bytes 2 B read + 4 B atomicMin + 4+2 B convert per pixel = 3.7 MB/frame"}
    del vol
    if rank == 0 and world == 1 and args.ctr_cpu_frames > 0:
        out["cpu_baseline"] = ctrlgrid_cpu_leg(sc, traj, d_frames, args.ctr_cpu_frames)
    return out


def ctrlgrid_cpu_leg(sc, traj, d_frames, n_cpu: int):
    """The verbatim reference binary in ControlGrid mode (--pose_traj --seg_traj --ctr --num) on
    the first fragments of the same scene; per-frame wall time of Execute(), which includes its serial Reproject
    (IntegrateApp.cpp:228-269)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    if not os.path.exists(pyoracle.REF_INTEGRATE):
        return None
    I = sc["interval"]
    num = max(1, n_cpu // I)
    n = num * I
    with tempfile.TemporaryDirectory() as d:
        # (two frames fewer than the trajectories cover: the stream must end before the application's own exit condition
        # -- frame id reaching the trajectory's length, IntegrateApp.cpp:200 -- for the shim to report its time stamps)
        frames = d_frames[: n - 2].cpu().numpy().view(np.uint16)
        synth.write_depth_container(os.path.join(d, "depth.raw"), frames)
        synth.write_log(os.path.join(d, "seg.log"), sc["seg"][: n + 1])
        synth.write_log(os.path.join(d, "pose.log"), sc["pose"][:num])
        synth.write_ctr(os.path.join(d, "grids.ctr"), sc["grids"][:num])
        tfile = os.path.join(d, "stamps.txt")
        env = dict(os.environ, ER_SHIM_TIMING_FILE=tfile)
        subprocess.run([pyoracle.REF_INTEGRATE, "-oni", os.path.join(d, "depth.raw"), "--seg_traj", os.path.join(d, "seg.log"),
                        "--pose_traj", os.path.join(d, "pose.log"), "--ctr", os.path.join(d, "grids.ctr"), "--num", str(num),
                        "--resolution", str(sc["res"]), "--length", str(sc["length"]), "--interval", str(I),
                        "--save_to", os.path.join(d, "w.pcd")], cwd=d, env=env, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, check=True)
        per = np.diff(np.loadtxt(tfile))
    per = per[: n - 2]
    skip = min(5, len(per) // 4)
    fps = (len(per) - skip) / float(per[skip:].sum())
    return {"value": fps, "unit": "frames/s", "cores": min(8, os.cpu_count() or 1), "kind": "reference",
            "sample": "first %d frames (%d fragment(s)) of the same scene through the verbatim reference binary in --ctr mode, "
                      "per-frame wall time of Execute() incl. its serial Reproject; %d warm-up frames dropped" % (n, num, skip)}


# --------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--metric", default="tsdf", choices=["tsdf", "icp", "fragopt"],
                    help="which half of BASELINE.json's metric the printed line carries (the TSDF line always embeds `icp`)")
    ap.add_argument("--frames-per-step", type=int, default=SEQ_FRAMES,
                    help="frames handed to the library per call; must divide the 1000-frame sequence")
    ap.add_argument("--ref-frames-per-step", type=int, default=8)
    ap.add_argument("--ref-pairs-per-step", type=int, default=8)
    ap.add_argument("--cpu-frames", type=int, default=120, help="frames timed for cpu_baseline (0 = skip)")
    ap.add_argument("--fusion", type=int, default=32, help="frames per fused launch triple (1..32)")
    ap.add_argument("--no-icp", action="store_true")
    ap.add_argument("--no-fragopt", action="store_true")
    ap.add_argument("--fragopt-multi", action="store_true",
                    help="N > 1 only: run the FragmentOptimizer leg sharded over the ranks (pairs round-robin + one all-reduce of the matrix "
                         "per iteration).  Off by default: gloo-tested on the CPU, not yet run on several GPUs")
    ap.add_argument("--fragopt-fragments", type=int, default=50)
    ap.add_argument("--fragopt-points", type=int, default=200000)
    ap.add_argument("--fragopt-iterations", type=int, default=3)
    ap.add_argument("--no-ctr", action="store_true")
    ap.add_argument("--no-replica", action="store_true")
    ap.add_argument("--icp-fragments", type=int, default=50)
    ap.add_argument("--icp-points", type=int, default=200000)
    ap.add_argument("--ctr-frames", type=int, default=1000)
    ap.add_argument("--ctr-steps", type=int, default=3)
    ap.add_argument("--ctr-pool", type=int, default=4096)
    ap.add_argument("--ctr-cpu-frames", type=int, default=100, help="frames of the ControlGrid scene timed on the reference binary (0 = skip)")
    ap.add_argument("--e2e-chunk", type=int, default=200, help="frames per H2D/all-gather chunk of the sharded feed (N > 1)")
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
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_src = load_peaks()
    if args.metric == "icp":
        import icp_bench
        icp = icp_bench.run(local, rank, world, barrier, peak_gbs=peak, num_fragments=args.icp_fragments,
                            points=args.icp_points, clock_sampler=ClockSampler)
        if rank == 0:
            line = icp_bench.as_line(icp, args, world)
            sys.stdout.flush()
            os.write(real_stdout, (json.dumps(line) + "\n").encode())
        if world > 1:
            dist.destroy_process_group()
        return 0

    if args.metric == "fragopt":
        import fragopt_bench
        if world > 1:
            if rank == 0:
                os.write(real_stdout, (json.dumps({"metric": fragopt_bench.METRIC, "unavailable": "the FragmentOptimizer leg is single-GPU this round"}) + "\n").encode())
            dist.destroy_process_group()
            return 0
        fo = fragopt_bench.run(local, fragments=args.fragopt_fragments, points=args.fragopt_points, iterations=args.fragopt_iterations,
                               clock_sampler=ClockSampler, peak_gbs=peak)
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(fragopt_bench.as_line(fo, args, world)) + "\n").encode())
        return 0

    F = args.frames_per_step
    nsteps = args.warmup + args.steps
    need = min(SEQ_FRAMES, F * nsteps)
    assert need % F == 0, "--frames-per-step must divide the resident sequence so every step is one contiguous run"
    step_idx = [(np.arange(F) + s * F) % need for s in range(nsteps)]
    partitioned = world > 1

    def new_volume(part: bool, pool: int = 1024):
        v = capi.TSDFVolume(device=local, max_blocks=pool)
        v.set_fusion(args.fusion)
        if part:
            v.set_partition(rank, world)
        return v

    def timed_device_run(vol, d_frames, poses):
        """W warm-up + K timed steps with the frames resident; returns (max-over-ranks device ms, launches, clocks, wall)."""
        def step(s):
            idx = step_idx[s]
            vol.integrate_device(d_frames.data_ptr() + int(idx[0]) * BYTES_PER_FRAME, poses[idx], F)
        for s in range(args.warmup):
            step(s)
        vol.sync()
        sampler = ClockSampler(local)
        barrier()
        sampler.start()
        l0 = vol.launch_count()
        vol.event_record(0)
        t0 = time.perf_counter()
        for s in range(args.warmup, nsteps):
            step(s)
        vol.event_record(1)
        vol.sync()
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop()
        return rmax(vol.event_elapsed_ms(0, 1)), vol.launch_count() - l0, clocks, wall

    # ---------------- value: the configs[1] sequence, frames resident in HBM; N > 1: one volume, blocks partitioned
    poses = make_sequence(0, 1)
    d_frames = room_depth_device(poses[:need], dev)
    vol = new_volume(partitioned)
    dev_ms, launches, clocks, wall = timed_device_run(vol, d_frames, poses)
    total_frames = F * args.steps
    value = total_frames / (dev_ms * 1e-3)
    owned_blocks = vol.num_blocks()
    del vol

    # ---------------- union check (N > 1, untimed): the partitions together hold exactly the unpartitioned volume
    union = None
    if partitioned:
        vp = new_volume(True)
        vp.integrate_device(d_frames.data_ptr(), poses[:need], need)
        dg = torch.from_numpy(voxel_digest(vp.extract())).cuda()
        del vp
        dist.all_reduce(dg, op=dist.ReduceOp.SUM)
        ok = None
        if rank == 0:
            vs = new_volume(False)
            vs.integrate_device(d_frames.data_ptr(), poses[:need], need)
            single = voxel_digest(vs.extract())
            del vs
            ok = bool(digest_equal(dg.cpu().numpy(), single))
        union = {"voxels": int(dg[0].item()), "equal_to_unpartitioned_volume": ok,
                 "how": "order-independent 64-bit digest of (x, y, z, tsdf) rows of er_tsdf_extract after one pass over the "
                        "sequence: sum over ranks vs one unpartitioned volume on rank 0"}

    # ---------------- roofline: production kernels timed one by one (overlap off), updated voxels counted separately
    vol2 = new_volume(partitioned)
    for s in range(args.warmup):
        idx = step_idx[s]
        vol2.integrate_device(d_frames.data_ptr() + int(idx[0]) * BYTES_PER_FRAME, poses[idx], F)
    vol2.sync()
    t_touch = t_cls = t_int = 0.0
    for s in range(args.warmup, nsteps):
        idx = step_idx[s]
        a, b, c = vol2.profile_device(d_frames.data_ptr() + int(idx[0]) * BYTES_PER_FRAME, poses[idx], F)
        t_touch += a
        t_cls += b
        t_int += c
    del vol2
    # the statistics build of the kernels on one pass over the sequence in steady state (every pass updates the same voxels)
    vol2 = new_volume(partitioned)
    vol2.integrate_device(d_frames.data_ptr(), poses[:need], need)
    vol2.set_stats(True)
    vol2.sync()
    st0 = vol2.stats()
    vol2.integrate_device(d_frames.data_ptr(), poses[:need], need)
    st1 = vol2.stats()
    del vol2
    n_prof = F * args.steps
    fusion = args.fusion
    n_launch = args.steps * ((F + fusion - 1) // fusion)          # integrate_kernel launches in the profiled pass
    upd_pf = (st1["updated_voxels"] - st0["updated_voxels"]) / need           # per frame, this rank's blocks
    blocks_pf = (st1["touched_blocks"] - st0["touched_blocks"]) / need
    alg_bytes_per_launch = (16.0 * upd_pf + 2.0 * W * H) * n_prof / n_launch
    int_s_per_launch = t_int * 1e-3 / n_launch
    achieved = alg_bytes_per_launch / int_s_per_launch / 1e9
    full_pairs = (st1["full_pairs"] - st0["full_pairs"]) / need
    one_pairs = (st1["one_pairs"] - st0["one_pairs"]) / need
    roofline = {"bound": "hbm", "kernel": "integrate_kernel<false,4> (production instantiation)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": load_traffic("integrate_kernel"),
                "traffic_source": "ncu --set full capture of one 32-frame launch (profiles/traffic.json)",
                "algorithmic_bytes_per_launch": alg_bytes_per_launch, "us_per_launch": int_s_per_launch * 1e6,
                "frames_per_launch": n_prof / n_launch,
                "note": "algorithmic bytes = 16 B x updated voxels + 2 B x W x H per frame (SURVEY 8d); one launch applies a whole "
                        "batch of frames to register-resident voxels and consecutive batches find their blocks in L2, so DRAM "
                        "traffic is far below the algorithmic bytes: the kernel is instruction-issue bound (see `issue`), the "
                        "HBM fraction is reported because the metric asks for it",
                "issue": load_issue_stats(),
                "touch_us_per_frame": t_touch * 1e3 / n_prof, "classify_us_per_frame": t_cls * 1e3 / n_prof,
                "integrate_us_per_frame": t_int * 1e3 / n_prof, "serial_us_per_frame": (t_touch + t_cls + t_int) * 1e3 / n_prof,
                "pipelined_us_per_frame": dev_ms * 1e3 / total_frames,
                "updated_voxels_per_frame": upd_pf, "active_blocks_per_frame": blocks_pf,
                "one_voxels_per_frame": 64.0 * one_pairs, "full_voxels_per_frame": 64.0 * full_pairs,
                "dense_bytes_per_frame": 16.0 * 262144 * blocks_pf + 2.0 * W * H}

    # ---------------- e2e: host buffers through the C ABI, H2D inside the timed region, the voxel list back on the host
    # page-locked host frames from the library's allocator (er_host_alloc places them on the NUMA node of this rank's GPU)
    host_frames = capi.HostBuffer(need * BYTES_PER_FRAME)
    pinned = torch.from_numpy(host_frames.array(np.int16)).view(need, H, W)
    pinned.copy_(d_frames)
    vol3 = new_volume(partitioned)
    cap_rows = 0
    result = None

    if partitioned:
        # sharded feed: rank r copies its 1/N slice of every chunk over its own PCIe link, the ranks all-gather the chunk
        # over NVLink (NCCL), the library integrates it from HBM; two chunk buffers so that the copy/gather of chunk c + 1
        # runs while chunk c is integrated
        C = next((c for c in range(min(args.e2e_chunk, F), 0, -1) if F % c == 0 and c % world == 0), None)
        assert C, "--frames-per-step must be a multiple of the rank count"
        per = C // world
        gath = [torch.empty((C, H, W), dtype=torch.int16, device=dev) for _ in range(2)]
        mine = [torch.empty((per, H, W), dtype=torch.int16, device=dev) for _ in range(2)]
        feed_stream = torch.cuda.Stream(device=dev)

        def e2e_step(s):
            f0 = int(step_idx[s][0])
            nchunk = F // C
            evs = [None, None]

            def issue(c):                       # copy + gather chunk c into buffer c & 1 (asynchronous, feed stream)
                b = c & 1
                with torch.cuda.stream(feed_stream):
                    lo = f0 + c * C + rank * per
                    mine[b].copy_(pinned[lo:lo + per], non_blocking=True)
                    dist.all_gather_into_tensor(gath[b].view(torch.uint8), mine[b].view(torch.uint8))   # (NCCL has no int16)
                    evs[b] = torch.cuda.Event()
                    evs[b].record(feed_stream)

            issue(0)
            for c in range(nchunk):
                if c + 1 < nchunk:
                    if c >= 1:
                        vol3.inputs_consumed()  # chunk c - 1 has been read: its buffer may be overwritten by chunk c + 1
                    issue(c + 1)
                evs[c & 1].synchronize()
                vol3.integrate_device(gath[c & 1].data_ptr(), poses[f0 + c * C: f0 + (c + 1) * C], C)
            vol3.inputs_consumed()
        h2d_per_step = F * BYTES_PER_FRAME // world
        feed = "sharded: 1/N of each %d-frame chunk per rank over PCIe, NCCL all-gather over NVLink, er_tsdf_integrate_device" % C
    else:
        def e2e_step(s):
            idx = step_idx[s]
            vol3.integrate_ptr(host_frames.ptr + int(idx[0]) * BYTES_PER_FRAME, poses[idx], F)
        h2d_per_step = F * BYTES_PER_FRAME
        feed = "er_tsdf_integrate_batch from page-locked host memory (er_host_alloc, NUMA-local to the GPU)"

    for s in range(args.warmup):
        e2e_step(s)
    cap_rows = int(vol3.extract_count() * 1.03) + 1024
    # the step's result -- SaveWorld's voxel list -- goes to one of two page-locked buffers through the non-blocking
    # extraction: its device->host copy overlaps the next step's host->device copies (PCIe is full duplex); every step's
    # rows are home before the timed region ends
    result = [capi.HostBuffer(cap_rows * 16), capi.HostBuffer(cap_rows * 16)]
    vol3.extract_begin(result[0], cap_rows)
    n_rows = vol3.extract_wait()
    barrier()
    t0 = time.perf_counter()
    for s in range(args.warmup, nsteps):
        e2e_step(s)
        if s > args.warmup:
            n_rows = vol3.extract_wait()                  # the previous step's rows (normally home by now)
        vol3.extract_begin(result[s & 1], cap_rows)
    n_rows = vol3.extract_wait()
    assert n_rows <= cap_rows, "result buffer too small: %d rows" % n_rows
    barrier()
    e2e_s = rmax(time.perf_counter() - t0)
    e2e = {"value": total_frames / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int(h2d_per_step),
           "d2h_bytes_per_step": int(cap_rows * 16 + 8), "result_rows": int(n_rows), "feed": feed,
           "result": "er_tsdf_extract_begin / _wait: every step's voxel list (x, y, z, tsdf) is copied to page-locked host memory, "
                     "overlapping the next step's input copies; the last one is awaited inside the timed region",
           "timing": "wall clock around the C-ABI calls, barrier+sync both sides, max over ranks"}
    del vol3, pinned
    for rb in result:
        rb.close()
    host_frames.close()

    # ---------------- replica mode of round 1 (N > 1): every rank its own volume and its own frame stream
    replica = None
    if partitioned and not args.no_replica:
        rposes = make_sequence(rank, world)
        rframes = room_depth_device(rposes[:need], dev)
        rv = new_volume(False)
        rms, _, _, _ = timed_device_run(rv, rframes, rposes)
        replica = {"value": total_frames * world / (rms * 1e-3), "unit": "frames/s", "scaling": "weak",
                   "sharding": "independent volume and frame stream per rank (round 1's mode)"}
        del rv, rframes

    del d_frames
    torch.cuda.empty_cache()

    # ---------------- configs[3]: ControlGrid mode
    ctr = None
    if not args.no_ctr:
        ctr = ctrlgrid_leg(args, local, rank, world, barrier, capi, torch, dist)
        torch.cuda.empty_cache()

    # ---------------- ICP (second half of the metric), then the cpu baseline (rank 0, N = 1 only; last, so that no GPU leg
    # starts from the idle clocks the GPU drops to while the host is busy)
    icp = None
    if not args.no_icp:
        import icp_bench
        icp = icp_bench.run(local, rank, world, barrier, peak_gbs=peak, num_fragments=args.icp_fragments,
                            points=args.icp_points, clock_sampler=ClockSampler)
    # ---------------- FragmentOptimizer --slac, single GPU; a failure here must not cost the line
    fragopt = None
    if not args.no_fragopt and world == 1:
        try:
            import fragopt_bench
            torch.cuda.empty_cache()
            fragopt = fragopt_bench.run(local, fragments=args.fragopt_fragments, points=args.fragopt_points,
                                        iterations=args.fragopt_iterations, clock_sampler=ClockSampler, peak_gbs=peak)
        except Exception as e:      # noqa: BLE001
            fragopt = {"error": "%s: %s" % (type(e).__name__, e)}
        torch.cuda.empty_cache()
    if not args.no_fragopt and world > 1 and args.fragopt_multi:
        import fragopt_bench
        torch.cuda.empty_cache()
        fragopt = fragopt_bench.run_sharded(local, rank, world, fragments=args.fragopt_fragments, points=args.fragopt_points,
                                            iterations=args.fragopt_iterations)
        torch.cuda.empty_cache()
    cpu = None
    if rank == 0 and world == 1 and args.cpu_frames > 0:
        cpu = cpu_baseline_leg(poses, args.cpu_frames)

    if rank == 0:
        line = {
            "metric": TSDF_METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if partitioned else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": F, "frames_per_launch": args.fusion, "block_pool": 1024,
                       "sharding": ("blocks partitioned: one volume, rank r owns the 64^3 blocks with (xi+yi+zi) %% %d == r; every "
                                    "rank sees every frame; no data-path collective" % world) if partitioned else "single volume",
                       "l2": "inputs larger than L2 (resident depth sequence %.0f MB + 2 GiB block pool); consecutive "
                             "frames legitimately reuse their active blocks through L2" % (need * BYTES_PER_FRAME / 1e6)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "wall_s_timed_region": wall, "blocks_owned_rank0": owned_blocks, "union_check": union, "replica": replica,
            "ctrlgrid": ctr, "icp": icp, "fragopt": fragopt,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


def load_issue_stats():
    """Instruction-issue figures of the integrate kernel from the committed ncu capture (profiles/traffic.json): the roof
    that actually binds it."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))["integrate_kernel"].get("issue")
    except Exception:
        return None


if __name__ == "__main__":
    sys.exit(main())
