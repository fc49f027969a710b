"""Shared scene builders and the host-logic emulation used by the parity tests.

`run_app` restates CIntegrateApp::{Init,Execute} (Integrate/IntegrateApp.cpp:43-79,190-226) on top of
any object exposing integrate(depth, T) / reproject(...) -- the C oracle or the CUDA path -- so the same
frame gating drives both sides of a comparison.
"""
from __future__ import annotations

import numpy as np

from elasticreconstruction_b200 import synth


def compose_traj(pose_traj, seg_traj, interval, mul):
    """traj[i*I+j] = pose[i] * seg[i*I+j]  (IntegrateApp.cpp:68-75)."""
    out = []
    for i in range(len(pose_traj)):
        for j in range(interval):
            out.append(mul(pose_traj[i], seg_traj[i * interval + j]))
    return np.array(out)


def run_app(vol, frames, traj, seg_traj=None, grids=None, res=8, length=3.0, interval=50,
            start_from=-1, end_at=100000000, frame_flags=None):
    """Drive `vol` the way CIntegrateApp::Execute does; returns the number of frames integrated."""
    done = 0
    num = 0 if grids is None else len(grids)
    for fid in range(1, len(frames) + 1):             # FrameID is 1-based (IntegrateApp.cpp:185)
        if frame_flags is not None and frame_flags[fid - 1] == -1:
            continue                                  # :196
        if fid >= len(traj):
            break                                     # :200
        if fid < start_from or fid > end_at:
            if fid > end_at:
                break
            continue                                  # :209-216
        depth = frames[fid - 1]
        if num > 0:
            if fid > interval * num:
                break                                 # :230-233
            chunk = (fid - 1) // interval
            depth = vol.reproject(depth, traj[fid - 1], traj[0], seg_traj[0], seg_traj[fid - 1],
                                  grids[chunk], res, length)
        vol.integrate(depth, traj[fid - 1])
        done += 1
    return done


# ----------------------------------------------------------------------------- scenes
def scene_wall(n=3):
    frames = synth.wavy_wall_depth(n)
    traj = synth.identity_trajectory(n + 1)
    return dict(name="wall_rigid", frames=frames, traj=traj)


def scene_room(n=4, total=40, bumps=False):
    P = synth.quantize_log(synth.loop_trajectory(total)[: n + 1])
    frames = synth.room_depth(P[:n], bumps=bumps)
    return dict(name="room_rigid", frames=frames, traj=P)


def scene_camera(n=2):
    """Non-default intrinsics and a short integration_trunc so ScaleDepth zeroes far pixels."""
    frames = synth.wavy_wall_depth(n)
    traj = synth.identity_trajectory(n + 1)
    traj[:, 0, 3] = np.linspace(0.0, 0.02, n + 1)
    traj = synth.quantize_log(traj)
    return dict(name="camera_trunc", frames=frames, traj=traj,
                camera=dict(fx=500.0, fy=510.0, cx=310.25, cy=245.75, icp_trunc=2.5, int_trunc=1.3))


def scene_ctr(num=2, interval=3, res=8, length=3.0):
    """ControlGrid mode: `num` fragments of `interval` frames, wavy wall seen from near the base pose."""
    n = num * interval
    frames = synth.wavy_wall_depth(n)
    seg, pose, frag_frames = [], [], []
    for c in range(num):
        Wc = synth.make_pose(synth.rot_axis([0.2, 1.0, 0.1], 0.02 * c), [0.01 * c, -0.005 * c, 0.004 * c])
        pose.append(Wc)
    W0inv = np.linalg.inv(pose[0])
    for c in range(num):
        frag_frames.append(W0inv @ pose[c])
        for j in range(interval):
            motion = synth.make_pose(synth.rot_axis([0.0, 1.0, 0.0], 0.004 * j), [0.003 * j, 0.0, 0.002 * j])
            seg.append(synth.BASEPOSE @ motion)
    seg.append(seg[-1])                               # seg_traj needs >= num*interval entries
    seg = synth.quantize_log(np.array(seg))
    pose = synth.quantize_log(np.array(pose))
    grids = synth.quantize_ctr(synth.make_control_grids(num, res, length, amp=0.01, frames=frag_frames))
    return dict(name="wall_ctr", frames=frames, seg_traj=seg, pose_traj=pose, grids=grids,
                res=res, length=length, interval=interval, num=num)


def random_scene(seed, n, cols=640, rows=480, holes=True):
    """Random camera poses inside the room (any orientation, sometimes hugging a wall), analytic room depth with
    random invalid holes, far spikes beyond integration_trunc and near spikes: stresses the conservative
    SKIP / ONE / FULL classification (camera-plane straddling bricks, image borders, depth discontinuities)."""
    rng = np.random.default_rng(seed)
    poses, frames = [], []
    for _ in range(n):
        c = rng.uniform(0.45, 2.55, size=3)
        R = synth.rot_axis(rng.normal(size=3), rng.uniform(0, np.pi))
        poses.append(synth.make_pose(R, c))
    poses = synth.quantize_log(np.array(poses))
    fx = 525.0 * cols / 640.0
    d = synth.room_depth(poses, w=cols, h=rows, fx=fx, fy=fx, cx=(cols - 1) / 2.0, cy=(rows - 1) / 2.0, bumps=True)
    if holes:
        for f in range(n):
            m = rng.random((rows, cols))
            d[f][m < 0.02] = 0                                   # dropouts
            y0, x0 = rng.integers(0, rows - 40), rng.integers(0, cols - 60)
            d[f][y0:y0 + 40, x0:x0 + 60] = 0                     # a rectangular hole
            y1, x1 = rng.integers(0, rows - 30), rng.integers(0, cols - 30)
            d[f][y1:y1 + 30, x1:x1 + 30] = rng.integers(300, 500)  # a near object (depth discontinuity)
            d[f][m > 0.9995] = 6000                              # beyond integration_trunc -> scaled depth 0
    return d, poses, (fx, fx, (cols - 1) / 2.0, (rows - 1) / 2.0)


ALL_SCENES = [scene_wall, scene_room, scene_camera, scene_ctr]


def write_scene(sc, d):
    """Materialise a scene as the files the reference CLI reads; returns its argv (without the binary)."""
    import os
    synth.write_depth_container(os.path.join(d, "depth.raw"), sc["frames"])
    args = ["-oni", os.path.join(d, "depth.raw"), "--save_to", os.path.join(d, "world.pcd")]
    if "camera" in sc:
        synth.write_camera(os.path.join(d, "camera.param"), **sc["camera"])
        args += ["--camera", os.path.join(d, "camera.param")]
    if "grids" in sc:
        synth.write_log(os.path.join(d, "seg.log"), sc["seg_traj"])
        synth.write_log(os.path.join(d, "pose.log"), sc["pose_traj"])
        synth.write_ctr(os.path.join(d, "grids.ctr"), sc["grids"])
        args += ["--seg_traj", os.path.join(d, "seg.log"), "--pose_traj", os.path.join(d, "pose.log"),
                 "--ctr", os.path.join(d, "grids.ctr"), "--num", str(sc["num"]),
                 "--resolution", str(sc["res"]), "--length", str(sc["length"]), "--interval", str(sc["interval"])]
    else:
        synth.write_log(os.path.join(d, "traj.log"), sc["traj"])
        args += ["--ref_traj", os.path.join(d, "traj.log")]
    return args


def run_scene(vol, sc, mul=None):
    """Integrate a scene into `vol` through run_app; returns frames done."""
    if "camera" in sc:
        c = sc["camera"]
        vol.set_camera(c["fx"], c["fy"], c["cx"], c["cy"], c["int_trunc"])
    if "grids" in sc:
        traj = compose_traj(sc["pose_traj"], sc["seg_traj"], sc["interval"], mul)
        return run_app(vol, sc["frames"], traj, sc["seg_traj"], sc["grids"], sc["res"], sc["length"], sc["interval"])
    return run_app(vol, sc["frames"], sc["traj"])


# ----------------------------------------------------------------------------------------------- FragmentOptimizer cases
FRAGOPT_CASES = {
    # name: (num, points per fragment, extra flags, pcd input, blacklist)
    "three": dict(num=3, n_points=6000, pcd=False, blacklist=None, flags=["--iteration", "2"], modes=("slac", "rigid")),
    "four_pcd_blacklist": dict(num=4, n_points=4000, pcd=True, blacklist=[3], flags=["--iteration", "2", "--weight", "0.5"],
                               modes=("slac", "rigid")),
    # the nonrigid system is dense here (num * 3 (res+1)^3 unknowns): resolution 4 keeps the CPU emulation's Cholesky small
    "three_res4": dict(num=3, n_points=6000, pcd=False, blacklist=None,
                       flags=["--iteration", "2", "--inner_iteration", "3", "--resolution", "4", "--weight", "2.0"], modes=("nonrigid",)),
}
FRAGOPT_CASE_MODES = [(n, m) for n, c in FRAGOPT_CASES.items() for m in c["modes"]]


def fragopt_case(name: str, out_dir: str):
    """Writes the inputs of one FragmentOptimizer case under out_dir/data/ and returns (dataset dict, argv without the
    mode flag).  The data set is deterministic (synth.make_fragopt_dataset)."""
    import os
    from elasticreconstruction_b200 import synth
    c = FRAGOPT_CASES[name]
    data = os.path.join(out_dir, "data") + os.sep
    d = synth.make_fragopt_dataset(data, num=c["num"], n_points=c["n_points"], pcd=c["pcd"])
    args = ["--rgbdslam", data + "traj.log", "--registration", data + "reg_output.log", "--dir", data, "--num", str(c["num"]),
            "--interval", "1", "--blacklistpair", "1000"] + list(c["flags"])
    if c["blacklist"]:
        with open(os.path.join(out_dir, "blacklist.txt"), "w") as f:
            f.write("# fragments to drop\n" + "".join("%d\n" % b for b in c["blacklist"]))
        args += ["--blacklist", os.path.join(out_dir, "blacklist.txt")]
    return d, args


def read_log_poses(path: str):
    import numpy as np
    L = open(path).read().split("\n")
    out, i = [], 0
    while i + 4 < len(L):
        out.append([[float(x) for x in L[i + 1 + r].split()] for r in range(4)])
        i += 5
    return np.array(out, dtype=np.float64)
