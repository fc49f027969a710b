"""Deterministic synthetic inputs for the two hot paths.

Everything here is plain numpy so the same bytes are produced in the build
container (where the verbatim reference runs) and on the GPU box.

* depth frames: 640x480 uint16 millimetres, 0 = invalid   (reference: IntegrateApp.cpp:170-188)
* trajectories: camera->world 4x4 doubles, written as ``.log`` (TSDFVolumeUnit.h:26-62)
* control lattices: ``.ctr`` text, (res+1)^3 points per fragment (ControlGrid.cpp:15-34)
* fragments: xyz + unit normals, ~200k points sampled on the room faces
"""
from __future__ import annotations

import math
import os
import struct

import numpy as np

W, H = 640, 480
FX, FY, CX, CY = 525.0, 525.0, 319.5, 239.5
SEED = 20260922
VOXEL = 3.0 / 512.0


# --------------------------------------------------------------------------- poses
def rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def rot_axis(axis, ang):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


def make_pose(R, t):
    T = np.eye(4, dtype=np.float64)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def loop_trajectory(n: int, center=(1.5, 1.5, 1.5), radius: float = 0.4, bob: float = 0.1) -> np.ndarray:
    """Rigid loop inside the room (default: the 3 m room of SURVEY 8d), looking outward.  Returns [n,4,4]."""
    out = np.empty((n, 4, 4), dtype=np.float64)
    for i in range(n):
        th = 2.0 * math.pi * i / n
        c = np.array([center[0] + radius * math.cos(th), center[1] + bob * math.sin(2 * th), center[2] + radius * math.sin(th)])
        # camera z axis (viewing direction) points outward from the room centre in the xz plane
        yaw = math.atan2(math.cos(th), math.sin(th))  # z-axis -> (cos th, 0, sin th)
        R = rot_y(yaw) @ rot_x(0.1 * math.sin(3 * th))
        out[i] = make_pose(R, c)
    return out


def identity_trajectory(n: int) -> np.ndarray:
    return np.tile(np.eye(4, dtype=np.float64), (n, 1, 1))


# --------------------------------------------------------------------------- depth
def wavy_wall_depth(n: int, w: int = W, h: int = H) -> np.ndarray:
    """Config-1 scene: d(u,v) = 1200 + 200 sin(u/80) cos(v/60) mm, slowly drifting per frame."""
    u = np.arange(w, dtype=np.float64)[None, :]
    v = np.arange(h, dtype=np.float64)[:, None]
    frames = np.empty((n, h, w), dtype=np.uint16)
    for f in range(n):
        d = 1200.0 + 200.0 * np.sin(u / 80.0 + 0.05 * f) * np.cos(v / 60.0) + 2.0 * f
        frames[f] = np.rint(d).astype(np.uint16)
    return frames


def room_depth(poses: np.ndarray, lo: float = 0.25, hi: float = 2.75, w: int = W, h: int = H,
               fx: float = FX, fy: float = FY, cx: float = CX, cy: float = CY,
               bumps: bool = False) -> np.ndarray:
    """Analytic z-depth of an axis-aligned room [lo,hi]^3 seen from camera->world `poses`."""
    poses = np.asarray(poses, dtype=np.float64).reshape(-1, 4, 4)
    u = (np.arange(w, dtype=np.float64) - cx) / fx
    v = (np.arange(h, dtype=np.float64) - cy) / fy
    dirs_c = np.stack([np.broadcast_to(u[None, :], (h, w)), np.broadcast_to(v[:, None], (h, w)),
                       np.ones((h, w))], axis=-1)  # z = 1 -> ray parameter == z-depth
    out = np.empty((len(poses), h, w), dtype=np.uint16)
    for i, T in enumerate(poses):
        d = dirs_c @ T[:3, :3].T
        o = T[:3, 3]
        with np.errstate(divide="ignore", invalid="ignore"):
            t_lo = (lo - o) / d
            t_hi = (hi - o) / d
        t = np.where(d > 0, t_hi, t_lo)
        t = np.where(d == 0, np.inf, t)
        tz = t.min(axis=-1)
        if bumps:
            p = o + d * tz[..., None]
            tz = tz + 0.01 * np.sin(7.0 * p[..., 0]) * np.cos(5.0 * p[..., 1]) * np.cos(6.0 * p[..., 2])
        mm = np.rint(tz * 1000.0)
        mm = np.where((mm > 0) & (mm < 65535), mm, 0)
        out[i] = mm.astype(np.uint16)
    return out


# --------------------------------------------------------------------------- files
def write_depth_container(path: str, frames: np.ndarray) -> None:
    """Raw depth container: int32 n,w,h then n*w*h uint16 (stand-in for .oni)."""
    frames = np.ascontiguousarray(frames, dtype="<u2")
    n, h, w = frames.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<iii", n, w, h))
        f.write(frames.tobytes())


def write_log(path: str, poses, ids=None) -> None:
    """`.log` writer, same text layout as RGBDTrajectory::SaveToFile (TSDFVolumeUnit.h:51-62)."""
    poses = np.asarray(poses, dtype=np.float64).reshape(-1, 4, 4)
    with open(path, "w") as f:
        for i, T in enumerate(poses):
            a, b, c = (i, i, i + 1) if ids is None else ids[i]
            f.write("%d\t%d\t%d\n" % (a, b, c))
            for r in range(4):
                f.write("%.8f %.8f %.8f %.8f\n" % tuple(T[r]))


def quantize_log(poses) -> np.ndarray:
    """Round poses to the 8 decimals the `.log` text carries, so in-memory == on-disk."""
    poses = np.asarray(poses, dtype=np.float64)
    return np.array([float("%.8f" % x) for x in poses.ravel()], dtype=np.float64).reshape(poses.shape)


def write_camera(path: str, fx=FX, fy=FY, cx=CX, cy=CY, icp_trunc=2.5, int_trunc=2.5) -> None:
    with open(path, "w") as f:
        for v in (fx, fy, cx, cy, icp_trunc, int_trunc):
            f.write("%.6f\n" % v)


# --------------------------------------------------------------------------- control grids
BASEPOSE = make_pose(np.eye(3), [1.5, 1.5, -0.3])


def regular_lattice(res: int, length: float) -> np.ndarray:
    """[(res+1)^3, 3] lattice, index i + j(res+1) + k(res+1)^2  (ControlGrid.h:41-43)."""
    ul = length / res
    r = np.arange(res + 1, dtype=np.float64) * ul
    k, j, i = np.meshgrid(r, r, r, indexing="ij")
    return np.stack([i.ravel(), j.ravel(), k.ravel()], axis=-1)


def make_control_grids(num: int, res: int, length: float, amp: float = 0.01, frames=None) -> np.ndarray:
    """`num` lattices, float32 [num,(res+1)^3,3]: the regular lattice of fragment c mapped into the
    fragment-0 volume frame by `frames[c]` (4x4, default identity; FragmentOptimizer/OptApp.cpp:735-750
    Pose2Ctr form) plus a smooth displacement <= amp."""
    base = regular_lattice(res, length)
    out = np.empty((num,) + base.shape, dtype=np.float32)
    for c in range(num):
        ph = 0.37 * c
        disp = amp * np.stack([np.sin(2 * np.pi * base[:, 1] / 3.0 + ph),
                               np.sin(2 * np.pi * base[:, 2] / 3.0 + 2 * ph),
                               np.sin(2 * np.pi * base[:, 0] / 3.0 + 3 * ph)], axis=-1)
        pts = base
        if frames is not None:
            F = np.asarray(frames[c], dtype=np.float64)
            pts = base @ F[:3, :3].T + F[:3, 3]
        out[c] = (pts + disp).astype(np.float32)
    return out


def write_ctr(path: str, grids: np.ndarray) -> None:
    with open(path, "w") as f:
        for g in np.asarray(grids).reshape(-1, 3):
            f.write("%.10f %.10f %.10f\n" % tuple(float(x) for x in g))


def quantize_ctr(grids: np.ndarray) -> np.ndarray:
    g = np.asarray(grids)
    return np.array([np.float32(float("%.10f" % float(x))) for x in g.ravel()], dtype=np.float32).reshape(g.shape)


# --------------------------------------------------------------------------- fragments
def make_fragment(seed: int, n_target: int = 200_000, lo: float = 0.25, hi: float = 2.75,
                  pitch: float = VOXEL, jitter: float = 0.001):
    """Points on three faces of the room meeting at a corner, sampled on a `pitch` lattice with
    uniform jitter, analytic inward normals.  Returns float32 xyz [n,3], normals [n,3]."""
    rng = np.random.default_rng(SEED + 7919 * seed)
    # choose a square patch side so that 3 faces * (side/pitch)^2 ~= n_target
    m = int(round(math.sqrt(n_target / 3.0)))
    side = m * pitch
    a = lo + np.arange(m) * pitch
    g0, g1 = np.meshgrid(a, a, indexing="ij")
    g0 = g0.ravel()
    g1 = g1.ravel()
    pts, nrm = [], []
    for axis in range(3):
        p = np.empty((m * m, 3))
        o = [ax for ax in range(3) if ax != axis]
        p[:, axis] = lo
        p[:, o[0]] = g0
        p[:, o[1]] = g1
        # gentle bumps so point-to-plane is well conditioned in all 6 dof
        p[:, axis] += 0.02 * np.sin(5.0 * p[:, o[0]]) * np.cos(4.0 * p[:, o[1]])
        n = np.zeros((m * m, 3))
        n[:, axis] = 1.0
        n[:, o[0]] = -0.02 * 5.0 * np.cos(5.0 * p[:, o[0]]) * np.cos(4.0 * p[:, o[1]])
        n[:, o[1]] = 0.02 * 4.0 * np.sin(5.0 * p[:, o[0]]) * np.sin(4.0 * p[:, o[1]])
        n /= np.linalg.norm(n, axis=1, keepdims=True)
        pts.append(p)
        nrm.append(n)
    pts = np.concatenate(pts)
    nrm = np.concatenate(nrm)
    pts = pts + rng.uniform(-jitter, jitter, size=pts.shape)
    keep = rng.random(len(pts)) < (0.99 + 0.01 * rng.random())
    del side
    return pts[keep].astype(np.float32), nrm[keep].astype(np.float32)


def fragment_frame(i: int) -> np.ndarray:
    """Local frame F_i of fragment i (local -> world): a small rigid motion about the room corner, so that every
    fragment stays inside its own [0,3]^3 volume (FragmentOptimizer/PointCloud.h:107-110).  Fragment 0 is the world."""
    if i == 0:
        return np.eye(4)
    rng = np.random.default_rng(SEED + 15485863 * i + 3)
    R = rot_axis(rng.normal(size=3), math.radians(8.0) * rng.uniform(0.2, 1.0))
    c = np.array([1.0, 1.0, 1.0])
    t = rng.uniform(-0.08, 0.08, size=3)
    return make_pose(R, c - R @ c + t)


def make_local_fragment(i: int, n_target: int = 200_000):
    """Fragment i expressed in its own frame: p_local = F_i^-1 p_world, normals rotated alike (float32)."""
    xyz, nrm = make_fragment(i, n_target)
    Finv = np.linalg.inv(fragment_frame(i))
    xyz = (xyz.astype(np.float64) @ Finv[:3, :3].T + Finv[:3, 3]).astype(np.float32)
    nrm = (nrm.astype(np.float64) @ Finv[:3, :3].T).astype(np.float32)
    return xyz, nrm


def pair_init(i: int, j: int, max_rot_deg: float = 1.5, max_trans: float = 0.015) -> np.ndarray:
    """Initial transform of pair (target i, source j): true relative pose F_i^-1 F_j times a perturbation that
    leaves nearest neighbours within reg_dist = 0.03 in the basin."""
    T_true = np.linalg.inv(fragment_frame(i)) @ fragment_frame(j)
    return T_true @ perturbation(1000 * i + j, max_rot_deg, max_trans)


def perturbation(seed: int, max_rot_deg: float = 1.5, max_trans: float = 0.015) -> np.ndarray:
    rng = np.random.default_rng(SEED + 104729 * seed + 1)
    axis = rng.normal(size=3)
    ang = math.radians(max_rot_deg) * rng.uniform(0.3, 1.0)
    t = rng.normal(size=3)
    t = t / np.linalg.norm(t) * max_trans * rng.uniform(0.3, 1.0)
    # rotate about the room corner region so the lever arm stays small
    c = np.array([1.0, 1.0, 1.0])
    R = rot_axis(axis, ang)
    return make_pose(R, c - R @ c + t)


def write_pcd_xyzrgbnormal(path: str, xyz: np.ndarray, nrm: np.ndarray) -> None:
    """Binary PCD with the field set the reference loads as PointXYZRGBNormal (CorresApp.cpp:86-97)."""
    n = len(xyz)
    rec = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgb", "<f4"),
                             ("normal_x", "<f4"), ("normal_y", "<f4"), ("normal_z", "<f4"), ("curvature", "<f4")])
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    rec["normal_x"], rec["normal_y"], rec["normal_z"] = nrm[:, 0], nrm[:, 1], nrm[:, 2]
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n"
           "FIELDS x y z rgb normal_x normal_y normal_z curvature\nSIZE 4 4 4 4 4 4 4 4\n"
           "TYPE F F F F F F F F\nCOUNT 1 1 1 1 1 1 1 1\nWIDTH %d\nHEIGHT 1\n"
           "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (n, n))
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(rec.tobytes())


def read_pcd_xyzi(path: str) -> np.ndarray:
    """Reader for the `world.pcd` the integrators write (binary PointXYZI, TSDFVolume.cpp:104-132)."""
    with open(path, "rb") as f:
        n = None
        while True:
            line = f.readline()
            if not line:
                raise ValueError("no DATA line in " + path)
            s = line.decode("ascii", "replace").strip()
            if s.startswith("POINTS"):
                n = int(s.split()[1])
            if s.startswith("DATA"):
                if s.split()[1] != "binary":
                    raise ValueError("expected DATA binary")
                break
        return np.frombuffer(f.read(16 * n), dtype="<f4").reshape(n, 4).copy()


def ensure_dir(p: str) -> str:
    os.makedirs(p, exist_ok=True)
    return p


# --------------------------------------------------------------------------- FragmentOptimizer inputs
def fragopt_initial_pose(i: int, max_rot_deg: float = 0.3, max_trans: float = 0.004) -> np.ndarray:
    """Initial pose of fragment i in the fragment-0 volume frame: the true frame F_i times a small error that the
    optimisation has to remove (fragment 0 stays the identity: it is the anchor, OptApp.cpp:459-464)."""
    if i == 0:
        return np.eye(4)
    return fragment_frame(i) @ perturbation(5000 + i, max_rot_deg, max_trans)


def write_xyzn(path: str, xyz: np.ndarray, nrm: np.ndarray) -> None:
    """`cloud_bin_xyzn_<i>.xyzn` as BuildCorrespondence --save_xyzn writes it (CorresApp.cpp:100-108): %.6f x 6."""
    with open(path, "w") as f:
        for p, n in zip(np.asarray(xyz, dtype=np.float64), np.asarray(nrm, dtype=np.float64)):
            f.write("%.6f %.6f %.6f %.6f %.6f %.6f\n" % (p[0], p[1], p[2], n[0], n[1], n[2]))


def make_fragopt_dataset(out_dir: str, num: int = 4, n_points: int = 20_000, dist: float = 0.015,
                         normal_cos: float = 0.866, pcd: bool = False, pairs=None, stride: int = 1):
    """Everything `FragmentOptimizer` reads (FragmentOptimizer/FragmentOptimizer.cpp:52-76), for `num` synthetic
    fragments: `traj.log` (one pose per fragment, use --interval 1) such that OptApp.cpp:55-70 yields the initial
    poses of `fragopt_initial_pose`; `cloud_bin_xyzn_<i>.xyzn` (or `cloud_bin_<i>.pcd`); `reg_output.log` with one record
    per pair (frame_ = number of correspondences) and `corres_<i>_<j>.txt` ("target_index source_index", ascending source
    index: CorresApp.cpp:157,175-184) found by a 1-NN search under the TRUE relative pose with the distance / normal tests
    of FindCorrespondence.  Returns a dict with the arrays (points as written, i.e. rounded to the file's 6 decimals)."""
    from scipy.spatial import cKDTree
    ensure_dir(out_dir)
    clouds = []
    for i in range(num):
        xyz, nrm = make_local_fragment(i, n_points)
        if pcd:
            write_pcd_xyzrgbnormal(os.path.join(out_dir, "cloud_bin_%d.pcd" % i), xyz, nrm)
        else:
            write_xyzn(os.path.join(out_dir, "cloud_bin_xyzn_%d.xyzn" % i), xyz, nrm)
            # what sscanf("%f") gives back from the %.6f text
            xyz = np.array([["%.6f" % v for v in row] for row in xyz.astype(np.float64)], dtype=np.float64).astype(np.float32)
            nrm = np.array([["%.6f" % v for v in row] for row in nrm.astype(np.float64)], dtype=np.float64).astype(np.float32)
        clouds.append((xyz, nrm))
    basepose = BASEPOSE
    traj = [np.linalg.inv(basepose) @ fragopt_initial_pose(i) @ basepose for i in range(num)]
    write_log(os.path.join(out_dir, "traj.log"), traj)
    if pairs is None:
        pairs = [(i, j) for i in range(num) for j in range(i + 1, num)]
    recs, poses, corres = [], [], {}
    for (i, j) in pairs:
        T = np.linalg.inv(fragment_frame(i)) @ fragment_frame(j)
        src = clouds[j][0].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
        sn = clouds[j][1].astype(np.float64) @ T[:3, :3].T
        d, idx = cKDTree(clouds[i][0].astype(np.float64)).query(src)
        ok = (d < dist) & (np.einsum("ij,ij->i", sn, clouds[i][1][idx].astype(np.float64)) > normal_cos)
        s = np.nonzero(ok)[0][::stride]
        c = np.stack([idx[s], s], axis=1).astype(np.int32)
        corres[(i, j)] = c
        with open(os.path.join(out_dir, "corres_%d_%d.txt" % (i, j)), "w") as f:
            f.write("".join("%d %d\n" % (a, b) for a, b in c))
        recs.append((i, j, len(c)))
        poses.append(T)
    write_log(os.path.join(out_dir, "reg_output.log"), poses, ids=recs)
    return {"clouds": clouds, "traj": quantize_log(np.array(traj)), "pairs": pairs, "corres": corres,
            "reg_poses": quantize_log(np.array(poses))}
