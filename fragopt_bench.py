"""FragmentOptimizer leg of bench.py (the stage between BuildCorrespondence and Integrate --ctr):
SLAC iterations over ALL correspondences of ALL fragment pairs of BASELINE configs[2]'s fragment set -- 50 synthetic
fragments (~200k points each), 1225 pairs, the correspondences found on the device by er_icp_correspondences_batch with
FindCorrespondence's thresholds (CorresApp.cpp:145-161) under the true relative poses.

A "step" is one iteration of COptApp::OptimizeSLAC (OptApp.cpp:454-664): assembly of the dense normal equations
(er_fragopt_slac_assemble: one kernel over all correspondences), the lattice regulariser (729 polar rotations, host),
the dense Cholesky solve on the device (er_fragopt_slac_solve), pose / lattice update and the re-evaluation of every
point (er_fragopt_update_points_pn).

  value         correspondences/s through whole iterations (rows x iterations / wall time of the iterations; inputs resident)
  kernel        the assembly kernel alone: rows/s from its CUDA events, and the FP64 rate that implies
                (1830 + 61 multiply-adds per row) against the FP64 FMA peak of the chip (the roof of this kernel)
  e2e           the same from host buffers: point upload + GetCoordinate, correspondence upload + sort, the iterations
  cpu_baseline  the VERBATIM reference binary (oracle/_ref/FragmentOptimizer_ref, its 8 OpenMP threads) on a bounded
                sample (4 fragments x 20k points, 6 pairs), "Neat Optimization" time per iteration

Single GPU this round (the pair list shards like the ICP one, with one all-reduce of the 49 MB matrix)."""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile
import time

import numpy as np

from elasticreconstruction_b200 import capi, synth

ROOT = os.path.dirname(os.path.abspath(__file__))
RES, LENGTH = 8, 3.0
METRIC = "correspondences/sec through SLAC iterations (FragmentOptimizer --slac)"
WORKLOAD = "FragmentOptimizer --slac: %d synthetic fragments ~%dk pts, %d pairs, %d correspondences, %d iterations"
FMA_PER_ROW = 60 * 61 // 2 + 61          # upper triangle of the 60 x 60 outer product + b * row + b * b
FP64_PEAK_TFLOPS = 37.2                  # 148 SMs x 64 FP64 FMA/clk x 1.965 GHz x 2; ncu agrees (FP64 pipe 24.8 % busy at 0.48 warp-DFMA/clk/SM,
#                                          profiles/r2_slac_assemble_kernel.txt); MEASURED_PEAKS.json has no FP64 entry


def _traffic_per_row():
    """DRAM bytes per correspondence of slac_assemble_kernel from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        import json
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["slac_assemble_kernel"]
        return float(t["dram_bytes_per_launch"]) / float(t["units_per_launch"])
    except Exception:
        return None


def lattice_nodes():
    r1 = RES + 1
    unit = LENGTH / RES
    k, j, i = np.meshgrid(np.arange(r1), np.arange(r1), np.arange(r1), indexing="ij")
    return np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1), unit


def neighbour_table():
    """[729, 6] neighbour node index or -1, in the reference's order (i-1, i+1, j-1, j+1, k-1, k+1; OptApp.cpp:594-611)."""
    ijk, _ = lattice_nodes()
    r1 = RES + 1
    out = -np.ones((len(ijk), 6), dtype=np.int64)
    for t, (ax, d) in enumerate([(0, -1), (0, 1), (1, -1), (1, 1), (2, -1), (2, 1)]):
        q = ijk.copy()
        q[:, ax] += d
        ok = (q[:, ax] >= 0) & (q[:, ax] <= RES)
        out[ok, t] = (q[ok, 0] + q[ok, 1] * r1 + q[ok, 2] * r1 * r1)
    return out


def base_matrix(num, default_weight):
    """baseJJ * default_weight + the six ones (OptApp.cpp:449-464, 813-858), upper triangle, dense."""
    nb = neighbour_table()
    M = 6 * num + 3 * len(nb)
    B = np.zeros((M, M))
    off = 6 * num
    for u in range(len(nb)):
        for v in nb[u][nb[u] >= 0]:
            for x in range(3):
                a, b = off + 3 * u + x, off + 3 * v + x
                B[a, a] += default_weight
                B[b, b] += default_weight
                B[min(a, b), max(a, b)] -= default_weight
    anchor = off + 3 * ((RES // 2) + (RES // 2) * (RES + 1))
    for x in range(3):
        B[anchor + x, anchor + x] += default_weight
    for k in range(6):
        B[k, k] += 1.0
    return B


def regulariser(ictr, ctr, num, default_weight, nb):
    """The regulariser loop of OptimizeSLAC (OptApp.cpp:581-635), vectorised: one 3x3 SVD per lattice node."""
    off = 6 * num
    P0 = ictr[off:].reshape(-1, 3)
    P1 = ctr[off:].reshape(-1, 3)
    valid = nb >= 0
    nbc = np.where(valid, nb, 0)
    d0 = (P0[:, None, :] - P0[nbc]) * valid[..., None]
    d1 = (P1[:, None, :] - P1[nbc]) * valid[..., None]
    Cm = np.einsum("nti,ntj->nij", d0, d1)
    U, _, Vt = np.linalg.svd(Cm)
    V = np.swapaxes(Vt, 1, 2)
    R = V @ np.swapaxes(U, 1, 2)
    neg = np.linalg.det(R) < 0
    if neg.any():
        U2 = U.copy()
        U2[neg, :, 2] *= -1
        R = V @ np.swapaxes(U2, 1, 2)
    anchor = (RES // 2) + (RES // 2) * (RES + 1)
    R[anchor] = np.eye(3)
    bx = (d1 - np.einsum("nij,ntj->nti", R, d0)) * valid[..., None]
    Jb = np.zeros_like(ictr)
    lat = Jb[off:].reshape(-1, 3)
    lat += bx.sum(axis=1) * default_weight
    np.subtract.at(lat, nbc[valid], bx[valid] * default_weight)
    return Jb, float(default_weight * np.sum(bx * bx))


def rot_zyx(r):
    cx, sx, cy, sy, cz, sz = np.cos(r[0]), np.sin(r[0]), np.cos(r[1]), np.sin(r[1]), np.cos(r[2]), np.sin(r[2])
    return np.array([[cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx],
                     [sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx],
                     [-sy, cy * sx, cy * cx]])


class SlacLoop:
    """OptimizeSLAC's loop over the C ABI (the Python twin of apps/FragmentOptimizer.cpp::optimize_slac)."""

    def __init__(self, g: "capi.FragOpt", ipose, weight=1.0):
        self.g, self.num = g, g.num
        self.M = g.M
        self.w = self.num * weight
        ijk, unit = lattice_nodes()
        self.ictr = np.zeros(self.M)
        self.ictr[6 * self.num:] = (ijk * unit).reshape(-1)
        self.ctr = self.ictr.copy()
        self.pose = np.array(ipose, dtype=np.float64).copy()
        self.nb = neighbour_table()
        self.base = base_matrix(self.num, self.w)
        self.first = True
        g.update_pose(self.pose)
        self.scores = []

    def expand(self):
        lat = self.ctr[6 * self.num:].reshape(-1, 3)
        return np.concatenate([lat @ self.pose[l][:3, :3].T + self.pose[l][:3, 3] for l in range(self.num)]).reshape(-1)

    def step(self):
        rot_t = np.array([self.pose[l][:3, :3].T for l in range(self.num)])
        Jb, score = self.g.slac_assemble(rot_t)
        base_jb, reg = regulariser(self.ictr, self.ctr, self.num, self.w, self.nb)
        x = -self.g.slac_solve(self.base if self.first else None, Jb + base_jb)
        self.first = False
        self.ctr[6 * self.num:] += x[6 * self.num:]
        for l in range(self.num):
            T = np.eye(4)
            T[:3, :3] = rot_zyx(x[6 * l: 6 * l + 3])
            T[:3, 3] = x[6 * l + 3: 6 * l + 6]
            self.pose[l] = T @ self.pose[l]
        self.g.update_points_pn(self.expand())
        self.scores.append((score, reg))
        return score


def device_correspondences(frags, pairs, device):
    """FindCorrespondence (dist 0.015, normal cos 0.866) for every pair under its true relative pose, on the device."""
    clouds = [capi.Cloud(f[0], f[1], device=device, cell_size=0.03) for f in frags]
    T = np.array([np.linalg.inv(synth.fragment_frame(i)) @ synth.fragment_frame(j) for (i, j) in pairs])
    rows, _ = capi.icp_correspondences_batch(clouds, [p[0] for p in pairs], [p[1] for p in pairs], T, 0.015, 0.8660, want_info=False)
    for c in clouds:
        c.close()
    capi.lib().er_icp_release(device)
    return rows


def reference_sample(iterations: int, num: int = 4, points: int = 20000):
    """The verbatim reference binary on a bounded sample; returns (rows, seconds per iteration of 'Neat Optimization', threads)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    if not os.path.exists(pyoracle.REF_FRAGOPT):
        pyoracle.build()
    if not os.path.exists(pyoracle.REF_FRAGOPT):
        return None
    with tempfile.TemporaryDirectory() as d:
        data = os.path.join(d, "data") + os.sep
        ds = synth.make_fragopt_dataset(data, num=num, n_points=points)
        rows = sum(len(v) for v in ds["corres"].values())
        out = pyoracle.run_reference_fragopt(["--rgbdslam", data + "traj.log", "--registration", data + "reg_output.log", "--dir", data,
                                              "--num", str(num), "--interval", "1", "--blacklistpair", "0", "--slac",
                                              "--iteration", str(iterations)], cwd=d)
    ms = [float(l.split("took")[1].replace("ms.", "")) for l in out.splitlines() if l.startswith("Neat Optimization took")]
    if not ms:
        return None
    return rows, ms[0] * 1e-3 / iterations, min(8, os.cpu_count() or 1)


ALG_BYTES_PER_ROW = 16 + 5 * 16         # the sorted correspondence (int4) + five 16-byte gathers (cell/residuals, position, normal of pi; cell/residuals, position of pj)


def run(local: int, fragments: int = 50, points: int = 200_000, iterations: int = 3, warmup: int = 1, cpu: bool = True,
        clock_sampler=None, peak_gbs: float = 0.0):
    frags = [synth.make_local_fragment(i, points) for i in range(fragments)]
    pairs = [(i, j) for i in range(fragments) for j in range(i + 1, fragments)]
    corres = device_correspondences(frags, pairs, local)
    rows = int(sum(len(c) for c in corres))
    xyzn = [np.concatenate([f[0], f[1]], axis=1) for f in frags]
    ipose = np.array([synth.fragopt_initial_pose(i) for i in range(fragments)])

    # ---- e2e: host buffers -> GetCoordinate + upload, correspondence upload + sort, iterations
    t0 = time.perf_counter()
    g = capi.FragOpt(fragments, RES, LENGTH, device=local)
    for l in range(fragments):
        g.set_cloud(l, xyzn[l])
    g.set_pairs(pairs, corres)
    t_setup = time.perf_counter() - t0
    loop = SlacLoop(g, ipose)
    for _ in range(iterations):
        loop.step()
    e2e_s = time.perf_counter() - t0
    h2d = sum(x.nbytes for x in xyzn) + rows * 8
    first_scores = list(loop.scores)

    # ---- value: everything resident, `warmup` untimed iterations then `iterations` timed ones (the loop simply continues:
    # converged iterations cost the same as the first, every row is visited every time)
    for _ in range(warmup):
        loop.step()
    sampler = clock_sampler(local) if clock_sampler else None
    if sampler:
        sampler.start()
    kernel_ms, steps_ms = [], []
    l0 = g.launch_count()
    for _ in range(iterations):
        t1 = time.perf_counter()                    # every call of the loop returns synchronised
        loop.step()
        steps_ms.append((time.perf_counter() - t1) * 1e3)
        ms, r, flushes = g.assemble_stats()
        kernel_ms.append(ms)
    launches = g.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    k_ms = max(float(np.median(kernel_ms)), 1e-9)      # (the CPU emulation of the tests reports 0)
    value = rows * iterations / (sum(steps_ms) * 1e-3)
    tflops = 2.0 * FMA_PER_ROW * rows / (k_ms * 1e-3) / 1e12
    err = float(max(np.linalg.norm(loop.pose[l] - synth.fragment_frame(l)) for l in range(fragments)))
    err0 = float(max(np.linalg.norm(ipose[l] - synth.fragment_frame(l)) for l in range(fragments)))
    res = {
        "metric": METRIC, "value": value, "unit": "correspondences/s",
        "config": {"workload": WORKLOAD % (fragments, points // 1000, len(pairs), rows, iterations), "resolution": RES,
                   "matrix_size": g.M,
                   "l2": "inputs larger than L2: %.1f GB of sorted correspondences + %.2f GB of points per pass; the %.0f MB "
                         "matrix the flushes add to is meant to stay in L2" % (rows * 16 / 1e9, sum(len(x) for x in xyzn) * 64 / 1e9,
                                                                             g.M * g.M * 8 / 1e6)},
        "steps_ms": [round(x, 2) for x in steps_ms], "iterations": iterations, "warmup": warmup, "rows": rows,
        "gpu_launches": int(launches), "clocks": clocks,
        "kernel": {"name": "slac_assemble_kernel", "ms": k_ms, "all_ms": [round(x, 3) for x in kernel_ms], "rows_per_s": rows / (k_ms * 1e-3),
                   "flushes": int(flushes), "rows_per_flush": rows / max(1, flushes)},
        "roofline": {"bound": "fp64 fma", "kernel": "slac_assemble_kernel", "achieved": tflops, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                     "frac": tflops / FP64_PEAK_TFLOPS, "traffic": (_traffic_per_row() * rows) if _traffic_per_row() else None,
                     "traffic_note": "DRAM bytes of this launch, scaled per correspondence from the committed ncu capture (59 B per row)",
                     "algorithmic_flops_per_row": 2 * FMA_PER_ROW,
                     "hbm_view": {"algorithmic_bytes_per_row": ALG_BYTES_PER_ROW, "achieved": ALG_BYTES_PER_ROW * rows / (k_ms * 1e-3) / 1e9,
                                  "peak": peak_gbs or None, "unit": "GB/s",
                                  "frac": (ALG_BYTES_PER_ROW * rows / (k_ms * 1e-3) / 1e9 / peak_gbs) if peak_gbs else None},
                     "note": "1891 double multiply-adds per correspondence against ~100 B of gathers: the kernel's roof is FP64 issue, "
                             "not HBM; peak = 148 SMs x 64 FMA/clk x 1.965 GHz (MEASURED_PEAKS.json has no FP64 entry; ncu's FP64-pipe utilisation agrees with it)"},
        "e2e": {"value": rows * iterations / e2e_s, "unit": "correspondences/s", "h2d_bytes_per_step": int(h2d / iterations),
                "d2h_bytes_per_step": int(g.M * 8 * 2),
                "setup_s": t_setup, "note": "first %d iterations of the process incl. point upload + GetCoordinate, correspondence upload + "
                                            "device sort; wall clock" % iterations},
        "data_error_scores": [round(s[0], 4) for s in first_scores + loop.scores[len(first_scores):]][:8],
        "max_pose_error_vs_truth": err, "initial_pose_error": err0,
        "cpu_baseline": None,
    }
    g.close()
    if cpu:
        ref = reference_sample(iterations)
        if ref:
            r_rows, r_s, threads = ref
            res["cpu_baseline"] = {"value": r_rows / r_s, "unit": "correspondences/s", "cores": threads, "kind": "reference",
                                   "sample": "verbatim reference FragmentOptimizer --slac (CHOLMOD replaced by the vendored Eigen's "
                                             "SimplicialLLT) on 4 fragments x 20k points, 6 pairs, %d correspondences, %d iterations; "
                                             "its per-pair dense %dx%d matrices make its cost per correspondence grow with the fragment "
                                             "count" % (r_rows, iterations, 6 * 4 + 2187, 6 * 4 + 2187)}
    return res


def as_line(r: dict, args, world: int) -> dict:
    return {"metric": r["metric"], "value": r["value"], "unit": r["unit"], "n_gpus": world, "steps": r["iterations"], "warmup": r["warmup"],
            "ms_per_step": float(np.mean(r["steps_ms"])), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 points, f64 normal equations", "data": "synthetic", "config": r["config"], "roofline": r["roofline"],
            "cpu_baseline": r["cpu_baseline"], "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "clocks": r["clocks"],
            "kernel": r["kernel"], "max_pose_error_vs_truth": r["max_pose_error_vs_truth"], "initial_pose_error": r["initial_pose_error"],
            "data_error_scores": r["data_error_scores"]}


def run_reference(args):
    """bench.py --impl reference --metric fragopt: the verbatim reference binary on the bounded sample."""
    import json
    ref = reference_sample(max(1, args.steps))
    if not ref:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/FragmentOptimizer_ref is not built"}), flush=True)
        return 0
    rows, s, threads = ref
    v = rows / s
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "correspondences/s", "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": 0, "ms_per_step": s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32 points, f64 normal equations", "data": "synthetic",
                      "config": {"workload": "FragmentOptimizer --slac sample: 4 fragments x 20k points, 6 pairs, %d correspondences" % rows,
                                 "threads": threads, "host_cores": os.cpu_count()},
                      "cpu_baseline": {"value": v, "unit": "correspondences/s", "cores": threads, "kind": "reference",
                                       "sample": "%d iterations of the verbatim reference binary on the sample" % args.steps},
                      "e2e": {"value": v, "unit": "correspondences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}), flush=True)
    return 0


def base_aa_triplets(num):
    """COptApp::InitBaseAA (OptApp.cpp:765-811) as upper-triangle triplets: per fragment the lattice graph Laplacian, every edge
    visited from both ends; fragment 0's first node pinned."""
    nb = neighbour_table()
    nper = 3 * len(nb)
    u, t = np.nonzero(nb >= 0)
    v = nb[u, t]
    rows, cols, vals = [], [], []
    for l in range(num):
        for x in range(3):
            a, b = l * nper + 3 * u + x, l * nper + 3 * v + x
            rows += [a, b, np.minimum(a, b)]
            cols += [a, b, np.maximum(a, b)]
            vals += [np.ones(len(a)), np.ones(len(a)), -np.ones(len(a))]
    rows.append(np.arange(3)); cols.append(np.arange(3)); vals.append(np.ones(3))
    return np.concatenate(rows).astype(np.int32), np.concatenate(cols).astype(np.int32), np.concatenate(vals)


def nonrigid_probe(local: int, fragments: int = 8, points: int = 50_000, weight: float = 1.0):
    """One outer iteration of OptimizeNonrigid at a size whose dense system fits comfortably: assembly kernel, factorisation, one solve."""
    frags = [synth.make_local_fragment(i, points) for i in range(fragments)]
    pairs = [(i, j) for i in range(fragments) for j in range(i + 1, fragments)]
    corres = device_correspondences(frags, pairs, local)
    rows = int(sum(len(c) for c in corres))
    g = capi.FragOpt(fragments, RES, LENGTH, device=local)
    for l in range(fragments):
        g.set_cloud(l, np.concatenate([frags[l][0], frags[l][1]], axis=1))
    g.set_pairs(pairs, corres)
    ipose = np.array([synth.fragopt_initial_pose(i) for i in range(fragments)])
    ijk, unit = lattice_nodes()
    lat = ijk * unit
    ctr = np.concatenate([lat @ ipose[l][:3, :3].T + ipose[l][:3, 3] for l in range(fragments)]).reshape(-1)
    g.update_normals(ctr)
    r, c, v = base_aa_triplets(fragments)
    out = {"fragments": fragments, "rows": rows, "unknowns": fragments * g.nper, "dense_gb": (fragments * g.nper) ** 2 * 8 / 1e9}
    for rep in range(2):
        t0 = time.perf_counter()
        g.nonrigid_assemble(weight)
        t1 = time.perf_counter()
        ms, _, flushes = g.assemble_stats()
        g.nonrigid_factor(r, c, v)
        t2 = time.perf_counter()
        x = g.nonrigid_solve(np.cos(np.arange(fragments * g.nper) * 0.01))
        t3 = time.perf_counter()
        out["rep%d" % rep] = {"assemble_kernel_ms": ms, "assemble_call_ms": (t1 - t0) * 1e3, "factor_ms": (t2 - t1) * 1e3, "solve_ms": (t3 - t2) * 1e3,
                              "flushes": int(flushes), "rows_per_s": rows / (ms * 1e-3),
                              "fp64_tflops": 2.0 * (48 * 49 // 2) * rows / (ms * 1e-3) / 1e12, "x_finite": bool(np.isfinite(x).all())}
    g.close()
    return out


# ------------------------------------------------------------------------------------------------ several ranks (opt-in)
class ShardedSlacLoop(SlacLoop):
    """OptimizeSLAC with the pair list dealt over the ranks of a torch.distributed group: every rank holds all fragments and
    assembles the normal equations of its pairs; ONE all-reduce per iteration sums the matrices (49 MB for 50 fragments), Jb
    and the score; every rank then solves the same system and applies the same update.  The summed matrix reaches the
    solver as base + (sum - own): er_fragopt_slac_solve adds the rank's own assembled matrix on the device.

    `reduce(buf)` sums a float64 numpy vector over the ranks in place (NCCL through a device tensor in the bench, gloo in the
    CPU test).  Written after the round's GPU minutes were spent: exercised on the CPU (gloo, emulated library,
    tests/test_fragopt_cpu.py::test_sharded_loop_world2), not yet on several GPUs; bench.py runs it only with --fragopt-multi."""

    def __init__(self, g, ipose, reduce, weight=1.0):
        super().__init__(g, ipose, weight)
        self.reduce = reduce

    def step(self):
        rot_t = np.array([self.pose[l][:3, :3].T for l in range(self.num)])
        Jb, score = self.g.slac_assemble(rot_t)
        own = self.g.slac_jj()
        M = self.M
        buf = np.empty(M * M + M + 1)
        buf[:M * M] = own.ravel()
        buf[M * M:M * M + M] = Jb
        buf[-1] = score
        self.reduce(buf)
        total, Jb, score = buf[:M * M].reshape(M, M), buf[M * M:M * M + M].copy(), float(buf[-1])
        base_jb, reg = regulariser(self.ictr, self.ctr, self.num, self.w, self.nb)
        x = -self.g.slac_solve(self.base + (total - own), Jb + base_jb)
        self.ctr[6 * self.num:] += x[6 * self.num:]
        for l in range(self.num):
            T = np.eye(4)
            T[:3, :3] = rot_zyx(x[6 * l: 6 * l + 3])
            T[:3, 3] = x[6 * l + 3: 6 * l + 6]
            self.pose[l] = T @ self.pose[l]
        self.g.update_points_pn(self.expand())
        self.scores.append((score, reg))
        return score


def run_sharded(local: int, rank: int, world: int, fragments: int = 50, points: int = 200_000, iterations: int = 3, warmup: int = 1):
    """The leg on `world` ranks (bench.py --fragopt-multi): pairs round-robin, matrix all-reduce over NCCL."""
    import torch
    import torch.distributed as dist
    frags = [synth.make_local_fragment(i, points) for i in range(fragments)]
    pairs = [(i, j) for i in range(fragments) for j in range(i + 1, fragments)]
    mine = pairs[rank::world]
    corres = device_correspondences(frags, mine, local)
    rows = int(sum(len(c) for c in corres))
    g = capi.FragOpt(fragments, RES, LENGTH, device=local)
    for l in range(fragments):
        g.set_cloud(l, np.concatenate([frags[l][0], frags[l][1]], axis=1))
    g.set_pairs(mine, corres)
    dev = torch.device("cuda", local)

    def reduce(buf):
        t = torch.from_numpy(buf).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        buf[:] = t.cpu().numpy()

    loop = ShardedSlacLoop(g, np.array([synth.fragopt_initial_pose(i) for i in range(fragments)]), reduce)
    for _ in range(warmup):
        loop.step()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iterations):
        loop.step()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0, float(rows)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    total_rows = float(t[1].item())
    sec = float(tmax[0].item())
    ms, _, _ = g.assemble_stats()
    g.close()
    return {"metric": METRIC, "value": total_rows * iterations / sec, "unit": "correspondences/s", "ranks": world,
            "rows": int(total_rows), "rows_this_rank": rows, "iterations": iterations, "ms_per_iteration": sec * 1e3 / iterations,
            "assemble_kernel_ms_rank0": ms, "exchange_bytes_per_iteration": int((g.M * g.M + g.M + 1) * 8),
            "sharding": "pairs round-robin over %d ranks, every rank holds all fragments; one all-reduce of the matrix per iteration" % world,
            "data_error_scores": [round(s[0], 4) for s in loop.scores]}
