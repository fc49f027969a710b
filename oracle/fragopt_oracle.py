"""numpy restatement of the reference FragmentOptimizer's --slac and --rigid modes -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Follows, function by function (paths relative to the reference checkout):

  get_coordinate            PointCloud::GetCoordinate                    FragmentOptimizer/PointCloud.h:96-177
  update_normal / _point    PointCloud::UpdateNormal / UpdatePoint       PointCloud.h:57-68, 85-94
  update_pose               PointCloud::UpdatePose                       PointCloud.h:70-83
  slac_rows / slac_normal_equations   the pair loop of OptimizeSLAC      OptApp.cpp:473-570
  rigid_rows / rigid_normal_equations the pair loop of OptimizeRigid     OptApp.cpp:312-385
  base_jj, regulariser, get_rotation  InitBaseJJ, the regulariser loop, GetRotation   OptApp.cpp:813-858, 581-635, 860-882
  optimize_slac / optimize_rigid      OptimizeSLAC / OptimizeRigid       OptApp.cpp:418-676, 280-416

Pinned (tests/test_fragopt_cpu.py) against the VERBATIM reference binary oracle/_ref/FragmentOptimizer_ref (built by
`make -C oracle fragopt`) and against tests/golden/fragopt_golden.npz, which holds that binary's pose.log / output.ctr on
the committed synthetic data set.  float32 quantities are numpy float32 arrays, so every float product and sum rounds
like the C++ (no fused operations); double sums over correspondences are order-dependent at rounding level in the
reference itself (OpenMP critical section), hence tolerances instead of bit equality there.

The normal equations are formed as triu(S^T S) with S the rows scattered into dense vectors: for one row this is exactly
the reference's folding rule (equal indices inside a row add val_i^2, val_j^2 and 2 val_i val_j to the diagonal,
OptApp.cpp:548-557) -- written this way on purpose, so that it shares nothing with the tile logic of the CUDA kernel.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


class Lattice:
    def __init__(self, resolution: int, length: float):
        self.res = int(resolution)
        self.r1 = self.res + 1
        self.nper = self.r1 ** 3 * 3
        self.unit_f = f32(f32(length) / f32(self.res))        # PointCloud::unit_length_ (float)
        self.unit_d = float(length) / self.res                 # COptApp::unit_length_ (double)
        t = np.arange(8)
        self.corner_off = ((t >> 2) & 1) + ((t >> 1) & 1) * self.r1 + (t & 1) * self.r1 * self.r1

    def node(self, i, j, k):
        return i + j * self.r1 + k * self.r1 * self.r1


def _vals(r):
    """val_[8] (PointCloud.h:127-134); r = [n,3] float32."""
    x0, y0, z0 = f32(1) - r[:, 0], f32(1) - r[:, 1], f32(1) - r[:, 2]
    x1, y1, z1 = r[:, 0], r[:, 1], r[:, 2]
    return np.stack([x0 * y0 * z0, x0 * y0 * z1, x0 * y1 * z0, x0 * y1 * z1,
                     x1 * y0 * z0, x1 * y0 * z1, x1 * y1 * z0, x1 * y1 * z1], axis=1).astype(f32)


def _nvals(r, n0):
    """nval_[8] (PointCloud.h:139-170); n0 = the loaded normal divided by unit_length_."""
    x0, y0, z0 = f32(1) - r[:, 0], f32(1) - r[:, 1], f32(1) - r[:, 2]
    x1, y1, z1 = r[:, 0], r[:, 1], r[:, 2]
    a, b, c = n0[:, 0], n0[:, 1], n0[:, 2]
    return np.stack([
        -a * y0 * z0 - b * x0 * z0 - c * x0 * y0,
        -a * y0 * z1 - b * x0 * z1 + c * x0 * y0,
        -a * y1 * z0 + b * x0 * z0 - c * x0 * y1,
        -a * y1 * z1 + b * x0 * z1 + c * x0 * y1,
        a * y0 * z0 - b * x1 * z0 - c * x1 * y0,
        a * y0 * z1 - b * x1 * z1 + c * x1 * y0,
        a * y1 * z0 + b * x1 * z0 - c * x1 * y1,
        a * y1 * z1 + b * x1 * z1 + c * x1 * y1], axis=1).astype(f32)


class Cloud:
    """One PointCloud: what GetCoordinate derives plus the mutable p_ / n_."""

    def __init__(self, L: Lattice, xyzn):
        xyzn = np.ascontiguousarray(xyzn, dtype=f32).reshape(-1, 6)
        q = xyzn[:, :3] / L.unit_f
        corner = np.floor(q).astype(np.int64)
        bad = np.any((corner < 0) | (corner >= L.res), axis=1)
        if bad.any():                                          # "Point out of bound": loading stops there
            stop = int(np.argmax(bad))
            xyzn, q, corner = xyzn[:stop], q[:stop], corner[:stop]
        self.L = L
        self.r = (q - corner.astype(f32)).astype(f32)
        self.base = corner[:, 0] + corner[:, 1] * L.r1 + corner[:, 2] * L.r1 * L.r1
        self.idx = (self.base[:, None] + L.corner_off[None, :]) * 3            # idx_[8]
        self.n0 = (xyzn[:, 3:6] / L.unit_f).astype(f32)
        self.val = _vals(self.r)
        self.nval = _nvals(self.r, self.n0)
        self.p = xyzn[:, :3].copy()
        self.n = self.n0.copy()                                # PointCloud.h:172-174: the DIVIDED normal

    def __len__(self):
        return len(self.p)

    def update_pose(self, T):
        m = np.asarray(T, dtype=np.float64).astype(f32)       # pose.cast<float>()
        p, n = self.p, self.n
        self.p = np.stack([((m[r, 0] * p[:, 0] + m[r, 1] * p[:, 1]) + m[r, 2] * p[:, 2]) + m[r, 3] * f32(1) for r in range(3)], axis=1)
        self.n = np.stack([((m[r, 0] * n[:, 0] + m[r, 1] * n[:, 1]) + m[r, 2] * n[:, 2]) + m[r, 3] * f32(0) for r in range(3)], axis=1)

    def update_normal(self, ctr):
        """ctr = this fragment's lattice block (nper doubles)."""
        c = np.asarray(ctr, dtype=np.float64).astype(f32)
        n = np.zeros((len(self), 3), dtype=f32)
        for i in range(3):
            acc = np.zeros(len(self), dtype=f32)
            for j in range(8):
                acc = acc + self.nval[:, j] * c[self.idx[:, j] + i]
            n[:, i] = acc
        ln = np.sqrt(n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1] + n[:, 2] * n[:, 2])
        with np.errstate(all="ignore"):
            self.n = (n / ln[:, None]).astype(f32)

    def warped(self, ctr):
        """UpdatePoint: float products, double sums."""
        c = np.asarray(ctr, dtype=np.float64).astype(f32)
        pos = np.zeros((len(self), 3), dtype=np.float64)
        for i in range(3):
            for j in range(8):
                pos[:, i] += (self.val[:, j] * c[self.idx[:, j] + i]).astype(np.float64)
        return pos

    def update_point_pn(self, ctr):
        self.update_normal(ctr)
        self.p = self.warped(ctr).astype(f32)


def rot_zyx(r):
    """AngleAxis(r[2], Z) * AngleAxis(r[1], Y) * AngleAxis(r[0], X) (OptApp.cpp:396-399)."""
    cx, sx, cy, sy, cz, sz = np.cos(r[0]), np.sin(r[0]), np.cos(r[1]), np.sin(r[1]), np.cos(r[2]), np.sin(r[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def increment(r6):
    T = np.eye(4)
    T[:3, :3] = rot_zyx(r6[:3])
    T[:3, 3] = r6[3:6]
    return T


def initial_poses(traj, rel2abs, interval, length):
    """COptApp::InitIPose (OptApp.cpp:49-72)."""
    base = np.eye(4)
    base[0, 3] = length / 2.0
    base[1, 3] = length / 2.0
    base[2, 3] = -0.3
    left = base @ np.linalg.inv(traj[0])
    return np.array([left @ traj[ii * interval] @ np.linalg.inv(base) for ii in rel2abs])


# ---------------------------------------------------------------------------------------------------------------- SLAC
def slac_rows(ci: Cloud, cj: Cloud, rows, i, j, num, rot_t):
    """Jacobian rows of the correspondences `rows` ([n,2]: index into ci, index into cj).  Returns idx [n,60], val [n,60], b [n]."""
    ii, jj = rows[:, 0], rows[:, 1]
    ppi, ppj, npi = ci.p[ii].astype(np.float64), cj.p[jj].astype(np.float64), ci.n[ii].astype(np.float64)
    diff = ppi - ppj
    b = (diff[:, 0] * npi[:, 0] + diff[:, 1] * npi[:, 1]) + diff[:, 2] * npi[:, 2]
    temp = np.cross(ppj, npi)
    n = len(rows)
    idx = np.empty((n, 60), dtype=np.int64)
    val = np.empty((n, 60), dtype=np.float64)
    idx[:, 0:6] = i * 6 + np.arange(6)
    idx[:, 6:12] = j * 6 + np.arange(6)
    val[:, 0:3], val[:, 3:6] = temp, npi
    val[:, 6:9], val[:, 9:12] = -temp, -npi
    dTi = npi @ rot_t[i].T                                     # pose_rot_t_[i] * npi
    dTj = -(npi @ rot_t[j].T)
    for ll in range(8):
        for xyz in range(3):
            idx[:, 12 + ll * 3 + xyz] = 6 * num + ci.idx[ii, ll] + xyz
            val[:, 12 + ll * 3 + xyz] = ci.val[ii, ll].astype(np.float64) * dTi[:, xyz]
            idx[:, 36 + ll * 3 + xyz] = 6 * num + cj.idx[jj, ll] + xyz
            val[:, 36 + ll * 3 + xyz] = cj.val[jj, ll].astype(np.float64) * dTj[:, xyz]
    return idx, val, b


def _accumulate(JJ, Jb, idx, val, b, M, chunk=1024):
    for s in range(0, len(b), chunk):
        e = min(len(b), s + chunk)
        S = np.zeros((e - s, M))
        np.add.at(S, (np.arange(e - s)[:, None], idx[s:e]), val[s:e])
        JJ += S.T @ S
        Jb += S.T @ b[s:e]


def slac_normal_equations(clouds, pairs, corres, num, rot_t, M):
    """Data term of one SLAC iteration: upper triangle of JJ (dense M x M), Jb, score."""
    JJ = np.zeros((M, M))
    Jb = np.zeros(M)
    score = 0.0
    for (i, j), rows in zip(pairs, corres):
        if len(rows) == 0:
            continue
        idx, val, b = slac_rows(clouds[i], clouds[j], rows, i, j, num, rot_t)
        _accumulate(JJ, Jb, idx, val, b, M)
        score += float(np.sum(b * b))
    return np.triu(JJ), Jb, score


def lattice_neighbours(L: Lattice, i, j, k):
    out = []
    if i > 0: out.append(L.node(i - 1, j, k))
    if i < L.res: out.append(L.node(i + 1, j, k))
    if j > 0: out.append(L.node(i, j - 1, k))
    if j < L.res: out.append(L.node(i, j + 1, k))
    if k > 0: out.append(L.node(i, j, k - 1))
    if k < L.res: out.append(L.node(i, j, k + 1))
    return out


def base_jj(L: Lattice, num, M, default_weight):
    """InitBaseJJ * default_weight, upper triangle, plus the six ones of OptApp.cpp:459-464."""
    B = np.zeros((M, M))
    off = 6 * num
    for i in range(L.r1):
        for j in range(L.r1):
            for k in range(L.r1):
                u = L.node(i, j, k) * 3 + off
                for v in lattice_neighbours(L, i, j, k):
                    v = v * 3 + off
                    for x in range(3):
                        B[u + x, u + x] += 1
                        B[v + x, v + x] += 1
                        B[u + x, v + x] += -1
                        B[v + x, u + x] += -1
    a = L.node(L.res // 2, L.res // 2, 0) * 3 + off
    for x in range(3):
        B[a + x, a + x] += 1
    B = np.triu(B) * default_weight
    for k in range(6):
        B[k, k] += 1.0
    return B


def get_rotation(dif, diff):
    """COptApp::GetRotation: dif / diff = [n,3] rest / current edge vectors."""
    Cm = dif.T @ diff
    U, _, Vt = np.linalg.svd(Cm)
    V = Vt.T
    R = V @ U.T
    if np.linalg.det(R) < 0:
        U = U.copy()
        U[:, 2] *= -1
        R = V @ U.T
    return R


def slac_regulariser(L: Lattice, num, M, ictr, ctr, default_weight):
    off = 6 * num
    Jb = np.zeros(M)
    score = 0.0
    for i in range(L.r1):
        for j in range(L.r1):
            for k in range(L.r1):
                u = L.node(i, j, k) * 3 + off
                nb = [v * 3 + off for v in lattice_neighbours(L, i, j, k)]
                d0 = np.array([ictr[u:u + 3] - ictr[v:v + 3] for v in nb])
                d1 = np.array([ctr[u:u + 3] - ctr[v:v + 3] for v in nb])
                R = np.eye(3) if (i == L.res // 2 and j == L.res // 2 and k == 0) else get_rotation(d0, d1)
                for t, v in enumerate(nb):
                    bx = d1[t] - R @ d0[t]
                    score += default_weight * float(bx @ bx)
                    Jb[u:u + 3] += bx * default_weight
                    Jb[v:v + 3] -= bx * default_weight
    return Jb, score


def optimize_slac(xyzn, pairs, corres, ipose, resolution=8, length=3.0, weight=1.0, iterations=5, num_before_blacklist=None,
                  trace=None):
    """OptimizeSLAC on in-memory inputs: xyzn = list of [n,6] float32 clouds, pairs = [(i, j)], corres = list of [m,2] int
    arrays, ipose = [num,4,4].  Returns (poses [num,4,4], expanded ctr [num*nper/3, 3])."""
    num = len(xyzn)
    L = Lattice(resolution, length)
    M = 6 * num + L.nper
    default_weight = (num_before_blacklist or num) * weight
    clouds = [Cloud(L, c) for c in xyzn]
    ictr = np.zeros(M)
    for i in range(L.r1):
        for j in range(L.r1):
            for k in range(L.r1):
                ictr[6 * num + L.node(i, j, k) * 3: 6 * num + L.node(i, j, k) * 3 + 3] = (i * L.unit_d, j * L.unit_d, k * L.unit_d)
    this_ctr = ictr.copy()
    pose = np.array(ipose, dtype=np.float64).copy()
    for l in range(num):
        clouds[l].update_pose(pose[l])
    B = base_jj(L, num, M, default_weight)

    def expand(ctr):
        lat = ctr[6 * num:].reshape(-1, 3)
        return np.concatenate([lat @ pose[l][:3, :3].T + pose[l][:3, 3] for l in range(num)]).reshape(-1)

    for itr in range(iterations):
        rot_t = np.array([pose[l][:3, :3].T for l in range(num)])
        JJ, Jb, score = slac_normal_equations(clouds, pairs, corres, num, rot_t, M)
        base_jb, regscore = slac_regulariser(L, num, M, ictr, this_ctr, default_weight)
        A = B + JJ
        A = A + np.triu(A, 1).T
        result = -np.linalg.solve(A, Jb + base_jb)
        if trace is not None:
            trace.append({"score": score, "regscore": regscore, "JJ": JJ, "Jb": Jb, "rot_t": rot_t,
                          "points": [(c.p.copy(), c.n.copy()) for c in clouds]})
        this_ctr[6 * num:] += result[6 * num:]
        for l in range(num):
            pose[l] = increment(result[6 * l: 6 * l + 6]) @ pose[l]
        ex = expand(this_ctr)
        for l in range(num):
            clouds[l].update_point_pn(ex[l * L.nper:(l + 1) * L.nper])
    return pose, expand(this_ctr).reshape(-1, 3)


# --------------------------------------------------------------------------------------------------------------- rigid
def rigid_rows(ci: Cloud, cj: Cloud, rows):
    ii, jj = rows[:, 0], rows[:, 1]
    n = len(rows)
    ppi = np.concatenate([ci.p[ii].astype(np.float64), np.ones((n, 1))], axis=1)
    ppj = np.concatenate([cj.p[jj].astype(np.float64), np.ones((n, 1))], axis=1)
    npi = np.concatenate([ci.n[ii].astype(np.float64), np.zeros((n, 1))], axis=1)
    d = ppi - ppj
    z, o = np.zeros(n), np.ones(n)

    def dot(a, b):
        return np.einsum("ij,ij->i", np.stack(a, axis=1) if isinstance(a, (list, tuple)) else a, b)

    b = dot(d, npi)
    val = np.empty((n, 12))
    val[:, 0] = dot([z, -ppi[:, 2], ppi[:, 1], o], npi) + dot([z, -npi[:, 2], npi[:, 1], z], d)
    val[:, 1] = dot([ppi[:, 2], z, -ppi[:, 0], o], npi) + dot([npi[:, 2], z, -npi[:, 0], z], d)
    val[:, 2] = dot([-ppi[:, 1], ppi[:, 0], z, o], npi) + dot([-npi[:, 1], npi[:, 0], z, z], d)
    val[:, 3:6] = npi[:, :3]
    val[:, 6] = -dot([z, -ppj[:, 2], ppj[:, 1], o], npi)
    val[:, 7] = -dot([ppj[:, 2], z, -ppj[:, 0], o], npi)
    val[:, 8] = -dot([-ppj[:, 1], ppj[:, 0], z, o], npi)
    val[:, 9:12] = -npi[:, :3]
    return val, b


def rigid_normal_equations(clouds, pairs, corres, num):
    """Data term of one rigid iteration WITHOUT the ones every pair adds to the first six diagonal entries."""
    n6 = 6 * num
    JJ = np.zeros((n6, n6))
    Jb = np.zeros(n6)
    score = 0.0
    for (i, j), rows in zip(pairs, corres):
        if len(rows) == 0:
            continue
        val, b = rigid_rows(clouds[i], clouds[j], rows)
        idx = np.concatenate([i * 6 + np.arange(6), j * 6 + np.arange(6)])
        H = val.T @ val
        JJ[np.ix_(idx, idx)] += H
        Jb[idx] += val.T @ b
        score += float(np.sum(b * b))
    return np.triu(JJ), Jb, score


def optimize_rigid(xyzn, pairs, corres, ipose, resolution=8, length=3.0, iterations=5, trace=None):
    num = len(xyzn)
    L = Lattice(resolution, length)
    clouds = [Cloud(L, c) for c in xyzn]
    pose = np.array(ipose, dtype=np.float64).copy()
    for l in range(num):
        clouds[l].update_pose(pose[l])
    for itr in range(iterations):
        JJ, Jb, score = rigid_normal_equations(clouds, pairs, corres, num)
        for k in range(6):
            JJ[k, k] += len(pairs)                            # mat_adder.Add(k, k, 1) once per pair (OptApp.cpp:323-325)
        if trace is not None:
            trace.append({"score": score, "JJ": JJ.copy(), "Jb": Jb.copy()})
        A = JJ + np.triu(JJ, 1).T
        result = -np.linalg.solve(A, Jb)
        for l in range(num):
            T = increment(result[6 * l: 6 * l + 6])
            pose[l] = T @ pose[l]
            clouds[l].update_pose(T)
    lat = np.array([[i * L.unit_d, j * L.unit_d, k * L.unit_d] for k in range(L.r1) for j in range(L.r1) for i in range(L.r1)])
    ctr = np.concatenate([lat @ pose[l][:3, :3].T + pose[l][:3, 3] for l in range(num)])
    return pose, ctr


# ------------------------------------------------------------------------------------------------------------ nonrigid
def nonrigid_rows(ci: Cloud, cj: Cloud, rows, weight):
    """OptApp.cpp:177-190: idx1/val1 on fragment i's lattice block, idx2/val2 on fragment j's (24 entries each)."""
    ii, jj = rows[:, 0], rows[:, 1]
    n = len(rows)
    idx1 = np.empty((n, 24), dtype=np.int64)
    idx2 = np.empty((n, 24), dtype=np.int64)
    v1 = np.empty((n, 24))
    v2 = np.empty((n, 24))
    ni = ci.n[ii].astype(np.float64)
    for t in range(3):
        for ll in range(8):
            idx1[:, t * 8 + ll] = ci.idx[ii, ll] + t
            idx2[:, t * 8 + ll] = cj.idx[jj, ll] + t
            v1[:, t * 8 + ll] = ci.val[ii, ll].astype(np.float64) * weight * ni[:, t]
            v2[:, t * 8 + ll] = -cj.val[jj, ll].astype(np.float64) * weight * ni[:, t]
    return idx1, v1, idx2, v2


def nonrigid_data_matrix(clouds, pairs, corres, nper, weight, chunk=2048):
    """Sum of mati + matj + matij over all pairs (OptApp.cpp:159-206), as the solver sees it: the upper triangle.  matij
    (rows of fragment i, columns of fragment j) lies in the upper triangle only when i < j."""
    num = len(clouds)
    N = num * nper
    A = np.zeros((N, N))
    for (i, j), rows in zip(pairs, corres):
        if len(rows) == 0:
            continue
        idx1, v1, idx2, v2 = nonrigid_rows(clouds[i], clouds[j], rows, weight)
        for s in range(0, len(rows), chunk):
            e = min(len(rows), s + chunk)
            S1 = np.zeros((e - s, nper))
            S2 = np.zeros((e - s, nper))
            np.add.at(S1, (np.arange(e - s)[:, None], idx1[s:e]), v1[s:e])
            np.add.at(S2, (np.arange(e - s)[:, None], idx2[s:e]), v2[s:e])
            A[i * nper:(i + 1) * nper, i * nper:(i + 1) * nper] += S1.T @ S1
            A[j * nper:(j + 1) * nper, j * nper:(j + 1) * nper] += S2.T @ S2
            A[i * nper:(i + 1) * nper, j * nper:(j + 1) * nper] += S1.T @ S2
    return np.triu(A)


def base_aa(L: Lattice, num):
    """COptApp::InitBaseAA (OptApp.cpp:765-811), upper triangle."""
    N = num * L.nper
    B = np.zeros((N, N))
    for l in range(num):
        off = l * L.nper
        for i in range(L.r1):
            for j in range(L.r1):
                for k in range(L.r1):
                    u = L.node(i, j, k) * 3 + off
                    for v in lattice_neighbours(L, i, j, k):
                        v = v * 3 + off
                        for x in range(3):
                            B[u + x, u + x] += 1
                            B[v + x, v + x] += 1
                            B[min(u, v) + x, max(u, v) + x] += -1
        if l == 0:
            for x in range(3):
                B[x, x] += 1
    return B


def nonrigid_rhs(L: Lattice, num, ictr, ctr):
    """The inner iteration's Ab (OptApp.cpp:221-259)."""
    Ab = np.zeros(num * L.nper)
    for l in range(num):
        off = l * L.nper
        for i in range(L.r1):
            for j in range(L.r1):
                for k in range(L.r1):
                    u = L.node(i, j, k) * 3 + off
                    nb = [v * 3 + off for v in lattice_neighbours(L, i, j, k)]
                    d0 = np.array([ictr[u:u + 3] - ictr[v:v + 3] for v in nb])
                    d1 = np.array([ctr[u:u + 3] - ctr[v:v + 3] for v in nb])
                    R = get_rotation(d0, d1)
                    for t, v in enumerate(nb):
                        bx = R @ d0[t]
                        Ab[u:u + 3] += bx
                        Ab[v:v + 3] -= bx
    return Ab


def optimize_nonrigid(xyzn, pairs, corres, ipose, resolution=8, length=3.0, weight=1.0, iterations=5, inner_iterations=10, trace=None):
    """OptimizeNonrigid (OptApp.cpp:120-278).  Returns the lattices [num*nper/3, 3]."""
    from scipy.linalg import cho_factor, cho_solve
    num = len(xyzn)
    L = Lattice(resolution, length)
    clouds = [Cloud(L, c) for c in xyzn]
    lat = np.array([[i * L.unit_d, j * L.unit_d, k * L.unit_d] for k in range(L.r1) for j in range(L.r1) for i in range(L.r1)])
    ctr = np.concatenate([lat @ np.asarray(ipose[l])[:3, :3].T + np.asarray(ipose[l])[:3, 3] for l in range(num)]).reshape(-1)
    ictr = ctr.copy()
    B = base_aa(L, num)
    for itr in range(iterations):
        for l in range(num):
            clouds[l].update_normal(ctr[l * L.nper:(l + 1) * L.nper])
        A = nonrigid_data_matrix(clouds, pairs, corres, L.nper, weight)
        if trace is not None:
            trace.append({"A": A, "normals": [c.n.copy() for c in clouds]})
        F = cho_factor((A + B).T, lower=True)                 # the stored upper triangle, read as the lower one of the transpose
        for m in range(inner_iterations):
            ctr = cho_solve(F, nonrigid_rhs(L, num, ictr, ctr))
    return ctr.reshape(-1, 3)
