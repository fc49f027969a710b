"""CPU check of the ICP path's nearest-neighbour search (no GPU): the product's own source
elasticreconstruction_b200/csrc/er_nn_math.h (grid geometry, cell assignment, candidate comparison, pruned walk),
compiled by g++ (tests/emul/nn_emul.cpp), against a brute-force search in the same float arithmetic
(((dx*dx + dy*dy) + dz*dz), ties to the smallest index -- flann::L2_Simple as restated in oracle/icp_oracle.c).

Both searches are exact by construction only inside the radius the grid was built for, which is all the gates use
(count_inliers, determineCorrespondences, FindCorrespondence reject anything farther): queries whose true neighbour is
farther than the radius must come back as "farther than the radius or nothing"."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emul", "nn_emul.cpp")
HDRS = [os.path.join(ROOT, "elasticreconstruction_b200", "csrc", h) for h in ("er_nn_math.h", "er_emul.h")]
OUT = os.path.join(ROOT, "tests", "emul", "_build", "libnn_emul.so")
f32 = np.float32


@pytest.fixture(scope="module")
def nn():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(p) for p in [SRC] + HDRS):
        subprocess.check_call(["g++", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-fno-fast-math",
                               "-I", os.path.dirname(HDRS[0]), SRC, "-o", OUT])
    L = C.CDLL(OUT)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.emul_grid_build.restype = C.c_void_p
    L.emul_grid_build.argtypes = [fp, C.c_int, C.c_double, C.c_int]
    L.emul_grid_free.argtypes = [C.c_void_p]
    L.emul_grid_radius.restype = C.c_double
    L.emul_grid_radius.argtypes = [C.c_void_p]
    L.emul_nn_search.argtypes = [C.c_void_p, fp, C.c_int, ip, fp]
    L.emul_nn_coop.argtypes = [C.c_void_p, fp, C.c_int, C.c_float, ip, fp]
    return L


def brute(points, queries):
    idx = np.empty(len(queries), dtype=np.int64)
    d2 = np.empty(len(queries), dtype=np.float32)
    for i, q in enumerate(queries):
        with np.errstate(all="ignore"):
            dx, dy, dz = q[0] - points[:, 0], q[1] - points[:, 1], q[2] - points[:, 2]
            d = dx * dx + dy * dy + dz * dz                      # float32, ((xx + yy) + zz)
        d = np.where(np.isnan(d), np.float32(np.inf), d)
        k = int(np.argmin(d))                                    # first minimum = smallest index on ties
        idx[i], d2[i] = k, d[k]
    return idx, d2


def run_both(nn, points, queries, radius, rings=2):
    points = np.ascontiguousarray(points, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    h = nn.emul_grid_build(points.ctypes.data_as(fp), len(points), radius, rings)
    try:
        assert nn.emul_grid_radius(h) >= radius
        m = len(queries)
        i1, d1 = np.empty(m, dtype=np.int32), np.empty(m, dtype=np.float32)
        i2, d2 = np.empty(m, dtype=np.int32), np.empty(m, dtype=np.float32)
        nn.emul_nn_search(h, queries.ctypes.data_as(fp), m, i1.ctypes.data_as(ip), d1.ctypes.data_as(fp))
        max_d2 = radius * radius
        bd_init = np.float32(max_d2)
        if float(bd_init) < max_d2:
            bd_init = np.nextafter(bd_init, np.float32(np.inf))
        nn.emul_nn_coop(h, queries.ctypes.data_as(fp), m, C.c_float(bd_init), i2.ctypes.data_as(ip), d2.ctypes.data_as(fp))
    finally:
        nn.emul_grid_free(h)
    bi, bd = brute(points, queries)
    inside = bd.astype(np.float64) <= max_d2                      # the gates accept exactly these
    for name, (gi, gd) in (("nn_search", (i1, d1)), ("cooperative", (i2, d2))):
        assert np.array_equal(gi[inside], bi[inside]), name
        assert np.array_equal(gd[inside].view(np.uint32), bd[inside].view(np.uint32)), name
        out = ~inside
        # outside the radius: nothing, or something that the gate rejects as well
        assert np.all((gi[out] < 0) | (gd[out].astype(np.float64) > max_d2)), name
    return int(inside.sum()), int((~inside).sum())


def test_random_cloud_random_queries(nn):
    rng = np.random.default_rng(1)
    pts = rng.uniform(0, 1, size=(20000, 3)).astype(f32)
    q = np.concatenate([pts[rng.integers(0, len(pts), 1500)] + rng.normal(0, 0.01, (1500, 3)).astype(f32),
                        rng.uniform(-0.2, 1.2, size=(1500, 3)).astype(f32)])
    n_in, n_out = run_both(nn, pts, q, 0.03)
    assert n_in > 1000 and n_out > 100


def test_lattice_with_exact_ties_and_duplicates(nn):
    """Points on a regular lattice (many equidistant candidates), every point stored twice, queries at lattice
    points, at cell centres of the lattice and exactly half-way between neighbours."""
    g = np.arange(0, 24, dtype=f32) * f32(0.0078125)             # 1/128: exact in float
    pts = np.stack(np.meshgrid(g, g, g[:6], indexing="ij"), -1).reshape(-1, 3)
    pts = np.concatenate([pts, pts])                             # duplicates: ties between different indices
    rng = np.random.default_rng(2)
    sel = pts[rng.integers(0, len(pts), 800)]
    half = f32(0.00390625)
    q = np.concatenate([sel, sel + np.array([half, 0, 0], dtype=f32), sel + np.array([half, half, 0], dtype=f32),
                        sel + np.array([half, half, half], dtype=f32)])
    n_in, _ = run_both(nn, pts, q, 0.03)
    assert n_in == len(q)


def test_queries_on_cell_boundaries_and_outside_the_grid(nn):
    rng = np.random.default_rng(3)
    pts = rng.uniform(0, 0.6, size=(8000, 3)).astype(f32)
    radius = 0.03
    cell = radius * 1.001 / 2
    lo = pts.min(axis=0)
    k = rng.integers(-3, 45, size=(2000, 3))
    on_edges = (lo.astype(np.float64) + k * cell).astype(f32)    # on (or one rounding away from) cell faces
    far = rng.uniform(-5, 5, size=(300, 3)).astype(f32)
    weird = np.array([[np.nan, 0.1, 0.1], [np.inf, 0.1, 0.1], [0.1, -np.inf, 0.1], [1e30, 1e30, 1e30], [0.3, 0.3, np.nan]], dtype=f32)
    n_in, n_out = run_both(nn, pts, np.concatenate([on_edges, far, weird]), radius)
    assert n_in > 500 and n_out > 300


@pytest.mark.parametrize("rings", [1, 2])
def test_surface_like_cloud_at_fragment_scale(nn, rings):
    """A jittered planar patch sampled like the benchmark fragments (3/512 m lattice), queries displaced by up to the
    full radius along and across the surface.  (rings = 1: only the per-thread walk supports it, the cooperative
    emulation handles it too.)"""
    rng = np.random.default_rng(4)
    u, v = np.meshgrid(np.arange(150), np.arange(150), indexing="ij")
    pts = np.stack([u.ravel() * (3.0 / 512), v.ravel() * (3.0 / 512), 0.5 + 0.02 * np.sin(u.ravel() / 9.0)], -1)
    pts = (pts + rng.uniform(-0.001, 0.001, pts.shape)).astype(f32)
    q = pts[rng.integers(0, len(pts), 4000)] + rng.uniform(-0.03, 0.03, (4000, 3)).astype(f32)
    n_in, n_out = run_both(nn, pts, q, 0.03, rings)
    assert n_in > 2000


def test_empty_and_single_point_clouds(nn):
    q = np.array([[0, 0, 0], [0.01, 0, 0], [1, 1, 1]], dtype=f32)
    run_both(nn, np.array([[0.0, 0.0, 0.0]], dtype=f32), q, 0.03)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    h = nn.emul_grid_build(np.zeros((1, 3), dtype=f32).ctypes.data_as(fp), 0, 0.03, 2)
    i1, d1 = np.empty(3, dtype=np.int32), np.empty(3, dtype=np.float32)
    nn.emul_nn_search(h, q.ctypes.data_as(fp), 3, i1.ctypes.data_as(ip), d1.ctypes.data_as(fp))
    nn.emul_grid_free(h)
    assert np.all(i1 == -1)
