"""CPU check of the integrate kernel's per-voxel arithmetic (no GPU): the product's own source
elasticreconstruction_b200/csrc/er_voxel_math.h is compiled by g++ against its emulation layer
(tests/emul/voxel_emul.cpp) and must reproduce the oracle's IntegrateVolumeUnit bit for bit -- on every block
the reference touches, frame by frame -- whatever the last bits of the approximate reciprocal are.

This pins the exact path's float/double evaluation order, the shared-reciprocal division, the rounding of the pixel
coordinates and the truncation thresholds on the CPU; the GPU parity tests then only have to show that the device
executes the same operations (tests/test_tsdf_gpu.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers
from elasticreconstruction_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emul", "voxel_emul.cpp")
HDR = os.path.join(ROOT, "elasticreconstruction_b200", "csrc", "er_voxel_math.h")
OUT = os.path.join(ROOT, "tests", "emul", "_build", "libvoxel_emul.so")


@pytest.fixture(scope="module")
def emul():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-fno-fast-math",
                               "-I", os.path.dirname(HDR), SRC, "-o", OUT])
    L = C.CDLL(OUT)
    fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
    L.emul_integrate_block.restype = C.c_longlong
    L.emul_integrate_block.argtypes = [fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_double, C.c_double,
                                       dp, dp, C.c_int, fp, fp]
    L.emul_set_rcp_bias.argtypes = [C.c_int]
    L.emul_thresholds.argtypes = [C.c_double, fp]
    L.emul_set_pixel_form.argtypes = [C.c_int]
    L.emul_pixel.restype = C.c_int
    L.emul_pixel.argtypes = [C.c_int] + [C.c_float] * 4 + [C.c_int, C.c_int] + [C.c_float] * 3
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def run_scene(oracle, emul, sc, voxel=None):
    cam = sc.get("camera", dict(fx=525.0, fy=525.0, cx=319.5, cy=239.5, int_trunc=2.5))
    ovol = oracle.OracleTSDF()
    ovol.set_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["int_trunc"])
    ul, trunc = (3.0 / 512.0, 0.03) if voxel is None else voxel
    if voxel is not None:
        ovol.set_voxel(ul, trunc)
    blocks = {}
    updated_o = updated_e = 0
    traj = sc["traj"]
    for f in range(len(sc["frames"])):
        if f + 1 >= len(traj):
            break
        depth = sc["frames"][f]
        T = np.ascontiguousarray(traj[f], dtype=np.float64)
        Tinv = np.ascontiguousarray(capi.mat4_inverse(T), dtype=np.float64)      # the product's host inverse
        scaled = np.ascontiguousarray(ovol.scale_depth(depth), dtype=np.float32)
        ovol.integrate(depth, T)
        updated_o += ovol.last_updated_voxels()
        for key in ovol.last_touched():
            if int(key) not in blocks:
                blocks[int(key)] = (np.zeros(64 ** 3, dtype=np.float32), np.zeros(64 ** 3, dtype=np.float32))
            s, w = blocks[int(key)]
            updated_e += emul.emul_integrate_block(_fp(scaled), 640, 480, cam["fx"], cam["fy"], cam["cx"], cam["cy"], ul, trunc,
                                                   _dp(T), _dp(Tinv), int(key), _fp(s), _fp(w))
    assert sorted(blocks) == sorted(int(k) for k in ovol.block_keys())
    assert updated_e == updated_o and updated_o > 100000
    for key, (s, w) in blocks.items():
        so, wo = ovol.read_block(key)
        assert np.array_equal(w, wo.ravel())
        assert np.array_equal(s.view(np.uint32), so.ravel().view(np.uint32)), "block %d" % key
    return updated_o


@pytest.mark.parametrize("bias", [0, -1, 1, -2, 2])
def test_exact_path_arithmetic_matches_oracle_room(oracle, emul, bias):
    """The loop-trajectory room (oblique views, every image region, voxels behind / in / in front of the band)."""
    emul.emul_set_rcp_bias(bias)
    try:
        run_scene(oracle, emul, helpers.scene_room(n=3))
    finally:
        emul.emul_set_rcp_bias(0)


def test_exact_path_arithmetic_matches_oracle_other_scenes(oracle, emul):
    """Wavy wall seen from the origin (identity poses), non-default intrinsics with a short integration_trunc, and
    12 mm voxels with a 5 cm truncation (BASELINE configs[0] geometry)."""
    run_scene(oracle, emul, helpers.scene_wall(n=2))
    run_scene(oracle, emul, helpers.scene_camera(n=2))
    run_scene(oracle, emul, helpers.scene_wall(n=2), voxel=(0.012, 0.05))


def test_threshold_helper_reproduces_the_double_comparisons(emul):
    """thr_neg / thr_one are the smallest floats whose double value passes `>= -trunc` / `>= trunc`
    (TSDFVolume.cpp:87-88 compare and divide in double)."""
    out = np.zeros(3, dtype=np.float32)
    for trunc in (0.03, 0.05, 0.036, 0.1, 1.0 / 3.0, 0.25):
        emul.emul_thresholds(C.c_double(trunc), _fp(out))
        lo, hi, eps = out
        assert float(lo) >= -trunc and float(np.nextafter(lo, np.float32(-np.inf))) < -trunc
        assert float(hi) >= trunc and float(np.nextafter(hi, np.float32(-np.inf))) < trunc
        assert abs(float(eps) - (trunc + 1e-4)) < 1e-7


# ------------------------------------------------------------------------------------------------ second pixel form
def ref_pixel(fx, fy, cx, cy, cols, rows, t0, t1, t2):
    """TSDFVolume.cpp:77-81 in numpy float32/float64 (the oracle's expression, one voxel)."""
    f = np.float32
    with np.errstate(all="ignore"):
        if not (f(t2) > 0):
            return -1
        px = f(f(f(t0) * f(fx)) / f(t2)) + f(cx)
        py = f(f(f(t1) * f(fy)) / f(t2)) + f(cy)
        rx, ry = np.floor(np.float64(px) + 0.5), np.floor(np.float64(py) + 0.5)
    if not (np.isfinite(rx) and np.isfinite(ry)):
        return -1                                   # (int) of NaN / inf is INT_MIN on x86: fails the bound test
    if 0 <= rx < cols and 0 <= ry < rows:
        return int(ry) * cols + int(rx)
    return -1


def test_both_pixel_forms_equal_the_reference_on_edge_cases(emul):
    """Pixel coordinates a few ulps either side of every rounding boundary (k + 0.5), of the image border (-0.5,
    dim - 0.5), depths inside and outside the fast-division range, huge / infinite / NaN operands."""
    rng = np.random.default_rng(20260922)
    cam = (525.0, 525.0, 319.5, 239.5, 640, 480)
    f32 = np.float32
    cases = []
    for _ in range(20000):
        t2 = f32(10.0 ** rng.uniform(-4.5, 6.5))
        kx = rng.integers(-2, 643)
        ky = rng.integers(-2, 483)
        # aim at the boundary kx - 0.5 (between pixels kx - 1 and kx), then step a few ulps around it
        t0 = f32((kx - 0.5 - cam[2]) * float(t2) / cam[0])
        t1 = f32((ky - 0.5 - cam[3]) * float(t2) / cam[1])
        for _ in range(int(rng.integers(0, 6))):
            t0 = np.nextafter(t0, f32(rng.choice([-np.inf, np.inf])))
        for _ in range(int(rng.integers(0, 6))):
            t1 = np.nextafter(t1, f32(rng.choice([-np.inf, np.inf])))
        cases.append((t0, t1, t2))
    special = [0.0, -0.0, 1e-3, 9.99e-4, 1e6, 1.0001e6, 1e-30, 1e30, 3e38, -1.0, np.inf, -np.inf, np.nan, 1.0, 2.5]
    for a in special:
        for b in special:
            for c in special:
                cases.append((f32(a), f32(b), f32(c)))
    bad = 0
    for (t0, t1, t2) in cases:
        want = ref_pixel(*cam, t0, t1, t2)
        for form in (1, 2):
            got = emul.emul_pixel(form, *[C.c_float(x) for x in cam[:4]], cam[4], cam[5], C.c_float(t0), C.c_float(t1), C.c_float(t2))
            if got != want:
                bad += 1
                assert bad < 5, (form, float(t0), float(t1), float(t2), got, want)
    assert bad == 0


@pytest.mark.parametrize("bias", [0, -2, 2])
def test_second_pixel_form_matches_oracle_on_scenes(oracle, emul, bias):
    emul.emul_set_rcp_bias(bias)
    emul.emul_set_pixel_form(2)
    try:
        run_scene(oracle, emul, helpers.scene_room(n=3))
        run_scene(oracle, emul, helpers.scene_camera(n=2))
    finally:
        emul.emul_set_rcp_bias(0)
        emul.emul_set_pixel_form(1)


# ------------------------------------------------------------------------------------------------ Reproject
def test_reproject_pixel_matches_oracle(oracle, emul):
    """CIntegrateApp::Reproject (ControlGrid warp + forward projection + z-buffer minimum): reproject_pixel from the
    product header, run over whole images on the CPU, against the oracle's Reproject -- the ControlGrid scene of the
    golden set (all frames, both fragments) and a stronger random lattice displacement."""
    emul.emul_reproject.argtypes = [C.POINTER(C.c_ushort), C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                    C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_int, C.c_float,
                                    C.POINTER(C.c_ushort)]
    sc = helpers.scene_ctr(num=2, interval=3)
    traj = helpers.compose_traj(sc["pose_traj"], sc["seg_traj"], sc["interval"], capi.mat4_mul)
    ovol = oracle.OracleTSDF()
    rng = np.random.default_rng(5)
    changed = 0
    for variant in range(2):
        for f in range(len(sc["frames"])):
            chunk = f // sc["interval"]
            ctr = np.ascontiguousarray(sc["grids"][chunk], dtype=np.float32)
            if variant == 1:
                ctr = (ctr + rng.normal(0, 0.004, ctr.shape)).astype(np.float32)      # a rougher lattice
            Ti, T0, A0, Ai = traj[f], traj[0], sc["seg_traj"][0], sc["seg_traj"][f]
            want = ovol.reproject(sc["frames"][f], Ti, T0, A0, Ai, ctr, sc["res"], sc["length"])
            adj = capi.mat4_mul(capi.mat4_mul(capi.mat4_inverse(Ti), T0), capi.mat4_inverse(A0))   # reproject_on_device
            depth = np.ascontiguousarray(sc["frames"][f], dtype=np.uint16)
            got = np.zeros_like(depth)
            emul.emul_reproject(depth.ctypes.data_as(C.POINTER(C.c_ushort)), 640, 480, 525.0, 525.0, 319.5, 239.5,
                                _dp(np.ascontiguousarray(Ai, dtype=np.float64)), _dp(np.ascontiguousarray(adj, dtype=np.float64)),
                                _fp(ctr.ravel()), sc["res"], C.c_float(sc["length"]), got.ctypes.data_as(C.POINTER(C.c_ushort)))
            assert np.array_equal(got, want), "variant %d frame %d: %d pixels differ" % (variant, f, int((got != want).sum()))
            changed += int((want != depth).sum())
    assert changed > 100000                      # the warp really moves pixels


def test_division_by_1000_is_exact(emul):
    """ScaleDepth's d * lambda / 1000.f (TSDFVolume.cpp:28) is computed as multiply + two FMAs on the device
    (scale_depth_div1000): equal to the IEEE division for EVERY float the fast path accepts, and for values outside it
    (which fall back to the division)."""
    emul.emul_check_div1000.restype = C.c_longlong
    emul.emul_check_div1000.argtypes = [C.c_float, C.c_float]
    assert emul.emul_check_div1000(0.5, 262144.0) == 0           # all 159 383 552 floats of the fast path
    assert emul.emul_check_div1000(0.0, 1e-30) == 0              # zero and denormals / tiny values
    assert emul.emul_check_div1000(262144.0, 262160.0) == 0      # just past the fast path
    assert emul.emul_check_div1000(1e-3, 0.5) == 0

