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
