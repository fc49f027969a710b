"""CPU property test of the conservative SKIP / ONE / FULL classification (no GPU).

classify_ball (elasticreconstruction_b200/csrc/er_voxel_math.h, the function classify_kernel calls) is compiled by
g++ together with the exact per-voxel path and run on every block the reference touches, for random camera poses
inside the room -- any orientation, walls at grazing angles, invalid holes, depth discontinuities, far spikes:
a sub-brick classified SKIP must contain no voxel the exact path updates, one classified ONE only voxels it updates
with tsdf == 1.  Any violation would show up on the GPU as a wrong voxel; here far more views are affordable."""
import ctypes as C

import numpy as np
import pytest

import helpers
from elasticreconstruction_b200 import capi
from test_voxel_math import emul, _fp, _dp  # noqa: F401  (fixture + pointer helpers)


def check_frames(oracle, emul, d, poses, cam, ul=3.0 / 512.0, trunc=0.03, cols=640, rows=480):
    emul.emul_check_classification.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                               C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int,
                                               C.POINTER(C.c_longlong)]
    ovol = oracle.OracleTSDF(cols, rows)
    ovol.set_camera(cam[0], cam[1], cam[2], cam[3], 2.5)
    tot = np.zeros(6, dtype=np.int64)
    for f in range(len(d)):
        T = np.ascontiguousarray(poses[f], dtype=np.float64)
        Tinv = np.ascontiguousarray(capi.mat4_inverse(T), dtype=np.float64)
        scaled = np.ascontiguousarray(ovol.scale_depth(d[f]), dtype=np.float32)
        ovol.integrate(d[f], T)
        for key in ovol.last_touched():
            out = np.zeros(6, dtype=np.int64)
            emul.emul_check_classification(_fp(scaled), cols, rows, cam[0], cam[1], cam[2], cam[3], ul, trunc, _dp(T), _dp(Tinv),
                                           int(key), out.ctypes.data_as(C.POINTER(C.c_longlong)))
            assert out[3] == 0, "frame %d block %d: %d sub-bricks classified SKIP/ONE against the exact path" % (f, key, out[3])
            tot += out
    return tot


@pytest.mark.parametrize("seed,holes", [(11, True), (12, True), (13, False), (14, False)])
def test_skip_and_one_never_contradict_the_exact_path_random_views(oracle, emul, seed, holes):
    d, poses, cam = helpers.random_scene(seed, 3, holes=holes)
    tot = check_frames(oracle, emul, d, poses, cam)
    # the test is not vacuous: the classes occur, and most sub-bricks avoid the exact path.  (With 2 % random pixel
    # dropouts nearly every footprint contains an invalid pixel, which rules ONE out: only the clean views see it.)
    assert tot[0] > 1000 and tot[2] > 1000
    if not holes:
        assert tot[1] > 1000
    assert tot[0] + tot[1] > tot[2]


def test_skip_and_one_on_the_benchmark_trajectory(oracle, emul):
    sc = helpers.scene_room(n=3)
    cam = (525.0, 525.0, 319.5, 239.5)
    tot = check_frames(oracle, emul, sc["frames"], sc["traj"], cam)
    assert tot[3] == 0 and tot[2] > 1000
    # slack of the classification on this scene (idle / all-one exact-path visits)
    assert tot[4] + tot[5] < tot[2]


def test_skip_and_one_odd_image_size(oracle, emul):
    d, poses, cam = helpers.random_scene(7, 2, 330, 247, holes=True)
    check_frames(oracle, emul, d, poses, cam, cols=330, rows=247)


def test_tighter_distance_bounds_are_still_conservative(oracle, emul):
    """classify_ball<TIGHT> (box extent along the viewing ray instead of the ball radius; not used by the kernels yet)
    must stay conservative, and it must actually remove exact-path work."""
    emul.emul_set_tight_classification.argtypes = [C.c_int]
    d, poses, cam = helpers.random_scene(13, 3, holes=False)
    base = check_frames(oracle, emul, d, poses, cam)
    emul.emul_set_tight_classification(1)
    try:
        tight = check_frames(oracle, emul, d, poses, cam)
        d2, poses2, cam2 = helpers.random_scene(12, 2, holes=True)
        check_frames(oracle, emul, d2, poses2, cam2)
    finally:
        emul.emul_set_tight_classification(0)
    assert tight[3] == 0 and tight[2] < base[2]


# ------------------------------------------------------------------------------------------------ touch kernel keys
def check_touch(emul, d, poses, cam, cols=640, rows=480, ul=3.0 / 512.0):
    emul.emul_check_touch.argtypes = [C.POINTER(C.c_ushort), C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_double,
                                      C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    tot = np.zeros(5, dtype=np.int64)
    for f in range(len(d)):
        depth = np.ascontiguousarray(d[f], dtype=np.uint16)
        T = np.ascontiguousarray(poses[f], dtype=np.float64)
        out = np.zeros(5, dtype=np.int64)
        emul.emul_check_touch(depth.ctypes.data_as(C.POINTER(C.c_ushort)), cols, rows, cam[0], cam[1], cam[2], cam[3], ul, _dp(T),
                              out.ctypes.data_as(C.POINTER(C.c_longlong)))
        assert out[4] == 0, "frame %d: %d block keys differ from the reference" % (f, out[4])
        tot += out
    return tot


def test_touch_block_keys_match_the_reference(emul):
    """Float estimate + exact fallback per pixel, and the tile-level frustum shortcut, against the reference's double
    arithmetic: random views with holes and spikes, the benchmark trajectory, an odd image size, and a coarse voxel
    size (block boundaries elsewhere)."""
    d, poses, cam = helpers.random_scene(21, 12)
    t1 = check_touch(emul, d, poses, cam)
    sc = helpers.scene_room(n=6)
    t2 = check_touch(emul, sc["frames"], sc["traj"], (525.0, 525.0, 319.5, 239.5))
    d3, poses3, cam3 = helpers.random_scene(22, 4, 330, 247)
    check_touch(emul, d3, poses3, cam3, 330, 247)
    check_touch(emul, d[:4], poses[:4], cam, ul=0.012)
    for t in (t1, t2):
        assert t[0] > 1000000                    # valid pixels
        assert 0 < t[1] < 0.05 * t[0]            # the exact fallback is rare but does occur
        assert t[3] > 0.5 * t[2]                 # most tiles take the one-key shortcut
