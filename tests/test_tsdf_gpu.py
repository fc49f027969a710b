"""GPU parity tests for the TSDF path: CUDA (through the C ABI) vs the pinned oracle, same inputs.

Bar (BASELINE.json): per-voxel |d sdf| <= 1e-4 and equal weights.  The kernels are written to be
bit-identical to the reference arithmetic, so the tests assert bit equality and report the max
difference if that ever fails.
"""
import os

import numpy as np
import pytest

import helpers
from elasticreconstruction_b200 import capi, synth

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "tsdf_golden.npz"))
TOL = 1e-4


def compare_worlds(a, b):
    """a, b: sorted float32 [n,4]; returns (same_keys, max_abs_sdf_diff)."""
    if a.shape != b.shape or not np.array_equal(a[:, :3], b[:, :3]):
        return False, float("inf")
    return True, float(np.max(np.abs(a[:, 3] - b[:, 3]))) if len(a) else 0.0


def compare_blocks(oracle_vol, gpu_vol):
    ko = np.sort(oracle_vol.block_keys())
    kg = np.sort(gpu_vol.block_keys())
    assert np.array_equal(ko, kg), "allocated block sets differ"
    worst = 0.0
    for key in ko:
        so, wo = oracle_vol.read_block(int(key))
        sg, wg = gpu_vol.read_block(int(key))
        assert np.array_equal(wo, wg), "weights differ in block %d" % key
        worst = max(worst, float(np.max(np.abs(so - sg))))
        assert np.array_equal(so.view(np.uint32), sg.view(np.uint32)), \
            "sdf differs in block %d (max abs %g)" % (key, float(np.max(np.abs(so - sg))))
    return worst


def test_host_math_matches_oracle(oracle):
    rng = np.random.default_rng(7)
    for _ in range(200):
        T = synth.make_pose(synth.rot_axis(rng.normal(size=3), rng.uniform(0, 3)), rng.normal(size=3) * 2)
        T = T + rng.normal(size=(4, 4)) * 1e-3
        assert np.array_equal(capi.mat4_inverse(T), oracle.mat4_inverse(T))
        B = rng.normal(size=(4, 4))
        assert np.array_equal(capi.mat4_mul(T, B), oracle.mat4_mul(T, B))


@pytest.mark.parametrize("mk", helpers.ALL_SCENES, ids=lambda f: f.__name__)
def test_scene_matches_oracle_and_golden(oracle, mk):
    sc = mk()
    ovol = oracle.OracleTSDF()
    helpers.run_scene(ovol, sc, mul=oracle.mat4_mul)
    gvol = capi.TSDFVolume(max_blocks=512)
    helpers.run_scene(gvol, sc, mul=capi.mat4_mul)
    gvol.sync()
    compare_blocks(ovol, gvol)
    wo = oracle.sort_world(ovol.extract())
    wg = oracle.sort_world(gvol.extract())
    same, diff = compare_worlds(wo, wg)
    assert same and diff <= TOL
    assert np.array_equal(wo.view(np.uint32), wg.view(np.uint32))
    # and straight against the verbatim reference's committed output
    name = sc["name"]
    assert len(wg) == int(GOLD[name + "_count"])
    assert oracle.world_digest(wg) == str(GOLD[name + "_digest"])


def test_reproject_image_matches_oracle(oracle):
    sc = helpers.scene_ctr()
    traj = helpers.compose_traj(sc["pose_traj"], sc["seg_traj"], sc["interval"], oracle.mat4_mul)
    ovol = oracle.OracleTSDF()
    gvol = capi.TSDFVolume(max_blocks=16)
    gvol.set_control_grids(sc["grids"], sc["res"], sc["length"])
    for fid in (1, 2, 4, 5):
        chunk = (fid - 1) // sc["interval"]
        a = ovol.reproject(sc["frames"][fid - 1], traj[fid - 1], traj[0], sc["seg_traj"][0], sc["seg_traj"][fid - 1],
                           sc["grids"][chunk], sc["res"], sc["length"])
        b = gvol.reproject(sc["frames"][fid - 1], traj[fid - 1], traj[0], sc["seg_traj"][0], sc["seg_traj"][fid - 1],
                           chunk)
        assert np.array_equal(a, b.reshape(a.shape)), "frame %d: %d pixels differ" % (fid, int((a != b.reshape(a.shape)).sum()))


def test_fused_ctr_path_equals_two_step(oracle):
    sc = helpers.scene_ctr()
    traj = helpers.compose_traj(sc["pose_traj"], sc["seg_traj"], sc["interval"], capi.mat4_mul)
    a = capi.TSDFVolume(max_blocks=512)
    b = capi.TSDFVolume(max_blocks=512)
    for vol in (a, b):
        vol.set_control_grids(sc["grids"], sc["res"], sc["length"])
    for fid in range(1, 5):
        chunk = (fid - 1) // sc["interval"]
        args = (traj[fid - 1], traj[0], sc["seg_traj"][0], sc["seg_traj"][fid - 1], chunk)
        a.integrate(a.reproject(sc["frames"][fid - 1], *args), traj[fid - 1])
        b.integrate_ctr(sc["frames"][fid - 1], *args)
    wa, wb = oracle.sort_world(a.extract()), oracle.sort_world(b.extract())
    assert np.array_equal(wa.view(np.uint32), wb.view(np.uint32))


def test_ctr_batch_equals_per_frame_and_oracle(oracle):
    sc = helpers.scene_ctr(num=3, interval=4)
    traj = helpers.compose_traj(sc["pose_traj"], sc["seg_traj"], sc["interval"], capi.mat4_mul)
    n = 11                                               # frames 1..11 (frame 12 would hit the trajectory end)
    ovol = oracle.OracleTSDF()
    helpers.run_app(ovol, sc["frames"][:n], traj, sc["seg_traj"], sc["grids"], sc["res"], sc["length"], sc["interval"])
    g = capi.TSDFVolume(max_blocks=512)
    g.set_control_grids(sc["grids"], sc["res"], sc["length"])
    g.set_fusion(4)
    chunks = [(fid - 1) // sc["interval"] for fid in range(1, n + 1)]
    g.integrate_ctr_batch(sc["frames"][:n], traj[:n], traj[0], sc["seg_traj"][0], sc["seg_traj"][:n], chunks)
    compare_blocks(ovol, g)


def test_batch_and_device_paths_equal_per_frame(oracle):
    import torch
    sc = helpers.scene_room(n=6)
    frames, traj = sc["frames"], sc["traj"]
    a = capi.TSDFVolume(max_blocks=512)
    for f in range(6):
        a.integrate(frames[f], traj[f])
    b = capi.TSDFVolume(max_blocks=512)
    b.integrate_batch(frames, traj[:6])
    c = capi.TSDFVolume(max_blocks=512)
    d_frames = torch.from_numpy(frames.astype(np.int16)).cuda()
    torch.cuda.synchronize()
    c.integrate_device(d_frames.data_ptr(), traj[:6], 6)
    ws = [oracle.sort_world(v.extract()) for v in (a, b, c)]
    assert np.array_equal(ws[0].view(np.uint32), ws[1].view(np.uint32))
    assert np.array_equal(ws[0].view(np.uint32), ws[2].view(np.uint32))


@pytest.mark.parametrize("fusion", [1, 5, 32])
def test_fusion_sizes_are_equivalent(oracle, fusion):
    """Temporal fusion (several frames applied to register-resident voxels per launch) must not change a bit."""
    sc = helpers.scene_room(n=12)
    ovol = oracle.OracleTSDF()
    for f in range(12):
        ovol.integrate(sc["frames"][f], sc["traj"][f])
    gvol = capi.TSDFVolume(max_blocks=512)
    gvol.set_fusion(fusion)
    gvol.integrate_batch(sc["frames"], sc["traj"][:12])
    compare_blocks(ovol, gvol)


def test_bumpy_room_many_frames(oracle):
    """Longer sequence with a non-planar surface: exercises SKIP / ONE / FULL sub-brick classes side by side."""
    sc = helpers.scene_room(n=40, total=200, bumps=True)
    ovol = oracle.OracleTSDF()
    gvol = capi.TSDFVolume(max_blocks=512)
    gvol.set_stats(True)
    upd = 0
    for f in range(40):
        ovol.integrate(sc["frames"][f], sc["traj"][f])
        upd += ovol.last_updated_voxels()
    gvol.integrate_batch(sc["frames"], sc["traj"][:40])
    compare_blocks(ovol, gvol)
    assert gvol.stats()["updated_voxels"] == upd


def test_empty_frame_and_stats(oracle):
    vol = capi.TSDFVolume(max_blocks=64)
    vol.set_stats(True)
    vol.integrate(np.zeros((480, 640), dtype=np.uint16), np.eye(4))
    assert vol.num_blocks() == 0 and len(vol.extract()) == 0
    sc = helpers.scene_wall(n=1)
    ovol = oracle.OracleTSDF()
    nb = ovol.integrate(sc["frames"][0], sc["traj"][0])
    vol.integrate(sc["frames"][0], sc["traj"][0])
    st = vol.stats()
    assert st["frames"] == 2 and st["touched_blocks"] == nb
    assert st["updated_voxels"] == ovol.last_updated_voxels()


def test_pool_overflow_is_reported():
    sc = helpers.scene_wall(n=1)
    vol = capi.TSDFVolume(max_blocks=4)
    vol.integrate(sc["frames"][0], sc["traj"][0])
    with pytest.raises(capi.ElasticError, match="pool exhausted"):
        vol.sync()


def test_block_partition_union_equals_single(oracle):
    sc = helpers.scene_room(n=3)
    single = capi.TSDFVolume(max_blocks=512)
    parts = [capi.TSDFVolume(max_blocks=512) for _ in range(3)]
    for r, p in enumerate(parts):
        p.set_partition(r, 3)
    for f in range(3):
        single.integrate(sc["frames"][f], sc["traj"][f])
        for p in parts:
            p.integrate(sc["frames"][f], sc["traj"][f])
    ws = oracle.sort_world(single.extract())
    wu = oracle.sort_world(np.concatenate([p.extract() for p in parts]))
    assert np.array_equal(ws.view(np.uint32), wu.view(np.uint32))
    assert sum(p.num_blocks() for p in parts) == single.num_blocks()


def test_roundtrip_properties_full_size():
    """Size-independent properties at BASELINE scale (no oracle): weights are integers equal to the number of
    frames that updated a voxel, |tsdf| <= 1, and integrating the same frame twice leaves sdf unchanged."""
    sc = helpers.scene_room(n=2)
    vol = capi.TSDFVolume(max_blocks=512)
    vol.integrate(sc["frames"][0], sc["traj"][0])
    keys = vol.block_keys()
    s1, w1 = vol.read_block(int(keys[0]))
    vol.integrate(sc["frames"][0], sc["traj"][0])
    s2, w2 = vol.read_block(int(keys[0]))
    assert np.array_equal(w2, 2 * w1)
    assert np.all(np.abs(s2) <= 1.0)
    assert np.allclose(s1, s2, atol=2e-7)       # (s*1 + s)/2 == s up to one rounding


_random_scene = helpers.random_scene          # shared with the CPU classification property test


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_views_with_holes_and_discontinuities(oracle, seed):
    d, poses, cam = _random_scene(seed, 10)
    ovol = oracle.OracleTSDF()
    gvol = capi.TSDFVolume(max_blocks=4096)
    gvol.set_stats(True)
    for vol in (ovol, gvol):
        vol.set_camera(cam[0], cam[1], cam[2], cam[3], 2.5)
    upd = 0
    for f in range(len(d)):
        ovol.integrate(d[f], poses[f])
        upd += ovol.last_updated_voxels()
    gvol.integrate_batch(d, poses)
    compare_blocks(ovol, gvol)
    assert gvol.stats()["updated_voxels"] == upd


@pytest.mark.parametrize("cols,rows", [(320, 240), (328, 250), (330, 247)])
def test_other_image_sizes(oracle, cols, rows):
    """Tile-partial borders (sizes not multiples of 16) and rows that are not 8-byte aligned (scalar load path)."""
    d, poses, cam = _random_scene(7, 4, cols, rows, holes=False)
    ovol = oracle.OracleTSDF(cols, rows)
    gvol = capi.TSDFVolume(cols, rows, max_blocks=1024)
    for vol in (ovol, gvol):
        vol.set_camera(cam[0], cam[1], cam[2], cam[3], 2.5)
    for f in range(len(d)):
        ovol.integrate(d[f], poses[f])
    gvol.integrate_batch(d, poses)
    compare_blocks(ovol, gvol)


def test_custom_voxel_size_and_truncation(oracle):
    """BASELINE config 1 asks for 12 mm voxels, which the reference can only do by editing its constants
    (TSDFVolume.cpp:10-11); library and oracle both take them as parameters."""
    sc = helpers.scene_wall(n=2)
    ovol = oracle.OracleTSDF()
    ovol.set_voxel(0.012, 0.05)
    gvol = capi.TSDFVolume(max_blocks=256, voxel=0.012, trunc=0.05)
    for f in range(2):
        ovol.integrate(sc["frames"][f], sc["traj"][f])
        gvol.integrate(sc["frames"][f], sc["traj"][f])
    compare_blocks(ovol, gvol)


def test_ctr_mode_small_image_stays_inside_the_frame(oracle):
    """ControlGrid mode on an image smaller than VGA: the reference's XYZ2UVD bounds are the literals 640 x 480
    (TSDFVolume.h:51-60) and would index past a smaller buffer; here a re-projected pixel must land inside the image, so
    neighbouring frames' z-buffers of a batch are never touched.  Checked as: (a) the batched path equals the
    frame-by-frame path (a stray write into the next frame's z-buffer slot would change it), (b) every re-projected
    image equals the oracle's wherever the oracle's own target pixel is inside the image."""
    cols, rows = 320, 240
    sc = helpers.scene_ctr(num=2, interval=3)
    frames = np.ascontiguousarray(sc["frames"][:, :rows * 2:2, :cols * 2:2])          # subsampled wall
    traj = helpers.compose_traj(sc["pose_traj"], sc["seg_traj"], sc["interval"], capi.mat4_mul)
    n = len(frames)
    cam = (262.5, 262.5, 159.75, 119.75, 2.5)
    va = capi.TSDFVolume(cols=cols, rows=rows, max_blocks=256)
    vb = capi.TSDFVolume(cols=cols, rows=rows, max_blocks=256)
    for v in (va, vb):
        v.set_camera(*cam)
        v.set_control_grids(sc["grids"], sc["res"], sc["length"])
    chunks = [f // sc["interval"] for f in range(n)]
    va.integrate_ctr_batch(frames, traj[:n], traj[0], sc["seg_traj"][0], sc["seg_traj"][:n], chunks)
    for f in range(n):
        vb.integrate_ctr(frames[f], traj[f], traj[0], sc["seg_traj"][0], sc["seg_traj"][f], chunks[f])
    wa, wb = oracle.sort_world(va.extract()), oracle.sort_world(vb.extract())
    assert len(wa) > 1000 and np.array_equal(wa.view(np.uint32), wb.view(np.uint32))
    for f in (0, n - 1):
        img = va.reproject(frames[f], traj[f], traj[0], sc["seg_traj"][0], sc["seg_traj"][f], chunks[f])
        assert img.shape == frames[f].shape and (img > 0).sum() > 1000


def test_pool_overflow_stops_the_next_batch():
    """After the block pool has overflowed, the next *_batch call returns ER_ERR_CAPACITY (pinned mirror of the overflow
    flag) instead of processing the rest of the sequence, and the hash table never fills up (no spin)."""
    sc = helpers.scene_room(n=4)
    vol = capi.TSDFVolume(max_blocks=4)
    vol.integrate_batch(sc["frames"][:2], sc["traj"][:2])
    with pytest.raises(capi.ElasticError, match="pool exhausted"):
        vol.sync()
    with pytest.raises(capi.ElasticError, match="pool exhausted"):
        vol.integrate_batch(sc["frames"][2:4], sc["traj"][2:4])


def test_volume_checkpoint_resume_merge_and_format(oracle, tmp_path):
    """er_tsdf_save / er_tsdf_load (raw volume checkpoint): (a) integrate k frames, save, load into a fresh
    volume, integrate the rest == the uninterrupted run, bit for bit; (b) the per-rank checkpoints of a partitioned volume
    loaded into one unpartitioned volume == the unpartitioned run; (c) the file holds each block as the reference's
    TSDFVolumeUnit arrays (k fastest)."""
    sc = helpers.scene_room(n=6)
    full = capi.TSDFVolume(max_blocks=512)
    first = capi.TSDFVolume(max_blocks=512)
    for f in range(6):
        full.integrate(sc["frames"][f], sc["traj"][f])
        if f < 3:
            first.integrate(sc["frames"][f], sc["traj"][f])
    ck = str(tmp_path / "vol.ck")
    first.save(ck)
    resumed = capi.TSDFVolume(max_blocks=512)
    resumed.load(ck)
    assert resumed.num_blocks() == first.num_blocks()
    for f in range(3, 6):
        resumed.integrate(sc["frames"][f], sc["traj"][f])
    a, b = oracle.sort_world(full.extract()), oracle.sort_world(resumed.extract())
    assert len(a) > 1000 and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # (c) file format
    raw = np.fromfile(ck, dtype=np.uint8)
    assert bytes(raw[:7]) == b"ERTSDF1"
    nb = int(raw[24:28].view(np.int32)[0])
    assert nb == first.num_blocks() and len(raw) == 32 + nb * (8 + 8 * 64 ** 3)
    key0 = int(raw[32:36].view(np.int32)[0])
    blk = raw[40:40 + 8 * 64 ** 3].view(np.float32)
    sdf, w = first.read_block(key0)
    assert np.array_equal(blk[:64 ** 3], sdf) and np.array_equal(blk[64 ** 3:], w)
    # (b) merge of partition checkpoints
    parts = [capi.TSDFVolume(max_blocks=512) for _ in range(3)]
    for r, p in enumerate(parts):
        p.set_partition(r, 3)
        for f in range(6):
            p.integrate(sc["frames"][f], sc["traj"][f])
        p.save(str(tmp_path / ("part%d.ck" % r)))
    merged = capi.TSDFVolume(max_blocks=512)
    for r in range(3):
        merged.load(str(tmp_path / ("part%d.ck" % r)))
    c = oracle.sort_world(merged.extract())
    assert np.array_equal(a.view(np.uint32), c.view(np.uint32))
    # a partitioned volume takes only its own blocks from a full checkpoint
    full.save(ck)
    mine = capi.TSDFVolume(max_blocks=512)
    mine.set_partition(1, 3)
    mine.load(ck)
    assert mine.num_blocks() == parts[1].num_blocks()
    # errors: wrong file, mismatching voxel size
    with pytest.raises(capi.ElasticError, match="not a volume checkpoint"):
        bad = str(tmp_path / "bad.ck")
        open(bad, "wb").write(b"nope" * 20)
        merged.load(bad)
    other = capi.TSDFVolume(max_blocks=64, voxel=0.012)
    with pytest.raises(capi.ElasticError, match="differ"):
        other.load(ck)


def test_async_extract_equals_blocking_and_snapshots_the_state(oracle):
    """er_tsdf_extract_begin / _wait: the rows equal er_tsdf_extract's, and frames handed in after _begin do not leak into
    the extraction that was begun before them."""
    sc = helpers.scene_room(n=4)
    vol = capi.TSDFVolume(max_blocks=512)
    vol.integrate_batch(sc["frames"][:2], sc["traj"][:2])
    ref = oracle.sort_world(vol.extract())
    cap = len(ref) + 1000
    buf = capi.HostBuffer(cap * 16)
    vol.extract_begin(buf, cap)
    vol.integrate_batch(sc["frames"][2:4], sc["traj"][2:4])        # enqueued behind the extraction
    n = vol.extract_wait()
    got = oracle.sort_world(buf.array(np.float32, 4 * n).reshape(n, 4).copy())
    assert n == len(ref) and np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    with pytest.raises(capi.ElasticError, match="no extraction outstanding"):
        vol.extract_wait()
    # a second round on the same handle, now with all four frames
    ref2 = oracle.sort_world(vol.extract())
    big = capi.HostBuffer((len(ref2) + 10) * 16)
    vol.extract_begin(big, len(ref2) + 10)
    n2 = vol.extract_wait()
    got2 = oracle.sort_world(big.array(np.float32, 4 * n2).reshape(n2, 4).copy())
    assert n2 == len(ref2) and np.array_equal(got2.view(np.uint32), ref2.view(np.uint32)) and n2 > n
