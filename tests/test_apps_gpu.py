"""GPU tests of the two drop-in executables: same files in, same files out as the reference.

Integrate is compared with the VERBATIM reference binary (oracle/_ref/Integrate_ref, which travels to the GPU box
as a built file) on identical input files; BuildCorrespondence against a Python restatement of CCorresApp's control
flow driven by the restated ICP oracle (PCL is not available, parity unpinned -- see oracle/icp_oracle.c)."""
import os
import subprocess

import numpy as np
import pytest

import helpers
from elasticreconstruction_b200 import build, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "elasticreconstruction_b200", "bin")


@pytest.fixture(scope="module", autouse=True)
def apps():
    build.build_library()
    build.build_apps()


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180)
    return p.returncode, p.stdout


@pytest.mark.parametrize("mk", helpers.ALL_SCENES, ids=lambda f: f.__name__)
def test_integrate_cli_matches_reference_binary(oracle, tmp_path, mk):
    sc = mk()
    d = str(tmp_path)
    args = helpers.write_scene(sc, d)
    mine_pcd = os.path.join(d, "mine.pcd")
    my_args = [a if a != os.path.join(d, "world.pcd") else mine_pcd for a in args]
    rc, out = run([os.path.join(BIN, "Integrate")] + my_args, d)
    assert rc == 0, out
    mine = oracle.sort_world(synth.read_pcd_xyzi(mine_pcd))
    if os.path.exists(oracle.REF_INTEGRATE):
        ref_out, _ = oracle.run_reference_integrate(args, cwd=d)
        ref = oracle.sort_world(synth.read_pcd_xyzi(os.path.join(d, "world.pcd")))
        line = [l for l in ref_out.splitlines() if l.startswith("Total ")]
        assert line and line[0] in out                      # "Total N frames processed."
        assert ("%d voxel points have been written." % len(ref)) in out
    else:
        vol = oracle.OracleTSDF()
        helpers.run_scene(vol, sc, mul=oracle.mat4_mul)
        ref = oracle.sort_world(vol.extract())
    assert mine.shape == ref.shape
    assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32))


def test_integrate_cli_12mm_voxels(oracle, tmp_path):
    """BASELINE.json configs[0] (10 frames, identity .log, 12 mm voxels): --voxel_size against the reference binary
    built with its unit_length_ constant rewritten to 0.012 (oracle/Makefile ref12), else against the oracle."""
    sc = helpers.scene_wall(n=10)
    d = str(tmp_path)
    args = helpers.write_scene(sc, d)
    mine_pcd = os.path.join(d, "mine.pcd")
    my_args = [a if a != os.path.join(d, "world.pcd") else mine_pcd for a in args] + ["--voxel_size", "0.012"]
    rc, out = run([os.path.join(BIN, "Integrate")] + my_args, d)
    assert rc == 0 and "Total 10 frames processed." in out, out
    mine = oracle.sort_world(synth.read_pcd_xyzi(mine_pcd))
    if os.path.exists(oracle.REF_INTEGRATE_12MM):
        oracle.run_reference_integrate(args, cwd=d, binary=oracle.REF_INTEGRATE_12MM)
        ref = oracle.sort_world(synth.read_pcd_xyzi(os.path.join(d, "world.pcd")))
    else:
        vol = oracle.OracleTSDF()
        vol.set_voxel(0.012, 0.03)
        helpers.run_scene(vol, sc, mul=oracle.mat4_mul)
        ref = oracle.sort_world(vol.extract())
    assert mine.shape == ref.shape
    assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32))


def test_integrate_cli_help_and_bad_source(tmp_path):
    rc, out = run([os.path.join(BIN, "Integrate")], str(tmp_path))
    assert rc == 0 and "--ref_traj <log_file>" in out            # no arguments: help, rc 0 (Integrate.cpp:37-39)
    rc, out = run([os.path.join(BIN, "Integrate"), "-oni", "missing.raw"], str(tmp_path))
    assert rc != 0 and "Can't open depth source" in out          # Integrate.cpp:60


def test_integrate_cli_start_end_window(oracle, tmp_path):
    sc = helpers.scene_room(n=6)
    d = str(tmp_path)
    args = helpers.write_scene(sc, d) + ["--start_from", "2", "--end_at", "4"]
    rc, out = run([os.path.join(BIN, "Integrate")] + args, d)
    assert rc == 0 and "Total 5 frames processed." in out         # frame 5 > end_at sets exit (IntegrateApp.cpp:209-216)
    vol = oracle.OracleTSDF()
    helpers.run_app(vol, sc["frames"], sc["traj"], start_from=2, end_at=4)
    ref = oracle.sort_world(vol.extract())
    mine = oracle.sort_world(synth.read_pcd_xyzi(os.path.join(d, "world.pcd")))
    assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32))


# ------------------------------------------------------------------------------------------------ BuildCorrespondence
def emulate_corres_app(oracle, clouds, pairs, Tinit, num, reg_dist=0.03, reg_ratio=0.25, reg_num=40000,
                       dist_thresh=0.015, normal_thresh=0.8660, registration=True):
    """CCorresApp::{Registration,FindCorrespondence} (CorresApp.cpp:112-319) on top of the oracle primitives."""
    recs = []
    for (i, j), T in zip(pairs, Tinit):
        frame, T = num, np.array(T, dtype=np.float64)
        if registration:
            cnt = oracle.icp_count_inliers(clouds[i], clouds[j], T, reg_dist)
            r1, r2 = cnt / len(clouds[i]), cnt / len(clouds[j])
            if cnt >= reg_num or (r1 > reg_ratio and r2 > reg_ratio):
                frame = cnt
                T = oracle.icp_align(clouds[i], clouds[j], T, reg_dist, 20, 1e-6, False)[0]
            else:
                frame = -1
        corr, info = None, None
        if frame != -1:
            corr, info = oracle.icp_correspondences(clouds[i], clouds[j], T, dist_thresh, normal_thresh)
            if len(corr) / frame < 0.5:
                frame = -1 if reg_num > 0 else len(corr)
            else:
                frame = len(corr)
        recs.append(dict(i=i, j=j, frame=frame, T=T, corr=corr, info=info))
    return recs


def parse_log(path):
    L = [l for l in open(path) if not l.startswith("#")]
    out = []
    for k in range(len(L) // 5):
        ids = [int(x) for x in L[5 * k].split()]
        T = np.array([[float(x) for x in L[5 * k + 1 + r].split()] for r in range(4)])
        out.append((ids, T))
    return out


def test_buildcorrespondence_cli_matches_emulation(oracle, tmp_path):
    d = str(tmp_path / "data")
    os.makedirs(d)
    n_frag = 4
    frags = [synth.make_local_fragment(i, 30000) for i in range(n_frag)]
    for i, (x, n) in enumerate(frags):
        synth.write_pcd_xyzrgbnormal(os.path.join(d, "cloud_bin_%d.pcd" % i), x, n)
    pairs = [(i, j) for i in range(n_frag) for j in range(i + 1, n_frag)]
    Tinit = [synth.quantize_log(synth.pair_init(i, j)) for (i, j) in pairs]
    Tinit[-1] = synth.quantize_log(Tinit[-1] @ synth.make_pose(np.eye(3), [2.0, 0, 0]))      # hopeless pair -> rejected
    synth.write_log(os.path.join(d, "pairs.log"), np.array(Tinit), ids=[(i, j, n_frag) for (i, j) in pairs])
    cwd = str(tmp_path / "run")
    os.makedirs(cwd)
    rc, out = run([os.path.join(BIN, "BuildCorrespondence"), "--reg_traj", os.path.join(d, "pairs.log"), "--registration",
                   "--reg_num", "5000", "--output_information", "--save_xyzn"], cwd)
    assert rc == 0, out
    clouds = [oracle.OracleCloud(x, n, 0.03) for (x, n) in frags]
    recs = emulate_corres_app(oracle, clouds, pairs, Tinit, n_frag, reg_num=5000)
    got = parse_log(os.path.join(cwd, "reg_output.log"))           # written to the CWD (CorresApp.cpp:323)
    assert len(got) == len(recs)
    assert sum(r["frame"] == -1 for r in recs) >= 1 and sum(r["frame"] > 0 for r in recs) >= 4
    info_lines = open(os.path.join(cwd, "reg_output.info")).read().split("\n")
    for k, (r, (ids, T)) in enumerate(zip(recs, got)):
        assert ids[:2] == [r["i"], r["j"]]
        assert ids[2] == r["frame"], (ids, r["frame"])
        assert np.linalg.norm(T - r["T"]) <= 1e-5 + 4e-8        # 1e-5 bar + the %.8f text rounding
        fn = os.path.join(d, "corres_%d_%d.txt" % (r["i"], r["j"]))
        if r["corr"] is not None:
            mine = np.loadtxt(fn, dtype=np.int64).reshape(-1, 2)
            # the correspondence set depends on the pose in the last float bits; require near-identity of the sets
            a = set(map(tuple, mine.tolist())); b = set(map(tuple, r["corr"].tolist()))
            assert len(a ^ b) <= max(4, len(b) // 2000), (len(a), len(b), len(a ^ b))
            blk = np.array([[float(x) for x in info_lines[7 * k + 1 + q].split()] for q in range(6)])
            assert abs(blk[0, 0] - len(mine)) < 0.5
            assert np.allclose(blk, r["info"], rtol=2e-3, atol=1.0)
        else:
            assert not os.path.exists(fn)
    x0 = np.loadtxt(os.path.join(d, "cloud_bin_xyzn_0.xyzn"))
    assert x0.shape == (len(frags[0][0]), 6) and np.allclose(x0[:, :3], frags[0][0], atol=5e-7)


def test_buildcorrespondence_traj_mode_creates_candidates(tmp_path):
    """--traj/--num mode: candidate pairs from odometry + volume overlap (CorresApp.cpp:40-69)."""
    d = str(tmp_path / "data")
    os.makedirs(d)
    n_frag, interval = 3, 2
    for i in range(n_frag):
        x, n = synth.make_local_fragment(i, 5000)
        synth.write_pcd_xyzrgbnormal(os.path.join(d, "cloud_bin_%d.pcd" % i), x, n)
    poses = [synth.make_pose(synth.rot_axis([0, 1, 0], 0.01 * k), [0.01 * k, 0, 0]) for k in range(n_frag * interval)]
    synth.write_log(os.path.join(d, "traj.log"), synth.quantize_log(np.array(poses)))
    cwd = str(tmp_path / "run")
    os.makedirs(cwd)
    rc, out = run([os.path.join(BIN, "BuildCorrespondence"), "--traj", os.path.join(d, "traj.log"), "--num", str(n_frag),
                   "--interval", str(interval)], cwd)
    assert rc == 0 and "3 initial matching candidates are created." in out, out
    got = parse_log(os.path.join(cwd, "reg_output.log"))
    assert [g[0][:2] for g in got] == [[0, 1], [0, 2], [1, 2]]


def test_integrate_cli_skips_flagged_frames(oracle, tmp_path):
    """A trajectory record with frame_ == -1 is skipped (IntegrateApp.cpp:196)."""
    sc = helpers.scene_room(n=5)
    d = str(tmp_path)
    synth.write_depth_container(os.path.join(d, "depth.raw"), sc["frames"])
    ids = [(i, i, -1 if i in (1, 3) else i + 1) for i in range(len(sc["traj"]))]
    synth.write_log(os.path.join(d, "traj.log"), sc["traj"], ids=ids)
    rc, out = run([os.path.join(BIN, "Integrate"), "-oni", os.path.join(d, "depth.raw"), "--ref_traj",
                   os.path.join(d, "traj.log"), "--save_to", os.path.join(d, "w.pcd")], d)
    assert rc == 0, out
    vol = oracle.OracleTSDF()
    helpers.run_app(vol, sc["frames"], sc["traj"], frame_flags=[f for (_, _, f) in ids])
    ref = oracle.sort_world(vol.extract())
    mine = oracle.sort_world(synth.read_pcd_xyzi(os.path.join(d, "w.pcd")))
    assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32))


def test_buildcorrespondence_blacklist_and_redux(oracle, tmp_path):
    """--blacklist drops every pair touching a listed fragment (CorresApp.cpp:222-229); --redux replaces the ICP
    result of listed pairs by the given transform (CorresApp.cpp:283-293) without running ICP on them."""
    d = str(tmp_path / "data")
    os.makedirs(d)
    n_frag = 3
    frags = [synth.make_local_fragment(i, 20000) for i in range(n_frag)]
    for i, (x, n) in enumerate(frags):
        synth.write_pcd_xyzrgbnormal(os.path.join(d, "cloud_bin_%d.pcd" % i), x, n)
    pairs = [(0, 1), (0, 2), (1, 2)]
    Tinit = [synth.quantize_log(synth.pair_init(i, j)) for (i, j) in pairs]
    synth.write_log(os.path.join(d, "pairs.log"), np.array(Tinit), ids=[(i, j, n_frag) for (i, j) in pairs])
    with open(os.path.join(d, "black.txt"), "w") as f:
        f.write("# fragments to ignore\n2\n")
    marker = synth.quantize_log(Tinit[0] @ synth.make_pose(np.eye(3), [0.001, 0.0, 0.0]))
    synth.write_log(os.path.join(d, "redux.log"), np.array([marker]), ids=[(0, 1, 3)])
    cwd = str(tmp_path / "run")
    os.makedirs(cwd)
    rc, out = run([os.path.join(BIN, "BuildCorrespondence"), "--reg_traj", os.path.join(d, "pairs.log"), "--registration",
                   "--reg_num", "2000", "--blacklist", os.path.join(d, "black.txt"), "--redux", os.path.join(d, "redux.log")], cwd)
    assert rc == 0, out
    got = parse_log(os.path.join(cwd, "reg_output.log"))
    assert [g[0][:2] for g in got] == [[0, 1], [0, 2], [1, 2]]
    assert got[1][0][2] == -1 and got[2][0][2] == -1               # blacklisted pairs are rejected
    assert np.allclose(got[0][1], marker, atol=5e-9)                # redux transform taken verbatim, no ICP
    assert got[0][0][2] > 0
    assert not os.path.exists(os.path.join(d, "corres_0_2.txt")) and os.path.exists(os.path.join(d, "corres_0_1.txt"))


def _device_count():
    import torch
    return torch.cuda.device_count()


def test_integrate_cli_devices_matches_reference_binary(oracle, tmp_path):
    """--devices 0,1[,..]: ONE volume whose 64^3 blocks are partitioned over the GPUs in-process (er_tsdf_set_partition,
    one host thread per device); the concatenated world.pcd equals the verbatim reference binary's, rigid and ControlGrid
    mode (TSDFVolume.cpp:45-63, IntegrateApp.cpp:217-225).  Skips below 2 devices."""
    nd = _device_count()
    if nd < 2:
        pytest.skip("needs >= 2 CUDA devices")
    devs = ",".join(str(k) for k in range(min(nd, 4)))
    for mk in (helpers.scene_room, helpers.scene_ctr):
        sc = mk()
        d = str(tmp_path / mk.__name__)
        os.makedirs(d)
        args = helpers.write_scene(sc, d)
        mine_pcd = os.path.join(d, "mine.pcd")
        my_args = [a if a != os.path.join(d, "world.pcd") else mine_pcd for a in args] + ["--devices", devs]
        rc, out = run([os.path.join(BIN, "Integrate")] + my_args, d)
        assert rc == 0, out
        mine = oracle.sort_world(synth.read_pcd_xyzi(mine_pcd))
        if os.path.exists(oracle.REF_INTEGRATE):
            oracle.run_reference_integrate(args, cwd=d)
            ref = oracle.sort_world(synth.read_pcd_xyzi(os.path.join(d, "world.pcd")))
        else:
            vol = oracle.OracleTSDF()
            helpers.run_scene(vol, sc, mul=oracle.mat4_mul)
            ref = oracle.sort_world(vol.extract())
        assert mine.shape == ref.shape and np.array_equal(mine.view(np.uint32), ref.view(np.uint32))


def test_buildcorrespondence_cli_devices_equals_single_device(tmp_path):
    """--devices: the gate, the ICP share and the correspondence pass are dealt round-robin over the GPUs; every output
    file equals the single-device run's byte for byte (pairs are independent, CorresApp.cpp:121,220).  Skips below 2."""
    nd = _device_count()
    if nd < 2:
        pytest.skip("needs >= 2 CUDA devices")
    d = str(tmp_path / "data")
    os.makedirs(d)
    n_frag = 4
    for i in range(n_frag):
        x, n = synth.make_local_fragment(i, 30000)
        synth.write_pcd_xyzrgbnormal(os.path.join(d, "cloud_bin_%d.pcd" % i), x, n)
    pairs = [(i, j) for i in range(n_frag) for j in range(i + 1, n_frag)]
    Tinit = [synth.quantize_log(synth.pair_init(i, j)) for (i, j) in pairs]
    synth.write_log(os.path.join(d, "pairs.log"), np.array(Tinit), ids=[(i, j, n_frag) for (i, j) in pairs])
    outs = {}
    for tag, extra in (("one", []), ("many", ["--devices", ",".join(str(k) for k in range(min(nd, 4)))])):
        cwd = str(tmp_path / tag)
        os.makedirs(cwd)
        rc, out = run([os.path.join(BIN, "BuildCorrespondence"), "--reg_traj", os.path.join(d, "pairs.log"), "--registration",
                       "--reg_num", "5000", "--output_information"] + extra, cwd)
        assert rc == 0, out
        files = {f: open(os.path.join(cwd, f), "rb").read() for f in ("reg_output.log", "reg_output.info")}
        for (i, j) in pairs:
            fn = os.path.join(d, "corres_%d_%d.txt" % (i, j))
            files["corres_%d_%d" % (i, j)] = open(fn, "rb").read() if os.path.exists(fn) else None
            if os.path.exists(fn):
                os.remove(fn)
        outs[tag] = files
    assert outs["one"].keys() == outs["many"].keys()
    for k in outs["one"]:
        if k == "reg_output.info":          # double atomics across CTAs: last-digit differences in the %.8f text are possible
            a = np.array([float(x) for x in outs["one"][k].split()]); b = np.array([float(x) for x in outs["many"][k].split()])
            assert a.shape == b.shape and np.allclose(a, b, rtol=1e-9, atol=1e-6)
        else:
            assert outs["one"][k] == outs["many"][k], k


def test_integrate_cli_checkpoint_resume(oracle, tmp_path):
    """--save_volume after frames 1..3, --load_volume + --start_from 4 for the rest: the final world.pcd equals the
    uninterrupted run's (the reference has no checkpoint; its world.pcd is the only state it can write)."""
    sc = helpers.scene_room(n=6)
    d = str(tmp_path)
    synth.write_depth_container(os.path.join(d, "depth.raw"), sc["frames"])
    synth.write_log(os.path.join(d, "traj.log"), sc["traj"])
    base = [os.path.join(BIN, "Integrate"), "-oni", os.path.join(d, "depth.raw"), "--ref_traj", os.path.join(d, "traj.log")]
    rc, out = run(base + ["--save_to", os.path.join(d, "full.pcd")], d)
    assert rc == 0, out
    rc, out = run(base + ["--save_to", os.path.join(d, "a.pcd"), "--end_at", "3", "--save_volume", os.path.join(d, "vol.ck")], d)
    assert rc == 0 and os.path.getsize(os.path.join(d, "vol.ck")) > 1 << 20, out
    rc, out = run(base + ["--save_to", os.path.join(d, "b.pcd"), "--start_from", "4", "--load_volume", os.path.join(d, "vol.ck")], d)
    assert rc == 0, out
    full = oracle.sort_world(synth.read_pcd_xyzi(os.path.join(d, "full.pcd")))
    resumed = oracle.sort_world(synth.read_pcd_xyzi(os.path.join(d, "b.pcd")))
    assert len(full) > 1000 and np.array_equal(full.view(np.uint32), resumed.view(np.uint32))
