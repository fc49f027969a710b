"""The three drop-in executables chained through their files ON THE GPU (sorted last on purpose: it was written after the
round's GPU minutes were spent and has only run in its CPU form, tests/test_pipeline_chain.py).  bin/BuildCorrespondence
-> bin/FragmentOptimizer -> bin/Integrate --ctr, the verbatim reference FragmentOptimizer / Integrate binaries run on the
same intermediate files where oracle/_ref travelled to the box."""
import os
import subprocess

import numpy as np
import pytest

import helpers
from elasticreconstruction_b200 import synth
from test_fragopt_cpu import CTR_TOL, POSE_TOL

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "elasticreconstruction_b200", "bin")
N_FRAG = 3


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    return p.returncode, p.stdout


def test_chain_on_the_gpu(oracle, tmp_path):
    data = str(tmp_path / "data") + os.sep
    os.makedirs(data)
    frags = [synth.make_local_fragment(i, 20000) for i in range(N_FRAG)]
    for i, (x, n) in enumerate(frags):
        synth.write_pcd_xyzrgbnormal(data + "cloud_bin_%d.pcd" % i, x, n)
    pairs = [(i, j) for i in range(N_FRAG) for j in range(i + 1, N_FRAG)]
    synth.write_log(data + "pairs.log", np.array([synth.quantize_log(synth.pair_init(i, j)) for (i, j) in pairs]),
                    ids=[(i, j, N_FRAG) for (i, j) in pairs])
    s1 = str(tmp_path / "stage1")
    os.makedirs(s1)
    rc, out = run([os.path.join(BIN, "BuildCorrespondence"), "--reg_traj", data + "pairs.log", "--registration", "--reg_num", "0", "--save_xyzn"], s1)
    assert rc == 0, out
    traj = [np.linalg.inv(synth.BASEPOSE) @ synth.fragopt_initial_pose(i) @ synth.BASEPOSE for i in range(N_FRAG)]
    synth.write_log(data + "traj.log", traj)
    args2 = ["--slac", "--rgbdslam", data + "traj.log", "--registration", os.path.join(s1, "reg_output.log"), "--dir", data,
             "--num", str(N_FRAG), "--interval", "1", "--blacklistpair", "1000", "--iteration", "3"]
    s2 = str(tmp_path / "stage2")
    os.makedirs(s2)
    rc, out = run([os.path.join(BIN, "FragmentOptimizer")] + args2, s2)
    assert rc == 0, out
    pose = helpers.read_log_poses(os.path.join(s2, "pose.log"))
    ctr = np.loadtxt(os.path.join(s2, "output.ctr"))
    assert pose.shape == (N_FRAG, 4, 4) and ctr.shape == (N_FRAG * 729, 3)
    start = [np.linalg.norm(synth.fragopt_initial_pose(i) - synth.fragment_frame(i)) for i in range(1, N_FRAG)]
    end = [np.linalg.norm(pose[i] - synth.fragment_frame(i)) for i in range(1, N_FRAG)]
    assert max(end) < 0.5 * min(start)
    if os.path.exists(oracle.REF_FRAGOPT):
        s2r = str(tmp_path / "stage2_ref")
        os.makedirs(s2r)
        oracle.run_reference_fragopt(args2, cwd=s2r)
        assert np.abs(pose - helpers.read_log_poses(os.path.join(s2r, "pose.log"))).max() <= POSE_TOL
        assert np.abs(ctr - np.loadtxt(os.path.join(s2r, "output.ctr"))).max() <= CTR_TOL
    interval = 3
    sc = helpers.scene_ctr(num=N_FRAG, interval=interval)
    s3 = str(tmp_path / "stage3")
    os.makedirs(s3)
    synth.write_depth_container(os.path.join(s3, "depth.raw"), sc["frames"])
    synth.write_log(os.path.join(s3, "seg.log"), sc["seg_traj"])
    common = ["-oni", os.path.join(s3, "depth.raw"), "--seg_traj", os.path.join(s3, "seg.log"), "--pose_traj", os.path.join(s2, "pose.log"),
              "--ctr", os.path.join(s2, "output.ctr"), "--num", str(N_FRAG), "--resolution", "8", "--length", "3.0", "--interval", str(interval)]
    rc, out = run([os.path.join(BIN, "Integrate")] + common + ["--save_to", os.path.join(s3, "mine.pcd")], s3)
    assert rc == 0, out
    mine = oracle.sort_world(synth.read_pcd_xyzi(os.path.join(s3, "mine.pcd")))
    assert len(mine) > 50000
    if os.path.exists(oracle.REF_INTEGRATE):
        oracle.run_reference_integrate(common + ["--save_to", os.path.join(s3, "ref.pcd")], cwd=s3)
        ref = oracle.sort_world(synth.read_pcd_xyzi(os.path.join(s3, "ref.pcd")))
        assert mine.shape == ref.shape and np.array_equal(mine[:, :3], ref[:, :3]) and np.abs(mine[:, 3] - ref[:, 3]).max() <= 1e-4
