"""The three drop-in executables chained through their FILES, on the CPU (no GPU):

    BuildCorrespondence  --reg_traj pairs.log --registration --save_xyzn     -> reg_output.log, corres_<i>_<j>.txt, cloud_bin_xyzn_<i>.xyzn
    FragmentOptimizer    --slac --registration reg_output.log --dir ...      -> pose.log, output.ctr
    Integrate            --pose_traj pose.log --ctr output.ctr --seg_traj .. -> world.pcd

apps/*.cpp are linked against the CPU stand-ins of the library (tests/emul/stub_elastic.cpp: oracle-backed;
tests/emul/fragopt_emul.cpp: the kernels' own arithmetic header), so this checks what the GPU tests cannot check here: that
each stage's output files are what the NEXT stage -- ours and the reference's -- reads.  The verbatim reference
FragmentOptimizer and Integrate binaries are run on the files our upstream stages wrote and must agree with our downstream
stages (lattices / poses to the printed digits, world.pcd bit for bit)."""
import os
import subprocess

import numpy as np
import pytest

import helpers
from elasticreconstruction_b200 import synth
from test_apps_host_logic import stub_apps  # noqa: F401  (fixture)
from test_fragopt_cpu import CTR_TOL, POSE_TOL, emul  # noqa: F401  (fixture)

N_FRAG = 3


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    return p.returncode, p.stdout


def test_buildcorrespondence_fragmentoptimizer_integrate_chain(oracle, stub_apps, emul, tmp_path):  # noqa: F811
    if not (os.path.exists(oracle.REF_FRAGOPT) and os.path.exists(oracle.REF_INTEGRATE)):
        pytest.skip("the reference binaries are not built (needs /root/reference)")
    data = str(tmp_path / "data") + os.sep
    os.makedirs(data)
    # ---- stage 1: BuildCorrespondence on PCD fragments
    frags = [synth.make_local_fragment(i, 5000) for i in range(N_FRAG)]
    for i, (x, n) in enumerate(frags):
        synth.write_pcd_xyzrgbnormal(data + "cloud_bin_%d.pcd" % i, x, n)
    pairs = [(i, j) for i in range(N_FRAG) for j in range(i + 1, N_FRAG)]
    synth.write_log(data + "pairs.log", np.array([synth.quantize_log(synth.pair_init(i, j)) for (i, j) in pairs]),
                    ids=[(i, j, N_FRAG) for (i, j) in pairs])
    s1 = str(tmp_path / "stage1")
    os.makedirs(s1)
    rc, out = run([stub_apps["BuildCorrespondence"], "--reg_traj", data + "pairs.log", "--registration", "--reg_num", "0", "--save_xyzn"], s1)
    assert rc == 0, out
    for (i, j) in pairs:
        assert os.path.getsize(data + "corres_%d_%d.txt" % (i, j)) > 1000
    for i in range(N_FRAG):
        assert os.path.exists(data + "cloud_bin_xyzn_%d.xyzn" % i)
    # ---- stage 2: FragmentOptimizer reads stage 1's files -- ours and the reference's, side by side
    traj = [np.linalg.inv(synth.BASEPOSE) @ synth.fragopt_initial_pose(i) @ synth.BASEPOSE for i in range(N_FRAG)]
    synth.write_log(data + "traj.log", traj)
    args2 = ["--slac", "--rgbdslam", data + "traj.log", "--registration", os.path.join(s1, "reg_output.log"), "--dir", data,
             "--num", str(N_FRAG), "--interval", "1", "--blacklistpair", "1000", "--iteration", "3"]
    s2, s2r = str(tmp_path / "stage2"), str(tmp_path / "stage2_ref")
    os.makedirs(s2)
    os.makedirs(s2r)
    rc, out = run([emul["exe"]] + args2, s2)
    assert rc == 0, out
    assert out.count("Read " + data + "corres_") == len(pairs)
    oracle.run_reference_fragopt(args2, cwd=s2r)
    pose, pose_r = helpers.read_log_poses(os.path.join(s2, "pose.log")), helpers.read_log_poses(os.path.join(s2r, "pose.log"))
    ctr, ctr_r = np.loadtxt(os.path.join(s2, "output.ctr")), np.loadtxt(os.path.join(s2r, "output.ctr"))
    assert pose.shape == (N_FRAG, 4, 4) and np.abs(pose - pose_r).max() <= POSE_TOL
    assert ctr.shape == (N_FRAG * 729, 3) and np.abs(ctr - ctr_r).max() <= CTR_TOL
    start = [np.linalg.norm(synth.fragopt_initial_pose(i) - synth.fragment_frame(i)) for i in range(1, N_FRAG)]
    end = [np.linalg.norm(pose[i] - synth.fragment_frame(i)) for i in range(1, N_FRAG)]
    assert max(end) < 0.5 * min(start)                       # the optimisation did its job on ICP-refined correspondences
    # ---- stage 3: Integrate --ctr reads stage 2's pose.log and output.ctr -- ours and the reference's, side by side
    interval = 3
    sc = helpers.scene_ctr(num=N_FRAG, interval=interval)
    s3 = str(tmp_path / "stage3")
    os.makedirs(s3)
    synth.write_depth_container(os.path.join(s3, "depth.raw"), sc["frames"])
    synth.write_log(os.path.join(s3, "seg.log"), sc["seg_traj"])
    common = ["-oni", os.path.join(s3, "depth.raw"), "--seg_traj", os.path.join(s3, "seg.log"), "--pose_traj", os.path.join(s2, "pose.log"),
              "--ctr", os.path.join(s2, "output.ctr"), "--num", str(N_FRAG), "--resolution", "8", "--length", "3.0", "--interval", str(interval)]
    rc, out = run([stub_apps["Integrate"]] + common + ["--save_to", os.path.join(s3, "mine.pcd")], s3)
    assert rc == 0, out
    oracle.run_reference_integrate(common + ["--save_to", os.path.join(s3, "ref.pcd")], cwd=s3)
    mine = oracle.sort_world(synth.read_pcd_xyzi(os.path.join(s3, "mine.pcd")))
    ref = oracle.sort_world(synth.read_pcd_xyzi(os.path.join(s3, "ref.pcd")))
    assert len(ref) > 50000 and mine.shape == ref.shape and np.array_equal(mine.view(np.uint32), ref.view(np.uint32))
