"""FragmentOptimizer path on the GPU (B200): the CUDA kernels of csrc/er_fragopt.cu through the C ABI against the numpy
oracle (which is pinned against the verbatim reference binary, tests/test_fragopt_cpu.py), and the drop-in executable
bin/FragmentOptimizer against the reference binary's pose.log / output.ctr (tests/golden/fragopt_golden.npz; live too
where oracle/_ref/FragmentOptimizer_ref travelled to the box).

(The file name sorts after test_tsdf_gpu.py / test_icp_gpu.py on purpose: the driver runs the suite with -x, and a failure in this
newest path must not keep the two headline paths' tests from running.  The three GPU sessions of round 2 ran it as
tests/test_fragopt_gpu.py -- that is the name in tools/gpu_r2_fragopt*.sh and profiles/r2_notes.md.)"""
import os
import subprocess

import numpy as np
import pytest

import helpers
from test_fragopt_cpu import CTR_TOL, POSE_TOL, accumulate_check, api_check, check_executable, nonrigid_api_check

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "elasticreconstruction_b200", "bin", "FragmentOptimizer")
GOLDEN = os.path.join(ROOT, "tests", "golden", "fragopt_golden.npz")


def test_kernels_against_the_oracle(tmp_path):
    from elasticreconstruction_b200 import capi
    api_check(lambda num: capi.FragOpt(num, 8, 3.0, device=0), tmp_path)


def test_nonrigid_kernels_against_the_oracle(tmp_path):
    from elasticreconstruction_b200 import capi
    nonrigid_api_check(lambda num, res: capi.FragOpt(num, res, 3.0, device=0), tmp_path)


@pytest.mark.parametrize("name,mode", helpers.FRAGOPT_CASE_MODES)
def test_executable_matches_reference_binary(oracle, tmp_path, name, mode):
    check_executable(EXE, oracle, np.load(GOLDEN), str(tmp_path), name, mode)


def _device_count():
    import ctypes
    try:
        n = ctypes.c_int(0)
        rt = ctypes.CDLL("libcudart.so.12")
        return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


def test_accumulate_on_one_device(tmp_path):
    """er_fragopt_slac_accumulate between two handles (here on the same device; across devices in the test below)."""
    from elasticreconstruction_b200 import capi
    accumulate_check(lambda num, k: capi.FragOpt(num, 8, 3.0, device=0), tmp_path)


@pytest.mark.parametrize("mode", ["slac", "rigid"])
def test_executable_on_two_devices(tmp_path, mode):
    if _device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from elasticreconstruction_b200 import capi
    accumulate_check(lambda num, k: capi.FragOpt(num, 8, 3.0, device=k % 2), tmp_path)
    golden = np.load(GOLDEN)
    d = str(tmp_path)
    _, args = helpers.fragopt_case("three", d)
    p = subprocess.run([EXE] + args + ["--" + mode, "--devices", "0,1"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stdout
    assert np.abs(helpers.read_log_poses(os.path.join(d, "pose.log")) - golden["three_%s_pose" % mode]).max() <= POSE_TOL
    assert np.abs(np.loadtxt(os.path.join(d, "output.ctr")) - golden["three_%s_ctr" % mode]).max() <= CTR_TOL


def test_sample_file_and_sorted_order_independence(tmp_path):
    """--write_xyzn_sample writes a readable binary_compressed PCD; two runs give the same poses to rounding (the flush
    atomics are the only order-dependent step)."""
    d = str(tmp_path)
    _, args = helpers.fragopt_case("three", d)
    outs = []
    for k in range(2):
        p = subprocess.run([EXE] + args + ["--slac", "--write_xyzn_sample", "50"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True, timeout=600)
        assert p.returncode == 0, p.stdout
        outs.append(helpers.read_log_poses(os.path.join(d, "pose.log")))
    assert np.abs(outs[0] - outs[1]).max() <= 2.5e-8          # one unit of the 8 printed decimals
    raw = open(os.path.join(d, "sample.pcd"), "rb").read()
    assert b"DATA binary_compressed\n" in raw
    n = int([l for l in raw.split(b"\n") if l.startswith(b"POINTS")][0].split()[1])
    body = raw[raw.index(b"DATA binary_compressed\n") + len(b"DATA binary_compressed\n"):]
    csize, usize = np.frombuffer(body[:8], dtype="<u4")
    assert usize == n * 32 and len(body) == 8 + csize and n > 100


def test_no_correspondence_rows_outside_the_clouds():
    from elasticreconstruction_b200 import capi
    g = capi.FragOpt(2, 8, 3.0, device=0)
    xyzn = np.array([[1, 1, 1, 0, 0, 1], [2, 2, 2, 0, 1, 0]], dtype=np.float32)
    g.set_cloud(0, xyzn)
    g.set_cloud(1, xyzn)
    with pytest.raises(capi.ElasticError, match="outside their clouds"):
        g.set_pairs([(0, 1)], [np.array([[0, 5]], dtype=np.int32)])
