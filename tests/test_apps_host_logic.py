"""The two executables' HOST logic on the CPU (no GPU): apps/Integrate.cpp and apps/BuildCorrespondence.cpp are linked
against tests/emul/stub_elastic.cpp, a stand-in for libelastic_b200 that implements the C ABI on top of the oracle.
Checks argument handling, frame gating, batching, ControlGrid mode, the --devices sharding and merging and the file
formats -- Integrate's world.pcd against the VERBATIM reference binary's.  What the CUDA library computes is the
business of tests/test_apps_gpu.py."""
import os
import subprocess

import numpy as np
import pytest

import helpers
from elasticreconstruction_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "emul", "_build")


@pytest.fixture(scope="module")
def stub_apps(oracle):
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(ROOT, "tests", "emul", "stub_elastic.cpp"), os.path.join(ROOT, "elasticreconstruction_b200", "csrc", "er_host_math.cpp")]
    apps = {}
    for name in ("Integrate", "BuildCorrespondence"):
        exe = os.path.join(OUT, name + "_stub")
        app_src = os.path.join(ROOT, "apps", name + ".cpp")
        deps = srcs + [app_src, os.path.join(ROOT, "apps", "er_io.h"), oracle.LIB_PATH]
        if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(p) for p in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                                   "-I", os.path.join(ROOT, "apps"), "-I", os.path.join(ROOT, "elasticreconstruction_b200", "csrc"),
                                   "-I", "/usr/local/cuda/include", app_src] + srcs +
                                  ["-L", os.path.dirname(oracle.LIB_PATH), "-loracle", "-Wl,-rpath," + os.path.dirname(oracle.LIB_PATH),
                                   "-o", exe])
        apps[name] = exe
    return apps


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    return p.returncode, p.stdout


def world(oracle, path):
    return oracle.sort_world(synth.read_pcd_xyzi(path))


@pytest.mark.parametrize("mk", helpers.ALL_SCENES, ids=lambda f: f.__name__)
@pytest.mark.parametrize("devices", [None, "0,1,2"])
def test_integrate_host_logic_matches_reference_binary(oracle, stub_apps, tmp_path, mk, devices):
    if not os.path.exists(oracle.REF_INTEGRATE):
        pytest.skip("oracle/_ref/Integrate_ref not built")
    sc = mk()
    d = str(tmp_path)
    args = helpers.write_scene(sc, d)
    mine_pcd = os.path.join(d, "mine.pcd")
    my_args = [a if a != os.path.join(d, "world.pcd") else mine_pcd for a in args]
    if devices:
        my_args += ["--devices", devices]
    rc, out = run([stub_apps["Integrate"]] + my_args, d)
    assert rc == 0, out
    ref_out, _ = oracle.run_reference_integrate(args, cwd=d)
    ref = world(oracle, os.path.join(d, "world.pcd"))
    mine = world(oracle, mine_pcd)
    line = [l for l in ref_out.splitlines() if l.startswith("Total ")]
    assert line and line[0] in out
    assert ("%d voxel points have been written." % len(ref)) in out
    assert mine.shape == ref.shape and np.array_equal(mine.view(np.uint32), ref.view(np.uint32))


def test_integrate_host_logic_window_flags_and_voxel_size(oracle, stub_apps, tmp_path):
    sc = helpers.scene_room(n=6)
    d = str(tmp_path)
    args = helpers.write_scene(sc, d) + ["--start_from", "2", "--end_at", "4"]
    rc, out = run([stub_apps["Integrate"]] + args, d)
    assert rc == 0 and "Total 5 frames processed." in out
    vol = oracle.OracleTSDF()
    helpers.run_app(vol, sc["frames"], sc["traj"], start_from=2, end_at=4)
    assert np.array_equal(world(oracle, os.path.join(d, "world.pcd")).view(np.uint32), oracle.sort_world(vol.extract()).view(np.uint32))
    # frame_ == -1 records are skipped (IntegrateApp.cpp:196)
    ids = [(i, i, -1 if i in (1, 3) else i + 1) for i in range(len(sc["traj"]))]
    synth.write_log(os.path.join(d, "traj.log"), sc["traj"], ids=ids)
    rc, out = run([stub_apps["Integrate"], "-oni", os.path.join(d, "depth.raw"), "--ref_traj", os.path.join(d, "traj.log"),
                   "--save_to", os.path.join(d, "flag.pcd"), "--devices", "0,1"], d)
    assert rc == 0, out
    vol = oracle.OracleTSDF()
    helpers.run_app(vol, sc["frames"], sc["traj"], frame_flags=[i[2] for i in ids])
    assert np.array_equal(world(oracle, os.path.join(d, "flag.pcd")).view(np.uint32), oracle.sort_world(vol.extract()).view(np.uint32))
    # --voxel_size against the reference built with its constant rewritten
    if os.path.exists(oracle.REF_INTEGRATE_12MM):
        sw = helpers.scene_wall(n=4)
        d2 = os.path.join(d, "w")
        os.makedirs(d2)
        a2 = helpers.write_scene(sw, d2)
        oracle.run_reference_integrate(a2, cwd=d2, binary=oracle.REF_INTEGRATE_12MM)
        ref = world(oracle, os.path.join(d2, "world.pcd"))
        rc, out = run([stub_apps["Integrate"]] + [a if not a.endswith("world.pcd") else os.path.join(d2, "mine.pcd") for a in a2] +
                      ["--voxel_size", "0.012"], d2)
        assert rc == 0, out
        assert np.array_equal(world(oracle, os.path.join(d2, "mine.pcd")).view(np.uint32), ref.view(np.uint32))


def test_integrate_help_and_bad_source(stub_apps, tmp_path):
    rc, out = run([stub_apps["Integrate"]], str(tmp_path))
    assert rc == 0 and "--ref_traj <log_file>" in out and "--devices" in out
    rc, out = run([stub_apps["Integrate"], "-oni", "missing.raw"], str(tmp_path))
    assert rc != 0 and "Can't open depth source" in out


def test_buildcorrespondence_host_logic_single_and_multi_device(oracle, stub_apps, tmp_path):
    """Same files from one device and from three: the pair list is dealt round-robin and merged back in place.  The
    single-device output is checked against the Python restatement of CCorresApp's control flow."""
    import test_apps_gpu as gpu_tests                      # parse_log / emulate_corres_app (no GPU needed to import)
    d = str(tmp_path / "data")
    os.makedirs(d)
    n_frag = 4
    frags = [synth.make_local_fragment(i, 6000) for i in range(n_frag)]
    for i, (x, n) in enumerate(frags):
        synth.write_pcd_xyzrgbnormal(os.path.join(d, "cloud_bin_%d.pcd" % i), x, n)
    pairs = [(i, j) for i in range(n_frag) for j in range(i + 1, n_frag)]
    Tinit = [synth.quantize_log(synth.pair_init(i, j)) for (i, j) in pairs]
    Tinit[-1] = synth.quantize_log(Tinit[-1] @ synth.make_pose(np.eye(3), [2.0, 0, 0]))      # hopeless pair -> rejected
    synth.write_log(os.path.join(d, "pairs.log"), np.array(Tinit), ids=[(i, j, n_frag) for (i, j) in pairs])
    outs = {}
    for tag, extra in (("one", []), ("three", ["--devices", "0,1,2"])):
        cwd = str(tmp_path / tag)
        os.makedirs(cwd)
        rc, out = run([stub_apps["BuildCorrespondence"], "--reg_traj", os.path.join(d, "pairs.log"), "--registration", "--reg_num", "1000",
                       "--output_information"] + extra, cwd)
        assert rc == 0, out
        outs[tag] = (open(os.path.join(cwd, "reg_output.log")).read(), open(os.path.join(cwd, "reg_output.info")).read(),
                     {f: open(os.path.join(d, f)).read() for f in sorted(os.listdir(d)) if f.startswith("corres_")})
    assert outs["one"] == outs["three"]
    clouds = [oracle.OracleCloud(x, n, 0.03) for (x, n) in frags]
    recs = gpu_tests.emulate_corres_app(oracle, clouds, pairs, Tinit, n_frag, reg_num=1000)
    got = gpu_tests.parse_log(str(tmp_path / "one" / "reg_output.log"))
    assert len(got) == len(recs) and sum(r["frame"] == -1 for r in recs) >= 1 and sum(r["frame"] > 0 for r in recs) >= 3
    for r, (ids, T) in zip(recs, got):
        assert ids[:2] == [r["i"], r["j"]] and ids[2] == r["frame"]
        assert np.linalg.norm(T - r["T"]) <= 4e-8              # same oracle underneath: only the %.8f text rounding
        fn = os.path.join(d, "corres_%d_%d.txt" % (r["i"], r["j"]))
        if r["corr"] is not None:
            assert np.array_equal(np.loadtxt(fn, dtype=np.int64).reshape(-1, 2), r["corr"])
        else:
            assert not os.path.exists(fn)
