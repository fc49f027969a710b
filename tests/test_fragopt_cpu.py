"""FragmentOptimizer path on the CPU (no GPU):

* the numpy restatement oracle/fragopt_oracle.py against the VERBATIM reference binary's results
  (tests/golden/fragopt_golden.npz, written by tests/golden/make_golden_fragopt.py; and live where oracle/_ref exists);
* the product's kernel arithmetic (csrc/er_fragopt_math.h) re-enacted thread by thread by tests/emul/fragopt_emul.cpp:
  point set-up and updates bit for bit against the oracle, the assembled SLAC / rigid normal equations within rounding
  of the oracle's, independent of how the correspondences are split over CTAs;
* apps/FragmentOptimizer.cpp linked against that emulation: pose.log / output.ctr against the reference binary's.

Tolerances: the reference prints poses with 8 and lattices with 10 decimals, and sums its per-pair matrices in arrival
order, so two runs of the reference itself differ at rounding level; 2e-8 / 2e-9 absolute is the print resolution plus
that noise.  (north_star's bar for a pose is 1e-5 Frobenius.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "elasticreconstruction_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "emul", "_build")
GOLDEN = os.path.join(ROOT, "tests", "golden", "fragopt_golden.npz")
POSE_TOL, CTR_TOL = 2e-8, 2e-9


def _stale(target, deps):
    return not os.path.exists(target) or os.path.getmtime(target) < max(os.path.getmtime(p) for p in deps)


@pytest.fixture(scope="module")
def emul():
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libfragopt_emul.so")
    srcs = [os.path.join(ROOT, "tests", "emul", "fragopt_emul.cpp"), os.path.join(CSRC, "er_host_math.cpp")]
    deps = srcs + [os.path.join(CSRC, "er_fragopt_math.h"), os.path.join(ROOT, "include", "elastic_b200.h")]
    if _stale(lib, deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                               "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-I", "/usr/local/cuda/include"] + srcs + ["-o", lib])
    exe = os.path.join(OUT, "FragmentOptimizer_emul")
    app = os.path.join(ROOT, "apps", "FragmentOptimizer.cpp")
    if _stale(exe, [app, os.path.join(ROOT, "apps", "er_io.h"), lib]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.join(ROOT, "apps"), app, "-o", exe, "-L", OUT, "-lfragopt_emul", "-Wl,-rpath," + OUT])
    return {"lib": C.CDLL(lib), "exe": exe}


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def case_arrays(name, d):
    import fragopt_oracle as fo
    c = helpers.FRAGOPT_CASES[name]
    keep = [i for i in range(c["num"]) if not (c["blacklist"] and i in c["blacklist"])]
    rel = {a: r for r, a in enumerate(keep)}
    xyzn = [np.concatenate([d["clouds"][i][0], d["clouds"][i][1]], axis=1) for i in keep]
    # InitCorrespondences: blacklisted ids and ids >= the REDUCED fragment count are dropped (OptApp.cpp:104-106)
    pairs = [(i, j) for (i, j) in d["pairs"] if i in rel and j in rel and i < len(keep) and j < len(keep) and len(d["corres"][(i, j)]) >= 1000]
    corres = [d["corres"][p] for p in pairs]
    ipose = fo.initial_poses(d["traj"], keep, 1, 3.0)
    return xyzn, [(rel[i], rel[j]) for (i, j) in pairs], corres, ipose, c


def flag(c, name, default):
    return type(default)(c["flags"][c["flags"].index(name) + 1]) if name in c["flags"] else default


@pytest.mark.parametrize("name,mode", helpers.FRAGOPT_CASE_MODES)
def test_numpy_oracle_matches_reference_binary(golden, tmp_path, name, mode):
    import fragopt_oracle as fo
    d, _ = helpers.fragopt_case(name, str(tmp_path))
    xyzn, pairs, corres, ipose, c = case_arrays(name, d)
    it = flag(c, "--iteration", 5)
    if mode == "slac":
        pose, ctr = fo.optimize_slac(xyzn, pairs, corres, ipose, iterations=it, weight=flag(c, "--weight", 1.0), num_before_blacklist=c["num"])
    elif mode == "rigid":
        pose, ctr = fo.optimize_rigid(xyzn, pairs, corres, ipose, iterations=it)
    else:
        pose = None
        ctr = fo.optimize_nonrigid(xyzn, pairs, corres, ipose, resolution=flag(c, "--resolution", 8), weight=flag(c, "--weight", 1.0),
                                   iterations=it, inner_iterations=flag(c, "--inner_iteration", 10))
    if pose is not None:
        assert np.abs(pose - golden["%s_%s_pose" % (name, mode)]).max() <= POSE_TOL
    assert np.abs(ctr - golden["%s_%s_ctr" % (name, mode)]).max() <= CTR_TOL


def check_executable(exe, oracle, golden, d, name, mode):
    """Shared with tests/test_z_fragopt_gpu.py: run `exe` on a case and compare its files with the reference binary's."""
    _, args = helpers.fragopt_case(name, d)
    p = subprocess.run([exe] + args + ["--" + mode, "--bench_json", os.path.join(d, "t.json")], cwd=d, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0, p.stdout
    pose = helpers.read_log_poses(os.path.join(d, "pose.log")) if mode != "nonrigid" else None
    ctr = np.loadtxt(os.path.join(d, "output.ctr"))
    if pose is not None:
        assert np.abs(pose - golden["%s_%s_pose" % (name, mode)]).max() <= POSE_TOL
    assert np.abs(ctr - golden["%s_%s_ctr" % (name, mode)]).max() <= CTR_TOL
    scores = [float(l.split(" is ")[1].lstrip(": ")) for l in p.stdout.splitlines() if "score is" in l]
    assert np.allclose(scores, golden["%s_%s_scores" % (name, mode)], atol=0.0101 if mode != "nonrigid" else 0.000101)   # printed digits
    if os.path.exists(oracle.REF_FRAGOPT):                                                   # and live, where the reference is built
        if pose is not None:
            os.rename(os.path.join(d, "pose.log"), os.path.join(d, "mine_pose.log"))
        oracle.run_reference_fragopt(args + ["--" + mode], cwd=d)
        if pose is not None:
            assert np.abs(pose - helpers.read_log_poses(os.path.join(d, "pose.log"))).max() <= POSE_TOL
        assert np.abs(ctr - np.loadtxt(os.path.join(d, "output.ctr"))).max() <= CTR_TOL


@pytest.mark.parametrize("name,mode", helpers.FRAGOPT_CASE_MODES)
def test_emulated_executable_matches_reference_binary(oracle, emul, golden, tmp_path, name, mode):
    check_executable(emul["exe"], oracle, golden, str(tmp_path), name, mode)


def test_emulated_kernels_against_the_oracle(emul, tmp_path):
    """API level: the emulated kernels (same header as the CUDA ones) against the numpy restatement."""
    from elasticreconstruction_b200 import capi
    api_check(lambda num: capi.FragOpt(num, 8, 3.0, library=emul["lib"]), tmp_path, vary_emulated_ctas=True)


def api_check(make, tmp_path, vary_emulated_ctas=False):
    """Shared with tests/test_z_fragopt_gpu.py: `make(num)` returns a capi.FragOpt on the library under test."""
    import fragopt_oracle as fo
    d, _ = helpers.fragopt_case("three", str(tmp_path))
    xyzn, pairs, corres, ipose, _ = case_arrays("three", d)
    num = len(xyzn)
    L = fo.Lattice(8, 3.0)
    M = 6 * num + L.nper
    clouds = [fo.Cloud(L, c) for c in xyzn]
    g = make(num)
    for l in range(num):
        assert g.set_cloud(l, xyzn[l]) == len(xyzn[l])
    g.set_pairs(pairs, corres)
    # loading: positions and the divided normal
    for l in range(num):
        pts = g.points(l)
        assert np.array_equal(pts[:, :3], clouds[l].p) and np.array_equal(pts[:, 3:], clouds[l].n)
    # UpdatePose, bit for bit
    g.update_pose(ipose)
    for l in range(num):
        clouds[l].update_pose(ipose[l])
        pts = g.points(l)
        assert np.array_equal(pts[:, :3].view(np.uint32), clouds[l].p.view(np.uint32))
        assert np.array_equal(pts[:, 3:].view(np.uint32), clouds[l].n.view(np.uint32))
    # SLAC normal equations
    rot_t = np.array([ipose[l][:3, :3].T for l in range(num)])
    JJ, Jb, score = fo.slac_normal_equations(clouds, pairs, corres, num, rot_t, M)
    gJb, gscore = g.slac_assemble(rot_t)
    gJJ = g.slac_jj()
    scale = np.abs(JJ).max()
    assert np.array_equal(gJJ, np.triu(gJJ))
    assert np.abs(gJJ - JJ).max() <= 1e-11 * scale
    assert np.abs(gJb - Jb).max() <= 1e-11 * np.abs(Jb).max() and abs(gscore - score) <= 1e-11 * score
    _, rows, flushes = g.assemble_stats()
    assert rows == sum(len(c) for c in corres) and 0 < flushes < rows / 4          # the sorted order keeps segments long
    # the split of the sorted list over CTAs does not matter
    for ctas in (("1", "3", "64") if vary_emulated_ctas else ()):
        os.environ["ER_FRAGOPT_EMUL_CTAS"] = ctas
        try:
            g2 = make(num)
        finally:
            del os.environ["ER_FRAGOPT_EMUL_CTAS"]
        for l in range(num):
            g2.set_cloud(l, xyzn[l])
        g2.set_pairs(pairs, corres)
        g2.update_pose(ipose)
        Jb2, s2 = g2.slac_assemble(rot_t)
        assert np.abs(g2.slac_jj() - JJ).max() <= 1e-11 * scale and np.abs(Jb2 - Jb).max() <= 1e-11 * np.abs(Jb).max()
    # solve against numpy
    B = fo.base_jj(L, num, M, num * 1.0)
    rhs = Jb + 0.01 * np.cos(np.arange(M))
    A = B + JJ
    x = g.slac_solve(B, rhs)
    assert np.abs(x - np.linalg.solve(A + np.triu(A, 1).T, rhs)).max() <= 1e-9 * max(1.0, np.abs(x).max())
    # UpdateAllPointPN under a warped lattice, bit for bit
    lat = np.array([[i * L.unit_d, j * L.unit_d, k * L.unit_d] for k in range(L.r1) for j in range(L.r1) for i in range(L.r1)])
    lat = lat + 0.004 * np.sin(3.0 * lat[:, [1, 2, 0]])
    ex = np.concatenate([lat @ ipose[l][:3, :3].T + ipose[l][:3, 3] for l in range(num)]).reshape(-1)
    g.update_points_pn(ex)
    for l in range(num):
        clouds[l].update_point_pn(ex[l * L.nper:(l + 1) * L.nper])
        pts = g.points(l)
        assert np.array_equal(pts[:, :3].view(np.uint32), clouds[l].p.view(np.uint32))
        assert np.array_equal(pts[:, 3:].view(np.uint32), clouds[l].n.view(np.uint32))
        warped = g.points(l, ctr=ex)[:, :3]
        assert np.array_equal(warped, clouds[l].warped(ex[l * L.nper:(l + 1) * L.nper]).astype(np.float32))
    # rigid normal equations on the updated points
    rJJ, rJb, rscore = fo.rigid_normal_equations(clouds, pairs, corres, num)
    qJJ, qJb, qscore = g.rigid_assemble()
    assert np.abs(qJJ - rJJ).max() <= 1e-11 * np.abs(rJJ).max() and np.abs(qJb - rJb).max() <= 1e-11 * np.abs(rJb).max()
    assert abs(qscore - rscore) <= 1e-11 * rscore


def nonrigid_api_check(make, tmp_path):
    """Shared with the GPU tests: the nonrigid data matrix, factorisation and solves against the oracle."""
    import fragopt_oracle as fo
    d, _ = helpers.fragopt_case("three_res4", str(tmp_path))
    xyzn, pairs, corres, ipose, _ = case_arrays("three_res4", d)
    # a pair listed with idx0 > idx1 keeps only its diagonal blocks (its cross block lies in the lower triangle the solver never reads)
    pairs = [pairs[0], (pairs[1][1], pairs[1][0]), pairs[2]]
    corres = [corres[0], corres[1][:, ::-1].copy(), corres[2]]
    num, res, weight = len(xyzn), 4, 2.0
    L = fo.Lattice(res, 3.0)
    N = num * L.nper
    clouds = [fo.Cloud(L, c) for c in xyzn]
    g = make(num, res)
    for l in range(num):
        g.set_cloud(l, xyzn[l])
    g.set_pairs(pairs, corres)
    lat = np.array([[i * L.unit_d, j * L.unit_d, k * L.unit_d] for k in range(L.r1) for j in range(L.r1) for i in range(L.r1)])
    ctr = np.concatenate([lat @ ipose[l][:3, :3].T + ipose[l][:3, 3] for l in range(num)]).reshape(-1)
    g.update_normals(ctr)
    for l in range(num):
        clouds[l].update_normal(ctr[l * L.nper:(l + 1) * L.nper])
        assert np.array_equal(g.points(l)[:, 3:].view(np.uint32), clouds[l].n.view(np.uint32))
    A = fo.nonrigid_data_matrix(clouds, pairs, corres, L.nper, weight)
    g.nonrigid_assemble(weight)
    gA = g.nonrigid_aa()
    assert gA.shape == (N, N) and np.array_equal(gA, np.triu(gA))
    assert np.abs(gA - A).max() <= 1e-11 * np.abs(A).max()
    i1, j1 = pairs[1]
    assert np.all(gA[j1 * L.nper:(j1 + 1) * L.nper, i1 * L.nper:(i1 + 1) * L.nper] == 0)          # the dropped cross block
    B = fo.base_aa(L, num)
    r, c = np.nonzero(B)
    g.nonrigid_factor(r, c, B[r, c])
    full = A + B
    full = full + np.triu(full, 1).T
    for k in range(2):
        rhs = np.cos(0.37 * (k + 1) * np.arange(N))
        x = g.nonrigid_solve(rhs)
        assert np.abs(x - np.linalg.solve(full, rhs)).max() <= 1e-8 * max(1.0, np.abs(x).max())


def test_emulated_nonrigid_kernels_against_the_oracle(emul, tmp_path):
    from elasticreconstruction_b200 import capi
    nonrigid_api_check(lambda num, res: capi.FragOpt(num, res, 3.0, library=emul["lib"]), tmp_path)


def test_out_of_bound_point_stops_loading(emul):
    """PointCloud.cpp:37-40: the first point outside [0, length)^3 ends the cloud (it stays as a zero-filled record)."""
    from elasticreconstruction_b200 import capi
    g = capi.FragOpt(1, 8, 3.0, library=emul["lib"])
    xyzn = np.array([[1, 1, 1, 0, 0, 1], [2, 2, 2, 0, 1, 0], [3.0, 1, 1, 1, 0, 0], [1, 2, 1, 1, 0, 0]], dtype=np.float32)
    assert g.set_cloud(0, xyzn) == 3
    xyzn[2, 0] = -0.001
    g2 = capi.FragOpt(1, 8, 3.0, library=emul["lib"])
    assert g2.set_cloud(0, xyzn) == 3
    xyzn[2, 0] = 2.999
    g3 = capi.FragOpt(1, 8, 3.0, library=emul["lib"])
    assert g3.set_cloud(0, xyzn) == 4


def test_executable_help_and_missing_inputs(emul, tmp_path):
    p = subprocess.run([emul["exe"]], stdout=subprocess.PIPE, text=True)
    assert p.returncode == 0 and "--blasklist <blacklist_file>" in p.stdout and "--slac" in p.stdout
    p = subprocess.run([emul["exe"], "--num", "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=str(tmp_path))
    assert p.returncode == 1 and "trajectory is empty" in p.stdout          # the default mode is nonrigid; nothing to optimise without inputs


def test_bench_loop_on_the_emulated_library_matches_reference_binary(emul, golden, tmp_path):
    """fragopt_bench.SlacLoop (the loop bench.py times, with its vectorised regulariser) over the emulated library."""
    import importlib.util
    from elasticreconstruction_b200 import capi
    spec = importlib.util.spec_from_file_location("fragopt_bench", os.path.join(ROOT, "fragopt_bench.py"))
    fb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fb)
    d, _ = helpers.fragopt_case("three", str(tmp_path))
    xyzn, pairs, corres, ipose, _ = case_arrays("three", d)
    g = capi.FragOpt(len(xyzn), 8, 3.0, library=emul["lib"])
    for l in range(len(xyzn)):
        g.set_cloud(l, xyzn[l])
    g.set_pairs(pairs, corres)
    loop = fb.SlacLoop(g, ipose)
    for _ in range(2):
        loop.step()
    assert np.abs(loop.pose - golden["three_slac_pose"]).max() <= POSE_TOL
    assert np.abs(loop.expand().reshape(-1, 3) - golden["three_slac_ctr"]).max() <= CTR_TOL


@pytest.mark.parametrize("mode", ["slac", "rigid"])
def test_emulated_executable_with_three_devices(oracle, emul, golden, tmp_path, mode):
    """--devices: every "device" holds all fragments and a share of the pairs; the sums are added in device order."""
    d = str(tmp_path)
    _, args = helpers.fragopt_case("three", d)
    p = subprocess.run([emul["exe"]] + args + ["--" + mode, "--devices", "0,1,2", "--bench_json", os.path.join(d, "t.json")], cwd=d,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0, p.stdout
    assert '"devices": 3' in open(os.path.join(d, "t.json")).read()
    assert np.abs(helpers.read_log_poses(os.path.join(d, "pose.log")) - golden["three_%s_pose" % mode]).max() <= POSE_TOL
    assert np.abs(np.loadtxt(os.path.join(d, "output.ctr")) - golden["three_%s_ctr" % mode]).max() <= CTR_TOL


def accumulate_check(make, tmp_path):
    """Shared with the GPU tests: two handles with half of the pairs each, accumulated, equal one handle with all pairs."""
    d, _ = helpers.fragopt_case("three", str(tmp_path))
    xyzn, pairs, corres, ipose, _ = case_arrays("three", d)
    num = len(xyzn)
    rot_t = np.array([ipose[l][:3, :3].T for l in range(num)])
    hs = [make(num, k) for k in range(3)]
    shares = [(pairs, corres), (pairs[:1], corres[:1]), (pairs[1:], corres[1:])]
    out = []
    for g, (pp, cc) in zip(hs, shares):
        for l in range(num):
            g.set_cloud(l, xyzn[l])
        g.set_pairs(pp, cc)
        g.update_pose(ipose)
        out.append(g.slac_assemble(rot_t))
    full = hs[0].slac_jj()
    hs[1].slac_accumulate(hs[2])
    both = hs[1].slac_jj()
    assert np.abs(both - full).max() <= 1e-11 * np.abs(full).max()
    assert np.abs(out[1][0] + out[2][0] - out[0][0]).max() <= 1e-11 * np.abs(out[0][0]).max()
    assert abs(out[1][1] + out[2][1] - out[0][1]) <= 1e-11 * out[0][1]


def test_emulated_accumulate(emul, tmp_path):
    from elasticreconstruction_b200 import capi
    accumulate_check(lambda num, k: capi.FragOpt(num, 8, 3.0, library=emul["lib"]), tmp_path)


def _read_sample_pcd(path):
    """x y z nx ny nz rows of a sample.pcd: `DATA binary` (the reference build's stand-in writer, 32-byte records) or
    `DATA binary_compressed` (the product: fields one after the other, LZF -- literal runs are all its writer emits)."""
    raw = open(path, "rb").read()
    head, _, body = raw.partition(b"DATA ")
    kind, _, body = body.partition(b"\n")
    n = int([l for l in head.split(b"\n") if l.startswith(b"POINTS")][0].split()[1])
    if kind.strip() == b"binary":
        a = np.frombuffer(body[: 32 * n], dtype="<f4").reshape(n, 8)
        return a[:, [0, 1, 2, 4, 5, 6]]
    csize, usize = np.frombuffer(body[:8], dtype="<u4")
    data, out, p = body[8: 8 + csize], bytearray(), 0
    while p < len(data):
        ctrl = data[p]
        assert ctrl < 32, "back-references are not expected from the product's writer"
        out += data[p + 1: p + 2 + ctrl]
        p += 2 + ctrl
    assert len(out) == usize == 32 * n
    a = np.frombuffer(bytes(out), dtype="<f4").reshape(8, n).T
    return a[:, [0, 1, 2, 4, 5, 6]]


FRAGOPT_VARIANTS = {
    # name: (case, mode, extra flags, files to compare besides output.ctr)
    "slac_sample": ("three", "slac", ["--write_xyzn_sample", "7"], ["sample.pcd"]),
    "rigid_sample_save_to": ("three", "rigid", ["--write_xyzn_sample", "11", "--save_to", "rigid.ctr"], ["sample.pcd"]),
    "nonrigid_intermediate_files": ("three_res4", "nonrigid", ["--write_xyzn_sample", "9"],
                                    ["itr0.ctr", "itr1.ctr", "itr0_inner0_out.ctr", "itr1_inner2_out.ctr", "sample.pcd"]),
    "slac_init_ctr": ("three", "slac", ["--init_ctr", "INIT_CTR"], []),
    "nonrigid_ipose": ("three_res4", "nonrigid", ["--ipose", "IPOSE"], []),
    "slac_resolution5_length": ("three", "slac", ["--resolution", "5", "--length", "3.2", "--weight", "3"], []),
    "default_mode_is_nonrigid": ("three_res4", "", [], []),
}


@pytest.mark.parametrize("variant", list(FRAGOPT_VARIANTS))
def test_emulated_executable_flag_variants_against_live_reference(oracle, emul, tmp_path, variant):
    """Flags beyond the golden cases, the emulated executable against the reference binary run side by side."""
    if not os.path.exists(oracle.REF_FRAGOPT):
        pytest.skip("oracle/_ref/FragmentOptimizer_ref not built (needs /root/reference)")
    case, mode, extra, files = FRAGOPT_VARIANTS[variant]
    d = str(tmp_path)
    data, args = helpers.fragopt_case(case, d)
    c = helpers.FRAGOPT_CASES[case]
    res = flag(c, "--resolution", 8)
    if "INIT_CTR" in extra:                                   # a gently bent lattice (nper / 3 rows)
        r1 = res + 1
        lat = np.array([[i, j, k] for k in range(r1) for j in range(r1) for i in range(r1)], dtype=np.float64) * (3.0 / res)
        lat += 0.003 * np.sin(2.0 * lat[:, [2, 0, 1]])
        np.savetxt(os.path.join(d, "init.ctr"), lat, fmt="%.10f")
        extra = [os.path.join(d, "init.ctr") if x == "INIT_CTR" else x for x in extra]
    if "IPOSE" in extra:
        from elasticreconstruction_b200 import synth
        poses = [synth.fragment_frame(i) @ synth.perturbation(7000 + i, 0.2, 0.003) for i in range(c["num"])]
        synth.write_log(os.path.join(d, "ipose.log"), poses)
        extra = [os.path.join(d, "ipose.log") if x == "IPOSE" else x for x in extra]
    argv = args + extra + (["--" + mode] if mode else [])
    ctr_name = extra[extra.index("--save_to") + 1] if "--save_to" in extra else "output.ctr"
    out = {}
    for who, run in (("mine", lambda: subprocess.run([emul["exe"]] + argv, cwd=os.path.join(d, "mine"), stdout=subprocess.PIPE,
                                                     stderr=subprocess.STDOUT, text=True, timeout=900)),
                     ("ref", lambda: oracle.run_reference_fragopt(argv, cwd=os.path.join(d, "ref")))):
        os.makedirs(os.path.join(d, who))
        r = run()
        if who == "mine":
            assert r.returncode == 0, r.stdout
        out[who] = os.path.join(d, who)
    a, b = np.loadtxt(os.path.join(out["mine"], ctr_name)), np.loadtxt(os.path.join(out["ref"], ctr_name))
    assert a.shape == b.shape and np.abs(a - b).max() <= CTR_TOL
    if mode in ("slac", "rigid"):
        assert np.abs(helpers.read_log_poses(os.path.join(out["mine"], "pose.log")) -
                      helpers.read_log_poses(os.path.join(out["ref"], "pose.log"))).max() <= POSE_TOL
    for f in files:
        if f.endswith(".ctr"):
            assert np.abs(np.loadtxt(os.path.join(out["mine"], f)) - np.loadtxt(os.path.join(out["ref"], f))).max() <= CTR_TOL, f
        else:
            pm, pr = _read_sample_pcd(os.path.join(out["mine"], f)), _read_sample_pcd(os.path.join(out["ref"], f))
            assert pm.shape == pr.shape and len(pm) > 100
            assert np.abs(pm[:, :3] - pr[:, :3]).max() <= 1e-6 and np.abs(pm[:, 3:] - pr[:, 3:]).max() <= 1e-5


@pytest.mark.parametrize("mode", ["slac", "rigid"])
def test_emulated_executable_on_ragged_inputs_against_live_reference(oracle, emul, tmp_path, mode):
    """Inputs the reference tolerates silently: a rejected pair (frame -1), a pair below --blacklistpair, a pair whose corres file is
    missing (it still counts as a pair: OptimizeRigid adds its six ones), comment and blank lines in .xyzn and corres files (a
    line that does not scan repeats the previous point / pair, because sscanf leaves its targets untouched)."""
    if not os.path.exists(oracle.REF_FRAGOPT):
        pytest.skip("oracle/_ref/FragmentOptimizer_ref not built (needs /root/reference)")
    from elasticreconstruction_b200 import synth
    d = str(tmp_path)
    data = os.path.join(d, "data") + os.sep
    ds = synth.make_fragopt_dataset(data, num=4, n_points=4000)
    # a comment line and a blank line in one cloud and one corres file
    p = data + "cloud_bin_xyzn_1.xyzn"
    L = open(p).read().split("\n")
    L.insert(5, "# a comment")
    L.insert(9, "")
    open(p, "w").write("\n".join(L))
    p = data + "corres_0_1.txt"
    L = open(p).read().split("\n")
    L.insert(3, "# comment")
    L.insert(7, "")
    open(p, "w").write("\n".join(L))
    os.remove(data + "corres_1_2.txt")                         # missing file: an accepted pair without correspondences (every
    #                                                            fragment stays connected: an unconstrained fragment makes the
    #                                                            system singular -- garbage from the reference, an error from us)
    recs = []
    for k, (i, j) in enumerate(ds["pairs"]):
        n = len(ds["corres"][(i, j)])
        recs.append((i, j, -1 if (i, j) == (0, 3) else (50 if (i, j) == (2, 3) else n)))    # rejected / below the threshold
    synth.write_log(data + "reg_output.log", ds["reg_poses"], ids=recs)
    args = ["--rgbdslam", data + "traj.log", "--registration", data + "reg_output.log", "--dir", data, "--num", "4", "--interval", "1",
            "--blacklistpair", "1000", "--iteration", "2", "--" + mode]
    os.makedirs(os.path.join(d, "mine"))
    os.makedirs(os.path.join(d, "ref"))
    p = subprocess.run([emul["exe"]] + args, cwd=os.path.join(d, "mine"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0, p.stdout
    ref_out = oracle.run_reference_fragopt(args, cwd=os.path.join(d, "ref"))
    assert p.stdout.count("correspondences.") == ref_out.count("correspondences.") == 4          # four accepted pairs, one of them empty
    a, b = np.loadtxt(os.path.join(d, "mine", "output.ctr")), np.loadtxt(os.path.join(d, "ref", "output.ctr"))
    assert np.abs(a - b).max() <= CTR_TOL
    assert np.abs(helpers.read_log_poses(os.path.join(d, "mine", "pose.log")) - helpers.read_log_poses(os.path.join(d, "ref", "pose.log"))).max() <= POSE_TOL


def test_bench_leg_end_to_end_on_the_emulated_library(emul, monkeypatch):
    """fragopt_bench.run() (the leg bench.py embeds) with the library replaced by the emulation and the device-side
    correspondence search by a cKDTree: every line of the leg's Python executes, and its record has the contract's keys."""
    import importlib.util
    from scipy.spatial import cKDTree
    from elasticreconstruction_b200 import capi, synth
    spec = importlib.util.spec_from_file_location("fragopt_bench", os.path.join(ROOT, "fragopt_bench.py"))
    fb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fb)

    def corres(frags, pairs, device):
        out = []
        for (i, j) in pairs:
            T = np.linalg.inv(synth.fragment_frame(i)) @ synth.fragment_frame(j)
            d, idx = cKDTree(frags[i][0].astype(np.float64)).query(frags[j][0].astype(np.float64) @ T[:3, :3].T + T[:3, 3])
            s = np.nonzero(d < 0.015)[0]
            out.append(np.stack([idx[s], s], axis=1).astype(np.int32))
        return out

    real = capi.FragOpt
    monkeypatch.setattr(fb, "device_correspondences", corres)
    monkeypatch.setattr(fb.capi, "FragOpt", lambda num, res, length, device=0: real(num, res, length, library=emul["lib"]))
    r = fb.run(0, fragments=3, points=3000, iterations=2, warmup=1, cpu=os.path.exists(os.path.join(ROOT, "oracle", "_ref", "FragmentOptimizer_ref")))
    assert r["rows"] > 5000 and r["value"] > 0 and r["e2e"]["value"] > 0 and r["roofline"]["bound"] == "fp64 fma"
    assert r["roofline"]["traffic"] and abs(r["roofline"]["traffic"] / r["rows"] - 59.45) < 0.1
    assert r["max_pose_error_vs_truth"] < r["initial_pose_error"]
    line = fb.as_line(r, None, 1)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "roofline", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in line
    if r["cpu_baseline"]:
        assert r["cpu_baseline"]["kind"] == "reference" and r["cpu_baseline"]["value"] > 0


def _sharded_worker(rank, world, port, libpath, q):
    import ctypes
    import importlib.util
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tempfile
    import torch
    import torch.distributed as dist
    from elasticreconstruction_b200 import capi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = importlib.util.spec_from_file_location("fragopt_bench", os.path.join(ROOT, "fragopt_bench.py"))
    fb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fb)
    with tempfile.TemporaryDirectory() as d:
        data, _ = helpers.fragopt_case("three", d)
    xyzn, pairs, corres, ipose, _ = case_arrays("three", data)
    lib = ctypes.CDLL(libpath)
    g = capi.FragOpt(len(xyzn), 8, 3.0, library=lib)
    for l in range(len(xyzn)):
        g.set_cloud(l, xyzn[l])
    g.set_pairs(pairs[rank::world], corres[rank::world])

    def reduce(buf):
        t = torch.from_numpy(buf)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    loop = fb.ShardedSlacLoop(g, ipose, reduce)
    for _ in range(2):
        loop.step()
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, loop.pose, loop.expand().reshape(-1, 3)))


def test_sharded_loop_world2(emul, golden):
    """fragopt_bench.ShardedSlacLoop on two gloo ranks (pairs dealt round-robin, one all-reduce of the matrix per iteration) against
    the reference binary's result."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    lib = os.path.join(OUT, "libfragopt_emul.so")
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, lib, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, pose, ctr in res:
        assert np.abs(pose - golden["three_slac_pose"]).max() <= POSE_TOL
        assert np.abs(ctr - golden["three_slac_ctr"]).max() <= CTR_TOL


def test_emulated_kernels_with_the_diagonal_jb_variant(tmp_path, golden):
    """-DER_FO_JB_DIAG=1 (compile-time option of csrc/er_fragopt_math.h, default off, not yet measured on a GPU): the diagonal tiles'
    idle accumulators sum b * row instead of eight extra threads.  Same API check and the same end result as the default."""
    import importlib.util
    from elasticreconstruction_b200 import capi
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libfragopt_emul_jbdiag.so")
    srcs = [os.path.join(ROOT, "tests", "emul", "fragopt_emul.cpp"), os.path.join(CSRC, "er_host_math.cpp")]
    if _stale(lib, srcs + [os.path.join(CSRC, "er_fragopt_math.h")]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-DER_FO_JB_DIAG=1",
                               "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-I", "/usr/local/cuda/include"] + srcs + ["-o", lib])
    L = C.CDLL(lib)
    api_check(lambda num: capi.FragOpt(num, 8, 3.0, library=L), tmp_path, vary_emulated_ctas=True)
    spec = importlib.util.spec_from_file_location("fragopt_bench", os.path.join(ROOT, "fragopt_bench.py"))
    fb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fb)
    d, _ = helpers.fragopt_case("three", str(tmp_path / "again"))
    xyzn, pairs, corres, ipose, _ = case_arrays("three", d)
    g = capi.FragOpt(len(xyzn), 8, 3.0, library=L)
    for l in range(len(xyzn)):
        g.set_cloud(l, xyzn[l])
    g.set_pairs(pairs, corres)
    loop = fb.SlacLoop(g, ipose)
    for _ in range(2):
        loop.step()
    assert np.abs(loop.pose - golden["three_slac_pose"]).max() <= POSE_TOL
    assert np.abs(loop.expand().reshape(-1, 3) - golden["three_slac_ctr"]).max() <= CTR_TOL
