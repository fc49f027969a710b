"""GPU parity tests for the ICP path: CUDA (through the C ABI) vs the restated oracle on the same inputs.
Bar (BASELINE.json): final pose within 1e-5 Frobenius of the CPU path with the same convergence rule."""
import numpy as np
import pytest

from elasticreconstruction_b200 import capi, synth

pytestmark = pytest.mark.gpu
POSE_TOL = 1e-5


@pytest.fixture(scope="module")
def frags(oracle):
    out = []
    for i in range(4):
        xyz, nrm = synth.make_fragment(i, 30000)
        out.append((xyz, nrm, oracle.OracleCloud(xyz, nrm, 0.03), capi.Cloud(xyz, nrm, cell_size=0.03)))
    return out


def test_count_inliers_equal(oracle, frags):
    for s in range(1, 4):
        P = synth.perturbation(s)
        a = oracle.icp_count_inliers(frags[0][2], frags[s][2], P, 0.03)
        b = capi.icp_count_inliers(frags[0][3], frags[s][3], P, 0.03)
        assert a == b and a > 1000


@pytest.mark.parametrize("fixed", [False, True])
def test_align_matches_oracle(oracle, frags, fixed):
    for s in range(1, 4):
        P = synth.perturbation(10 + s)
        To, ito, fo, _ = oracle.icp_align(frags[0][2], frags[s][2], P, 0.03, 20, 1e-6, fixed)
        Tg, itg, fg = capi.icp_align(frags[0][3], frags[s][3], P, 0.03, 20, 1e-6, fixed)
        assert ito == itg
        assert np.linalg.norm(To - Tg) <= POSE_TOL, (np.linalg.norm(To - Tg), ito)
        assert abs(fo - fg) <= 1e-9 + 1e-6 * fo


def test_align_batch_matches_singles(oracle, frags):
    clouds = [f[3] for f in frags]
    tgt, src, Ti = [], [], []
    for i in range(4):
        for j in range(i + 1, 4):
            tgt.append(i); src.append(j); Ti.append(synth.perturbation(100 + 4 * i + j))
    T, its, fit = capi.icp_align_batch(clouds, tgt, src, np.array(Ti), fixed_iters=False)
    for p in range(len(tgt)):
        To, ito, fo, _ = oracle.icp_align(frags[tgt[p]][2], frags[src[p]][2], Ti[p])
        assert its[p] == ito and np.linalg.norm(T[p] - To) <= POSE_TOL


def test_no_correspondences_keeps_guess(frags):
    far = synth.make_pose(np.eye(3), [5.0, 0.0, 0.0])
    T, it, fit = capi.icp_align(frags[0][3], frags[1][3], far)
    assert it == 0 and np.array_equal(T, far.astype(np.float32).astype(np.float64))


def test_correspondences_and_information(oracle, frags):
    P = synth.perturbation(7, max_rot_deg=0.1, max_trans=0.002)
    po, io = oracle.icp_correspondences(frags[0][2], frags[1][2], P, 0.015, 0.8660)
    pg, ig = capi.icp_correspondences(frags[0][3], frags[1][3], P, 0.015, 0.8660)
    assert np.array_equal(po, pg)
    assert np.allclose(io, ig, rtol=1e-10, atol=1e-6)
    assert ig[0, 0] == len(pg)


def test_empty_and_tiny_clouds():
    xyz, nrm = synth.make_fragment(0, 3000)
    big = capi.Cloud(xyz, nrm)
    tiny = capi.Cloud(xyz[:2], nrm[:2])
    T, it, fit = capi.icp_align(big, tiny, np.eye(4))
    assert it == 0                                    # fewer than 3 correspondences: loop breaks (icp.hpp)
    pairs, info = capi.icp_correspondences(big, tiny, np.eye(4))
    assert len(pairs) == 2
    with pytest.raises(capi.ElasticError):
        capi.icp_align(big, tiny, np.eye(4), max_dist=0.5)        # radius larger than the grid cell


def test_partial_overlap_and_large_offsets(oracle):
    """Only part of the source has neighbours within the gate, and the first iterations have large residuals (the
    pruned neighbourhood search must still return the exact nearest neighbour)."""
    xyz0, n0 = synth.make_fragment(0, 40000)
    xyz1, n1 = synth.make_fragment(1, 40000)
    keep = xyz1[:, 0] + xyz1[:, 1] < 2.2                       # drop a corner region of the source
    xyz1, n1 = xyz1[keep], n1[keep]
    xyz1 = xyz1 + np.where(xyz1[:, 2:3] > 1.0, np.float32(0.08), np.float32(0.0))   # push a slab out of reach
    to, so = oracle.OracleCloud(xyz0, n0, 0.03), oracle.OracleCloud(xyz1, n1, 0.03)
    tg, sg = capi.Cloud(xyz0, n0, cell_size=0.03), capi.Cloud(xyz1, n1, cell_size=0.03)
    for seed in (21, 22):
        P = synth.perturbation(seed, max_rot_deg=2.5, max_trans=0.02)
        assert oracle.icp_count_inliers(to, so, P, 0.03) == capi.icp_count_inliers(tg, sg, P, 0.03)
        To, ito, fo, _ = oracle.icp_align(to, so, P, 0.03, 20, 1e-6, False)
        Tg, itg, fg = capi.icp_align(tg, sg, P, 0.03, 20, 1e-6, False)
        assert ito == itg and np.linalg.norm(To - Tg) <= POSE_TOL
        po, io = oracle.icp_correspondences(to, so, To, 0.015, 0.8660)
        pg, ig = capi.icp_correspondences(tg, sg, To, 0.015, 0.8660)
        assert np.array_equal(po, pg) and np.allclose(io, ig, rtol=1e-10, atol=1e-6)
