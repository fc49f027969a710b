"""The order-independent voxel digest bench.py uses for its union check (partitions together == unpartitioned volume):
CPU unit test, including the reduction over ranks with gloo (world_size 2)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _voxels(seed, n):
    rng = np.random.default_rng(seed)
    v = np.empty((n, 4), dtype=np.float32)
    v[:, :3] = rng.integers(-500, 500, size=(n, 3))
    v[:, 3] = rng.uniform(-0.98, 0.98, size=n)
    return v


def test_digest_is_order_and_partition_independent():
    v = _voxels(1, 100000)
    single = bench.voxel_digest(v)
    perm = np.random.default_rng(2).permutation(len(v))
    assert np.array_equal(bench.voxel_digest(v[perm]), single)
    owner = (v[:, 0].astype(np.int64) + v[:, 1].astype(np.int64) + v[:, 2].astype(np.int64)) % 3
    total = sum(bench.voxel_digest(v[owner == r]) for r in range(3))
    assert bench.digest_equal(total, single)
    w = v.copy()
    w[12345, 3] = np.nextafter(w[12345, 3], np.float32(1))        # one tsdf value off by one ulp
    assert not bench.digest_equal(bench.voxel_digest(w), single)
    assert not bench.digest_equal(bench.voxel_digest(v[:-1]), single)
    assert bench.digest_equal(bench.voxel_digest(v[:0]) + single, single)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    v = _voxels(7, 50000)
    owner = (np.abs(v[:, 0]).astype(np.int64) + np.abs(v[:, 1]).astype(np.int64)) % world
    dg = torch.from_numpy(bench.voxel_digest(v[owner == rank]))
    dist.all_reduce(dg, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put(bool(bench.digest_equal(dg.numpy(), bench.voxel_digest(v))))
    dist.destroy_process_group()


def test_digest_sum_over_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_ctrlgrid_scene_is_consistent():
    """The configs[3] inputs bench.py generates: pose[c] * seg[c*I + j] reproduces the world trajectory (what
    CIntegrateApp::Init composes, IntegrateApp.cpp:68-75), seg_traj has the extra entry Init needs, every fragment has a
    lattice, and the lattice of fragment 0 is the regular one up to the <= 1 cm displacement."""
    from elasticreconstruction_b200 import synth
    sc = bench.ctrlgrid_scene(200, interval=50)
    assert sc["num"] == 4 and len(sc["seg"]) == 201 and sc["grids"].shape == (4, 9 ** 3, 3)
    for f in range(200):
        T = sc["pose"][f // 50] @ sc["seg"][f]
        assert np.abs(T - sc["world"][f]).max() < 5e-7              # quantised to %.8f like the .log files
    reg = synth.regular_lattice(8, 3.0)
    assert np.abs(sc["grids"][0] - reg).max() <= 0.0100001
    # camera inside the 6 m room, looking at walls within integration range for most pixels
    assert np.all(sc["world"][:, :3, 3] > 0.9) and np.all(sc["world"][:, :3, 3] < 5.1)
