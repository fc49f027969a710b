"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): pair sharding + result gather, and the TSDF
block-partition rule.  The GPU kernels are not involved; the GPU-side equivalence (union of partitions == single
volume) is covered by tests/test_tsdf_gpu.py::test_block_partition_union_equals_single."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from elasticreconstruction_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = sharding.shard_indices(n_pairs, rank, world)
    # fake "alignment result" of pair p: a matrix that encodes p, so the reassembled order can be verified
    local = np.stack([np.concatenate([np.full(16, float(p)), [20.0]]) for p in idx]) if len(idx) else np.zeros((0, 17))
    full = sharding.gather_pair_results(local, n_pairs, rank, world)
    rows = np.full((3 + 2 * rank, 4), float(rank), dtype=np.float32)          # uneven voxel extractions
    allrows = sharding.gather_rows(rows, rank, world)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, full, allrows))


def test_pair_sharding_and_gather_world2():
    world, n_pairs = 2, 7                          # odd count: shards of 4 and 3 (uneven all_gather path)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, full, allrows in res:
        assert full.shape == (n_pairs, 17)
        assert np.array_equal(full[:, 0], np.arange(n_pairs, dtype=np.float64))
        assert np.all(full[:, 16] == 20.0)
        assert allrows.shape == (3 + 5, 4) and np.array_equal(allrows[:3], np.zeros((3, 4))) and np.all(allrows[3:] == 1.0)


def test_shards_cover_everything_once():
    sys.path.insert(0, ROOT)
    from elasticreconstruction_b200 import sharding
    for world in (1, 2, 3, 8):
        seen = np.concatenate([sharding.shard_indices(1225, r, world) for r in range(world)])
        assert np.array_equal(np.sort(seen), np.arange(1225))
        sizes = [len(sharding.shard_indices(1225, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    owners = {sharding.block_owner(x, y, z, 4) for x in range(250, 262) for y in range(250, 262) for z in range(250, 262)}
    assert owners == {0, 1, 2, 3}
