"""Multi-GPU plumbing (one process per GPU, torch.distributed): the two hot paths shard without any data-path
collective.

* ICP: fragment pairs are independent (each writes only its own record, CorresApp.cpp:220-318).  Pairs are dealt
  round-robin to the ranks; every rank holds all fragments; at the end ONE all_gather of the 16 doubles + iteration
  count per pair reassembles reg_output.log's records in their original order.
* TSDF: a voxel update reads only the depth image and the pose (TSDFVolume.cpp:75-94), so blocks partition freely;
  rank r owns the blocks with (xi + yi + zi) % world == r (er_tsdf_set_partition) and every rank sees every frame.
  The final world.pcd is the concatenation of the ranks' extractions (gather_rows).
Works with the nccl backend on GPUs and with gloo on CPU (the unit tests run it at world_size 2 under gloo).
"""
from __future__ import annotations

import numpy as np


def shard_indices(n: int, rank: int, world: int) -> np.ndarray:
    """Indices of the work items (pairs) rank `rank` processes: round-robin, like schedule(dynamic) evens out cost."""
    return np.arange(rank, n, world)


def block_owner(xi: int, yi: int, zi: int, world: int) -> int:
    """Owner rank of TSDF block (xi, yi, zi); must match the test in touch_kernel (er_tsdf.cu)."""
    return (xi + yi + zi) % world


def gather_pair_results(local: np.ndarray, n_total: int, rank: int, world: int, device=None) -> np.ndarray:
    """all_gather per-pair rows (e.g. [n_local, 17] = pose + iterations) and restore the original pair order."""
    local = np.ascontiguousarray(local, dtype=np.float64)
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    width = local.shape[1]
    counts = [len(shard_indices(n_total, r, world)) for r in range(world)]
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    bufs = [torch.empty((c, width), dtype=torch.float64, device=dev) for c in counts]
    dist.all_gather(bufs, torch.from_numpy(local).to(dev)) if len(set(counts)) == 1 else _all_gather_uneven(bufs, local, dev, rank, world)
    out = np.empty((n_total, width), dtype=np.float64)
    for r in range(world):
        out[shard_indices(n_total, r, world)] = bufs[r].cpu().numpy()
    return out


def _all_gather_uneven(bufs, local, dev, rank, world):
    """all_gather needs equal shapes: pad every shard to the longest, then trim."""
    import torch
    import torch.distributed as dist
    m = max(b.shape[0] for b in bufs)
    width = bufs[0].shape[1]
    mine = torch.zeros((m, width), dtype=torch.float64, device=dev)
    mine[: local.shape[0]] = torch.from_numpy(local).to(dev)
    padded = [torch.empty((m, width), dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(padded, mine)
    for r in range(world):
        bufs[r].copy_(padded[r][: bufs[r].shape[0]])


def gather_rows(local: np.ndarray, rank: int, world: int, device=None) -> np.ndarray:
    """Concatenate variable-length float32 row blocks (extracted voxels) from all ranks on every rank."""
    local = np.ascontiguousarray(local, dtype=np.float32)
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    ns = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(ns, n)
    ns = [int(x.item()) for x in ns]
    m = max(ns)
    width = local.shape[1]
    mine = torch.zeros((m, width), dtype=torch.float32, device=dev)
    mine[: local.shape[0]] = torch.from_numpy(local).to(dev)
    padded = [torch.empty((m, width), dtype=torch.float32, device=dev) for _ in range(world)]
    dist.all_gather(padded, mine)
    return np.concatenate([padded[r][: ns[r]].cpu().numpy() for r in range(world)], axis=0)
