"""Multi-GPU plumbing: one process per GPU (torchrun), static geometry replicated on every rank, configurations /
edges sharded in contiguous blocks, and ONE exchange step at the end -- a gather of the per-configuration result
bytes (and optional distances).  There is no collective on the data path of the kernels themselves.

torch.distributed is only plumbing here: NCCL over NVLink when the results are wanted device-resident on every rank,
gloo for the host-side tests."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous block [lo, hi) of rank `rank` when n units are dealt in blocks of ceil(n / world)"""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank %d / world %d" % (rank, world))
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def interleaved_indices(n: int, rank: int, world: int, block: int = 256) -> np.ndarray:
    """block-cyclic assignment for edges: early-exit makes their cost very uneven, so neighbouring blocks of `block`
    edges go to different ranks"""
    idx = np.arange(n)
    return idx[(idx // block) % world == rank]


def gather_results(local: np.ndarray, n_total: int, index: Optional[np.ndarray] = None, group=None, device=None) -> np.ndarray:
    """all-gathers per-unit results.  `local` holds this rank's block (contiguous sharding, index=None) or the values at
    `index` (block-cyclic).  Works with any initialised torch.distributed backend; `device` = 'cuda' routes through NCCL."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-n_total // world)
    if index is None:
        pad = np.zeros((per,) + local.shape[1:], dtype=local.dtype)
        pad[:len(local)] = local
        t = torch.from_numpy(pad)
        if device is not None:
            t = t.to(device)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t, group=group)
        full = torch.cat(outs)[:n_total]
        return full.cpu().numpy()
    # block-cyclic: exchange (index, value) pairs padded to the largest shard
    cnt = torch.tensor([len(index)], dtype=torch.int64)
    if device is not None:
        cnt = cnt.to(device)
    cnts = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    m = int(max(int(c.item()) for c in cnts))
    pi = np.full(m, -1, dtype=np.int64)
    pi[:len(index)] = index
    pv = np.zeros((m,) + local.shape[1:], dtype=local.dtype)
    pv[:len(local)] = local
    ti, tv = torch.from_numpy(pi), torch.from_numpy(pv)
    if device is not None:
        ti, tv = ti.to(device), tv.to(device)
    oi = [torch.empty_like(ti) for _ in range(world)]
    ov = [torch.empty_like(tv) for _ in range(world)]
    dist.all_gather(oi, ti, group=group)
    dist.all_gather(ov, tv, group=group)
    full = np.zeros((n_total,) + local.shape[1:], dtype=local.dtype)
    for a, b in zip(oi, ov):
        a, b = a.cpu().numpy(), b.cpu().numpy()
        ok = a >= 0
        full[a[ok]] = b[ok]
    return full


class ShardedRunner:
    """Runs a per-shard batch function on this rank's block and gathers the result on every rank.

    `feasible_fn(Q_block) -> uint8 array` is normally `Engine.feasible_batch` of this rank's replica;
    `visible_fn(A_block, B_block) -> uint8 array` is `Engine.edges_visible_batch(..., return_nchecks=False)`."""

    def __init__(self, feasible_fn: Optional[Callable] = None, visible_fn: Optional[Callable] = None, group=None, device=None):
        self.feasible_fn, self.visible_fn, self.group, self.device = feasible_fn, visible_fn, group, device

    def _rw(self):
        import torch.distributed as dist
        return dist.get_rank(self.group), dist.get_world_size(self.group)

    def feasible_batch(self, Q: np.ndarray) -> np.ndarray:
        rank, world = self._rw()
        lo, hi = shard_range(len(Q), rank, world)
        local = np.asarray(self.feasible_fn(Q[lo:hi]), dtype=np.uint8)
        return gather_results(local, len(Q), group=self.group, device=self.device)

    def visible_batch(self, A: np.ndarray, B: np.ndarray, block: int = 256) -> np.ndarray:
        rank, world = self._rw()
        idx = interleaved_indices(len(A), rank, world, block)
        local = np.asarray(self.visible_fn(A[idx], B[idx]), dtype=np.uint8)
        return gather_results(local, len(A), index=idx, group=self.group, device=self.device)
