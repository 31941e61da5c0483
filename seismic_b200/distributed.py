"""Multi-GPU batch search: index replicated on every GPU, queries split across ranks, ONE gather of result tuples.

Queries are independent units of the reference's search (`&self`, no shared state; the reference itself runs
`par_iter` over queries, src/pylib/mod.rs:1129-1145), so there is no data-path collective: each rank searches its
contiguous slice of the batch and the `(ids, scores, counts)` tuples are gathered on `dst` with a single
torch.distributed collective (NCCL over NVLink for CUDA tensors, gloo for the CPU tests)."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_queries: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced ranges: the first `n % world` ranks get one extra query."""
    base, extra = divmod(n_queries, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_queries(offsets: np.ndarray, comps: np.ndarray, values: np.ndarray, rank: int, world: int):
    lo, hi = shard_bounds(len(offsets) - 1, rank, world)
    o = offsets[lo:hi + 1].astype(np.uint64)
    return (o - o[0]), comps[int(o[0]):int(o[-1])], values[int(o[0]):int(o[-1])]


def pack_results(ids: torch.Tensor, scores: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """[nq, 3k+1] int32: ids (u64 as two i32), score bits, count — one tensor so that one collective moves it."""
    nq = ids.shape[0]
    return torch.cat([ids.contiguous().view(torch.int32).view(nq, -1), scores.contiguous().view(torch.int32),
                      counts.contiguous().view(torch.int32).view(nq, 1)], dim=1)


def unpack_results(packed: torch.Tensor, k: int):
    nq = packed.shape[0]
    ids = packed[:, : 2 * k].contiguous().view(torch.int64).view(nq, k)
    scores = packed[:, 2 * k: 3 * k].contiguous().view(torch.float32)
    counts = packed[:, 3 * k].contiguous()
    return ids, scores, counts


def gather_results(ids: torch.Tensor, scores: torch.Tensor, counts: torch.Tensor, n_total: int, dst: int = 0,
                   group=None):
    """Gather per-rank result tuples (shards made by shard_bounds) on `dst`, in input order.
    Shards differ by at most one row, so every rank pads to the largest shard and ONE dist.gather moves everything."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    k = ids.shape[1]
    rows = -(-n_total // world)
    packed = pack_results(ids, scores, counts)
    if packed.shape[0] < rows:
        packed = torch.cat([packed, packed.new_zeros((rows - packed.shape[0], packed.shape[1]))], dim=0)
    out = [torch.empty_like(packed) for _ in range(world)] if rank == dst else None
    dist.gather(packed, out, dst=dst, group=group)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        parts.append(out[r][: hi - lo])
    return unpack_results(torch.cat(parts, dim=0), k)


def sharded_batch_search(search_fn: Callable, offsets, comps, values, k: int, device: Optional[torch.device] = None,
                         dst: int = 0, group=None):
    """search_fn(offsets, comps, values) -> (ids[nq,k] u64, scores[nq,k] f32, counts[nq] u32) numpy arrays for the
    local shard (e.g. GpuIndex.batch_search bound to its parameters).  Returns the full result on `dst`, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    o, c, v = shard_queries(offsets, comps, values, rank, world)
    ids, scores, counts = search_fn(o, c, v)
    dev = device if device is not None else torch.device("cpu")
    t_ids = torch.from_numpy(ids.view(np.int64)).to(dev)
    t_sc = torch.from_numpy(scores).to(dev)
    t_cnt = torch.from_numpy(counts.view(np.int32)).to(dev)
    res = gather_results(t_ids, t_sc, t_cnt, len(offsets) - 1, dst=dst, group=group)
    if res is None:
        return None
    g_ids, g_sc, g_cnt = res
    return (g_ids.cpu().numpy().view(np.uint64), g_sc.cpu().numpy(), g_cnt.cpu().numpy().view(np.uint32))
