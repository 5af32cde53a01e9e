"""Iteration sharding across the GPUs of one box (one process per GPU, torch.distributed).

Bootstrap iterations are independent (replicate() at R/bootRanges.R:90), so rank k of W takes a
contiguous slice of 1..R.  Philox draws are keyed by the GLOBAL iteration index, hence the
gathered result is bit-identical for every W.  No data-path collective: only the small
per-iteration vectors are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests); the
[R/W x G] count tiles stay on their rank unless asked for.
"""
from __future__ import annotations

import numpy as np


def shard_iterations(R: int, world: int, rank: int) -> tuple[int, int]:
    """(first global iteration, count) of rank's slice: iterations {r : floor(r * W / R) == rank}."""
    lo = -(-rank * R // world)          # ceil(rank * R / W)
    hi = -(-(rank + 1) * R // world)
    return lo, hi - lo


def gather_iteration_vectors(local: dict, R: int, group=None, device=None) -> dict:
    """All-gathers per-iteration int64 vectors (dict name -> [n_local]) into [R] on every rank."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    names = sorted(local)
    n_max = max(shard_iterations(R, world, k)[1] for k in range(world))
    mine = torch.zeros((len(names), n_max), dtype=torch.int64, device=device)
    for i, k in enumerate(names):
        v = torch.as_tensor(np.ascontiguousarray(local[k], dtype=np.int64))
        mine[i, : v.numel()] = v.to(mine.device)
    out = torch.empty((world * len(names), n_max), dtype=torch.int64, device=device)  # ranks concatenated on dim 0
    dist.all_gather_into_tensor(out, mine, group=group)
    out = out.cpu().numpy().reshape(world, len(names), n_max)
    res = {}
    for i, k in enumerate(names):
        res[k] = np.concatenate([out[w, i, : shard_iterations(R, world, w)[1]] for w in range(world)])
    return res


def boot_counts_sharded(run_shard, R: int, group=None, device=None) -> dict:
    """run_shard(iter0, n) -> dict with 'iter_totals', 'iter_nranges' ([n] each) and optionally
    'feature_counts' ([n, G], kept local).  Returns the gathered [R] vectors plus this rank's tile."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    iter0, n = shard_iterations(R, world, rank)
    res = run_shard(iter0, n)
    out = gather_iteration_vectors({k: res[k] for k in ("iter_totals", "iter_nranges")}, R, group, device)
    out["local_iter0"], out["local_n"] = iter0, n
    out["local_feature_counts"] = res.get("feature_counts")
    return out
