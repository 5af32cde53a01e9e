"""Iteration sharding across the GPUs of one box (one process per GPU, torch.distributed).

Bootstrap iterations are independent (replicate() at R/bootRanges.R:90), so rank k of W takes a
contiguous slice of 1..R.  Philox draws are keyed by the GLOBAL iteration index, hence the
gathered result is bit-identical for every W.  No data-path collective: only the small
per-iteration vectors are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests); the
[R/W x G] count tiles stay on their rank unless asked for (gather_count_matrix: one all-gather of
the tiles, device to device over NVLink, straight out of the engine's buffers), and per-feature
moment vectors are all-reduced (allreduce_moments).
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


class DeviceView:
    """Zero-copy view of a device buffer owned by libnrb200 (``torch.as_tensor(DeviceView(...), device="cuda")``)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def resident_tensors(engine, R: int, G: int | None = None) -> dict:
    """Torch views (no copy) of the results the last run_counts_resident() left on the GPU: iter_nranges and
    iter_totals [R] int64, feature_counts [R, G] int32.  Valid until the engine's next run.  The engine queues its
    kernels on ITS stream: create it on the torch stream that consumes the views
    (``Engine(device, stream=torch.cuda.current_stream().cuda_stream)``), or synchronise in between."""
    import torch
    nr_p, tot_p, cnt_p = engine.resident_pointers()
    out = {"iter_nranges": torch.as_tensor(DeviceView(nr_p, (R,), "<i8"), device="cuda"),
           "iter_totals": torch.as_tensor(DeviceView(tot_p, (R,), "<i8"), device="cuda")}
    G = engine.n_query if G is None else G
    if cnt_p and R and G:
        out["feature_counts"] = torch.as_tensor(DeviceView(cnt_p, (R, G), "<i4"), device="cuda")
    return out


def gather_count_matrix(local_tile, R: int, group=None):
    """All-gather of the ranks' [n_local, G] count tiles (torch tensors: CUDA for NCCL, CPU for gloo) into the
    [R, G] matrix on every rank, rows in global iteration order (ncclAllGather of the [R/W x G] tiles).
    The tiles go out as they are; only when R is not a multiple of the world size is a shard padded by one row."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    spans = [shard_iterations(R, world, k) for k in range(world)]
    n_max = max(n for _, n in spans)
    G = local_tile.shape[1]
    if local_tile.shape[0] != spans[rank][1]:
        raise ValueError(f"rank {rank} holds {local_tile.shape[0]} iterations, its shard has {spans[rank][1]}")
    mine = local_tile.contiguous()
    if mine.shape[0] != n_max:
        mine = torch.cat([mine, mine.new_zeros((n_max - mine.shape[0], G))])
    out = torch.empty((world * n_max, G), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    if all(n == n_max for _, n in spans):
        return out  # rank-major shards are already the global order
    return torch.cat([out[k * n_max: k * n_max + n] for k, (_, n) in enumerate(spans)])


def allreduce_moments(*vectors, group=None):
    """Sum over ranks, in place, of per-feature moment vectors (sum, sum of squares, #{>= observed}): torch tensors on
    the collective's device (ncclAllReduce of 3 * G * 8 bytes).  One collective for all of them."""
    import torch
    import torch.distributed as dist
    flat = torch.cat([v.reshape(-1) for v in vectors])
    dist.all_reduce(flat, group=group)
    pos = 0
    for v in vectors:
        v.copy_(flat[pos: pos + v.numel()].reshape(v.shape))
        pos += v.numel()
    return vectors


def boot_counts_sharded(run_shard, R: int, group=None, device=None, gather_matrix: bool = False) -> dict:
    """run_shard(iter0, n) -> dict with 'iter_totals', 'iter_nranges' ([n] each) and optionally
    'feature_counts' ([n, G]).  Returns the gathered [R] vectors plus this rank's tile; with gather_matrix the
    whole [R, G] matrix as 'feature_counts' on every rank as well."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    iter0, n = shard_iterations(R, world, rank)
    res = run_shard(iter0, n)
    out = gather_iteration_vectors({k: res[k] for k in ("iter_totals", "iter_nranges")}, R, group, device)
    out["local_iter0"], out["local_n"] = iter0, n
    out["local_feature_counts"] = res.get("feature_counts")
    if gather_matrix:
        tile = torch.as_tensor(res["feature_counts"])
        if device is not None:
            tile = tile.to(device)
        out["feature_counts"] = gather_count_matrix(tile, R, group).cpu().numpy()
    return out
