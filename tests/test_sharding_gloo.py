"""CPU, world_size 2 over gloo: the iteration sharding used at N > 1 GPUs.

The compute inside each rank is the oracle here (no GPU in this container); what is under test is
the host logic: slice arithmetic, global-iteration Philox keys, and the gather."""
import os
import socket

import numpy as np
import pytest

from nullranges_b200.sharding import shard_iterations


def test_shard_iterations_partition():
    for R in (0, 1, 7, 1000, 1001):
        for W in (1, 2, 3, 8):
            spans = [shard_iterations(R, W, k) for k in range(W)]
            assert sum(n for _, n in spans) == R
            pos = 0
            for lo, n in spans:
                assert lo == pos
                pos += n
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


def _worker(rank, world, port, R, q):
    import torch.distributed as dist
    from nullranges_b200.sharding import boot_counts_sharded
    from oracle import oracle as O
    from tests import helpers as H
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = H.random_case(seed=77, n=400, n_seq=3, n_query=30)
    plan, y, qi, _ = H.oracle_objects(case, 0)
    out = boot_counts_sharded(lambda i0, n: O.boot_counts(plan, y, qi, n, seed=99, iter0=i0), R, gather_matrix=True)
    import torch
    from nullranges_b200.sharding import allreduce_moments
    fc = torch.as_tensor(out["local_feature_counts"]).to(torch.int64)
    m1, m2 = fc.sum(0), (fc * fc).sum(0)
    allreduce_moments(m1, m2)
    q.put((rank, out["iter_totals"], out["iter_nranges"], out["local_iter0"], out["local_feature_counts"], out["feature_counts"],
           m1.numpy(), m2.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("R", [9, 10])
def test_sharded_equals_unsharded_gloo(R):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    from tests import helpers as H
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, R, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = H.random_case(seed=77, n=400, n_seq=3, n_query=30)
    plan, y, qi, _ = H.oracle_objects(case, 0)
    ref = O.boot_counts(plan, y, qi, R, seed=99, iter0=0)
    full = ref["feature_counts"].astype(np.int64)
    for rank, tot, nr, i0, fc, matrix, m1, m2 in got:
        assert np.array_equal(tot, ref["iter_totals"]) and np.array_equal(nr, ref["iter_nranges"])
        assert np.array_equal(fc, ref["feature_counts"][i0 : i0 + fc.shape[0]])
        assert np.array_equal(matrix, ref["feature_counts"])          # all-gathered tiles, global iteration order
        assert np.array_equal(m1, full.sum(0)) and np.array_equal(m2, (full * full).sum(0))  # all-reduced moments
