"""BASELINE configs 2 (sharded, matrix gathered), 4 and 5 on N GPUs of one box (torchrun, one process per GPU, NCCL).

  cfg 2: R = 1000 iterations sharded over the ranks; every rank ends up with the whole [R x 20k] count matrix on its
         GPU: one all-gather of the [R/N x G] tiles, device to device, straight out of the engine's buffers.
  cfg 4: 10M points, withinChrom=TRUE, 200k enhancers, R = 10^4 iterations sharded over the ranks;
         per-feature moments (sum, sum^2, #>=observed) all-reduced, per-iteration totals all-gathered.
  cfg 5: blockLength in {1e5, 2e5, 5e5, 1e6} x R = 1000 on 1M ranges: the four block lengths are spread
         over the ranks (rank k takes L_b[k % 4] and 1/ceil(N/4) of its iterations); variance of the
         per-iteration totals per block length (the vignette's variance-vs-L_b diagnostic).

usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_multi.py --config 4
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from nullranges_b200 import Draws, Engine, _lib, synth
from nullranges_b200.sharding import allreduce_moments, gather_count_matrix, gather_iteration_vectors, resident_tensors, shard_iterations

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--iters", type=int, default=0)
ap.add_argument("--group", type=int, default=0, help="cfg 2 in ONE process through a device-group handle of this many GPUs: "
                "nrb200_run_counts_gathered, the count kernels store their tiles into every member's matrix over NVLink")
args = ap.parse_args()
if args.group:
    # ---- cfg 2, strong scaling, device-resident consumers: R = 1000 iterations sharded over the group, every GPU ends
    # up with the whole [R x 20k] matrix.  No collective: the all-gather is the count kernel's store (peer pointers).
    from nullranges_b200.sharding import DeviceView
    R = args.iters or 1000
    y, x = synth.atac_peaks(1_000_000, seed=3), synth.genes(20_000, seed=2)
    G = len(x)
    eng = Engine(list(range(args.group)) if args.group > 1 else 0)
    eng.set_ranges(y.seqlengths, y.seqid, y.start, y.end)
    eng.set_query(x.seqid, x.start, x.end)
    eng.set_plan(_lib.UNSEG_ACROSS, 500_000)
    for s_ in range(3):
        eng.run_counts_gathered(Draws(n_iter=R, seed=100 + s_), want_stats=False)
    times, kernel_ms = [], []
    for s_ in range(10):
        t0 = time.perf_counter()
        eng.run_counts_gathered(Draws(n_iter=R, seed=s_), want_stats=False)   # returns when every member's stream has drained
        times.append(time.perf_counter() - t0)
    compute_ms = []
    for s_ in range(5):   # device time of the slowest member (its kernels + push), from a separate pass with event timing on
        st = eng.run_counts_gathered(Draws(n_iter=R, seed=s_), want_stats=True)
        kernel_ms.append(st["total_ms"])
        compute_ms.append(st["kernel_ms"])
    chks = []
    for k in range(max(args.group, 1)):
        d, _, _, cnt_p = eng.gathered_pointers(k)
        full = torch.as_tensor(DeviceView(cnt_p, (R, G), "<i4"), device=f"cuda:{d}")
        chks.append([int(full.sum()), int((full.to(torch.int64) * torch.arange(1, G + 1, device=full.device)).sum() % (1 << 61))])
    ms = 1e3 * sorted(times)[len(times) // 2]
    print(json.dumps({"config": "cfg2 R=1000 sharded over a device group, [R x 20k] matrix on every GPU (stored by the count kernels over NVLink)",
                      "n_gpus": max(args.group, 1), "iters": R, "ms_per_1000_iterations_wall_median": ms, "ms_wall_min": 1e3 * min(times),
                      "iters_per_s": R / ms * 1e3, "slowest_gpu_device_ms_median": sorted(kernel_ms)[len(kernel_ms) // 2], "of_it_compute_ms": sorted(compute_ms)[len(compute_ms) // 2],
                      "push": "cudaMemcpyPeerAsync" if os.environ.get("NRB200_PUSH_MEMCPY") else "push_tile_kernel (16-byte peer stores)",
                      "matrix_bytes": R * G * 4, "checksum": chks[0], "members_agree": all(c == chks[0] for c in chks)}))
    eng.close()
    sys.exit(0)
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
stream = torch.cuda.Stream(device=dev)   # one stream for the engine's kernels AND the collectives that read its buffers
torch.cuda.set_stream(stream)
eng = Engine(local, stream=stream.cuda_stream)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


if args.config == 2:
    R = args.iters or 1000
    y, x = synth.atac_peaks(1_000_000, seed=3), synth.genes(20_000, seed=2)
    eng.set_ranges(y.seqlengths, y.seqid, y.start, y.end)
    eng.set_query(x.seqid, x.start, x.end)
    B = eng.set_plan(_lib.UNSEG_ACROSS, 500_000)
    G = len(x)
    iter0, n = shard_iterations(R, world, rank)

    def step(seed):
        eng.run_counts_resident(Draws(n_iter=n, seed=seed, iter0=iter0), True, False)
        tile = resident_tensors(eng, n, G)["feature_counts"]
        return gather_count_matrix(tile, R) if world > 1 else tile

    for s_ in range(3):
        step(100 + s_)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    steps = 10
    ev[0].record()
    for s_ in range(steps):
        full = step(s_)
    ev[1].record()
    barrier()
    ms = ev[0].elapsed_time(ev[1]) / steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # every rank holds the same matrix; its checksum does not depend on the number of GPUs
    chk = torch.tensor([int(full.sum()), int((full.to(torch.int64) * torch.arange(1, G + 1, device=dev)).sum() % (1 << 61))], device=dev)
    same = chk.clone()
    if world > 1:
        dist.all_reduce(same, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": "cfg2 R=1000 sharded, [R x 20k] matrix all-gathered on every GPU", "n_gpus": world, "iters": R,
                          "ms_per_1000_iterations": float(t[0]), "iters_per_s": R / float(t[0]) * 1e3,
                          "matrix_bytes": int(full.numel() * 4), "checksum": [int(v) for v in chk.tolist()],
                          "ranks_agree": bool((same == chk).all())}))
elif args.config == 4:
    R = args.iters or 10_000
    y, x = synth.points(10_000_000, seed=6), synth.enhancers(200_000, seed=7)
    eng.set_ranges(y.seqlengths, y.seqid, y.start, y.end)
    eng.set_query(x.seqid, x.start, x.end)
    B = eng.set_plan(_lib.UNSEG_WITHIN, 500_000)
    obs, obs_total = eng.count_observed()
    iter0, n = shard_iterations(R, world, rank)
    eng.run_moments(Draws(n_iter=min(n, 64), seed=7, iter0=iter0), observed=obs)  # warm-up
    if world > 1:  # first collectives set up the communicator: keep that out of the timing
        dist.all_reduce(torch.zeros(8, device=dev))
        gather_iteration_vectors({"iter_totals": np.zeros(n, np.int64)}, R, device=dev)
    barrier()
    t0 = time.perf_counter()
    m = eng.run_moments(Draws(n_iter=n, seed=7, iter0=iter0), observed=obs)
    mom = torch.from_numpy(np.stack([m["sum"], m["sumsq"], m["n_ge_observed"]])).to(dev)
    if world > 1:
        allreduce_moments(mom)                                 # 3 x G x 8 bytes over NVLink
        tot = gather_iteration_vectors({"iter_totals": m["iter_totals"], "iter_nranges": m["iter_nranges"]}, R, device=dev)
    else:
        tot = {"iter_totals": m["iter_totals"], "iter_nranges": m["iter_nranges"]}
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt, m["stats"]["kernel_ms"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    mom = mom.cpu().numpy()
    if rank == 0:
        mean = mom[0] / R
        var = mom[1] / R - mean ** 2
        z = (obs - mean) / np.sqrt(np.maximum(var, 1e-12))
        print(json.dumps({"config": "cfg4 10M points withinChrom vs 200k enhancers, per-feature moments", "n_gpus": world, "iters": R,
                          "wall_s": float(t[0]), "iters_per_s": R / float(t[0]), "max_rank_kernel_ms": float(t[1]),
                          "observed_total": int(obs_total), "bootstrap_mean_total": float(tot["iter_totals"].mean()),
                          "bootstrap_sd_total": float(tot["iter_totals"].std(ddof=1)),
                          "moments_checksum": [int(mom[0].sum()), int(mom[1].sum()), int(mom[2].sum())],
                          "max_abs_feature_z": float(np.abs(z).max()), "blocks": int(B)}))
else:
    R = args.iters or 1000
    y, x = synth.atac_peaks(1_000_000, seed=3), synth.genes(20_000, seed=2)
    eng.set_ranges(y.seqlengths, y.seqid, y.start, y.end)
    eng.set_query(x.seqid, x.start, x.end)
    LBS = [100_000, 200_000, 500_000, 1_000_000]
    per = max(1, world // 4)                                   # ranks sharing one block length
    mine = [LBS[rank % 4]] if world >= 4 else LBS[rank::world]
    res = {}
    if world > 1:
        dist.all_reduce(torch.zeros(8, device=dev))
    barrier()
    t0 = time.perf_counter()
    for Lb in mine:
        eng.set_plan(_lib.UNSEG_ACROSS, Lb)
        sub = (rank // 4) if world >= 4 else 0
        iter0, n = shard_iterations(R, per, sub)
        r = eng.run_counts(Draws(n_iter=n, seed=5, iter0=iter0), want_counts=False)
        res[Lb] = (iter0, r["iter_totals"])
    barrier()
    dt = time.perf_counter() - t0
    # gather every (L_b, slice) on rank 0
    payload = [None] * world
    if world > 1:
        dist.all_gather_object(payload, res)
    else:
        payload = [res]
    if rank == 0:
        out = {}
        for Lb in LBS:
            parts = sorted((p[Lb] for p in payload if Lb in p), key=lambda a: a[0])
            tot = np.concatenate([a[1] for a in parts])
            assert tot.size == R
            out[str(Lb)] = {"mean_total": float(tot.mean()), "var_total": float(tot.var(ddof=1))}
        print(json.dumps({"config": "cfg5 blockLength sweep x R=1000 on 1M ranges", "n_gpus": world, "wall_s": dt,
                          "iters_per_s": 4 * R / dt, "by_block_length": out}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
eng.close()
