#!/usr/bin/env python
"""bench.py -- bootstrap iterations/sec of the fused bootRanges -> countOverlaps path.

Workload (BASELINE.json configs[1], SURVEY 8d cfg 2): unsegmented across-chromosome block
bootstrap of 1M synthetic ATAC-like peaks on hg38 chr1-22, blockLength = 5e5, against 20k
synthetic genes; per-iteration totals AND the [R x 20k] per-gene count matrix.
One "step" = R = 1000 bootstrap iterations per GPU.

  value   iterations/s, device-resident (ranges, query and plan already in HBM; block draws are
          Philox on the device, so a step has no input copy), CUDA events, max over ranks; one process per GPU
  e2e     iterations/s through the public API bootCounts(y, x, ..., gpus=N) from PAGEABLE host arrays into a pageable
          caller-owned int32 matrix (R's vectors are pageable): every step uploads y and x, builds the indexes, runs
          R * N iterations sharded over the N GPUs and delivers ONE [R * N x G] host matrix.  e2e_fresh_result: the
          same with the result allocated inside every call; e2e_pinned: from / into pinned buffers.  Under torchrun
          rank 0 drives the N devices through one device-group handle (the product's multi-GPU path); the other
          ranks wait on the host.
  roofline  the dominant kernel (boot_rank_count_kernel) against the minimal DRAM bytes of the rank path;
          roofline_l2: the same kernel against what bounds it -- 32-byte sectors through the L2 -- with the ceiling
          measured on the device in this run; roofline_model_8d: SURVEY 8d's byte model (which the rank path does
          not follow: it never reads the sampled ranges); roofline_stream: the streaming kernel that does
  checks  N-GPU result == 1-GPU result (per-iteration totals of the same global iterations; count matrix of the
          device group against one GPU)
  cpu_baseline  the CPU oracle (plain-C port of the reference algorithm) on this box's cores

--impl reference: the reference's CPU implementation of the same path.  The reference is pure
R and R is not installed, so this arm times the oracle port (oracle/nr_oracle.c) with all host
threads on a bounded sample (iterations per step reduced; same ranges, query and block length).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# torchrun pins OMP_NUM_THREADS to 1 for every rank unless the caller set it.  The engine's host side (window lists,
# staging copies) is OpenMP: give each rank its share of the cores instead.  Must happen before numpy / torch load.
_lws = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
if _lws > 1 and os.environ.get("OMP_NUM_THREADS") == "1":
    try:
        _cores = len(os.sched_getaffinity(0))
    except AttributeError:
        _cores = os.cpu_count() or 1
    # rank 0 may use all cores: in the end-to-end leg it alone works (it drives all N GPUs through one device-group handle
    # while the other ranks sleep in a host barrier); the others keep to their share
    os.environ["OMP_NUM_THREADS"] = str(max(1, _cores if os.environ.get("RANK", "0") == "0" else _cores // _lws))

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_STEP = 1000
L_B = 500_000
WORKLOAD = "cfg2: unsegmented across-chrom bootstrap, 1M ATAC-like peaks (hg38 chr1-22), blockLength=5e5, R=1000/step, countOverlaps vs 20k genes -> [R] totals + [R x 20k] counts"
METRIC = "bootstrap iters/sec @1M ranges (sample+shift+countOverlaps)"


def make_inputs():
    from nullranges_b200 import synth
    return synth.atac_peaks(1_000_000, seed=3), synth.genes(20_000, seed=2)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------
# CPU legs (oracle)
# ---------------------------------------------------------------------------------------------
def cpu_objects(y, x):
    from oracle import oracle as O
    plan = O.Plan(O.UNSEG_ACROSS, y.seqlengths, L_B)
    yi = O.Index(len(y.seqlengths), y.seqid, y.start, y.end)
    qi = O.Index(len(y.seqlengths), x.seqid, x.start, x.end)
    return O, plan, yi, qi


def cpu_sample(O, plan, yi, qi, n_iter, threads, iter0=0):
    t0 = time.perf_counter()
    O.boot_counts(plan, yi, qi, n_iter, seed=2024, iter0=iter0, nthreads=threads)
    return time.perf_counter() - t0


def config_dict():
    """The workload description, identical in both arms (the driver compares the two `config` objects)."""
    return {"workload": WORKLOAD, "iters_per_step_per_gpu": R_STEP, "ranges": 1_000_000, "features": 20_000, "block_length": L_B,
            "rng": "philox4x32-10, counter = (global iteration, block), seed 2024",
            "l2": "flushed between timed steps (256 MiB fill, outside the step events); inside a step the index is re-read by design"}


def run_reference(args):
    """The reference arm: oracle port of the R algorithm, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    y, x = make_inputs()
    O, plan, yi, qi = cpu_objects(y, x)
    cores = host_cores()
    # size a step to ~2 s of CPU work
    probe = cpu_sample(O, plan, yi, qi, cores, cores)
    per_step = max(cores, int(round(2.0 / max(probe / cores, 1e-6) / cores)) * cores)
    per_step = min(per_step, R_STEP)
    for w in range(args.warmup):
        cpu_sample(O, plan, yi, qi, per_step, cores, iter0=w * per_step)
    t = 0.0
    for s in range(args.steps):
        t += cpu_sample(O, plan, yi, qi, per_step, cores, iter0=(args.warmup + s) * per_step)
    value = per_step * args.steps / t
    sample = f"{per_step} iterations/step of the same workload (of {R_STEP}), Philox draws, {cores} pthreads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "iter/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config_dict(),
        "note": "the reference is pure R (not installable here: no R); this is the plain-C oracle port of its algorithm, faster than real R",
        "cpu_baseline": {"value": value, "unit": "iter/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every ~2 ms via NVML while running."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.002)

    def stop(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = (int(v) for v in out.strip().split(","))
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "note": "idle sample (nvml unavailable)"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def pinned(shape, dtype):
    import torch
    t = torch.empty(shape, dtype=dtype, pin_memory=True)
    return t, t.numpy()


def checksum(a) -> int:
    """Order-sensitive 61-bit checksum of an integer array (sum of value * (position + 1) mod 2^61 - 1)."""
    v = np.asarray(a, dtype=np.int64).ravel()
    m = (1 << 61) - 1
    w = (np.arange(1, v.size + 1, dtype=np.int64) % 1_000_003)
    return int((int((v % 1_000_003 * w).sum()) + int(v.sum())) % m)


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from nullranges_b200 import Draws, Engine, _lib, bootCounts
    from nullranges_b200.sharding import resident_tensors, shard_iterations

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libnrb200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    # Keep stdout to the one JSON line: libraries (NCCL with NCCL_DEBUG=VERSION set on the box, for one) write
    # there on their own.  File descriptor 1 points at stderr until the line is ready.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    host_pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_pg = dist.new_group(backend="gloo")  # host-side waits: a rank parked here does not spin a kernel on its GPU

    def host_barrier():
        if world > 1:
            dist.barrier(group=host_pg)

    y, x = make_inputs()
    G, R = len(x), R_STEP
    stream = torch.cuda.Stream()
    eng = Engine(local, stream=stream.cuda_stream)
    eng.set_ranges(y.seqlengths, y.seqid, y.start, y.end, None)
    eng.set_query(x.seqid, x.start, x.end, None)
    B = eng.set_plan(_lib.UNSEG_ACROSS, L_B)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step_resident(step, want_stats):
        # global iteration index: rank-major shards of R iterations per step
        d = Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=2024, iter0=(step * world + rank) * R)
        return eng.run_counts_resident(d, want_counts=True, want_stats=want_stats)

    pending = []

    def gather_small():
        """All-gather of this step's [2 x R] per-iteration vectors, asynchronous: it overlaps the next
        step's kernels; wait_gathers() (inside the next step's timed bracket) accounts for it."""
        if world == 1:
            return
        views = resident_tensors(eng, R, 0)
        mine = torch.stack([views["iter_nranges"], views["iter_totals"]])  # staging copy: the engine reuses its buffers
        out = torch.empty((world * 2, R), dtype=mine.dtype, device="cuda")
        pending.append((dist.all_gather_into_tensor(out, mine, async_op=True), out))

    def wait_gathers():
        while pending:
            work, out = pending.pop(0)
            work.wait()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)  # samples from the warm-up to the end of the accounting pass (same kernels throughout)
    clocks.start()
    with torch.cuda.stream(stream):
        for w in range(args.warmup):
            step_resident(w, False)
            wait_gathers()
            gather_small()
        wait_gathers()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        t_wall = time.perf_counter()
        for s in range(args.steps):  # no host synchronisation inside the timed region: the queue runs ahead
            flush.fill_(s & 0xFF)  # L2 flush between timed steps (outside the per-step events)
            ev[s][0].record(stream)
            step_resident(args.warmup + s, False)
            wait_gathers()          # the previous step's gather ran under this step's kernels
            gather_small()
            if s == args.steps - 1:
                wait_gathers()      # the last one has nothing to hide behind
            ev[s][1].record(stream)
        barrier()
        t_wall = time.perf_counter() - t_wall
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        # accounting pass over the SAME steps (same global iterations, hence the same sampled ranges):
        # per-kernel CUDA-event times, launches and H of the byte model.  Not part of `value`.
        kernel_ms, hits, launches, rows_ms, count_ms = 0.0, 0, 0, 0.0, 0.0
        for s in range(args.steps):
            flush.fill_(s & 0xFF)
            st = step_resident(args.warmup + s, True)
            kernel_ms += st["kernel_ms"]
            rows_ms += st["rows_kernel_ms"]
            count_ms += st["count_kernel_ms"]
            hits += st["n_hits"]
            launches += st["n_launches"]
            n_windows = st["n_windows"]
        clk = clocks.stop()
        clk["window"] = "warm-up + timed region + accounting pass (the timed region alone lasts a few ms)"

        # ---- result check 1: the world's GPUs together reproduce what one GPU computes.  Global iterations [0, R) are
        # split over the ranks exactly as the product splits them (shard_iterations); the gathered per-iteration
        # totals must equal rank 0's own run of all R iterations.
        lo, n_mine = shard_iterations(R, world, rank)
        eng.run_counts_resident(Draws(n_iter=n_mine, rng_mode=_lib.RNG_PHILOX, seed=2024, iter0=lo), want_counts=True, want_stats=False)
        mine = resident_tensors(eng, n_mine, 0)["iter_totals"].clone()
        n_max = max(shard_iterations(R, world, k)[1] for k in range(world))
        pad = torch.zeros(n_max, dtype=torch.int64, device="cuda")
        pad[:n_mine] = mine
        if world > 1:
            allv = torch.empty(world * n_max, dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(allv, pad)
            allv = allv.cpu().numpy().reshape(world, n_max)
            sharded = np.concatenate([allv[k, : shard_iterations(R, world, k)[1]] for k in range(world)])
        else:
            sharded = pad.cpu().numpy()[:n_mine]
        eng.run_counts_resident(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=2024, iter0=0), want_counts=True, want_stats=False)
        alone = resident_tensors(eng, R, 0)["iter_totals"].cpu().numpy()
        checks = {"iter_totals_checksum_N_gpus": checksum(sharded), "iter_totals_checksum_1_gpu": checksum(alone),
                  "n_gpu_equals_1_gpu": bool(np.array_equal(sharded, alone))}

    # ---- the streaming kernel on the same workload (the path that really reads the 8*H bytes)
    stream_ms, stream_hits, stream_steps = 0.0, 0, 3
    l2_peak = None
    if rank == 0:
        eng_s = Engine(local, stream=stream.cuda_stream, force_stream=True)
        eng_s.set_ranges(y.seqlengths, y.seqid, y.start, y.end, None)
        eng_s.set_query(x.seqid, x.start, x.end, None)
        eng_s.set_plan(_lib.UNSEG_ACROSS, L_B)
        with torch.cuda.stream(stream):
            for s_ in range(1 + stream_steps):
                st = eng_s.run_counts_resident(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=2024, iter0=s_ * R), True, True)
                if s_:
                    stream_ms += st["kernel_ms"]
                    stream_hits += st["n_hits"]
        eng_s.close()
        l2_peak = eng.measure_l2_sector_rate(64 << 20)  # the count kernel's ceiling, measured on this GPU now
    torch.cuda.synchronize()
    host_barrier()

    # ---- end to end through the public API, all N GPUs driven by rank 0 through ONE device-group handle; the other
    # ranks wait on the host (gloo), their GPUs idle.  Weak scaling: R iterations per GPU per step, one [R * N x G] result.
    e2e = {}
    if rank == 0:
        gpus = list(range(world)) if world > 1 else None
        Rt = R * world
        e2e_steps = max(2, min(args.steps, 5))
        y_pg, x_pg = make_inputs()  # plain numpy arrays: pageable, like R's vectors

        def timed(fn):
            fn(0)
            fn(1)
            t0 = time.perf_counter()
            for s_ in range(e2e_steps):
                res = fn(2 + s_)
            return (time.perf_counter() - t0) / e2e_steps, res

        # (a) pageable in, pageable out: caller-owned plain host buffers, as the C ABI takes them (R's vectors are pageable)
        out_pg = {"iter_totals": np.zeros(Rt, np.int64), "iter_nranges": np.zeros(Rt, np.int64), "feature_counts": np.zeros((Rt, G), np.int32)}
        t_pg, res_pg = timed(lambda s_: bootCounts(y_pg, x_pg, L_B, R=Rt, rng="philox", seed=2024, iter0=s_ * Rt, gpus=gpus, out=out_pg))
        # (a') the same with the result matrix allocated inside every call (fresh pages: the kernel zeroes and maps
        # R * G * 4 bytes while the call runs)
        t_fresh, res_fresh = timed(lambda s_: bootCounts(y_pg, x_pg, L_B, R=Rt, rng="philox", seed=2024, iter0=s_ * Rt, gpus=gpus))
        # (b) pinned in, pinned out
        keep = []

        def pinned_copy(a):
            t = torch.empty(a.shape, dtype={np.dtype(np.int32): torch.int32, np.dtype(np.int8): torch.int8}[a.dtype], pin_memory=True)
            t.numpy()[...] = a
            keep.append(t)
            return t.numpy()

        y_pin, x_pin = make_inputs()
        for gr in (y_pin, x_pin):
            gr.seqid, gr.start, gr.end = pinned_copy(gr.seqid), pinned_copy(gr.start), pinned_copy(gr.end)
        outs = {"iter_totals": torch.empty((Rt,), dtype=torch.int64, pin_memory=True), "iter_nranges": torch.empty((Rt,), dtype=torch.int64, pin_memory=True),
                "feature_counts": torch.empty((Rt, G), dtype=torch.int32, pin_memory=True)}
        out_np = {k: v.numpy() for k, v in outs.items()}
        t_pin, res_pin = timed(lambda s_: bootCounts(y_pin, x_pin, L_B, R=Rt, rng="philox", seed=2024, iter0=s_ * Rt, gpus=gpus, out=out_np))
        # (c) the same into a pinned uint16 matrix (nrb200_run_counts_u16): for callers that can take the counts as uint16
        out16 = dict(out_np)
        keep16 = torch.empty((Rt, G), dtype=torch.uint16, pin_memory=True)
        out16["feature_counts"] = keep16.numpy()
        t_u16, res_u16 = timed(lambda s_: bootCounts(y_pin, x_pin, L_B, R=Rt, rng="philox", seed=2024, iter0=s_ * Rt, gpus=gpus, out=out16,
                                                     counts_dtype=np.uint16))
        # (d) the matrix-free statistic: per-feature mean / sd / empirical p of the same counts over the same R * N iterations,
        # reduced on the GPUs (bootMoments: 3 * G * 8 bytes come back instead of R * G * 4) -- where the host does not bound the step
        from nullranges_b200 import bootMoments
        t_mom, res_mom = timed(lambda s_: bootMoments(y_pg, x_pg, L_B, R=Rt, rng="philox", seed=2024, iter0=s_ * Rt, gpus=gpus))
        last = 2 + e2e_steps - 1
        assert int(res_pg.feature_counts.sum()) == int(res_pg.iter_totals.sum())
        # result check 2: the device group's matrix == one GPU's matrix, for the first R iterations of the last step
        one = bootCounts(y_pin, x_pin, L_B, R=R, rng="philox", seed=2024, iter0=last * Rt, device=0)
        checks["e2e_matrix_checksum_group"] = checksum(res_pg.feature_counts[:R])
        checks["e2e_matrix_checksum_1_gpu"] = checksum(one.feature_counts)
        checks["e2e_group_equals_1_gpu"] = bool(np.array_equal(res_pg.feature_counts[:R], one.feature_counts) and
                                                np.array_equal(res_pin.feature_counts[:R], one.feature_counts) and
                                                np.array_equal(res_pg.iter_totals[:R], one.iter_totals))
        h2d = int(sum(a.nbytes for gr in (y_pg, x_pg) for a in (gr.seqid, gr.start, gr.end)))
        api = ("nullranges_b200.bootCounts(y, x, blockLength, R = 1000 * N, gpus = N): upload + index build + run + D2H every step, "
               "one [R x G] int32 host matrix; rank 0 drives the N GPUs through one device-group handle")
        e2e = {"e2e": {"value": Rt / t_pg, "unit": "iter/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": int(res_pg.stats["d2h_bytes"]),
                       "steps": e2e_steps, "ms_per_step": 1e3 * t_pg, "api": api,
                       "buffers": "pageable (plain numpy) inputs and a pageable caller-owned [R x G] int32 result, as the C ABI / an R caller hands them in; "
                                  "the matrix crosses PCIe as uint16 and the host widens it into the caller's int32 matrix"},
               "e2e_fresh_result": {"value": Rt / t_fresh, "unit": "iter/s", "ms_per_step": 1e3 * t_fresh,
                                    "buffers": "as e2e, but the result matrix is allocated inside every call (an R call allocates its result): "
                                               "the difference is the kernel zeroing and mapping the fresh pages"},
               "e2e_pinned": {"value": Rt / t_pin, "unit": "iter/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": int(res_pin.stats["d2h_bytes"]),
                              "steps": e2e_steps, "ms_per_step": 1e3 * t_pin, "buffers": "pinned inputs and pinned preallocated outputs ("
                                         + ("int32 on the wire: plain DMA into the caller's matrix, what the engine chooses from 4 GPUs on, where the host's memory is the bottleneck)"
                                            if int(res_pin.stats["d2h_bytes"]) >= 4 * Rt * G else
                                            "uint16 on the wire, widened into the caller's int32 matrix by the host threads: faster than DMA of the int32 matrix up to 2 GPUs)")}}
        e2e["e2e_pinned_uint16"] = {"value": Rt / t_u16, "unit": "iter/s", "d2h_bytes_per_step": int(res_u16.stats["d2h_bytes"]), "ms_per_step": 1e3 * t_u16,
                                    "buffers": "pinned inputs, pinned uint16 result (bootCounts(..., counts_dtype = uint16)): half the bytes into host memory; "
                                               "not what an R caller gets (R's integer is 32-bit) -- reported beside the headline, not as it",
                                    "overflow": bool(res_u16.stats.get("overflow", False))}
        e2e["e2e_moments"] = {"value": Rt / t_mom, "unit": "iter/s", "ms_per_step": 1e3 * t_mom, "d2h_bytes_per_step": int(3 * G * 8 + 2 * Rt * 8),
                              "api": "nullranges_b200.bootMoments(y, x, blockLength, R = 1000 * N, gpus = N): the same bootstrap and counts, reduced to per-feature "
                                     "mean / sd / z / empirical p on the GPUs (no [R x G] matrix crosses PCIe)"}
        assert np.array_equal(res_mom.sum, res_pg.feature_counts.sum(axis=0))  # the same counts, summed over the iterations
        assert np.array_equal(res_fresh.feature_counts, res_pg.feature_counts)
        assert np.array_equal(res_u16.feature_counts, res_pg.feature_counts)
        del res_fresh, res_u16, res_mom
        del res_pg, res_pin, one
    host_barrier()

    # ---- max over ranks
    t = torch.tensor([dev_ms, kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, kernel_ms_max = (float(v) for v in t.tolist())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
        groups = max(launches // 2, 1)  # one launch group = rows kernel + count kernel over <= 2048 iterations
        count_s = count_ms / groups * 1e-3
        W = int(n_windows)
        # the rank path per launch (R iterations): what MUST cross HBM -- the count matrix out, the draw table in (the rows
        # kernel left it in the L2 / HBM), the index (tables + records) once
        table_entries = int(np.sum(y.seqlengths)) >> 10  # one bucket per 2^10 positions at this size (index.cuh, build_rank_tables)
        index_bytes = 2 * 8 * table_entries + 16 * W + 8 * G
        min_dram = (4.0 * G + 8.0 * B) * R + index_bytes
        # what must cross the L2 -> SM crossbar, in 32-byte sectors: two table entries and half a record sector per
        # evaluation, 12 bytes per feature (list bounds, caller's position, count out), the tile's draw
        sectors = R * (2.5 * W + 0.375 * G + 0.25 * B)
        alg8d = 8.0 * hits + (16.0 * B + 4.0 * G + 16.0) * R * args.steps
        stream_alg = 8.0 * stream_hits + (16.0 * B + 4.0 * G + 16.0) * R * stream_steps
        stream_achieved = stream_alg / (stream_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            traffic, traffic_src = tj.get("boot_rank_loop_kernel", {}).get("dram_bytes_per_launch"), tj.get("source")
        except Exception:
            pass
        # CPU baseline: bounded sample on this box's cores
        cores = host_cores()
        O, plan, yi, qi = cpu_objects(y, x)
        n_cpu = max(cores, 16)
        probe = cpu_sample(O, plan, yi, qi, n_cpu, cores)
        n_cpu = int(min(R_STEP, max(n_cpu, round(12.0 / (probe / n_cpu)))))
        cpu_t = cpu_sample(O, plan, yi, qi, n_cpu, cores, iter0=100)
        n_one = max(4, n_cpu // (4 * cores))
        cpu_t1 = cpu_sample(O, plan, yi, qi, n_one, 1, iter0=100)
        line = {
            "metric": METRIC, "value": R * world * args.steps / (dev_ms * 1e-3), "unit": "iter/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config_dict(),
            "run": {"n_blocks": int(B), "windows_per_iteration": W,
                    "sharding": "iterations, rank-major; [2 x R] int64 per-iteration vectors all-gathered over NCCL each step" if world > 1 else "single GPU",
                    "host_threads": host_cores(), "wall_s_timed_region": t_wall},
            **e2e,
            "gpu_launches": int(launches),
            "checks": checks,
            "roofline": {"bound": "hbm", "kernel": "boot_rank_loop_kernel (the rank path's count kernel, dominant)",
                         "achieved": min_dram / count_s / 1e9, "peak": peak, "unit": "GB/s", "frac": min_dram / count_s / 1e9 / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": min_dram,
                         "bytes_model": "per launch of R iterations: (4*G + 8*B)*R (count matrix out, draw table in) + the index once (rank tables, window records)",
                         "kernel_ms_per_launch": {"count": count_ms / groups, "rows": rows_ms / groups, "both": kernel_ms / groups},
                         "kernel_share_of_step": (count_ms / args.steps) / (dev_ms / args.steps),
                         "timing": "CUDA events inside libnrb200 around each kernel, accounting pass over the same steps right after the timed region",
                         "note": "this kernel is not HBM-bound: it counts by rank differences in L2-resident tables (two scattered 8-byte probes per "
                                 "(feature, tile, iteration) evaluation); what bounds it is in roofline_l2"},
            "roofline_l2": {"bound": "32-byte sectors through the L2 (scattered 8-byte loads)", "achieved": sectors * 32.0 / count_s / 1e9,
                            "peak": l2_peak, "unit": "GB/s", "frac": sectors * 32.0 / count_s / 1e9 / l2_peak if l2_peak else None,
                            "peak_source": "nrb200_measure_l2_sector_rate (64 MiB buffer) on this GPU, this run",
                            "sectors_model": "R * (2.5 * windows + 0.375 * G + 0.25 * B): two table entries + a 16-byte record per evaluation, 12 bytes per feature, 8 bytes per tile",
                            "sectors_per_launch": sectors, "evaluations_per_launch": W * R,
                            "sm_cycles_per_evaluation": count_s * 148 * (clk.get("sm_mhz") or 1965) * 1e6 / max(W * R, 1)},
            "roofline_model_8d": {"bound": "hbm", "achieved": alg8d / (kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": alg8d / (kernel_ms * 1e-3) / 1e9 / peak,
                                  "note": "SURVEY 8d bytes (8*H + 16*B + 4*G + 16 per iteration) over rows + count kernel time; above 1 because the "
                                          "rank path never reads the H sampled ranges -- kept for comparison, not a bound"},
            "roofline_stream": {"bound": "hbm", "kernel": "boot_stream_kernel (NRB200_FLAG_FORCE_STREAM, same workload): streams the sampled ranges, the 8d model's kernel",
                                "achieved": stream_achieved, "peak": peak, "unit": "GB/s", "frac": stream_achieved / peak,
                                "kernel_ms_per_launch": stream_ms / stream_steps, "iters_per_s": R * stream_steps / (stream_ms * 1e-3)},
            "cpu_baseline": {"value": n_cpu / cpu_t, "unit": "iter/s", "cores": cores, "kind": "port",
                             "single_thread_value": n_one / cpu_t1, "single_thread_sample": f"{n_one} iterations on 1 thread (R itself is single-threaded)",
                             "sample": f"{n_cpu} iterations of the same workload (same ranges/query/blockLength, Philox draws), oracle/nr_oracle.c on {cores} pthreads"},
            "clocks": clk,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    host_barrier()
    if world > 1:
        dist.destroy_process_group()
    eng.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
