#!/usr/bin/env python
"""bench.py -- bootstrap iterations/sec of the fused bootRanges -> countOverlaps path.

Workload (BASELINE.json configs[1], SURVEY 8d cfg 2): unsegmented across-chromosome block
bootstrap of 1M synthetic ATAC-like peaks on hg38 chr1-22, blockLength = 5e5, against 20k
synthetic genes; per-iteration totals AND the [R x 20k] per-gene count matrix.
One "step" = R = 1000 bootstrap iterations (one nrb200_run_counts* call).

  value   iterations/s, device-resident (ranges, query and plan already in HBM; block draws are
          Philox on the device, so a step has no input copy), CUDA events, max over ranks
  e2e     iterations/s through the public API bootCounts(y, x, ...) from HOST arrays: every
          step uploads y and x, builds the indexes, runs, and copies totals + the count matrix
          back to pinned host memory
  roofline  algorithmic bytes (8*H + 16*B + 4*G + 16 per iteration) / kernel time
  cpu_baseline  the CPU oracle (plain-C port of the reference algorithm) on this box's cores

--impl reference: the reference's CPU implementation of the same path.  The reference is pure
R and R is not installed, so this arm times the oracle port (oracle/nr_oracle.c) with all host
threads on a bounded sample (iterations per step reduced; same ranges, query and block length).

Multi-GPU (torchrun): iterations shard across ranks with no data-path collective; the small
per-iteration vectors are all-gathered over NCCL at the end of each step (weak scaling: every
rank runs R iterations per step).
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
    os.environ["OMP_NUM_THREADS"] = str(max(1, _cores // _lws))

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
        "config": {"workload": WORKLOAD, "iters_per_step": per_step,
                   "note": "reference is pure R (not installable here: no R); this is the plain-C oracle port of its algorithm, faster than real R"},
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


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from nullranges_b200 import Draws, Engine, _lib, bootCounts
    from nullranges_b200.sharding import resident_tensors

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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

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

    # ---- the streaming kernel on the same workload (the path that really reads the 8*H bytes)
    stream_ms, stream_hits, stream_steps = 0.0, 0, 3
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

    # ---- end-to-end through the public API: pinned host arrays in, pinned host arrays out
    from nullranges_b200 import GRanges

    def pinned_copy(a):
        keep, view = pinned(a.shape, {np.dtype(np.int32): torch.int32, np.dtype(np.int8): torch.int8}[a.dtype])
        view[...] = a
        pinned_keep.append(keep)
        return view

    pinned_keep = []
    for gr in (y, x):
        gr.seqid, gr.start, gr.end = pinned_copy(gr.seqid), pinned_copy(gr.start), pinned_copy(gr.end)
    keep_t, out_tot = pinned((R,), torch.int64)
    keep_n, out_nr = pinned((R,), torch.int64)
    keep_c, out_cnt = pinned((R, G), torch.int32)
    e2e_steps = max(1, min(args.steps, 5))

    def step_e2e(step):
        return bootCounts(y, x, L_B, R=R, rng="philox", seed=2024, iter0=(step * world + rank) * R, device=local,
                          out={"iter_totals": out_tot, "iter_nranges": out_nr, "feature_counts": out_cnt})

    step_e2e(0)
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        res = step_e2e(1 + s)
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = int(y.seqid.nbytes + y.start.nbytes + y.end.nbytes + x.seqid.nbytes + x.start.nbytes + x.end.nbytes)
    d2h = int(out_tot.nbytes + out_nr.nbytes + out_cnt.nbytes)
    assert int(res.feature_counts.sum()) == int(res.iter_totals.sum())

    # ---- max over ranks
    t = torch.tensor([dev_ms, e2e_s, kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s, kernel_ms_max = (float(v) for v in t.tolist())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
        alg_bytes = 8.0 * hits + (16.0 * B + 4.0 * G + 16.0) * R * args.steps  # this rank, whole timed region
        groups = max(launches // 2, 1)  # one launch group = rows kernel + count kernel over <= 2048 iterations
        per_launch = alg_bytes / groups
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        stream_alg = 8.0 * stream_hits + (16.0 * B + 4.0 * G + 16.0) * R * stream_steps
        stream_achieved = stream_alg / (stream_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("rank_launch_group_dram_bytes_per_launch")
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
            "config": {"workload": WORKLOAD, "iters_per_step_per_gpu": R, "n_blocks": int(B), "rng": "philox4x32-10 on device",
                       "l2": "flushed between timed steps (256 MiB fill, outside the step events); inside a step the 12 MB range table is re-read by design",
                       "sharding": "iterations, rank-major; [2 x R] int64 per-iteration vectors all-gathered over NCCL each step" if world > 1 else "single GPU",
                       "host_threads_per_rank": int(os.environ.get("OMP_NUM_THREADS", "0") or 0) or host_cores(),
                       "wall_s_timed_region": t_wall},
            "e2e": {"value": R * world * e2e_steps / e2e_s, "unit": "iter/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "nullranges_b200.bootCounts(y, x, blockLength, R) from host arrays (upload + index build + run + D2H every step)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "boot_rank_count_kernel (dominant) + boot_block_rows_kernel: one launch of each per step",
                         "kernel_ms_per_launch": {"boot_rank_count_kernel": count_ms / groups, "boot_block_rows_kernel": rows_ms / groups,
                                                  "both": kernel_ms / groups},
                         "algorithmic_bytes_per_launch": per_launch, "peak_source": peak_src,
                         "kernel_share_of_step": kernel_ms / dev_ms if world == 1 else kernel_ms_max / dev_ms,
                         "timing": "kernel times: CUDA events inside libnrb200 around each kernel, accounting pass over the same steps right after the timed region",
                         "note": "bytes = SURVEY 8d model (8*H + 16*B + 4*G + 16 per iteration, H = sampled ranges). The rank path counts by "
                                 "rank differences and never reads the H ranges, so it legitimately exceeds 1.0 on this scale (SURVEY 8d); "
                                 "its own limit is L1TEX wavefronts of scattered table probes, see DESIGN.md 4a. roofline_stream is the kernel that does read them."},
            "roofline_own": {"bound": "l1tex wavefronts (scattered 128-byte-line accesses), not HBM",
                             "evaluations_per_launch": n_windows * R, "windows_per_iteration": n_windows,
                             "sm_cycles_per_evaluation": (count_ms / groups) * 1e-3 * 148 * (clk.get("sm_mhz") or 1965) * 1e6 / max(n_windows * R, 1),
                             "floor": "2 rank-table probes per evaluation = 2 line accesses at ~1 line/cycle/SM; measured ~3.4 accesses per evaluation (ncu l1tex sectors), see DESIGN.md 4a"},
            "roofline_stream": {"bound": "hbm", "kernel": "boot_count_kernel<0,1,0> (NRB200_FLAG_FORCE_STREAM, same workload)",
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
    if world > 1:
        dist.barrier()
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
