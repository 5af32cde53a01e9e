"""Development probe of the end-to-end bootCounts() step and of the resident kernels (cfg 2), one script with modes
(run on the GPU box; every line it prints is JSON):

  python scripts/e2e_probe.py step [--gpus N] [--dest pinned|pageable|fresh] [--reps K]
      wall clock of bootCounts(y, x, 5e5, R = 1000 * N) from host arrays: dest = pinned result buffers, pageable
      buffers reused across calls, or freshly allocated pageable buffers every call (what an R caller has)
  python scripts/e2e_probe.py kernels
      rows / count kernel ms per 1000 iterations, device-resident
  python scripts/e2e_probe.py sweep
      both of the above under the tuning knobs (NRB200_RPT, NRB200_WIRE16, NRB200_CHUNKS), one child process each
  python scripts/e2e_probe.py cfg4 [--gpus N]
      BASELINE configs[3] through bootMoments(..., gpus=N): wall clock of the whole call
  python scripts/e2e_probe.py cfg1
      BASELINE configs[0] end to end: the materialised bootRanges() table and the fused bootCounts(), R's RNG
  python scripts/e2e_probe.py timeline
      NRB200_TRACE_DEV timeline of one step (host queue time vs device arrival of every stage)
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L_B, R1 = 500_000, 1000


def inputs(pin: bool):
    import numpy as np
    from nullranges_b200 import synth
    y, x = synth.atac_peaks(1_000_000, seed=3), synth.genes(20_000, seed=2)
    keep = []
    if pin:
        import torch
        for gr in (y, x):
            for k in ("seqid", "start", "end"):
                a = getattr(gr, k)
                t = torch.empty(a.shape, dtype=torch.int32, pin_memory=True)
                t.numpy()[...] = a
                keep.append(t)
                setattr(gr, k, t.numpy())
    return y, x, keep


def engine_laps():
    """--detail: wraps every engine call the front ends make; returns the list that collects (call, start ms, duration ms),
    start relative to the first call since the list was last cleared."""
    from nullranges_b200 import engine as _e
    laps = []
    for name in ("set_ranges", "set_exclude", "clear_exclude", "set_plan", "set_query", "finish_ranges", "ranges_info", "count_observed",
                 "run_moments", "run_counts"):
        def wrap(f, name=name):
            def g(*a, **k):
                t = time.perf_counter()
                if not laps:
                    laps.append(("t0", t, 0.0))
                try:
                    return f(*a, **k)
                finally:
                    laps.append((name, round(1e3 * (t - laps[0][1]), 3), round(1e3 * (time.perf_counter() - t), 3)))
            return g
        setattr(_e.Engine, name, wrap(getattr(_e.Engine, name)))
    return laps


def step(args):
    import numpy as np
    import torch
    from nullranges_b200 import bootCounts
    y, x, keep = inputs(pin=args.dest == "pinned")
    R, G = R1 * args.gpus, len(x)
    out = None
    if args.dest == "pinned":
        ts = {k: torch.empty(s, dtype=d, pin_memory=True) for k, s, d in (("iter_totals", (R,), torch.int64), ("iter_nranges", (R,), torch.int64),
                                                                           ("feature_counts", (R, G), torch.int32))}
        out = {k: t.numpy() for k, t in ts.items()}
    elif args.dest == "pageable":
        out = {"iter_totals": np.zeros(R, np.int64), "iter_nranges": np.zeros(R, np.int64), "feature_counts": np.zeros((R, G), np.int32)}
    gpus = args.gpus if args.gpus > 1 else None
    times, check = [], None
    laps = engine_laps() if args.detail else []
    for rep in range(args.reps + 2):
        laps.clear()
        t0 = time.perf_counter()
        res = bootCounts(y, x, L_B, R=R, rng="philox", seed=2024, iter0=0, out=out, gpus=gpus)
        times.append(time.perf_counter() - t0)
        last = (t0, times[-1])
        check = int(res.iter_totals.sum()), int(res.feature_counts[:: max(R // 7, 1)].sum())
        del res
    times = sorted(times[2:])
    print(json.dumps({"mode": "step", "gpus": args.gpus, "dest": args.dest, "R": R, "ms_min": 1e3 * times[0], "ms_median": 1e3 * times[len(times) // 2],
                      "iters_per_s_median": R / times[len(times) // 2], "checksum": check,
                      "env": {k: v for k, v in os.environ.items() if k.startswith("NRB200_")},
                      **({"last_call_ms": round(1e3 * last[1], 3), "first_engine_call_at_ms": round(1e3 * (laps[0][1] - last[0]), 3),
                          "last_call_laps_ms": laps[1:]} if args.detail else {})}), flush=True)


def kernels(args):
    from nullranges_b200 import Draws, Engine, _lib
    y, x, _ = inputs(pin=False)
    e = Engine(0)
    e.set_ranges(y.seqlengths, y.seqid, y.start, y.end, None)
    e.set_query(x.seqid, x.start, x.end, None)
    e.set_plan(_lib.UNSEG_ACROSS, L_B)
    best = None
    for rep in range(8):
        st = e.run_counts_resident(Draws(n_iter=R1, rng_mode=_lib.RNG_PHILOX, seed=2024, iter0=rep * R1), True, True)
        if rep >= 2 and (best is None or st["kernel_ms"] < best["kernel_ms"]):
            best = st
    print(json.dumps({"mode": "kernels", "rows_ms": best["rows_kernel_ms"], "count_ms": best["count_kernel_ms"], "both_ms": best["kernel_ms"],
                      "windows": best["n_windows"], "env": {k: v for k, v in os.environ.items() if k.startswith("NRB200_")}}), flush=True)


def sweep(args):
    def child(mode, env, *extra):
        out = subprocess.run([sys.executable, __file__, mode, *extra], env=dict(os.environ, **env), capture_output=True, text=True)
        print(out.stdout.strip() or json.dumps({"failed": mode, "env": env, "stderr": out.stderr[-400:]}), flush=True)
    for rpt in ("1", "2", "4"):
        child("kernels", {"NRB200_RPT": rpt})
    for dest in ("pinned", "pageable", "fresh"):
        for wire in ("0", "1"):
            child("step", {"NRB200_WIRE16": wire}, "--dest", dest)
    for chunks in ("4", "16"):
        child("step", {"NRB200_CHUNKS": chunks}, "--dest", "fresh")


def timeline(args):
    os.environ["NRB200_TRACE_DEV"] = "1"
    os.environ["NRB200_TRACE"] = "1"
    args.reps = 1
    step(args)


def cfg4(args):
    """BASELINE configs[3]: 10M SNP-scale points vs 200k enhancers, withinChrom, R = 10^4, per-feature moments through
    bootMoments(..., gpus=N): wall clock of the whole call from host arrays (upload, index, lists, run, moments back)."""
    import numpy as np
    from nullranges_b200 import bootMoments, synth
    y, x = synth.points(10_000_000, seed=6), synth.enhancers(200_000, seed=7)
    gpus = args.gpus if args.gpus > 1 else None
    times, sig = [], None
    laps = engine_laps() if args.detail else []
    for rep in range(args.reps + 1):
        laps.clear()
        t0 = time.perf_counter()
        m = bootMoments(y, x, L_B, R=10_000, withinChrom=True, rng="philox", seed=1, gpus=gpus)
        last = time.perf_counter() - t0
        times.append(last)
        sig = (int(m.sum.sum()), int(m.sumsq.sum() % (1 << 61)), float(np.nansum(m.p_ge)))
        kernel_ms = m.stats["kernel_ms"]
    times = sorted(times[1:])
    print(json.dumps({"mode": "cfg4", "gpus": args.gpus, "R": 10_000, "wall_ms_min": 1e3 * times[0], "wall_ms_median": 1e3 * times[len(times) // 2],
                      "kernel_ms_slowest_gpu": kernel_ms, "signature": sig, "env": {k: v for k, v in os.environ.items() if k.startswith("NRB200_")},
                      **({"last_call_laps_ms": laps[1:], "last_call_ms": round(1e3 * last, 3)} if args.detail else {})}), flush=True)


def cfg1(args):
    """BASELINE configs[0] end to end: bootRanges(y, 5e5, R = 50) on 10k ranges -- the literal drop-in call, materialised
    BootRanges table with its iter column -- and the fused bootCounts() against 20k genes, under R's RNG (set.seed(1))."""
    from nullranges_b200 import bootCounts, bootRanges, synth
    y, x = synth.uniform_ranges(10_000, seed=1), synth.genes(20_000, seed=2)
    out = {"mode": "cfg1"}
    for rng in ("R", "philox"):
        t_tab, t_cnt, rows = [], [], 0
        for rep in range(args.reps + 2):
            t0 = time.perf_counter()
            br = bootRanges(y, L_B, R=50, rng=rng, seed=1)
            t1 = time.perf_counter()
            bc = bootCounts(y, x, L_B, R=50, rng=rng, seed=1)
            t2 = time.perf_counter()
            t_tab.append(t1 - t0)
            t_cnt.append(t2 - t1)
            rows = len(br.start)
        t_tab, t_cnt = sorted(t_tab[2:]), sorted(t_cnt[2:])
        out[rng] = {"rows": rows, "bootRanges_ms_median": 1e3 * t_tab[len(t_tab) // 2], "bootCounts_ms_median": 1e3 * t_cnt[len(t_cnt) // 2],
                    "iters_per_s_table": 50 / t_tab[len(t_tab) // 2], "iters_per_s_counts": 50 / t_cnt[len(t_cnt) // 2],
                    "totals_checksum": int(bc.iter_totals.sum())}
    out["note"] = ("rng = 'R' (validation mode): the 50 x 5,760 block draws are made on the host by the R-compatible RNG inside the timed call; "
                   "rng = 'philox' (production mode): drawn on the GPU")
    print(json.dumps(out), flush=True)


def calib(args):
    from nullranges_b200 import Engine
    e = Engine(0)
    print(json.dumps({"mode": "calib", "thp": open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(),
                      "l2_sector_rate_GBps": {f"{mb}MiB": e.measure_l2_sector_rate(mb << 20) for mb in (8, 32, 64, 256, 1024)}}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["step", "kernels", "sweep", "timeline", "calib", "cfg4", "cfg1"])
    ap.add_argument("--detail", action="store_true", help="step / cfg4: start and duration of every engine call of the last repetition")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--dest", default="pinned", choices=["pinned", "pageable", "fresh"])
    ap.add_argument("--reps", type=int, default=9)
    a = ap.parse_args()
    {"step": step, "kernels": kernels, "sweep": sweep, "timeline": timeline, "calib": calib, "cfg4": cfg4, "cfg1": cfg1}[a.mode](a)
