"""Device-resident throughput of every BASELINE.json config on one B200 (auxiliary to bench.py,
which reports configs[1]).  Prints one JSON line per config.  Run on the GPU box."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nullranges_b200 import synth, Engine, Draws, _lib

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
force_stream = bool(int(os.environ.get("FORCE_STREAM", "0"))) or "--stream" in sys.argv
only = set(a for a in sys.argv[1:] if not a.startswith("--"))
eng = Engine(0, force_stream=force_stream)


def report(name, R, B, G, st, extra=None):
    ms = st["kernel_ms"]
    alg = 8.0 * st["n_hits"] + (16.0 * B + 4.0 * G + 16.0) * R
    line = {"config": name, "path": "stream" if force_stream else "default", "iters": R, "kernel_ms": round(ms, 3),
            "iters_per_s": round(R / ms * 1e3, 1), "launches": st["n_launches"], "hits_per_iter": st["n_hits"] / max(R, 1),
            "algorithmic_GBps": round(alg / ms / 1e6, 1), "frac_of_measured_hbm": round(alg / ms / 1e6 / PEAK, 3)}
    if not force_stream and st.get("rows_kernel_ms"):
        line.update({"rows_kernel_ms": round(st["rows_kernel_ms"], 3), "count_kernel_ms": round(st["count_kernel_ms"], 3),
                     "windows_per_iter": st.get("n_windows")})
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def best(f, n=3):
    out = None
    for _ in range(n):
        st = f()
        if out is None or st["kernel_ms"] < out["kernel_ms"]:
            out = st
    return out


if not only or "1" in only:
    y, x = synth.uniform_ranges(10_000, seed=1), synth.genes(20_000, seed=2)
    eng.set_ranges(y.seqlengths, y.seqid, y.start, y.end); eng.clear_exclude(); eng.set_query(x.seqid, x.start, x.end)
    B = eng.set_plan(_lib.UNSEG_ACROSS, 500_000)
    report("cfg1 10k ranges, R=50", 50, B, len(x), best(lambda: eng.run_counts_resident(Draws(n_iter=50, seed=1), True, True)))

if not only or only & {"2", "3", "5"}:
    y, x = synth.atac_peaks(1_000_000, seed=3), synth.genes(20_000, seed=2)
    eng.set_ranges(y.seqlengths, y.seqid, y.start, y.end); eng.clear_exclude(); eng.set_query(x.seqid, x.start, x.end)
    if not only or "2" in only:
        B = eng.set_plan(_lib.UNSEG_ACROSS, 500_000)
        report("cfg2 1M ranges unsegmented, R=1000", 1000, B, len(x), best(lambda: eng.run_counts_resident(Draws(n_iter=1000, seed=1), True, True)))
    if not only or "5" in only:
        for Lb in (100_000, 200_000, 500_000, 1_000_000):
            B = eng.set_plan(_lib.UNSEG_ACROSS, Lb)
            res = eng.run_counts(Draws(n_iter=1000, seed=1), want_counts=False)
            st = best(lambda: eng.run_counts_resident(Draws(n_iter=1000, seed=1), True, True))
            report(f"cfg5 blockLength={Lb}, R=1000", 1000, B, len(x), st, {"n_blocks": B, "var_of_iter_totals": float(np.var(res["iter_totals"], ddof=1)),
                                                                         "mean_of_iter_totals": float(res["iter_totals"].mean())})
    if not only or "3" in only:
        seg, ex = synth.segmentation(seed=4), synth.exclude_regions(seed=5)
        eng.set_exclude(ex.seqid, ex.start, ex.end)
        B = eng.set_plan(_lib.SEG_PROP, 500_000, {"seqid": seg.seqid, "start": seg.start, "end": seg.end, "state": seg.mcols["state"]})
        report("cfg3 1M ranges segmented prop + exclude drop, R=1000", 1000, B, len(x),
               best(lambda: eng.run_counts_resident(Draws(n_iter=1000, seed=1), True, True)), {"n_blocks": B, "n_seg": len(seg), "n_exclude": len(ex)})
        eng.clear_exclude()

if not only or "4" in only:
    t0 = time.time()
    y, x = synth.points(10_000_000, seed=6), synth.enhancers(200_000, seed=7)
    eng.set_ranges(y.seqlengths, y.seqid, y.start, y.end); eng.clear_exclude(); eng.set_query(x.seqid, x.start, x.end)
    B = eng.set_plan(_lib.UNSEG_WITHIN, 500_000)
    obs, tot = eng.count_observed()
    R = 10_000 if not force_stream else 200
    m = eng.run_moments(Draws(n_iter=R, seed=1), observed=obs)
    report(f"cfg4 10M points withinChrom vs 200k enhancers, moments, R={R}", R, B, len(x), m["stats"],
           {"observed_total": int(tot), "mean_boot_total": float(m["iter_totals"].mean()), "setup_s": round(time.time() - t0, 1)})

if not only or "m" in only:
    # materialised BootRanges table (the literal drop-in return value): 24*H + 16*B bytes per iteration
    y = synth.atac_peaks(1_000_000, seed=3)
    eng.set_ranges(y.seqlengths, y.seqid, y.start, y.end); eng.clear_exclude(); eng.clear_query()
    B = eng.set_plan(_lib.UNSEG_ACROSS, 500_000)
    import ctypes as C
    from nullranges_b200._lib import Stats, ptr
    for R in (20, 100):
        best_ms, rows = None, 0
        for _ in range(3):
            offs = np.zeros(R + 1, np.int64); st = Stats(); d = Draws(n_iter=R, seed=1).to_c()
            eng._check(eng._lib.nrb200_run_materialize(eng._h, C.byref(d), ptr(offs, C.c_int64), C.byref(st)))
            rows = int(offs[-1])
            best_ms = st.kernel_ms if best_ms is None else min(best_ms, st.kernel_ms)
        alg = 24.0 * rows + 16.0 * B * R  # read 8*H, write 16*H (+ we also write seqid: 20*H)
        print(json.dumps({"config": f"materialise cfg2 ranges, R={R}", "rows": rows, "kernel_ms": round(best_ms, 3),
                          "iters_per_s": round(R / best_ms * 1e3, 1), "algorithmic_GBps": round(alg / best_ms / 1e6, 1),
                          "frac_of_measured_hbm": round(alg / best_ms / 1e6 / PEAK, 3),
                          "actual_bytes_GBps": round((28.0 * rows) / best_ms / 1e6, 1),
                          "written_GBps": round((20.0 * rows) / best_ms / 1e6, 1),
                          "note": "write-dominated: 20 of the 28 bytes per row are stores; a pure device fill (torch fill_ of 256 MiB) runs at ~3850 GB/s on this GPU"}), flush=True)
