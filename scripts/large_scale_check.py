"""Parity at sizes well beyond the BASELINE configs (run on the GPU box; needs ~8 GB of host memory):
  A. 1e8 width-1 points (equal widths: the sorted-ends index aliases the start index), withinChrom bootstrap
  B. 5e7 ATAC-like ranges (variable widths: the ends are radix-sorted), across-chromosome bootstrap
For each: rank path vs streaming path (two independent implementations) on R iterations, and both vs the CPU oracle
on one iteration; prints one JSON line per case."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nullranges_b200 import synth, Engine, Draws, _lib
from oracle import oracle as O

def run(name, y, x, kind, R):
    t0 = time.perf_counter()
    res = {}
    for path, force in (("rank", False), ("stream", True)):
        e = Engine(0, force_stream=force)
        t1 = time.perf_counter()
        e.set_ranges(y.seqlengths, y.seqid, y.start, y.end, None)
        e.set_query(x.seqid, x.start, x.end, None)
        e.set_plan(kind, 500_000)
        t_load = time.perf_counter() - t1
        r = e.run_counts(Draws(n_iter=R, seed=11))
        res[path] = (r, t_load)
        e.close()
    a, b = res["rank"][0], res["stream"][0]
    same = all(np.array_equal(a[k], b[k]) for k in ("iter_totals", "iter_nranges", "feature_counts"))
    C = len(y.seqlengths)
    plan = O.Plan(kind, y.seqlengths, 500_000, None)
    yi, qi = O.Index(C, y.seqid, y.start, y.end, None), O.Index(C, x.seqid, x.start, x.end, None)
    t1 = time.perf_counter()
    ref = O.boot_counts(plan, yi, qi, 1, seed=11, nthreads=1)
    t_cpu = time.perf_counter() - t1
    ok = all(np.array_equal(a[k][:1], ref[k]) for k in ("iter_totals", "iter_nranges", "feature_counts"))
    print(json.dumps({"case": name, "n_ranges": len(y), "n_query": len(x), "iterations": R, "rank_equals_stream": bool(same),
                      "rank_equals_oracle_iteration_0": bool(ok), "sum_counts_equals_totals": bool(a["feature_counts"].sum() == a["iter_totals"].sum()),
                      "set_ranges_s": round(res["rank"][1], 3), "rank_kernel_ms": round(a["stats"]["kernel_ms"], 3),
                      "stream_kernel_ms": round(b["stats"]["kernel_ms"], 3), "oracle_one_iteration_s": round(t_cpu, 2),
                      "mean_ranges_per_iteration": float(a["iter_nranges"].mean())}), flush=True)

x = synth.enhancers(200_000, seed=7)
run("1e8 points, withinChrom", synth.points(100_000_000, seed=6), x, _lib.UNSEG_WITHIN, 3)
run("5e7 ATAC-like ranges, across chromosomes", synth.atac_peaks(50_000_000, seed=3), synth.genes(20_000, seed=2), _lib.UNSEG_ACROSS, 3)
