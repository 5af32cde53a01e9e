"""Where an end-to-end bootCounts() step spends its time (run on the GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nullranges_b200 import synth, bootranges as BR, _lib, Draws
from nullranges_b200.engine import get_engine

y, x = synth.atac_peaks(), synth.genes()
R, G = 1000, len(x)
out = {k: torch.empty(s, dtype=d, pin_memory=True).numpy() for k, s, d in
       (("iter_totals", (R,), torch.int64), ("iter_nranges", (R,), torch.int64), ("feature_counts", (R, G), torch.int32))}
T = {}
def tick(name, t0):
    torch.cuda.synchronize()
    T.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
for rep in range(6):
    t0 = time.perf_counter(); srt = y.is_sorted(); tick("is_sorted", t0)
    eng = get_engine(0)
    t0 = time.perf_counter(); eng.set_ranges(y.seqlengths, y.seqid, y.start, y.end, None); tick("set_ranges", t0)
    t0 = time.perf_counter(); eng.clear_exclude(); eng.set_plan(_lib.UNSEG_ACROSS, 500000); tick("set_plan", t0)
    t0 = time.perf_counter(); qi = BR._side_arrays(x, y, "x"); eng.set_query(qi[0], qi[1], qi[2], None); tick("set_query", t0)
    d = Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=1, iter0=rep * R)
    t0 = time.perf_counter(); st = eng.run_counts_resident(d, True, True); tick("run_resident(incl prepare)", t0)
    t0 = time.perf_counter(); st = eng.run_counts_resident(d, True, True); tick("run_resident(again)", t0)
    t0 = time.perf_counter(); res = eng.run_counts(d, True, out=out); tick("run_counts(host out, pinned)", t0)
    t0 = time.perf_counter(); res = eng.run_counts(d, True); tick("run_counts(host out, pageable)", t0)
    t0 = time.perf_counter(); BR.bootCounts(y, x, 500000, R=R, rng="philox", seed=1, out=out); tick("bootCounts total", t0)
for k, v in T.items():
    print(f"{k:36s} min {min(v):8.3f} ms   median {sorted(v)[len(v)//2]:8.3f} ms")
print(st)
