"""End-to-end run_counts() into pageable (plain numpy, R-like) result buffers, fresh allocation every call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nullranges_b200 import synth, Engine, Draws, _lib, bootRanges
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
y, x = synth.atac_peaks(), synth.genes()
e = Engine(0)
e.set_ranges(y.seqlengths, y.seqid, y.start, y.end, None)
e.set_query(x.seqid, x.start, x.end, None)
e.set_plan(_lib.UNSEG_ACROSS, 500000)
ts = []
for rep in range(12):
    d = Draws(n_iter=1000, rng_mode=_lib.RNG_PHILOX, seed=1, iter0=rep * 1000)
    t0 = time.perf_counter(); r = e.run_counts(d, True); ts.append(time.perf_counter() - t0)
    s = int(r["feature_counts"][::97].sum()); del r
ts = sorted(ts[2:])
print(f"run_counts -> fresh pageable [1000 x 20k] int32: min {1e3*ts[0]:.2f} ms  median {1e3*ts[len(ts)//2]:.2f} ms")
ts = []
for rep in range(5):
    t0 = time.perf_counter(); br = bootRanges(y, 500_000, R=20, rng="philox", seed=1); ts.append(time.perf_counter() - t0); del br
print(f"bootRanges R=20 (20M rows, pageable columns): min {1e3*min(ts[1:]):.1f} ms")
