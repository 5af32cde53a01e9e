"""Wall-clock timeline of one end-to-end bootCounts() step (pinned host arrays in and out, as bench.py's e2e leg).
Every Engine method and front-end helper is wrapped with perf_counter stamps; nothing synchronises in between."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nullranges_b200 import synth, bootranges as BR
from nullranges_b200.engine import Engine

y, x = synth.atac_peaks(), synth.genes()
R, G = 1000, len(x)
keep = []
def pin(a):
    t = torch.empty(a.shape, dtype={np.dtype(np.int32): torch.int32, np.dtype(np.int8): torch.int8, np.dtype(np.int64): torch.int64}[a.dtype], pin_memory=True)
    keep.append(t); v = t.numpy(); v[...] = a; return v
for gr in (y, x):
    gr.seqid, gr.start, gr.end = pin(gr.seqid), pin(gr.start), pin(gr.end)
out = {k: pin(np.zeros(s, d)) for k, s, d in (("iter_totals", (R,), np.int64), ("iter_nranges", (R,), np.int64), ("feature_counts", (R, G), np.int32))}

log = []
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); log.append((name, t0, time.perf_counter())); return r
    setattr(obj, name, g)
for m in ("set_ranges", "ranges_info", "set_query", "set_exclude", "clear_exclude", "set_plan", "run_counts"):
    wrap(Engine, m)
for m in ("_side_arrays", "_setup", "_make_draws"):
    wrap(BR, m)
for rep in range(6):
    log.clear()
    t0 = time.perf_counter(); BR.bootCounts(y, x, 500000, R=R, rng="philox", seed=1, iter0=rep * R, out=out); t1 = time.perf_counter()
    if rep >= 3:
        print(f"--- rep {rep}: bootCounts {1e3 * (t1 - t0):.3f} ms")
        for name, a, b in sorted(log, key=lambda r: r[1]):
            print(f"  {1e3 * (a - t0):7.3f} -> {1e3 * (b - t0):7.3f}  ({1e3 * (b - a):6.3f})  {name}")
