import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from nullranges_b200 import synth, bootranges as BR
y, x = synth.atac_peaks(), synth.genes()
def pin(a):
    t = torch.empty(a.shape, dtype={np.dtype(np.int32): torch.int32, np.dtype(np.int8): torch.int8}[a.dtype], pin_memory=True)
    v = t.numpy(); v[...] = a; keep.append(t); return v
keep = []
for g in (y, x):
    g.seqid, g.start, g.end = pin(g.seqid), pin(g.start), pin(g.end)
R, G = 1000, len(x)
out = {k: torch.empty(s, dtype=d, pin_memory=True).numpy() for k, s, d in
       (("iter_totals", (R,), torch.int64), ("iter_nranges", (R,), torch.int64), ("feature_counts", (R, G), torch.int32))}
for _ in range(3):
    BR.bootCounts(y, x, 500000, R=R, rng="philox", seed=1, out=out)
t0 = time.perf_counter()
for _ in range(10):
    BR.bootCounts(y, x, 500000, R=R, rng="philox", seed=1, out=out)
print("ms per call", (time.perf_counter() - t0) * 100)
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    BR.bootCounts(y, x, 500000, R=R, rng="philox", seed=1, out=out)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
