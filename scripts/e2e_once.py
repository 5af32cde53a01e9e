"""Three end-to-end bootCounts() steps (pinned arrays), for profiling the per-call kernels (index build included)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nullranges_b200 import synth, bootranges as BR
y, x = synth.atac_peaks(), synth.genes()
R, G = 1000, len(x)
keep = []
def pin(a):
    t = torch.empty(a.shape, dtype={np.dtype(np.int32): torch.int32, np.dtype(np.int8): torch.int8, np.dtype(np.int64): torch.int64}[a.dtype], pin_memory=True)
    keep.append(t); v = t.numpy(); v[...] = a; return v
for gr in (y, x):
    gr.seqid, gr.start, gr.end = pin(gr.seqid), pin(gr.start), pin(gr.end)
out = {k: pin(np.zeros(s, d)) for k, s, d in (("iter_totals", (R,), np.int64), ("iter_nranges", (R,), np.int64), ("feature_counts", (R, G), np.int32))}
for rep in range(3):
    BR.bootCounts(y, x, 500000, R=R, rng="philox", seed=1, iter0=rep * R, out=out)
print("ok")
