import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NRB200_TRACE"] = "1"
import torch
from nullranges_b200 import synth, bootranges as BR
y, x = synth.atac_peaks(), synth.genes()
R, G = 1000, len(x)
out = {k: torch.empty(s, dtype=d, pin_memory=True).numpy() for k, s, d in
       (("iter_totals", (R,), torch.int64), ("iter_nranges", (R,), torch.int64), ("feature_counts", (R, G), torch.int32))}
for rep in range(3):
    sys.stderr.write(f"--- rep {rep}\n")
    t0 = time.perf_counter(); BR.bootCounts(y, x, 500000, R=R, rng="philox", seed=1, out=out)
    sys.stderr.write(f"bootCounts {1e3*(time.perf_counter()-t0):.3f} ms\n")
