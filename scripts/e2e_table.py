"""End-to-end time of the table path: bootRanges() on cfg-2 ranges (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nullranges_b200 import synth, bootRanges
y = synth.atac_peaks(1_000_000, seed=3)
for R in (5, 20):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        br = bootRanges(y, 500_000, R=R, rng="philox", seed=1)
        best = min(best, time.perf_counter() - t0)
    print(f"bootRanges 1M ranges R={R}: {len(br)} rows, {best*1e3:.1f} ms, {len(br)/best/1e6:.1f} M rows/s")
