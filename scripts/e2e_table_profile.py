"""Where bootRanges() (materialised table) spends its time end to end (run on the GPU box)."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NRB200_TRACE", "1")
from nullranges_b200 import synth, bootRanges
y = synth.atac_peaks(1_000_000, seed=3)
R = 20
for _ in range(2):
    bootRanges(y, 500_000, R=R, rng="philox", seed=1)
sys.stderr.write("---- profiled call\n")
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable(); br = bootRanges(y, 500_000, R=R, rng="philox", seed=1); pr.disable()
print(f"total {1e3 * (time.perf_counter() - t0):.1f} ms for {len(br)} rows")
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
