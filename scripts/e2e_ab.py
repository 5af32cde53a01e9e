"""A/B of the end-to-end step under tuning knobs, alternating on the SAME box: usage e2e_ab.py KEY=a,b [KEY=...]"""
import sys, os, subprocess, json
here = os.path.dirname(os.path.abspath(__file__))
CHILD = r'''
import sys, os, time
sys.path.insert(0, os.path.dirname(%r))
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
ts = []
for rep in range(40):
    t0 = time.perf_counter(); BR.bootCounts(y, x, 500000, R=R, rng="philox", seed=1, iter0=rep * R, out=out); ts.append(time.perf_counter() - t0)
ts = sorted(ts[5:])
print("%%.4f %%.4f" %% (1e3 * ts[0], 1e3 * ts[len(ts) // 2]))
''' % here
variants = [{}]
for arg in sys.argv[1:]:
    k, vs = arg.split("=")
    variants = [dict(v, **{k: val}) for v in variants for val in vs.split(",")]
for rnd in range(3):
    for v in variants:
        env = dict(os.environ, **v)
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print(rnd, v, "min/median ms:", r.stdout.strip() or r.stderr[-300:], flush=True)
