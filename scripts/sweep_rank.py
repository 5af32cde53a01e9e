"""Tuning sweep of the rank-path kernels on cfg 2 (run on the GPU box): kernel ms per 1000 iterations."""
import itertools, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    from nullranges_b200 import synth, Engine, Draws, _lib
    y, x = synth.atac_peaks(), synth.genes()
    e = Engine(0)
    e.set_ranges(y.seqlengths, y.seqid, y.start, y.end, None)
    e.set_plan(_lib.UNSEG_ACROSS, 500000)
    d = Draws(n_iter=1000, rng_mode=_lib.RNG_PHILOX, seed=1)
    a = min(e.run_counts_resident(d, False, True)["kernel_ms"] for _ in range(5))
    e.set_query(x.seqid, x.start, x.end, None)
    ab = min(e.run_counts_resident(d, True, True)["kernel_ms"] for _ in range(8))
    print(json.dumps({"A_ms": round(a, 4), "AB_ms": round(ab, 4), "B_ms": round(ab - a, 4)}))
    sys.exit(0)
combos = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or None  # e.g. 1,1,4 1,8,4
for rpt, minb, mult in combos or [(4, 1, 4), (4, 3, 4), (4, 4, 4), (2, 1, 4), (2, 4, 4), (1, 1, 4), (8, 1, 4),
                        (2, 1, 1), (2, 1, 2), (2, 1, 8), (4, 3, 1), (4, 3, 2), (1, 1, 1), (1, 1, 2)]:
    env = dict(os.environ, NRB200_RPT=str(rpt), NRB200_MINB=str(minb), NRB200_TABLE_MULT=str(mult))  # NRB200_MINB_ROWS passes through
    out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
    print(f"rpt={rpt} minb={minb} table_mult={mult}: {out.stdout.strip() or out.stderr[-300:]}", flush=True)
