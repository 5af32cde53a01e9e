"""Shared builders for the parity tests: random cases + engine-vs-oracle comparison."""
from __future__ import annotations

import numpy as np

from oracle import oracle as O

KINDS = {"across": 0, "within": 1, "seg_prop": 2, "seg_noprop": 3, "permute_within": 4, "permute_across": 5}


def random_case(seed, n=400, n_seq=3, max_len=5000, n_query=60, n_excl=0, stranded=False, wmax=120,
                L_b=500, n_seg=0, n_states=3, sort=True, points=False):
    """Small random genome with ranges, query, optional exclude and segmentation."""
    rng = np.random.default_rng(seed)
    seqlen = rng.integers(max(L_b, max_len // 3), max_len + 1, n_seq).astype(np.int64)

    def ranges(m, wmax):
        sid = rng.integers(0, n_seq, m).astype(np.int32)
        w = np.ones(m, np.int64) if points else rng.integers(1, wmax + 1, m)
        w = np.minimum(w, seqlen[sid] - 1)
        st = 1 + (rng.random(m) * (seqlen[sid] - w)).astype(np.int64)
        strand = rng.integers(0, 3, m).astype(np.int8) if stranded else None
        return sid, st.astype(np.int32), (st + w - 1).astype(np.int32), strand

    y = ranges(n, wmax)
    if sort:
        o = np.lexsort((y[2], y[1], y[0]))
        y = tuple(None if a is None else a[o] for a in y)
    q = ranges(n_query, wmax * 4) if n_query else None
    x = ranges(n_excl, wmax * 2) if n_excl else None
    seg = None
    if n_seg:
        sc, ss, se, st = [], [], [], []
        for c in range(n_seq):
            k = max(1, n_seg // n_seq)
            cuts = np.sort(rng.choice(np.arange(2, seqlen[c]), size=k - 1, replace=False)) if k > 1 else np.array([], np.int64)
            starts = np.concatenate(([1], cuts))
            ends = np.concatenate((cuts - 1, [seqlen[c]]))
            sc += [c] * k
            ss += starts.tolist()
            se += ends.tolist()
            st += rng.integers(1, n_states + 1, k).tolist()
        st[:n_states] = list(range(1, n_states + 1))  # every state present
        seg = {"chrom": np.array(sc, np.int32), "start": np.array(ss, np.int32), "end": np.array(se, np.int32),
               "state": np.array(st, np.int32)}
    return {"seqlen": seqlen, "y": y, "query": q, "excl": x, "seg": seg, "L_b": L_b}


def oracle_objects(case, kind):
    """(plan, y, query, exclude) for the oracle.  With case["trim"] the fourth item is the gaps() index and
    must be passed with exclude_mode=O.EXCLUDE_TRIM (see excl_mode())."""
    C = case["seqlen"].size
    plan = O.Plan(kind, case["seqlen"], case["L_b"], case["seg"] if kind in (2, 3) else None)
    y = O.Index(C, *case["y"])
    q = O.Index(C, *case["query"], maxgap=case.get("maxgap", -1), minoverlap=case.get("minoverlap", 0)) if case["query"] is not None else None
    x = None
    if case["excl"] is not None:
        x = O.gaps(C, case["seqlen"], *case["excl"]) if case.get("trim") else O.Index(C, *case["excl"])
    return plan, y, q, x


def excl_mode(case):
    return O.EXCLUDE_TRIM if case.get("trim") else O.EXCLUDE_DROP


def load_engine(eng, case, kind):
    eng.set_ranges(case["seqlen"], *case["y"])
    if case["query"] is not None:
        eng.set_query(*case["query"], maxgap=case.get("maxgap", -1), minoverlap=case.get("minoverlap", 0))
    else:
        eng.clear_query()
    if case["excl"] is not None:
        eng.set_exclude(*case["excl"], option="trim" if case.get("trim") else "drop")
    else:
        eng.clear_exclude()
    seg = case["seg"]
    segd = None if (seg is None or kind not in (2, 3)) else {"seqid": seg["chrom"], "start": seg["start"], "end": seg["end"], "state": seg["state"]}
    return eng.set_plan(kind, case["L_b"], segd)


def draw_R(plan, seed, R):
    """R-RNG block draws for R iterations.  seg_bootstrap_noprop legitimately fails (in the reference
    too: R/bootRanges.R:303) when a state draws no segment; such seeds are skipped."""
    last = None
    for k in range(50):
        try:
            return plan.draw_R(O.RRng(seed + 1000 * k), R)
        except ValueError as e:
            if "drew no segments" not in str(e):
                raise
            last = e
    raise last


def canon_rows(seqid, start, end, src, block, it):
    """Sort rows on (iter, block, src) -- the contract order made explicit."""
    o = np.lexsort((src, block, it))
    return np.stack([it[o], seqid[o], start[o], end[o], src[o], block[o]], axis=1)


def oracle_rows(plan, y, x, draws, R, mode=O.EXCLUDE_DROP):
    bc, bs, dt = draws
    out = []
    for r in range(R):
        rows = O.iteration_rows(plan, y, bc[r], bs[r], dt[r], exclude=x, exclude_mode=mode)
        out.append(canon_rows(rows["chrom"], rows["start"], rows["end"], rows["src"], rows["block"], np.full(rows["src"].size, r)))
    return np.concatenate(out) if out else np.empty((0, 6), np.int64)


def brute_counts(case, rows):
    """O(rows x G) countOverlaps with explicit strand rule; rows = (seqid, start, end, strand)."""
    q = case["query"]
    qs = q[3] if q[3] is not None else np.zeros(q[0].size, np.int8)
    counts = np.zeros(q[0].size, np.int64)
    for c, s, e, st in zip(*rows):
        if e < s and not O.zero_width_rule():   # THE ZERO-WIDTH RULE (oracle/nr_oracle.c)
            continue
        m = (q[0] == c) & (q[1] <= e) & (q[2] >= s) & ((qs == 0) | (st == 0) | (qs == st))
        counts += m
    return counts
