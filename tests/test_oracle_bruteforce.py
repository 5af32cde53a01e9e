"""CPU: the oracle's indexed implementation against a brute-force restatement in plain numpy.

The brute force shares no search code with the oracle: hits by full masks over y, exclusion by full
masks over the exclude set, counts by full masks over the query -- straight from the reference's
definitions (findOverlaps / shift / trim / width filter / overlapsAny / gaps + intersect /
countOverlaps with strand rule and maxgap)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H

ZERO_WIDTH_RULE = O.zero_width_rule()


def _compat(a, b):
    return (a == 0) | (b == 0) | (a == b)


def brute_iteration(case, kind, plan, bc, bs, dt):
    yc, ys, ye, ysd = case["y"]
    ysd = np.zeros(yc.size, np.int8) if ysd is None else ysd
    L = case["seqlen"]
    permute, drop_zero = kind >= 4, kind not in (1, 4)
    x = case["excl"]
    trim = bool(case.get("trim"))
    if x is not None:
        xc, xs, xe, xsd = x
        xsd = np.zeros(xc.size, np.int8) if xsd is None else xsd
    q = case["query"]
    qsd = np.zeros(q[0].size, np.int8) if q[3] is None else q[3]
    m1 = case.get("maxgap", -1) + 1
    mo = case.get("minoverlap", 0)
    counts = np.zeros(q[0].size, np.int64)
    rows = []
    for b in range(plan.B):
        c, s = bc[b], bs[b]
        e = s + plan.bait_width[b] - 1
        t = dt[b]
        tc, shift = plan.tile_chrom[t], plan.tile_start[t] - s
        hit = (yc == c) & ((ys >= s) & (ys <= e) if permute else (ys <= e) & (ye >= s))
        for i in np.flatnonzero(hit):                     # ascending y index
            s2, e2 = ys[i] + shift, ye[i] + shift
            if s2 > L[tc]:
                s2, e2 = L[tc] + 1, L[tc]
            else:
                s2, e2 = max(s2, 1), min(e2, L[tc])
            if e2 < s2 and drop_zero:
                continue
            pieces = [(s2, e2)]
            if x is not None and trim:
                pieces = []
                if e2 >= s2:
                    covered = np.zeros(L[tc] + 2, bool)   # gaps = positions of [1, L] no exclude range covers
                    for k in np.flatnonzero(xc == tc):
                        covered[max(xs[k], 1): min(xe[k], L[tc]) + 1] = True
                    pos = s2
                    while pos <= e2:                      # maximal uncovered runs inside [s2, e2], cut at gap borders
                        if covered[pos]:
                            pos += 1
                            continue
                        end = pos
                        while end + 1 <= e2 and not covered[end + 1]:
                            end += 1
                        pieces.append((pos, end))
                        pos = end + 1
            elif x is not None and (e2 >= s2 or ZERO_WIDTH_RULE):   # THE ZERO-WIDTH RULE (oracle/nr_oracle.c): 1 = plain test at width 0 too
                if np.any((xc == tc) & (xs <= e2) & (xe >= s2) & _compat(ysd[i], xsd)):
                    continue
            for ps, pe in pieces:
                rows.append((tc, ps, pe, i, b))
                if pe >= ps or (ZERO_WIDTH_RULE and mo == 0):
                    ov = np.minimum(pe, q[2]) - np.maximum(ps, q[1]) + 1
                    counts += (q[0] == tc) & (q[1] - m1 <= pe) & (q[2] + m1 >= ps) & _compat(ysd[i], qsd) & ((mo == 0) | (ov >= mo))
    return np.array(rows, np.int64).reshape(-1, 5), counts


@pytest.mark.parametrize("seed", range(150))
def test_oracle_equals_bruteforce(seed):
    rng = np.random.default_rng(50_000 + seed)
    kind = int(rng.integers(0, 6))
    n_seq = int(rng.integers(1, 4))
    case = H.random_case(seed=60_000 + seed, n=int(rng.integers(0, 160)), n_seq=n_seq, max_len=int(rng.choice([600, 1500])),
                         n_query=int(rng.integers(1, 30)), stranded=bool(rng.integers(0, 2)), wmax=int(rng.choice([1, 20, 150])),
                         L_b=int(rng.choice([40, 100, 300])), n_seg=2 * n_seq, n_states=int(rng.integers(1, 3)),
                         sort=bool(rng.integers(0, 2)) or kind in (0, 5))
    L = case["seqlen"]
    mode = rng.choice(["none", "drop", "trim"])
    if mode != "none":
        m = int(rng.integers(1, 12))
        sid = rng.integers(0, n_seq, m).astype(np.int32)
        st = (1 + rng.random(m) * (L[sid] - 1)).astype(np.int32)
        en = (st + rng.integers(0, 60, m)).astype(np.int32)
        sd = rng.integers(0, 3, m).astype(np.int8) if (mode == "drop" and rng.random() < 0.5) else None
        case["excl"], case["trim"] = (sid, st, en, sd), mode == "trim"
    case["maxgap"] = int(rng.choice([-1, 0, 15]))
    if case["maxgap"] == -1 and rng.random() < 0.5:
        case["minoverlap"] = int(rng.choice([1, 4, 30]))
    try:
        plan, y, q, x = H.oracle_objects(case, kind)
    except ValueError:
        return
    d = H.draw_R(plan, seed, 2)
    fused = O.boot_counts(plan, y, q, 2, exclude=x, exclude_mode=H.excl_mode(case), draws=d)
    for r in range(2):
        rows = O.iteration_rows(plan, y, d[0][r], d[1][r], d[2][r], exclude=x, exclude_mode=H.excl_mode(case))
        want_rows, want_counts = brute_iteration(case, kind, plan, d[0][r], d[1][r], d[2][r])
        got_rows = np.stack([rows["chrom"], rows["start"], rows["end"], rows["src"], rows["block"]], axis=1).astype(np.int64)
        # gaps split a range into ascending pieces; the brute force emits them in the same order
        assert np.array_equal(got_rows, want_rows), (kind, mode)
        assert np.array_equal(fused["feature_counts"][r], want_counts), (kind, mode)
        assert fused["iter_nranges"][r] == want_rows.shape[0] and fused["iter_totals"][r] == want_counts.sum()
