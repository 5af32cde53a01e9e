"""The schema of tests/golden/r_reference_golden.json (written by tests/golden/make_golden.R where R exists) and the two
ways this repo reproduces a case of it: the CPU oracle and the product (GPU).  Also an emulator that writes the same
schema from the oracle, so that the consumer (tests/test_r_golden.py) is exercised even while no R-made file exists --
that run proves the plumbing, not parity."""
from __future__ import annotations

import json
import os

import numpy as np

from oracle import oracle as O

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "r_reference_golden.json")
KIND = {("bootstrap", False): 0, ("bootstrap", True): 1, ("permute", True): 4, ("permute", False): 5}


def case_kind(c) -> int:
    if c.get("seg") is not None:
        return 2 if c["proportion_length"] else 3
    return KIND[(c["type"], bool(c["within_chrom"]))]


def _arr(d, k, dtype=np.int32):
    return np.asarray(d[k], dtype=dtype).reshape(-1)


def _side(d):
    if d is None:
        return None
    st = _arr(d, "strand", np.int8)
    return _arr(d, "seqid0"), _arr(d, "start"), _arr(d, "end"), (st if st.any() else None)


def oracle_case(c):
    """Runs one golden case on the oracle under R's RNG -> dict(rows = [n, 6] (iter, seqid, start, end, strand, src) in
    the oracle's row order, rows_per_iteration, counts [R, G])."""
    kind, R, C = case_kind(c), int(c["R"]), len(c["seqlengths"])
    seg = None
    if c.get("seg") is not None:
        seg = {"chrom": _arr(c["seg"], "seqid0"), "start": _arr(c["seg"], "start"), "end": _arr(c["seg"], "end"), "state": _arr(c["seg"], "state")}
    plan = O.Plan(kind, np.asarray(c["seqlengths"], dtype=np.int64), int(c["block_length"]), seg)
    ys = _side(c["y"])
    y = O.Index(C, *ys)
    q = O.Index(C, *_side(c["query"]))
    x, mode = None, O.EXCLUDE_DROP
    if c.get("exclude") is not None:
        xs = _side(c["exclude"])
        if c["exclude_option"] == "trim":
            x, mode = O.gaps(C, np.asarray(c["seqlengths"], dtype=np.int64), *xs), O.EXCLUDE_TRIM
        else:
            x = O.Index(C, *xs)
    d = plan.draw_R(O.RRng(int(c["seed"])), R)
    rows, per = [], []
    ystrand = ys[3] if ys[3] is not None else np.zeros(ys[0].size, np.int8)
    for r in range(R):
        t = O.iteration_rows(plan, y, d[0][r], d[1][r], d[2][r], exclude=x, exclude_mode=mode)
        per.append(t["src"].size)
        rows.append(np.stack([np.full(t["src"].size, r + 1), t["chrom"], t["start"], t["end"], ystrand[t["src"]], t["src"] + 1], axis=1))
    counts = O.boot_counts(plan, y, q, R, exclude=x, exclude_mode=mode, draws=d)["feature_counts"]
    return {"rows": np.concatenate(rows) if rows else np.empty((0, 6), np.int64), "rows_per_iteration": per, "counts": counts}


def golden_rows(c) -> np.ndarray:
    r = c["rows"]
    return np.stack([_arr(r, k, np.int64) for k in ("iter", "seqid0", "start", "end", "strand", "src_1based")], axis=1)


def canonical(rows: np.ndarray) -> np.ndarray:
    """Order-free form: sorted on (iter, seqid, start, end, strand, src)."""
    return rows[np.lexsort(tuple(rows[:, k] for k in (5, 4, 3, 2, 1, 0)))] if rows.size else rows


def product_case(c, gpus=None):
    """The same case through the product's front end (GPU): bootRanges() rows and bootCounts() counts, rng = 'R'."""
    from nullranges_b200 import GRanges, bootCounts, bootRanges
    lv, sl = list(c["seqlevels"]), dict(zip(c["seqlevels"], (int(v) for v in c["seqlengths"])))

    def gr(d, **mcols):
        if d is None:
            return None
        return GRanges(_arr(d, "seqid0"), _arr(d, "start"), end=_arr(d, "end"), strand=_arr(d, "strand", np.int8), seqlengths=sl, seqlevels=lv, **mcols)

    y = gr(c["y"], src=np.arange(1, len(np.atleast_1d(c["y"]["start"])) + 1))
    seg = gr(c["seg"], state=_arr(c["seg"], "state")) if c.get("seg") is not None else None
    kw = dict(seg=seg, exclude=gr(c.get("exclude")), excludeOption=c["exclude_option"], proportionLength=bool(c["proportion_length"]),
              type=c["type"], withinChrom=bool(c["within_chrom"]), rng="R", seed=int(c["seed"]), gpus=gpus)
    br = bootRanges(y, int(c["block_length"]), R=int(c["R"]), **kw)
    rows = np.stack([br.mcols["iter"].astype(np.int64), br.seqid, br.start, br.end, br.strand, br.mcols["src"]], axis=1).astype(np.int64)
    bc = bootCounts(y, gr(c["query"]), int(c["block_length"]), R=int(c["R"]), **kw)
    return {"rows": rows, "rows_per_iteration": np.bincount(rows[:, 0], minlength=int(c["R"]) + 1)[1:].tolist(), "counts": bc.feature_counts}


# ---- emulator: the reference's fixtures (tests/golden/make_golden.R builds the same ones in R) ---------------------
def _fixtures():
    gr = {"seqid0": [0] * 4 + [1] * 5 + [2] * 2, "start": [1, 101, 121, 201, 101, 201, 216, 231, 401, 1, 101],
          "width": [20, 5, 5, 30, 20, 5, 5, 5, 30, 80, 40]}
    gr["end"] = [s + w - 1 for s, w in zip(gr["start"], gr["width"])]
    gr["strand"] = [0] * 11
    sl3, lv3 = [300, 450, 200], ["chr1", "chr2", "chr3"]
    seg = {"seqid0": [0, 0, 1, 1, 2], "start": [1, 151, 1, 226, 1], "end": [150, 300, 225, 450, 200], "strand": [0] * 5, "state": [1, 2, 1, 2, 1]}
    q3 = {"seqid0": [0, 0, 1, 1, 2, 2], "start": [1, 120, 90, 300, 1, 150], "end": [60, 260, 240, 450, 90, 200], "strand": [0] * 6}
    ex3 = {"seqid0": [0, 1], "start": [40, 200], "end": [60, 260], "strand": [0, 0]}
    # sort(y) of test_trim_exclude.R: strand + before -, then start
    st = [1 + 2 * k for k in range(46)]
    sd = [1 if k % 2 == 0 else 2 for k in range(46)]
    o = sorted(range(46), key=lambda k: (sd[k], st[k]))
    y_trim = {"seqid0": [0] * 46, "start": [st[k] for k in o], "end": [st[k] + 9 for k in o], "strand": [sd[k] for k in o]}
    ex_trim = {"seqid0": [0], "start": [36], "end": [45], "strand": [0]}
    q_trim = {"seqid0": [0] * 4, "start": [1, 30, 50, 80], "end": [25, 60, 70, 100], "strand": [0, 1, 2, 0]}
    gr1 = {"seqid0": [0] * 5, "start": [1, 11, 21, 31, 41], "end": [5, 15, 25, 35, 45], "strand": [0] * 5}
    q1 = {"seqid0": [0] * 3, "start": [1, 20, 44], "end": [12, 30, 50], "strand": [0] * 3}

    def case(name, source, y, seed, L, R, query, lv, sl, type="bootstrap", within=False, prop=True, seg=None, exclude=None, opt="drop"):
        return {"name": name, "source": source, "seed": seed, "block_length": L, "R": R, "type": type, "within_chrom": within,
                "proportion_length": prop, "exclude_option": opt, "seqlevels": lv, "seqlengths": sl,
                "y": {k: v for k, v in y.items() if k != "width"}, "seg": seg, "exclude": exclude, "query": query}

    return [
        case("roxygen_example", "R/bootRanges.R:65-69", gr1, 1, 10, 1, q1, ["chr1"], [50]),
        case("boot_across", "tests/testthat/test_boot.R:67", gr, 1, 100, 5, q3, lv3, sl3),
        case("boot_within", "tests/testthat/test_boot.R:41", gr, 1, 100, 5, q3, lv3, sl3, within=True),
        case("permute_within", "tests/testthat/test_boot.R:27", gr, 1, 100, 5, q3, lv3, sl3, type="permute", within=True),
        case("permute_across", "tests/testthat/test_boot.R:54", gr, 1, 100, 5, q3, lv3, sl3, type="permute"),
        case("segmented_prop", "tests/testthat/test_boot.R:80-86", gr, 1, 100, 5, q3, lv3, sl3, seg=seg),
        case("segmented_noprop", "tests/testthat/test_boot.R:80-82 with proportionLength = FALSE", gr, 3, 100, 5, q3, lv3, sl3, seg=seg, prop=False),
        case("exclude_drop", "tests/testthat/test_boot.R:7-16 + exclude", gr, 2, 100, 5, q3, lv3, sl3, exclude=ex3),
        case("exclude_trim", "tests/testthat/test_trim_exclude.R:4-11", y_trim, 1, 20, 3, q_trim, ["chr1"], [100], exclude=ex_trim, opt="trim"),
    ]


def emulate(path: str) -> str:
    """Writes `path` in make_golden.R's schema with the ORACLE's outputs (provenance says so)."""
    cases = []
    for c in _fixtures():
        got = oracle_case(c)
        r = got["rows"]
        c["rows"] = {k: r[:, i].tolist() for i, k in enumerate(("iter", "seqid0", "start", "end", "strand", "src_1based"))}
        c["rows_per_iteration"] = [int(v) for v in got["rows_per_iteration"]]
        c["counts"] = got["counts"].tolist()
        cases.append(c)
    facts = {"zero_width_hits_range_past_end": O.zero_width_rule(), "zero_width_hits_range_inside": 0,
             "zero_width_overlapsAny_past_end": O.zero_width_rule(), "within_block_hit_order": [[1, 2, 4], [1, 3, 4, 5]]}
    json.dump({"provenance": {"R": None, "emulated_by": "tests/r_golden.py from oracle/nr_oracle.c -- NOT an R run"}, "facts": facts, "cases": cases},
              open(path, "w"))
    return path
