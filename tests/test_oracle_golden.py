"""CPU: the oracle against the golden vectors (tests/golden/, provenance in make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KNOWN = json.load(open(os.path.join(GOLD, "r_known_answers.json")))
WORKED = json.load(open(os.path.join(GOLD, "worked_examples.json")))


@pytest.mark.parametrize("case", KNOWN["runif"])
def test_r_runif_known_answers(case):
    assert np.allclose(O.RRng(case["seed"]).runif(case["n"]), case["value"], rtol=0, atol=5e-8)  # R prints 7 digits


@pytest.mark.parametrize("case", KNOWN["sample"])
def test_r_sample_known_answers(case):
    assert O.RRng(case["seed"]).sample(case["n"], case["size"], bool(case.get("replace", False))).tolist() == case["value"]


@pytest.mark.parametrize("case", KNOWN["philox4x32_10"])
def test_philox_known_answers(case):
    assert O.philox4x32_10(case["ctr"], case["key"]).tolist() == case["value"]


def test_runif_degenerate_interval_consumes_no_draw():
    g1, g2 = O.RRng(7), O.RRng(7)
    a = g1.runif(4, [1.0, 5.0, 2.0, 9.0], [1.0, 6.0, 2.0, 10.0])  # elements 0 and 2: min == max
    b = g2.runif(2, [5.0, 9.0], [6.0, 10.0])
    assert a[0] == 1.0 and a[2] == 2.0 and np.array_equal(a[[1, 3]], b)


def test_r_round_half_even():
    assert [O.r_round(v) for v in (0.5, 1.5, 2.5, -0.5, 2.4999, 3.5)] == [0.0, 2.0, 2.0, -0.0, 2.0, 4.0]


def test_sample_prob_walker_alias_switch():
    """sample(prob=) uses inversion on the sorted probabilities up to 200 'large' categories and
    Walker's alias above; both must return valid indices and respect zero weights."""
    for n in (5, 150, 260, 1000):
        p = np.ones(n)
        p[::7] = 0.0
        s = O.RRng(3).sample(n, 4000, replace=True, prob=p)
        assert s.min() >= 1 and s.max() <= n and not np.any(p[s - 1] == 0.0)
        freq = np.bincount(s, minlength=n + 1)[1:][p > 0]
        assert freq.std() / freq.mean() < 1.5  # roughly uniform over the eligible categories


def test_roxygen_example():
    w = WORKED["roxygen_example"]
    plan = O.Plan(O.UNSEG_ACROSS, w["seqlengths"], w["block_length"])
    y = O.Index(1, w["y"]["chrom"], w["y"]["start"], w["y"]["end"])
    d = plan.draw_R(O.RRng(w["seed"]), 1)
    assert d[1][0].tolist() == w["bait_start"]
    rows = O.iteration_rows(plan, y, d[0][0], d[1][0], d[2][0])
    assert rows["start"].tolist() == w["rows"]["start"] and rows["end"].tolist() == w["rows"]["end"]
    assert (rows["src"] + 1).tolist() == w["rows"]["src_1based"] and (rows["block"] + 1).tolist() == w["rows"]["block_1based"]


def test_test_boot_fixture_worked_example():
    w = WORKED["test_boot_fixture"]
    st = np.array(w["y"]["start"])
    plan = O.Plan(O.UNSEG_ACROSS, w["seqlengths"], w["block_length"])
    y = O.Index(3, w["y"]["chrom"], st, st + np.array(w["y"]["width"]) - 1)
    d = plan.draw_R(O.RRng(w["seed"]), 5)
    assert d[0][0].tolist() == w["iter1_bait_chrom"] and d[1][0].tolist() == w["iter1_bait_start"]
    rows = O.iteration_rows(plan, y, d[0][0], d[1][0], d[2][0])
    got = [list(t) for t in zip(rows["chrom"].tolist(), rows["start"].tolist(), rows["end"].tolist(),
                                (rows["src"] + 1).tolist(), (rows["block"] + 1).tolist())]
    assert got == w["iter1_rows"]
    per_iter = [O.iteration_rows(plan, y, d[0][r], d[1][r], d[2][r])["src"].size for r in range(5)]
    assert per_iter == w["rows_per_iteration"]


@pytest.mark.parametrize("kind", range(6))
def test_regression_vectors(kind):
    z = np.load(os.path.join(GOLD, f"regression_kind{kind}.npz"))
    case = H.random_case(seed=900 + kind, n=300, n_seq=3, n_query=40, n_excl=10 if kind % 2 else 0, stranded=kind in (1, 4),
                         n_seg=9, n_states=2)
    plan, y, q, x = H.oracle_objects(case, kind)
    d = H.draw_R(plan, 11, 4)
    assert np.array_equal(d[0], z["bait_chrom"]) and np.array_equal(d[1], z["bait_start"]) and np.array_equal(d[2], z["dst_tile"])
    ref = O.boot_counts(plan, y, q, 4, exclude=x, draws=d)
    for k in ("iter_nranges", "iter_totals", "feature_counts"):
        assert np.array_equal(ref[k], z[k])
    assert np.array_equal(H.oracle_rows(plan, y, x, d, 4), z["rows"])


def test_regression_philox():
    z = np.load(os.path.join(GOLD, "regression_philox.npz"))
    case = H.random_case(seed=950, n=300, n_seq=3, n_query=40)
    plan, y, q, x = H.oracle_objects(case, 0)
    d = plan.draw_philox(2024, 7, 5)
    assert np.array_equal(d[0], z["bait_chrom"]) and np.array_equal(d[1], z["bait_start"])
    ref = O.boot_counts(plan, y, q, 5, seed=2024, iter0=7, nthreads=2)
    for k in ("iter_nranges", "iter_totals", "feature_counts"):
        assert np.array_equal(ref[k], z[k])


def test_fused_counts_equal_two_step_workflow():
    """The fused path = bootRanges() then countOverlaps(x, boots[iter == r]) (vignette :497-503),
    and both equal an O(rows x G) brute force with the explicit strand rule."""
    for kind in (0, 1, 2, 4):
        case = H.random_case(seed=300 + kind, n=250, n_seq=3, n_query=30, n_excl=8, stranded=True, n_seg=9, n_states=2)
        plan, y, q, x = H.oracle_objects(case, kind)
        d = H.draw_R(plan, 5, 3)
        fused = O.boot_counts(plan, y, q, 3, exclude=x, draws=d)
        for r in range(3):
            rows = O.iteration_rows(plan, y, d[0][r], d[1][r], d[2][r], exclude=x)
            strand = case["y"][3][rows["src"]]
            cnt, tot = O.count_overlaps(q, rows["chrom"], rows["start"], rows["end"], strand)
            assert np.array_equal(cnt, fused["feature_counts"][r]) and tot == fused["iter_totals"][r]
            assert np.array_equal(H.brute_counts(case, (rows["chrom"], rows["start"], rows["end"], strand)), cnt)
            assert rows["src"].size == fused["iter_nranges"][r]


def test_maxgap_matches_gap_definition():
    """Oracle countOverlaps(maxgap = m) vs the definition: gap(a, b) = positions between them (adjacent: 0,
    overlapping: -1); overlap iff gap <= m."""
    rng = np.random.default_rng(3)
    qs = rng.integers(1, 900, 40); qe = qs + rng.integers(0, 30, 40)
    rs = rng.integers(1, 900, 200); re_ = rs + rng.integers(0, 30, 200)
    for m in (-1, 0, 1, 25):
        q = O.Index(1, np.zeros(40, np.int32), qs, qe, maxgap=m)
        cnt, tot = O.count_overlaps(q, np.zeros(200, np.int32), rs, re_)
        gap = np.maximum(qs[:, None] - re_[None, :], rs[None, :] - qe[:, None]) - 1
        want = (np.maximum(gap, -1) <= m).sum(axis=1)
        assert np.array_equal(cnt, want) and tot == want.sum()


def test_minoverlap_matches_definition():
    rng = np.random.default_rng(4)
    qs = rng.integers(1, 900, 40); qe = qs + rng.integers(0, 60, 40)
    rs = rng.integers(1, 900, 300); re_ = rs + rng.integers(0, 30, 300)
    for k in (1, 8, 20):
        q = O.Index(1, np.zeros(40, np.int32), qs, qe, minoverlap=k)
        cnt, _ = O.count_overlaps(q, np.zeros(300, np.int32), rs, re_)
        width = np.minimum(qe[:, None], re_[None, :]) - np.maximum(qs[:, None], rs[None, :]) + 1
        assert np.array_equal(cnt, (width >= k).sum(axis=1))
