"""GPU parity: the CUDA engine through the C ABI vs the CPU oracle, bit-exact (integer work)."""
import numpy as np
import pytest

from nullranges_b200 import Draws
from nullranges_b200 import _lib
from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

BOOT_KINDS = [0, 1, 2, 3]
ALL_KINDS = [0, 1, 2, 3, 4, 5]


def _check_counts(eng, case, kind, R, seed, philox):
    plan, y, q, x = H.oracle_objects(case, kind)
    B = H.load_engine(eng, case, kind)
    assert B == plan.B
    ts, tt, bw = eng.tiles()
    assert np.array_equal(ts, plan.tile_chrom) and np.array_equal(tt, plan.tile_start) and np.array_equal(bw, plan.bait_width)
    if philox:
        draws = Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=seed, iter0=5)
        try:
            ref = O.boot_counts(plan, y, q, R, exclude=x, exclude_mode=H.excl_mode(case), seed=seed, iter0=5)
        except ValueError as e:   # proportionLength = FALSE: a state without draws fails in the reference too
            assert "drew no segments" in str(e)
            from nullranges_b200 import EngineError
            with pytest.raises(EngineError, match="drew no segments"):
                eng.run_counts(draws)
            return None, None
    else:
        d = H.draw_R(plan, seed, R)
        draws = Draws(n_iter=R, rng_mode=_lib.RNG_HOST, bait_seqid=d[0], bait_start=d[1], dst_tile=d[2] if kind >= 4 else None)
        ref = O.boot_counts(plan, y, q, R, exclude=x, exclude_mode=H.excl_mode(case), draws=d)
    got = eng.run_counts(draws)
    assert np.array_equal(got["iter_nranges"], ref["iter_nranges"])
    assert np.array_equal(got["iter_totals"], ref["iter_totals"])
    if q is not None:
        assert np.array_equal(got["feature_counts"], ref["feature_counts"])
        assert np.array_equal(got["feature_counts"].sum(axis=1), got["iter_totals"])
    return got, ref


@pytest.mark.parametrize("kind", ALL_KINDS)
@pytest.mark.parametrize("stranded", [False, True])
@pytest.mark.parametrize("n_excl", [0, 25])
def test_counts_validation_mode(engine, kind, stranded, n_excl):
    case = H.random_case(seed=10 + kind, n=600, n_seq=4, n_query=80, n_excl=n_excl, stranded=stranded, n_seg=12,
                         sort=True)
    _check_counts(engine, case, kind, R=6, seed=3, philox=False)


@pytest.mark.parametrize("kind", ALL_KINDS)
@pytest.mark.parametrize("n_excl", [0, 25])
def test_counts_philox_mode(engine, kind, n_excl):
    case = H.random_case(seed=20 + kind, n=700, n_seq=5, n_query=90, n_excl=n_excl, n_seg=15)
    _check_counts(engine, case, kind, R=7, seed=2024, philox=True)


@pytest.mark.parametrize("kind", [1, 2, 3, 4])
def test_unsorted_y(engine, kind):
    case = H.random_case(seed=130 + kind, n=500, n_seq=3, n_query=50, stranded=True, n_seg=6, n_states=2, sort=False)
    _check_counts(engine, case, kind, R=4, seed=9, philox=False)


@pytest.mark.parametrize("kind", ALL_KINDS)
@pytest.mark.parametrize("n_excl", [0, 20])
def test_rows_match_oracle(engine, kind, n_excl):
    case = H.random_case(seed=40 + kind, n=500, n_seq=3, n_query=0, n_excl=n_excl, stranded=bool(n_excl), n_seg=9)
    plan, y, q, x = H.oracle_objects(case, kind)
    H.load_engine(engine, case, kind)
    R = 5
    d = H.draw_R(plan, 77, R)
    draws = Draws(n_iter=R, rng_mode=_lib.RNG_HOST, bait_seqid=d[0], bait_start=d[1], dst_tile=d[2] if kind >= 4 else None)
    rows = engine.run_materialize(draws)
    it = np.repeat(np.arange(R), np.diff(rows["iter_offsets"]))
    got = H.canon_rows(rows["seqid"], rows["start"], rows["end"], rows["src"], rows["block"], it)
    want = H.oracle_rows(plan, y, x, d, R)
    assert np.array_equal(got, want)
    # emitted order is already (iter, block, canonical y order) = (iter, block, src) for sorted y
    assert np.array_equal(rows["block"][np.argsort(it, kind="stable")], rows["block"])


def test_points_and_long_ranges(engine):
    for pts, wmax in ((True, 1), (False, 2500)):
        case = H.random_case(seed=51, n=800, n_seq=3, n_query=70, points=pts, wmax=wmax)
        for kind in (0, 1):
            _check_counts(engine, case, kind, R=5, seed=1, philox=False)


def test_edge_cases(engine):
    # empty y
    case = H.random_case(seed=60, n=0, n_query=10)
    got, _ = _check_counts(engine, case, 0, R=3, seed=1, philox=False)
    assert got["iter_nranges"].sum() == 0
    # no query: only iter_nranges
    case = H.random_case(seed=61, n=300, n_query=0)
    _check_counts(engine, case, 0, R=3, seed=1, philox=False)
    # a chromosome without ranges, single range, R = 0
    case = H.random_case(seed=62, n=1, n_seq=4, n_query=5)
    _check_counts(engine, case, 1, R=4, seed=1, philox=False)
    got, _ = _check_counts(engine, case, 0, R=0, seed=1, philox=True)
    assert got["iter_totals"].size == 0


def test_moments_and_observed(engine):
    case = H.random_case(seed=70, n=900, n_seq=4, n_query=120, stranded=True)
    plan, y, q, x = H.oracle_objects(case, 0)
    H.load_engine(engine, case, 0)
    R = 23
    ref = O.boot_counts(plan, y, q, R, seed=5, iter0=0)
    obs, tot = engine.count_observed()
    cy = case["y"]
    want_obs = H.brute_counts(case, (cy[0], cy[1], cy[2], cy[3]))
    assert np.array_equal(obs, want_obs) and tot == want_obs.sum()
    m = engine.run_moments(Draws(n_iter=R, seed=5), observed=obs, batch_iter=7)
    fc = ref["feature_counts"].astype(np.int64)
    assert np.array_equal(m["sum"], fc.sum(0))
    assert np.array_equal(m["sumsq"], (fc * fc).sum(0))
    assert np.array_equal(m["n_ge_observed"], (fc >= obs[None, :]).sum(0))
    assert np.array_equal(m["iter_totals"], ref["iter_totals"])


def test_features_beyond_sequence_end(engine):
    """Query features that start or end past seqlengths: trim() keeps ranges inside [1, L]."""
    case = H.random_case(seed=81, n=900, n_seq=3, n_query=40, wmax=400)
    sid, st, en, sd = case["query"]
    L = case["seqlen"][sid]
    st = st.copy(); en = en.copy()
    st[::4] = (L[::4] - 5).astype(np.int32); en[::4] = (L[::4] + 300).astype(np.int32)     # straddles the end
    st[1::4] = (L[1::4] + 1).astype(np.int32); en[1::4] = (L[1::4] + 50).astype(np.int32)  # wholly beyond
    case["query"] = (sid, st, en, sd)
    for kind in (0, 1):
        _check_counts(engine, case, kind, R=6, seed=4, philox=False)


def _synth_case(y, x):
    return {"seqlen": y.seqlengths, "y": (y.seqid, y.start, y.end, None), "query": (x.seqid, x.start, x.end, None),
            "excl": None, "seg": None, "L_b": 500_000}


def test_cfg1_validation_mode_full(engine):
    """BASELINE configs[0]: 10k ranges, hg38 seqlengths, blockLength 5e5, R = 50, 20k genes; block draws
    from R's RNG under set.seed(1): totals, [50 x 20k] counts and the whole table, bit-exact."""
    from nullranges_b200 import synth
    y, x = synth.uniform_ranges(10_000, seed=1), synth.genes(20_000, seed=2)
    case = _synth_case(y, x)
    got, ref = _check_counts(engine, case, 0, R=50, seed=1, philox=False)
    assert got["iter_nranges"].sum() > 400_000
    plan, yi, _, _ = H.oracle_objects(case, 0)
    d = H.draw_R(plan, 1, 50)
    rows = engine.run_materialize(Draws(n_iter=50, rng_mode=_lib.RNG_HOST, bait_seqid=d[0], bait_start=d[1]))
    it = np.repeat(np.arange(50), np.diff(rows["iter_offsets"]))
    assert np.array_equal(np.diff(rows["iter_offsets"]), got["iter_nranges"])
    assert np.array_equal(H.canon_rows(rows["seqid"], rows["start"], rows["end"], rows["src"], rows["block"], it),
                          H.oracle_rows(plan, yi, None, d, 50))


def test_cfg2_full_size(engine):
    """BASELINE configs[1] at full size: 1M ATAC-like peaks vs 20k genes, blockLength 5e5; 64 Philox
    iterations against the oracle, then R = 1000 through size-independent properties."""
    from nullranges_b200 import synth
    y, x = synth.atac_peaks(1_000_000, seed=3), synth.genes(20_000, seed=2)
    case = _synth_case(y, x)
    plan, yi, qi, _ = H.oracle_objects(case, 0)
    H.load_engine(engine, case, 0)
    R = 64
    got = engine.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=2024, iter0=0))
    ref = O.boot_counts(plan, yi, qi, R, seed=2024, iter0=0, nthreads=O.max_threads())
    for k in ("iter_nranges", "iter_totals", "feature_counts"):
        assert np.array_equal(got[k], ref[k]), k
    # R = 1000: the sharded result equals the one-shot result (counter-based draws), row sums = totals
    big = engine.run_counts(Draws(n_iter=1000, rng_mode=_lib.RNG_PHILOX, seed=2024, iter0=0))
    assert np.array_equal(big["feature_counts"][:R], got["feature_counts"])
    assert np.array_equal(big["feature_counts"].sum(axis=1), big["iter_totals"])
    part = engine.run_counts(Draws(n_iter=300, rng_mode=_lib.RNG_PHILOX, seed=2024, iter0=700))
    assert np.array_equal(part["feature_counts"], big["feature_counts"][700:])
    assert np.array_equal(part["iter_nranges"], big["iter_nranges"][700:])
    # bootstrap mean of the total tracks the observed total (block bootstrap preserves density)
    obs, tot = engine.count_observed()
    assert abs(big["iter_totals"].mean() / tot - 1.0) < 0.05


@pytest.mark.parametrize("L_b", [100_000, 200_000, 1_000_000])
def test_cfg5_block_length_sweep_full_size(engine, L_b):
    """BASELINE configs[4] (the vignette's variance-vs-L_b diagnostic, vignettes/bootRanges.Rmd:195-199) at full size:
    the 1M-range workload at the other three block lengths (5e5 is test_cfg2_full_size), 8 Philox iterations each
    against the oracle -- whole count matrix, per-iteration totals and rows; B = 28,760 / 14,385 / 2,887 blocks."""
    from nullranges_b200 import synth
    y, x = synth.atac_peaks(1_000_000, seed=3), synth.genes(20_000, seed=2)
    case = _synth_case(y, x)
    case["L_b"] = L_b
    plan, yi, qi, _ = H.oracle_objects(case, 0)
    assert H.load_engine(engine, case, 0) == plan.B == {100_000: 28_760, 200_000: 14_385, 1_000_000: 2_887}[L_b]
    R = 8
    got = engine.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=77, iter0=123))
    ref = O.boot_counts(plan, yi, qi, R, seed=77, iter0=123, nthreads=O.max_threads())
    for k in ("iter_nranges", "iter_totals", "feature_counts"):
        assert np.array_equal(got[k], ref[k]), k
    # R = 400 through size-independent properties: a later shard reproduces its rows; row sums are the totals
    big = engine.run_counts(Draws(n_iter=400, rng_mode=_lib.RNG_PHILOX, seed=77, iter0=123))
    assert np.array_equal(big["feature_counts"][:R], got["feature_counts"])
    assert np.array_equal(big["feature_counts"].sum(axis=1), big["iter_totals"])
    part = engine.run_counts(Draws(n_iter=50, rng_mode=_lib.RNG_PHILOX, seed=77, iter0=123 + 350))
    assert np.array_equal(part["feature_counts"], big["feature_counts"][350:])


@pytest.mark.parametrize("kind", BOOT_KINDS)
def test_dense_exclusion(engine, kind):
    """Many small, close exclude regions and ranges long enough to span several of them: the rank
    path's union-by-consecutive-overlaps must agree with the per-range test of the reference."""
    case = H.random_case(seed=90 + kind, n=900, n_seq=3, n_query=80, n_excl=0, wmax=400, n_seg=9, n_states=2)
    rng = np.random.default_rng(5)
    L = case["seqlen"]
    sid = rng.integers(0, 3, 150).astype(np.int32)
    st = (1 + rng.random(150) * (L[sid] - 40)).astype(np.int32)
    en = (st + rng.integers(0, 30, 150)).astype(np.int32)
    en[::10] = (L[sid[::10]] + 50).astype(np.int32)      # some reach past the sequence end
    st[5::25] = (L[sid[5::25]] + 1).astype(np.int32)     # some lie wholly beyond it
    en[5::25] = st[5::25] + 10
    case["excl"] = (sid, st, en, None)
    _check_counts(engine, case, kind, R=6, seed=8, philox=False)
    if kind < 3:
        _check_counts(engine, case, kind, R=5, seed=77, philox=True)
    # materialised rows too (the count pass of the table uses the same row counts)
    plan, y, q, x = H.oracle_objects(case, kind)
    H.load_engine(engine, case, kind)
    d = H.draw_R(plan, 3, 3)
    rows = engine.run_materialize(Draws(n_iter=3, rng_mode=_lib.RNG_HOST, bait_seqid=d[0], bait_start=d[1]))
    it = np.repeat(np.arange(3), np.diff(rows["iter_offsets"]))
    assert np.array_equal(H.canon_rows(rows["seqid"], rows["start"], rows["end"], rows["src"], rows["block"], it),
                          H.oracle_rows(plan, y, x, d, 3))


def test_cfg3_full_size(engine):
    """BASELINE configs[2]: segmented (3 states, L_s = 2e6), proportionLength = TRUE, exclude 'drop',
    1M ranges vs 20k genes; 24 Philox iterations against the oracle."""
    from nullranges_b200 import synth
    y, x = synth.atac_peaks(1_000_000, seed=3), synth.genes(20_000, seed=2)
    seg, ex = synth.segmentation(seed=4), synth.exclude_regions(seed=5)
    case = _synth_case(y, x)
    case["seg"] = {"chrom": seg.seqid, "start": seg.start, "end": seg.end, "state": seg.mcols["state"].astype(np.int32)}
    case["excl"] = (ex.seqid, ex.start, ex.end, None)
    plan, yi, qi, xi = H.oracle_objects(case, 2)
    assert H.load_engine(engine, case, 2) == plan.B
    R = 24
    got = engine.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=11, iter0=100))
    ref = O.boot_counts(plan, yi, qi, R, exclude=xi, seed=11, iter0=100, nthreads=O.max_threads())
    for k in ("iter_nranges", "iter_totals", "feature_counts"):
        assert np.array_equal(got[k], ref[k]), k
    assert got["iter_nranges"].mean() < 1_000_000  # exclusion removes ranges


def test_philox_moments_match_r_rng(engine):
    """Production-mode (Philox) draws vs validation-mode (R RNG) draws: same bootstrap distribution.
    Tolerances: mean of the per-iteration total within 4 standard errors of the difference; variance
    ratio within [0.6, 1.6]; per-feature bootstrap means correlated > 0.98; block-start histogram of the
    Philox draws uniform over the admissible starts (chi-square / dof < 1.5)."""
    case = H.random_case(seed=123, n=20_000, n_seq=4, max_len=400_000, n_query=300, wmax=300, L_b=20_000)
    plan, y, q, _ = H.oracle_objects(case, 0)
    H.load_engine(engine, case, 0)
    R = 600
    d = H.draw_R(plan, 42, R)
    a = engine.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_HOST, bait_seqid=d[0], bait_start=d[1]))
    b = engine.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=42))
    ta, tb = a["iter_totals"].astype(float), b["iter_totals"].astype(float)
    se = np.sqrt(ta.var(ddof=1) / R + tb.var(ddof=1) / R)
    assert abs(ta.mean() - tb.mean()) < 4 * se
    assert 0.6 < ta.var(ddof=1) / tb.var(ddof=1) < 1.6
    assert abs(a["iter_nranges"].mean() - b["iter_nranges"].mean()) < 4 * np.sqrt(
        a["iter_nranges"].var(ddof=1) / R + b["iter_nranges"].var(ddof=1) / R)
    fa, fb = a["feature_counts"].mean(0), b["feature_counts"].mean(0)
    assert np.corrcoef(fa, fb)[0, 1] > 0.98
    # distribution of the Philox block draws themselves
    pc, ps, _ = plan.draw_philox(42, 0, 2000)
    L = case["seqlen"].astype(float)
    share = np.bincount(pc.ravel(), minlength=4) / pc.size
    assert np.all(np.abs(share - L / L.sum()) < 0.008)         # sequences drawn in proportion to length (about 5 s.e.)
    c0 = ps[pc == 0]
    span = (np.ceil(L[0] / 20_000) - 1) * 20_000 + 1
    hist = np.histogram(c0, bins=20, range=(1, span + 1))[0]
    assert ((hist - hist.mean()) ** 2 / hist.mean()).sum() / 19 < 1.5


@pytest.mark.parametrize("kind", ALL_KINDS)
@pytest.mark.parametrize("stranded", [False, True])
def test_exclude_trim(engine, kind, stranded):
    """excludeOption = "trim" (R/bootRanges.R:110-115): ranges are cut to gaps(exclude); one row per
    (range, gap) overlap.  Counts, row counts and the table itself against the oracle."""
    case = H.random_case(seed=210 + kind, n=700, n_seq=3, n_query=70, n_excl=0, stranded=stranded, wmax=300, n_seg=9, n_states=2)
    rng = np.random.default_rng(6)
    L = case["seqlen"]
    sid = rng.integers(0, 3, 60).astype(np.int32)
    st = (1 + rng.random(60) * (L[sid] - 60)).astype(np.int32)
    en = (st + rng.integers(0, 50, 60)).astype(np.int32)
    en[::12] = (L[sid[::12]] + 20).astype(np.int32)
    case["excl"], case["trim"] = (sid, st, en, None), True
    _check_counts(engine, case, kind, R=5, seed=4, philox=False)
    if kind < 3:
        _check_counts(engine, case, kind, R=4, seed=99, philox=True)
    plan, y, q, x = H.oracle_objects(case, kind)
    H.load_engine(engine, case, kind)
    d = H.draw_R(plan, 8, 3)
    rows = engine.run_materialize(Draws(n_iter=3, rng_mode=_lib.RNG_HOST, bait_seqid=d[0], bait_start=d[1], dst_tile=d[2] if kind >= 4 else None))
    it = np.repeat(np.arange(3), np.diff(rows["iter_offsets"]))
    want = H.oracle_rows(plan, y, x, d, 3, mode=O.EXCLUDE_TRIM)
    got = np.stack([it, rows["seqid"], rows["start"], rows["end"], rows["src"], rows["block"]], axis=1)
    # (iter, block, src, start): the pieces of one range come out in ascending gap order
    key = lambda a: a[np.lexsort((a[:, 2], a[:, 4], a[:, 5], a[:, 0]))]
    assert np.array_equal(key(got), key(want))
    # no piece overlaps an exclude range (tests/testthat/test_trim_exclude.R:13)
    for c, s, e in zip(rows["seqid"], rows["start"], rows["end"]):
        assert not np.any((sid == c) & (st <= e) & (en >= s))


def test_trim_needs_unstranded_exclude(engine):
    from nullranges_b200 import EngineError
    case = H.random_case(seed=5, n=50, n_excl=5, stranded=True)
    engine.set_ranges(case["seqlen"], *case["y"])
    with pytest.raises(EngineError, match="strand"):
        engine.set_exclude(*case["excl"], option="trim")
    engine.clear_exclude()


def test_cfg2_validation_mode_R1000(engine):
    """BASELINE target: bit-exact overlap counts in validation mode for 1M ranges x R = 1000 -- block draws
    from R's RNG under set.seed(2024), whole [1000 x 20k] matrix, totals and row counts vs the oracle."""
    from nullranges_b200 import synth
    y, x = synth.atac_peaks(1_000_000, seed=3), synth.genes(20_000, seed=2)
    case = _synth_case(y, x)
    plan, yi, qi, _ = H.oracle_objects(case, 0)
    H.load_engine(engine, case, 0)
    R = 1000
    d = plan.draw_R(O.RRng(2024), R)
    got = engine.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_HOST, bait_seqid=d[0], bait_start=d[1]))
    ref = O.boot_counts(plan, yi, qi, R, draws=d, nthreads=O.max_threads())
    for k in ("iter_nranges", "iter_totals", "feature_counts"):
        assert np.array_equal(got[k], ref[k]), k


@pytest.mark.parametrize("kind", [0, 1, 2, 4, 5])
@pytest.mark.parametrize("maxgap", [0, 75])
def test_maxgap(engine, kind, maxgap):
    """countOverlaps(x, boots, maxgap = m): features within m positions of a bootstrapped range count."""
    case = H.random_case(seed=400 + kind, n=700, n_seq=3, n_query=90, n_excl=12 if kind == 2 else 0, stranded=kind == 1, n_seg=9,
                         n_states=2, wmax=60)
    case["maxgap"] = maxgap
    got, ref = _check_counts(engine, case, kind, R=5, seed=6, philox=False)
    case["maxgap"] = -1
    _, plain = _check_counts(engine, case, kind, R=5, seed=6, philox=False)
    assert ref["iter_totals"].sum() > plain["iter_totals"].sum()   # the gap really adds overlaps


def test_error_behaviour(engine):
    """The reference's stopifnot()s and the ABI's state checks come back as error codes with the reference's wording."""
    from nullranges_b200 import EngineError, GRanges, bootRanges, bootCounts
    L = np.array([1000, 800], np.int64)
    sid, st, en = np.array([0, 0, 1], np.int32), np.array([5, 50, 7], np.int32), np.array([9, 60, 20], np.int32)
    with pytest.raises(EngineError, match="is.na"):          # stopifnot(all(!is.na(chr_lens)))  R/bootRanges.R:81
        engine.set_ranges(np.array([1000, -1]), sid, st, en)
    engine.set_ranges(L, sid, st, en)
    with pytest.raises(EngineError, match="chrom_lens >= L_b"):   # R/bootUnsegmented.R:8
        engine.set_plan(_lib.UNSEG_ACROSS, 900)
    with pytest.raises(EngineError, match="blockLength"):
        engine.set_plan(_lib.UNSEG_ACROSS, 0)
    with pytest.raises(EngineError, match="state"):
        engine.set_plan(_lib.SEG_PROP, 100, {"seqid": [0], "start": [1], "end": [1000], "state": [0]})
    with pytest.raises(EngineError, match="seqlevels"):
        engine.set_query(np.array([2], np.int32), np.array([1], np.int32), np.array([5], np.int32))
    with pytest.raises(EngineError, match="maxgap"):
        engine.set_query(np.array([0], np.int32), np.array([1], np.int32), np.array([5], np.int32), maxgap=-2)
    engine.set_plan(_lib.UNSEG_ACROSS, 100)
    with pytest.raises(EngineError, match="out of range"):     # host draws are validated before they reach the device
        engine.run_counts(Draws(n_iter=1, rng_mode=_lib.RNG_HOST, bait_seqid=np.full((1, 18), 5, np.int32), bait_start=np.ones((1, 18), np.int32)))
    engine.set_plan(_lib.PERMUTE_ACROSS, 100)
    with pytest.raises(EngineError, match="dst_tile"):
        engine.run_counts(Draws(n_iter=1, rng_mode=_lib.RNG_HOST, bait_seqid=np.zeros((1, 18), np.int32), bait_start=np.ones((1, 18), np.int32)))
    # front end: the reference's own checks
    y = GRanges(["b", "a"], [5, 3], width=[2, 2], seqlengths={"a": 50, "b": 50}, seqlevels=["a", "b"])
    with pytest.raises(ValueError, match="sort"):              # all(y == GenomicRanges::sort(y))  R/bootUnsegmented.R:31
        bootRanges(y, 10)
    assert len(bootRanges(y, 10, withinChrom=True, seed=1)) >= 0   # within-chromosome accepts any order (:15-28)
    with pytest.raises(ValueError, match="should be one of"):
        bootRanges(y.sort(), 10, type="jackknife")
    with pytest.raises(ValueError, match="integer"):
        bootRanges(y.sort(), 10.5)
    empty = bootCounts(y.sort(), GRanges(["zz"], [1], width=[5], seqlevels=["zz"], seqlengths=[9]), 10, R=3, seed=1)
    assert empty.feature_counts.shape == (3, 1) and empty.feature_counts.sum() == 0   # features on unknown sequences count 0


@pytest.mark.parametrize("kind", ALL_KINDS)
def test_dispatch_rank_vs_stream(engine, kind):
    """Every sampler -- stranded, excluded, permuted -- runs on the rank path by default (stats.n_windows > 0);
    an engine created with force_stream reports none."""
    case = H.random_case(seed=600 + kind, n=300, n_seq=3, n_query=40, n_excl=10, stranded=True, n_seg=9, n_states=2)
    plan, y, q, x = H.oracle_objects(case, kind)
    H.load_engine(engine, case, kind)
    d = H.draw_R(plan, 2, 2)
    st = engine.run_counts_resident(Draws(n_iter=2, rng_mode=_lib.RNG_HOST, bait_seqid=d[0], bait_start=d[1],
                                          dst_tile=d[2] if kind >= 4 else None), True, True)
    assert (st["n_windows"] == 0) == engine.force_stream
    assert st["n_launches"] == (1 if engine.force_stream else 2)


@pytest.mark.parametrize("seed", range(240))
def test_fuzz_parity(engine, seed):
    """Random configurations across every option at once (sampler, strands of y / query / exclude set
    independently, exclusion mode, maxgap, unsorted input, points, ranges wider than a block, features past
    the sequence end): fused counts, row counts and the materialised table vs the oracle."""
    rng = np.random.default_rng(10_000 + seed)
    kind = int(rng.integers(0, 6))
    n_seq = int(rng.integers(1, 6))
    L_b = int(rng.choice([50, 200, 500, 1500]))
    case = H.random_case(seed=20_000 + seed, n=int(rng.integers(0, 900)), n_seq=n_seq, max_len=int(rng.choice([3000, 6000, 20000])),
                         n_query=int(rng.integers(1, 120)), n_excl=0, stranded=False, wmax=int(rng.choice([1, 30, 300, 2500])),
                         L_b=L_b, n_seg=3 * n_seq, n_states=int(rng.integers(1, 4)), sort=bool(rng.integers(0, 2)) or kind in (0, 5),
                         points=bool(rng.integers(0, 5) == 0))
    L = case["seqlen"]

    def restrand(t, p):
        return t if rng.random() > p else (t[0], t[1], t[2], rng.integers(0, 3, t[0].size).astype(np.int8))

    case["y"] = restrand(case["y"], 0.4)
    q = restrand(case["query"], 0.4)
    qs, qe = q[1].copy(), q[2].copy()
    far = rng.random(qs.size) < 0.1                                 # some features straddle / lie past the sequence end
    qe[far] = (L[q[0][far]] + rng.integers(0, 200, int(far.sum()))).astype(np.int32)
    qs[far] = np.minimum(qs[far], qe[far])
    case["query"] = (q[0], qs, qe, q[3])
    mode = rng.choice(["none", "drop", "trim"], p=[0.3, 0.4, 0.3])
    if mode != "none":
        m = int(rng.integers(1, 80))
        sid = rng.integers(0, n_seq, m).astype(np.int32)
        st = (1 + rng.random(m) * (L[sid] - 1)).astype(np.int32)
        en = (st + rng.integers(0, int(rng.choice([5, 60, 900])), m)).astype(np.int32)
        x = (sid, st, en, None)
        case["excl"] = x if mode == "trim" else restrand(x, 0.4)
        case["trim"] = mode == "trim"
    case["maxgap"] = int(rng.choice([-1, -1, 0, 40]))
    if case["maxgap"] == -1 and rng.random() < 0.4:
        case["minoverlap"] = int(rng.choice([1, 5, 25, 200]))
    try:
        plan, y, qi, xi = H.oracle_objects(case, kind)
    except ValueError as e:                                      # e.g. a state's scaled block length rounds to 0
        with pytest.raises(Exception):
            H.load_engine(engine, case, kind)
        return
    B = H.load_engine(engine, case, kind)
    assert B == plan.B
    R = 3
    d = H.draw_R(plan, 1 + seed, R)
    draws = Draws(n_iter=R, rng_mode=_lib.RNG_HOST, bait_seqid=d[0], bait_start=d[1], dst_tile=d[2] if kind >= 4 else None)
    ref = O.boot_counts(plan, y, qi, R, exclude=xi, exclude_mode=H.excl_mode(case), draws=d)
    got = engine.run_counts(draws)
    for k in ("iter_nranges", "iter_totals", "feature_counts"):
        assert np.array_equal(got[k], ref[k]), (k, kind, mode)
    try:
        ref = O.boot_counts(plan, y, qi, R, exclude=xi, exclude_mode=H.excl_mode(case), seed=seed, iter0=11)
    except ValueError as e:       # proportionLength = FALSE: a state without draws (the reference fails as well)
        assert "drew no segments" in str(e)
        from nullranges_b200 import EngineError
        with pytest.raises(EngineError, match="drew no segments"):
            engine.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=seed, iter0=11))
        ref = None
    if ref is not None:
        got = engine.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=seed, iter0=11))
        for k in ("iter_nranges", "iter_totals", "feature_counts"):
            assert np.array_equal(got[k], ref[k]), (k, kind, mode, "philox")
    rows = engine.run_materialize(draws)
    it = np.repeat(np.arange(R), np.diff(rows["iter_offsets"]))
    got_rows = np.stack([it, rows["seqid"], rows["start"], rows["end"], rows["src"], rows["block"]], axis=1)
    want = H.oracle_rows(plan, y, xi, d, R, mode=H.excl_mode(case))
    key = lambda a: a[np.lexsort((a[:, 2], a[:, 4], a[:, 5], a[:, 0]))] if a.size else a
    assert np.array_equal(key(got_rows), key(want)), (kind, mode)


def test_count_overlaps_and_segment_tile_counts(engine):
    """countOverlaps(query, subject, maxgap / minoverlap) on plain ranges and the tile counts of
    segmentDensity() (R/segmentDensity.R:53-78, roxygen example data) against the oracle."""
    from nullranges_b200 import GRanges, countOverlaps, segmentTileCounts
    rng = np.random.default_rng(1)
    n = 10_000                                                     # R/segmentDensity.R:40-47
    st = np.rint(np.concatenate([rng.uniform(1, 991, n // 4), rng.uniform(1001, 3991, n // 4),
                                 rng.uniform(4001, 4991, n // 4), rng.uniform(7001, 9991, n // 4)])).astype(np.int64)
    gr = GRanges(["chr1"] * n, np.sort(st), width=np.full(n, 10), seqlengths={"chr1": 10_000})
    exclude = GRanges(["chr1"], [5001], end=[6000], seqlengths={"chr1": 10_000})
    tiles, raw, std = segmentTileCounts(gr, L_s=100, exclude=exclude)
    assert len(tiles) == 90 and not np.any((tiles.start <= 6000) & (tiles.end >= 5001))
    qi = O.Index(1, tiles.seqid, tiles.start, tiles.end, minoverlap=8)
    want, _ = O.count_overlaps(qi, gr.seqid, gr.start, gr.end)
    assert np.array_equal(raw, want) and np.allclose(std, want / tiles.width * 100)
    plain = countOverlaps(tiles, gr)
    assert np.all(plain >= raw) and plain.sum() > raw.sum()          # minoverlap = 8 drops the shallow overlaps
    # random stranded ranges, each option
    case = H.random_case(seed=77, n=2000, n_seq=3, n_query=150, stranded=True, wmax=80)
    names = np.array(["a", "b", "c"])
    sl = {k: int(v) for k, v in zip(names, case["seqlen"])}
    y = GRanges(names[case["y"][0]], case["y"][1], end=case["y"][2], strand=case["y"][3], seqlengths=sl)
    q = GRanges(names[case["query"][0]], case["query"][1], end=case["query"][2], strand=case["query"][3], seqlengths=sl)
    for kw in ({}, {"maxgap": 25}, {"minoverlap": 12}):
        qi = O.Index(3, q.seqid, q.start, q.end, q.strand, **kw)
        want, _ = O.count_overlaps(qi, y.seqid, y.start, y.end, y.strand)
        assert np.array_equal(countOverlaps(q, y, **kw), want), kw
    with pytest.raises(ValueError):
        countOverlaps(q, y, maxgap=3, minoverlap=2)


@pytest.mark.parametrize("kind", [0, 1, 2, 4])
@pytest.mark.parametrize("minoverlap", [1, 20, 90])
def test_minoverlap(engine, kind, minoverlap):
    """countOverlaps(x, boots, minoverlap = k): the intersection of the trimmed range and the feature must be k wide."""
    case = H.random_case(seed=700 + kind, n=800, n_seq=3, n_query=90, n_excl=12 if kind == 2 else 0, stranded=kind == 1, n_seg=9,
                         n_states=2, wmax=120)
    if kind == 0:
        case["excl"], case["trim"] = (np.array([0, 1], np.int32), np.array([300, 900], np.int32), np.array([420, 1500], np.int32), None), True
    case["minoverlap"] = minoverlap
    got, ref = _check_counts(engine, case, kind, R=5, seed=6, philox=False)
    case["minoverlap"] = 0
    _, plain = _check_counts(engine, case, kind, R=5, seed=6, philox=False)
    assert np.array_equal(ref["iter_nranges"], plain["iter_nranges"])          # rows do not depend on the query
    if minoverlap > 1:
        assert ref["iter_totals"].sum() < plain["iter_totals"].sum()
    else:
        assert np.array_equal(ref["feature_counts"], plain["feature_counts"])   # minoverlap = 1 is plain overlap
    from nullranges_b200 import EngineError
    with pytest.raises(EngineError, match="cannot both"):
        engine.set_query(*case["query"], maxgap=5, minoverlap=3)


@pytest.mark.parametrize("kind", ALL_KINDS)
@pytest.mark.parametrize("variant", ["plain", "stranded", "drop", "trim", "maxgap", "minoverlap", "points", "unsorted"])
def test_weighted_sums(engine, kind, variant):
    """sum of a numeric column of y over the bootstrapped ranges that overlap each feature (rank path; double
    arithmetic: relative tolerance 1e-9 against a plain summation over the oracle's rows)."""
    from nullranges_b200 import EngineError
    sort = variant != "unsorted" or kind in (0, 5)
    case = H.random_case(seed=800 + kind, n=500, n_seq=3, n_query=60, stranded=variant == "stranded", n_excl=15 if variant == "drop" else 0,
                         n_seg=9, n_states=2, wmax=90, points=variant == "points", sort=sort)
    if variant == "trim":
        case["excl"], case["trim"] = (np.array([0, 1, 2], np.int32), np.array([300, 900, 50], np.int32), np.array([420, 1500, 80], np.int32), None), True
    if variant == "maxgap":
        case["maxgap"] = 30
    if variant == "minoverlap":
        case["minoverlap"] = 15
    plan, y, q, x = H.oracle_objects(case, kind)
    H.load_engine(engine, case, kind)
    w = np.random.default_rng(kind).normal(5.0, 2.0, case["y"][0].size)
    R = 3
    d = H.draw_R(plan, 4, R)
    draws = Draws(n_iter=R, rng_mode=_lib.RNG_HOST, bait_seqid=d[0], bait_start=d[1], dst_tile=d[2] if kind >= 4 else None)
    if engine.force_stream:
        engine.set_weights(w)
        with pytest.raises(EngineError, match="rank path"):
            engine.run_weighted(draws)
        return
    engine.set_weights(w)
    try:
        got = engine.run_weighted(draws)
    except EngineError as e:
        # zero-width rule 1 (a development build, device_views.cuh): the rank path steps aside where the two rules differ,
        # and the weighted statistic exists on the rank path only
        if O.zero_width_rule() == 1 and kind in (1, 4) and "rank path" in str(e):
            pytest.skip("zero-width rule 1: this input takes the streaming kernels, which have no weighted statistic")
        raise
    ref = O.boot_weighted(plan, y, q, w, R, exclude=x, exclude_mode=H.excl_mode(case), draws=d, maxgap=case.get("maxgap", -1),
                          minoverlap=case.get("minoverlap", 0))
    scale = np.abs(w).sum()
    assert np.allclose(got["feature_sums"], ref["feature_sums"], rtol=1e-9, atol=1e-9 * scale)
    assert np.allclose(got["iter_sums"], ref["iter_sums"], rtol=1e-9, atol=1e-9 * scale)
    # unit weights reproduce the counts exactly
    engine.set_weights(np.ones_like(w))
    ones = engine.run_weighted(draws)
    cnt = engine.run_counts(draws)
    assert np.array_equal(ones["feature_sums"], cnt["feature_counts"].astype(np.float64))


def test_many_iterations_cross_batches(engine):
    """R larger than one launch batch (2048) and than the D2H pipeline chunking: every entry point that loops
    over batches gives the same answer as the oracle / as one-shot runs."""
    case = H.random_case(seed=901, n=400, n_seq=3, n_query=37, n_excl=8, wmax=80)
    plan, y, q, x = H.oracle_objects(case, 0)
    H.load_engine(engine, case, 0)
    R = 5000
    ref = O.boot_counts(plan, y, q, R, exclude=x, seed=3, iter0=17, nthreads=O.max_threads())
    got = engine.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=3, iter0=17))
    for k in ("iter_nranges", "iter_totals", "feature_counts"):
        assert np.array_equal(got[k], ref[k]), k
    tot_only = engine.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=3, iter0=17), want_counts=False)
    assert np.array_equal(tot_only["iter_totals"], ref["iter_totals"]) and tot_only["feature_counts"] is None
    obs, _ = engine.count_observed()
    m = engine.run_moments(Draws(n_iter=R, seed=3, iter0=17), observed=obs, batch_iter=1700)
    fc = ref["feature_counts"].astype(np.int64)
    assert np.array_equal(m["sum"], fc.sum(0)) and np.array_equal(m["sumsq"], (fc * fc).sum(0))
    assert np.array_equal(m["n_ge_observed"], (fc >= obs[None, :]).sum(0))
    rows = engine.run_materialize(Draws(n_iter=2500, rng_mode=_lib.RNG_PHILOX, seed=3, iter0=17))
    assert np.array_equal(np.diff(rows["iter_offsets"]), ref["iter_nranges"][:2500])
    if not engine.force_stream:
        engine.set_weights(np.ones(case["y"][0].size))
        ws = engine.run_weighted(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=3, iter0=17))
        assert np.array_equal(ws["feature_sums"], ref["feature_counts"].astype(np.float64))
        assert np.array_equal(ws["iter_sums"], ref["iter_totals"].astype(np.float64))


def test_host_permutation_is_validated(engine):
    from nullranges_b200 import EngineError
    case = H.random_case(seed=5, n=100, n_seq=2, n_query=10)
    plan, *_ = H.oracle_objects(case, 5)
    H.load_engine(engine, case, 5)
    d = H.draw_R(plan, 1, 2)
    bad = d[2].copy()
    bad[1, 0] = bad[1, 1]          # two blocks sent to the same tile
    with pytest.raises(EngineError, match="permutation"):
        engine.run_counts(Draws(n_iter=2, rng_mode=_lib.RNG_HOST, bait_seqid=d[0], bait_start=d[1], dst_tile=bad))


@pytest.mark.parametrize("kind", [0, 1, 2, 5])
def test_many_sequences(engine, kind):
    """Hundreds of seqlevels (assemblies with contigs): per-sequence tables, draws in proportion to length,
    device permutations with a wide sequence field."""
    case = H.random_case(seed=950 + kind, n=3000, n_seq=300, max_len=4000, n_query=400, n_excl=40, stranded=kind == 1, L_b=700,
                         n_seg=600, n_states=3, wmax=150)
    _check_counts(engine, case, kind, R=3, seed=5, philox=False)
    _check_counts(engine, case, kind, R=3, seed=5, philox=True)


def test_cfg4_full_size(engine):
    """BASELINE configs[3] at full size: 10M width-1 points, withinChrom = TRUE, 200k enhancers, blockLength 5e5;
    8 Philox iterations against the oracle (counts and per-feature moments), plus a shard taken elsewhere in the
    iteration range (what a second GPU would compute)."""
    if engine.force_stream:
        pytest.skip("the streaming path at this size is covered by cfg 2 / cfg 3")
    from nullranges_b200 import synth
    y, x = synth.points(10_000_000, seed=6), synth.enhancers(200_000, seed=7)
    case = _synth_case(y, x)
    plan, yi, qi, _ = H.oracle_objects(case, 1)
    assert H.load_engine(engine, case, 1) == plan.B
    R = 8
    ref = O.boot_counts(plan, yi, qi, R, seed=7, iter0=5000, nthreads=O.max_threads())
    got = engine.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=7, iter0=5000))
    for k in ("iter_nranges", "iter_totals", "feature_counts"):
        assert np.array_equal(got[k], ref[k]), k
    obs, tot = engine.count_observed()
    m = engine.run_moments(Draws(n_iter=R, seed=7, iter0=5000), observed=obs, batch_iter=3)
    fc = ref["feature_counts"].astype(np.int64)
    assert np.array_equal(m["sum"], fc.sum(0)) and np.array_equal(m["sumsq"], (fc * fc).sum(0))
    assert np.array_equal(m["n_ge_observed"], (fc >= obs[None, :]).sum(0))
    assert abs(got["iter_totals"].mean() / tot - 1.0) < 0.01


def test_state_resets_with_new_ranges(engine):
    """nrb200_set_ranges() starts a new problem: query, exclude set, query options and weights of the previous one are gone."""
    from nullranges_b200 import EngineError
    case = H.random_case(seed=11, n=200, n_query=20, n_excl=5)
    H.load_engine(engine, case, 0)
    engine.set_weights(np.ones(200))
    engine.set_ranges(case["seqlen"], *H.random_case(seed=12, n=50)["y"])
    engine.set_plan(_lib.UNSEG_ACROSS, case["L_b"])
    with pytest.raises(EngineError, match="set_query"):
        engine.count_observed()
    engine.set_query(*case["query"])
    if not engine.force_stream:
        with pytest.raises(EngineError, match="set_weights"):
            engine.run_weighted(Draws(n_iter=1, seed=1))
    got = engine.run_counts(Draws(n_iter=2, seed=1))
    assert got["iter_nranges"].sum() > 0


def test_set_ranges_in_two_halves(engine):
    """nrb200_set_ranges_begin / _end: query, exclude and plan may be set in between; results equal the one-call form;
    y's errors surface at _end; running before _end is a state error."""
    from nullranges_b200 import EngineError
    case = H.random_case(seed=21, n=3000, n_query=200, n_excl=20)
    H.load_engine(engine, case, 0)
    want = engine.run_counts(Draws(n_iter=5, seed=3))
    sid, st, en, sd = case["y"]
    engine.set_ranges(case["seqlen"], sid, st, en, sd, defer=True)
    engine.set_exclude(*case["excl"])
    engine.set_plan(_lib.UNSEG_ACROSS, case["L_b"])
    engine.set_query(*case["query"])
    with pytest.raises(EngineError, match="set_ranges_end"):
        engine.run_counts(Draws(n_iter=1, seed=3))
    engine.finish_ranges()
    got = engine.run_counts(Draws(n_iter=5, seed=3))
    for k in ("iter_totals", "iter_nranges", "feature_counts"):
        np.testing.assert_array_equal(got[k], want[k])
    with pytest.raises(EngineError, match="set_ranges_begin"):
        engine.finish_ranges()
    bad = st.copy()
    bad[7] = 0
    engine.set_ranges(case["seqlen"], sid, bad, en, sd, defer=True)
    with pytest.raises(EngineError, match="start at position >= 1"):
        engine.finish_ranges()
    with pytest.raises(EngineError):
        engine.set_plan(_lib.UNSEG_ACROSS, case["L_b"])  # no ranges: a failed _end leaves the handle without y


def test_resident_tensors_view_the_engine_buffers():
    """sharding.resident_tensors(): zero-copy torch views of the device-resident results equal what run_counts copies out."""
    import torch
    from nullranges_b200 import Engine
    from nullranges_b200.sharding import resident_tensors
    case = H.random_case(seed=31, n=2000, n_query=150)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        eng = Engine(0, stream=stream.cuda_stream)
        H.load_engine(eng, case, 0)
        d = Draws(n_iter=7, seed=5)
        want = eng.run_counts(d)
        eng.run_counts_resident(d, True, False)
        views = resident_tensors(eng, 7)
        got = {k: v.cpu().numpy() for k, v in views.items()}
        eng.close()
    for k in ("iter_totals", "iter_nranges", "feature_counts"):
        np.testing.assert_array_equal(got[k], want[k])


def test_weighted_moments_equal_the_matrix_moments():
    """nrb200_run_weighted_moments(): per-feature sum / sum of squares / #{>= observed} of the weighted sums equal the
    moments of the [R, G] matrix nrb200_run_weighted() returns, across launch batches (R > NRB200_MAX_BATCH), and unit
    weights reproduce the integer moments of the counts."""
    from nullranges_b200 import Engine
    case = H.random_case(seed=41, n=3000, n_seq=3, n_query=90, wmax=120)
    eng = Engine(0)
    H.load_engine(eng, case, 0)
    w = np.random.default_rng(3).gamma(2.0, 3.0, case["y"][0].size)
    eng.set_weights(w)
    R = 2500  # two launch batches
    d = Draws(n_iter=R, seed=17)
    full = eng.run_weighted(d)
    fs = full["feature_sums"]
    obs = np.quantile(fs, 0.9, axis=0)
    got = eng.run_weighted_moments(d, observed=obs)
    assert np.allclose(got["sum"], fs.sum(0), rtol=1e-9)
    assert np.allclose(got["sumsq"], (fs * fs).sum(0), rtol=1e-9)
    assert np.array_equal(got["n_ge_observed"], (fs >= obs).sum(0))
    assert np.allclose(got["iter_sums"], full["iter_sums"], rtol=1e-12)
    eng.set_weights(np.ones_like(w))
    ones = eng.run_weighted_moments(d)
    mom = eng.run_moments(d)
    assert np.array_equal(ones["sum"], mom["sum"].astype(np.float64)) and np.array_equal(ones["sumsq"], mom["sumsq"].astype(np.float64))
    assert not ones["n_ge_observed"].any()
    eng.close()


def test_boot_moments_front_end():
    """bootMoments(): mean / sd / z / empirical p per feature equal what the [R, G] matrix of bootCounts() gives."""
    from nullranges_b200 import GRanges, bootCounts, bootMoments, countOverlaps
    rng = np.random.default_rng(5)
    lv, L = ["chr1", "chr2"], {"chr1": 60000, "chr2": 40000}
    n = 4000
    sq = np.sort(rng.integers(0, 2, n))
    st = rng.integers(1, 39000, n)
    y = GRanges(np.array(lv, dtype=object)[sq], st, width=rng.integers(20, 200, n), seqlengths=L, score=rng.gamma(2.0, 2.0, n))
    y = y[np.lexsort((y.end, y.start, y.seqid))]
    g = 50
    xs = np.sort(rng.integers(0, 2, g))
    x = GRanges(np.array(lv, dtype=object)[xs], rng.integers(1, 38000, g), width=rng.integers(100, 1500, g), seqlengths=L)
    R = 300
    bc = bootCounts(y, x, 5000, R=R, seed=9, weights="score")
    bm = bootMoments(y, x, 5000, R=R, seed=9)
    fc = bc.feature_counts.astype(np.float64)
    assert np.allclose(bm.mean, fc.mean(0)) and np.allclose(bm.sd, fc.std(0, ddof=1))
    obs = countOverlaps(x, y)
    assert np.array_equal(bm.observed, obs)
    assert np.allclose(bm.p_ge, ((fc >= obs).sum(0) + 1) / (R + 1))
    ok = bm.sd > 0
    assert np.allclose(bm.z[ok], ((obs - fc.mean(0)) / fc.std(0, ddof=1))[ok])
    fs = bc.feature_sums
    target = np.median(fs, axis=0)
    bw = bootMoments(y, x, 5000, R=R, seed=9, weights="score", observed=target)
    assert np.allclose(bw.mean, fs.mean(0), rtol=1e-9) and np.allclose(bw.sd, fs.std(0, ddof=1), rtol=1e-6)
    assert np.allclose(bw.p_ge, ((fs >= target).sum(0) + 1) / (R + 1))


# ---- wire format of a host-bound count matrix ----------------------------------------------------------------
# Plain numpy result arrays are pageable, so every run_counts() above already travelled as uint16 and was widened
# on the host (run.cuh).  Here: the int32 path with the format switched off, pinned destinations, and a count that
# does not fit uint16 (the chunk is redone in int32).

def test_host_matrix_in_more_batches_than_chunk_slots(monkeypatch):
    """A host-bound matrix leaves in at most 64 chunks (one copy event and one overflow-flag slot per chunk in flight): with
    small launch batches (NRB200_MAX_BATCH) and many iterations the chunk grows instead of the slots being reused."""
    from nullranges_b200 import Engine
    monkeypatch.setenv("NRB200_MAX_BATCH", "8")
    eng = Engine(0)
    try:
        case = H.random_case(seed=88, n=400, n_seq=3, n_query=9)
        plan, y, q, x = H.oracle_objects(case, 0)
        H.load_engine(eng, case, 0)
        R = 2000  # 250 batches of 8 iterations if the chunk did not grow
        got = eng.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=3))
        ref = O.boot_counts(plan, y, q, R, seed=3, iter0=0)
        for k in ("iter_nranges", "iter_totals", "feature_counts"):
            assert np.array_equal(got[k], ref[k]), k
    finally:
        eng.close()


@pytest.mark.parametrize("wire", ["0", "1", "2"])
@pytest.mark.parametrize("pinned", [False, True])
def test_wire_format_variants(wire, pinned, monkeypatch):
    import torch
    from nullranges_b200 import Engine
    monkeypatch.setenv("NRB200_WIRE16", wire)
    monkeypatch.setenv("NRB200_CHUNKS", "3")
    eng = Engine(0)
    try:
        case = H.random_case(seed=71, n=3000, n_seq=4, n_query=257, n_excl=15, n_seg=8)
        plan, y, q, x = H.oracle_objects(case, 2)
        H.load_engine(eng, case, 2)
        R = 200  # several chunks (the pipeline makes at least 64 iterations per chunk)
        out = None
        if pinned:
            keep = torch.empty((R, 257), dtype=torch.int32, pin_memory=True)
            out = {"feature_counts": keep.numpy()}
        got = eng.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=12, iter0=40), out=out)
        ref = O.boot_counts(plan, y, q, R, exclude=x, seed=12, iter0=40)
        for k in ("iter_nranges", "iter_totals", "feature_counts"):
            assert np.array_equal(got[k], ref[k]), k
        expect16 = wire != "0"  # "2" = where it pays: always for one GPU, pinned destination or not (run.cuh)
        assert got["stats"]["d2h_bytes"] == 2 * R * 8 + R * 257 * (2 if expect16 else 4)
    finally:
        eng.close()


def test_wire_format_overflow_falls_back_to_int32():
    """70,000 ranges inside one feature: the count exceeds 65535, the chunk's overflow flag goes up and the chunk is
    computed again with int32 output; the other chunks stay uint16."""
    from nullranges_b200 import Engine
    n = 70_000
    seqlen = np.array([5000, 4000], np.int64)
    sid = np.concatenate([np.zeros(n, np.int32), np.ones(500, np.int32)])
    rng = np.random.default_rng(2)
    st = np.concatenate([np.full(n, 100), np.sort(rng.integers(1, 3500, 500))]).astype(np.int32)
    en = (st + np.concatenate([np.full(n, 100), rng.integers(1, 300, 500)])).astype(np.int32)
    o = np.lexsort((en, st, sid))
    case = {"seqlen": seqlen, "y": (sid[o], st[o], en[o], None), "excl": None, "seg": None, "L_b": 1000,
            "query": (np.array([0, 0, 1], np.int32), np.array([1, 900, 10], np.int32), np.array([1000, 1500, 3000], np.int32), None)}
    plan, y, q, x = H.oracle_objects(case, 1)
    eng = Engine(0)
    try:
        H.load_engine(eng, case, 1)
        R = 150
        got = eng.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=3))
        ref = O.boot_counts(plan, y, q, R, seed=3)
        assert ref["feature_counts"].max() > 65535
        for k in ("iter_nranges", "iter_totals", "feature_counts"):
            assert np.array_equal(got[k], ref[k]), k
    finally:
        eng.close()


@pytest.mark.parametrize("pinned", [False, True])
def test_uint16_result(pinned):
    """nrb200_run_counts_u16: the matrix delivered as uint16 (pinned: DMA straight into the caller's buffer; pageable:
    staged and copied by the drain team); saturation is flagged, not hidden."""
    import torch
    from nullranges_b200 import Engine
    eng = Engine(0)
    try:
        case = H.random_case(seed=72, n=3000, n_seq=4, n_query=130, n_excl=10, n_seg=8)
        plan, y, q, x = H.oracle_objects(case, 0)
        H.load_engine(eng, case, 0)
        R = 150
        out = None
        if pinned:
            keep = torch.empty((R, 130), dtype=torch.uint16, pin_memory=True)
            out = {"feature_counts": keep.numpy()}
        got = eng.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=12, iter0=40), out=out, counts_dtype=np.uint16)
        ref = O.boot_counts(plan, y, q, R, exclude=x, seed=12, iter0=40)
        assert got["feature_counts"].dtype == np.uint16 and not got["overflow"]
        assert np.array_equal(got["feature_counts"], ref["feature_counts"]) and np.array_equal(got["iter_totals"], ref["iter_totals"])
        assert got["stats"]["d2h_bytes"] == 2 * R * 8 + R * 130 * 2
        # 70,000 ranges inside one feature: saturated and flagged; the totals stay exact
        n = 70_000
        sid = np.zeros(n, np.int32)
        st = np.full(n, 100, np.int32)
        big = {"seqlen": np.array([5000], np.int64), "y": (sid, st, st + 50, None), "excl": None, "seg": None, "L_b": 5000,
               "query": (np.array([0, 0], np.int32), np.array([1, 2000], np.int32), np.array([1000, 3000], np.int32), None)}
        H.load_engine(eng, big, 1)
        got = eng.run_counts(Draws(n_iter=4, rng_mode=_lib.RNG_PHILOX, seed=1), counts_dtype=np.uint16)
        assert got["overflow"] and got["feature_counts"].max() == 65535 and int(got["iter_totals"][0]) == n
    finally:
        eng.close()


@pytest.mark.parametrize("n_query,L_b", [(1500, 2000), (6000, 2000), (30000, 20000)])
def test_stream_kernel_large_query_slices(n_query, L_b):
    """The tile-group streaming kernel with query slices near and beyond its shared-memory capacity (2048 features):
    groups are cut down until their slice fits (more than 48 KB of shared memory needs the opt-in), and a single tile
    with a larger slice sends the plan to the block-centric kernel.  All three against the oracle."""
    from nullranges_b200 import Engine
    eng = Engine(0, force_stream=True)
    try:
        case = H.random_case(seed=333, n=4000, n_seq=3, max_len=60000, n_query=n_query, wmax=200, L_b=L_b, n_excl=12)
        plan, y, q, x = H.oracle_objects(case, 0)
        H.load_engine(eng, case, 0)
        R = 5
        got = eng.run_counts(Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=6, iter0=2))
        ref = O.boot_counts(plan, y, q, R, exclude=x, seed=6, iter0=2)
        for k in ("iter_nranges", "iter_totals", "feature_counts"):
            assert np.array_equal(got[k], ref[k]), k
    finally:
        eng.close()
