"""The reference's own tests (tests/testthat/test_boot.R, test_trim_exclude.R), translated.

They pin structural properties only.  Each runs twice: against the oracle on the CPU, and --
marked gpu -- against the product through its bootRanges() front end, reading like the R tests.
"""
import numpy as np
import pytest

from oracle import oracle as O

SEQLEN = [300, 450, 200]                                             # test_boot.R:15
CHROM = np.repeat([0, 1, 2], [4, 5, 2])                              # :7
START = np.array([1, 101, 121, 201, 101, 201, 216, 231, 401, 1, 101])  # :9-11
WIDTH = np.array([20, 5, 5, 30, 20, 5, 5, 5, 30, 80, 40])            # :12-14
L_B, R = 100, 5
SEG = {"chrom": np.array([0, 0, 1, 1, 2]), "start": np.array([1, 151, 1, 226, 1]), "end": np.array([150, 300, 225, 450, 200]),
       "state": np.array([1, 2, 1, 2, 1])}                           # :80-82


def _oracle_boot(kind, seed=1, seg=None):
    plan = O.Plan(kind, SEQLEN, L_B, seg)
    y = O.Index(3, CHROM, START, START + WIDTH - 1)
    d = plan.draw_R(O.RRng(seed), R)
    return [O.iteration_rows(plan, y, d[0][r], d[1][r], d[2][r]) for r in range(R)]


def _state_at(chrom, pos):
    for c, s, e, st in zip(SEG["chrom"], SEG["start"], SEG["end"], SEG["state"]):
        if c == chrom and s <= pos <= e:
            return st
    return 0


def _check_permute_within(iters):      # test_boot.R:27-35
    for chrom_new, origin in iters:
        assert np.array_equal(np.bincount(origin, minlength=3), [4, 5, 2])
        assert np.array_equal(np.bincount(chrom_new, minlength=3), [4, 5, 2])


def _check_bootstrap_within(iters):    # :41-48: one chromosome of origin per output chromosome
    for chrom_new, origin in iters:
        assert np.array_equal(chrom_new, origin)


def _check_permute_across(iters):      # :54-61
    assert [c.size for c, _ in iters] == [11] * R


def _check_bootstrap_across(iters):    # :67-74
    assert not all(c.size == 11 for c, _ in iters)


def test_oracle_unsegmented_properties():
    pick = lambda rows: [(r["chrom"], CHROM[r["src"]]) for r in rows]
    _check_permute_within(pick(_oracle_boot(O.PERMUTE_WITHIN)))
    _check_bootstrap_within(pick(_oracle_boot(O.UNSEG_WITHIN)))
    _check_permute_across(pick(_oracle_boot(O.PERMUTE_ACROSS)))
    _check_bootstrap_across(pick(_oracle_boot(O.UNSEG_ACROSS)))


def test_oracle_segmented_same_state_rate():   # :86-96
    same = total = 0
    for rows in _oracle_boot(O.SEG_PROP, seg=SEG):
        for c, s, e, src in zip(rows["chrom"], rows["start"], rows["end"], rows["src"]):
            orig = _state_at(CHROM[src], START[src])
            # join_overlap_left(seg): one row per overlapped segment
            for sc, ss, se, st in zip(SEG["chrom"], SEG["start"], SEG["end"], SEG["state"]):
                if sc == c and ss <= e and se >= s:
                    total += 1
                    same += int(st == orig)
    assert same / total > 0.5


# ---- the same, through the product ------------------------------------------------------------

def _gr():
    from nullranges_b200 import GRanges
    names = np.array(["chr1", "chr2", "chr3"])[CHROM]
    return GRanges(names, START, width=WIDTH, seqlengths={"chr1": 300, "chr2": 450, "chr3": 200}, chrom=CHROM + 1)


def _per_iter(br):
    return [(br.seqid[br.iter == r], br.mcols["chrom"][br.iter == r] - 1) for r in range(1, R + 1)]


@pytest.mark.gpu
def test_product_unsegmented_properties():
    from nullranges_b200 import bootRanges, set_seed
    gr = _gr()
    set_seed(1)
    _check_permute_within(_per_iter(bootRanges(gr, L_B, R, type="permute", withinChrom=True)))
    set_seed(1)
    _check_bootstrap_within(_per_iter(bootRanges(gr, L_B, R, type="bootstrap", withinChrom=True)))
    set_seed(1)
    _check_permute_across(_per_iter(bootRanges(gr, L_B, R, type="permute", withinChrom=False)))
    set_seed(1)
    br = bootRanges(gr, L_B, R, type="bootstrap", withinChrom=False)
    _check_bootstrap_across(_per_iter(br))
    # SURVEY 8c worked example: rows per iteration under set.seed(1)
    assert np.bincount(br.iter, minlength=R + 1)[1:].tolist() == [14, 12, 17, 17, 10]
    assert "block" not in br.mcols and "iter" in br.mcols      # R/bootRanges.R:122-126


@pytest.mark.gpu
def test_product_segmented_same_state_rate():
    from nullranges_b200 import GRanges, bootRanges, set_seed
    seg = GRanges(np.array(["chr1", "chr2", "chr3"])[SEG["chrom"]], SEG["start"], end=SEG["end"],
                  seqlengths={"chr1": 300, "chr2": 450, "chr3": 200}, state=SEG["state"])
    gr = _gr()
    gr.mcols["state"] = np.array([_state_at(c, s) for c, s in zip(CHROM, START)])
    set_seed(1)
    br = bootRanges(gr, L_B, R, seg)
    same = total = 0
    for c, s, e, orig in zip(br.seqid, br.start, br.end, br.mcols["state"]):
        for sc, ss, se, st in zip(SEG["chrom"], SEG["start"], SEG["end"], SEG["state"]):
            if sc == c and ss <= e and se >= s:
                total += 1
                same += int(st == orig)
    assert same / total > 0.5


@pytest.mark.gpu
def test_product_roxygen_example():
    """R/bootRanges.R:65-69 with the SURVEY 8c worked values."""
    from nullranges_b200 import GRanges, bootRanges, set_seed
    set_seed(1)
    gr = GRanges(["chr1"] * 5, np.arange(5) * 10 + 1, width=[5] * 5, seqlengths={"chr1": 50})
    br = bootRanges(gr, blockLength=10, storeBlockLength=True)
    assert br.start.tolist() == [5, 13, 25, 36, 39, 49] and br.end.tolist() == [9, 17, 29, 40, 43, 50]
    assert br.iter.tolist() == [1] * 6 and br.mcols["blockLength"].tolist() == [10] * 6


# ---- tests/testthat/test_trim_exclude.R ----------------------------------------------------------

def _trim_fixture():
    start = 1 + 2 * np.arange(46)                                  # :4
    strand = np.tile([1, 2], 23).astype(np.int8)                   # rep(c("+","-"), length = 46)
    order = np.lexsort((start + 9, start, np.array([2, 0, 1])[strand]))  # y <- sort(y): + before -
    return start[order], start[order] + 9, strand[order]


def test_oracle_trim_exclude():
    st, en, sd = _trim_fixture()
    plan = O.Plan(O.UNSEG_ACROSS, [100], 20)
    y = O.Index(1, np.zeros(46, np.int32), st, en, sd)
    g = O.gaps(1, [100], [0], [36], [45])
    assert (g.start.tolist(), g.end.tolist()) == ([1, 46], [35, 100])
    for seed in range(1, 6):
        d = plan.draw_R(O.RRng(seed), 1)
        rows = O.iteration_rows(plan, y, d[0][0], d[1][0], d[2][0], exclude=g, exclude_mode=O.EXCLUDE_TRIM)
        assert rows["src"].size > 0
        assert not np.any((rows["start"] <= 45) & (rows["end"] >= 36))   # expect_true(!any(overlapsAny(br, exclude)))


@pytest.mark.gpu
def test_product_trim_exclude():
    from nullranges_b200 import GRanges, bootRanges, set_seed
    st, en, sd = _trim_fixture()
    y = GRanges(["chr1"] * 46, st, end=en, strand=sd, seqlengths={"chr1": 100}, block=np.ones(46))
    exclude = GRanges(["chr1"], [36], end=[45], seqlengths={"chr1": 100})
    assert y.is_sorted()
    for seed in range(1, 6):
        set_seed(seed)
        br = bootRanges(y, blockLength=20, exclude=exclude, excludeOption="trim")
        assert len(br) > 0 and not np.any((br.start <= 45) & (br.end >= 36))
        assert "block" not in br.mcols                                # the input's own `block` column is dropped (:126)
        assert set(np.unique(br.strand)) <= {1, 2}                     # strands ride along with the pieces
    with pytest.raises(ValueError):
        bootRanges(y, blockLength=20, excludeOption="trim")          # strand(NULL): the reference errors too


@pytest.mark.gpu
@pytest.mark.parametrize("kind_args", [dict(type="bootstrap", withinChrom=True), dict(type="permute", withinChrom=True), dict(seg=True)])
def test_product_row_order_follows_given_y(kind_args):
    """An unsorted y (allowed within chromosomes and with seg): rows come out in the reference's order -- a block's
    hits by index into y as given; the permute variants chromosome by chromosome in y's own order."""
    from nullranges_b200 import GRanges, bootRanges, set_seed
    rng = np.random.default_rng(3)
    n = 60
    seqid = rng.integers(0, 3, n)
    start = (1 + rng.random(n) * (np.array(SEQLEN)[seqid] - 30)).astype(np.int64)
    gr = GRanges(np.array(["chr1", "chr2", "chr3"])[seqid], start, width=rng.integers(1, 25, n),
                 seqlengths={"chr1": 300, "chr2": 450, "chr3": 200}, idx=np.arange(n))
    assert not gr.is_sorted()
    kw = dict(kind_args)
    if kw.pop("seg", False):
        kw["seg"] = GRanges(np.array(["chr1", "chr2", "chr3"])[SEG["chrom"]], SEG["start"], end=SEG["end"],
                            seqlengths={"chr1": 300, "chr2": 450, "chr3": 200}, state=SEG["state"])
    set_seed(2)
    br = bootRanges(gr, L_B, 3, storeBlockStart=True, **kw)
    idx, it = br.mcols["idx"], br.iter
    for r in (1, 2, 3):
        m = it == r
        if kw.get("type") == "permute":
            key = np.stack([gr.seqid[idx[m]], idx[m]], axis=1)      # chromosome of origin, then y's own order
            assert np.array_equal(key, key[np.lexsort((key[:, 1], key[:, 0]))])
    if kw.get("type") != "permute":
        # compare with the oracle's (block, y-index) order directly
        from nullranges_b200 import Draws, _lib, get_engine
        eng = get_engine(0)
        kind = _lib.SEG_PROP if "seg" in kw else _lib.UNSEG_WITHIN
        seg = SEG if "seg" in kw else None
        plan = O.Plan(kind, SEQLEN, L_B, seg)
        y = O.Index(3, gr.seqid, gr.start, gr.end)
        d = plan.draw_R(O.RRng(2), 3)
        want = np.concatenate([np.stack([O.iteration_rows(plan, y, d[0][r], d[1][r], d[2][r])[k] for k in ("chrom", "start", "end", "src")], axis=1)
                               for r in range(3)])
        got = np.stack([br.seqid, br.start, br.end, idx], axis=1)
        assert np.array_equal(got, want)


# ---- vignette fixtures: toy segmentation + uniform toy ranges, one-region case ----------

TOY_SEG = {"chrom": np.array([0, 0, 0, 0, 1, 1, 1]), "start": np.array([1, 101, 201, 401, 1, 201, 301]),
           "end": np.array([100, 200, 400, 500, 200, 300, 400]), "state": np.array([1, 2, 1, 3, 3, 2, 1])}   # Rmd:796-804
TOY_LEN = [500, 400]


def _toy_ranges():
    """vignettes/bootRanges.Rmd:816-831 under set.seed(1), with R's own stream."""
    g = O.RRng(1)
    chrom = np.sort(g.sample(2, 200, replace=True) - 1)
    start = np.rint(g.runif(200, 1, 500 - 10 + 1)).astype(np.int64)
    end = start + 9
    keep = ~((chrom == 1) & (end > 400))
    chrom, start, end = chrom[keep], start[keep], end[keep]
    o = np.lexsort((end, start, chrom))
    chrom, start, end = chrom[o], start[o], end[o]
    state = np.zeros(chrom.size, np.int64)     # findOverlaps(gr, seg, type = "within", select = "first")
    for c, s, e, st in zip(TOY_SEG["chrom"], TOY_SEG["start"], TOY_SEG["end"], TOY_SEG["state"]):
        inside = (chrom == c) & (start >= s) & (end <= e) & (state == 0)
        state[inside] = st
    m = state > 0
    return chrom[m], start[m], end[m], state[m]


def _same_state_rate(rows, src_state):
    same = total = 0
    for c, s, e, src in zip(rows["chrom"], rows["start"], rows["end"], rows["src"]):
        for sc, ss, se, st in zip(TOY_SEG["chrom"], TOY_SEG["start"], TOY_SEG["end"], TOY_SEG["state"]):
            if sc == c and ss <= e and se >= s:
                total += 1
                same += int(st == src_state[src])
    return same / max(total, 1)


@pytest.mark.parametrize("kind,L_b", [(O.SEG_NOPROP, 25), (O.SEG_PROP, 50)])      # Rmd:841-858
def test_oracle_toy_segmentation(kind, L_b):
    chrom, start, end, state = _toy_ranges()
    assert 150 < chrom.size <= 200
    plan = O.Plan(kind, TOY_LEN, L_b, TOY_SEG)
    if kind == O.SEG_PROP:   # L_b_per_state = round(50 * c(400, 200, 300) / 900)
        assert [plan.state_block_length(s) for s in (1, 2, 3)] == [22, 11, 17]
    y = O.Index(2, chrom, start, end)
    d = plan.draw_R(O.RRng(1), 1)
    rows = O.iteration_rows(plan, y, d[0][0], d[1][0], d[2][0])
    assert rows["src"].size > 100
    assert np.all(rows["start"] >= 1) and np.all(rows["end"] <= np.array(TOY_LEN)[rows["chrom"]])
    assert _same_state_rate(rows, state) > 0.6      # "some ranges from adjacent states are allowed to be placed within different states"


@pytest.mark.gpu
@pytest.mark.parametrize("prop,L_b", [(False, 25), (True, 50)])
def test_product_toy_segmentation(prop, L_b):
    from nullranges_b200 import GRanges, bootRanges, set_seed
    chrom, start, end, state = _toy_ranges()
    names = np.array(["chr1", "chr2"])
    sl = {"chr1": 500, "chr2": 400}
    gr = GRanges(names[chrom], start, end=end, seqlengths=sl, state=state)
    seg = GRanges(names[TOY_SEG["chrom"]], TOY_SEG["start"], end=TOY_SEG["end"], seqlengths=sl, state=TOY_SEG["state"])
    set_seed(1)
    br = bootRanges(gr, blockLength=L_b, seg=seg, proportionLength=prop)
    plan = O.Plan(O.SEG_PROP if prop else O.SEG_NOPROP, TOY_LEN, L_b, TOY_SEG)
    d = plan.draw_R(O.RRng(1), 1)
    rows = O.iteration_rows(plan, O.Index(2, chrom, start, end), d[0][0], d[1][0], d[2][0])
    assert np.array_equal(br.seqid, rows["chrom"]) and np.array_equal(br.start, rows["start"]) and np.array_equal(br.end, rows["end"])
    assert np.array_equal(br.mcols["state"], state[rows["src"]])


def _one_region_case():
    """vignettes/bootRanges.Rmd:733-749 with synthetic peaks in place of the DHS data (not in the tree)."""
    rng = np.random.default_rng(9)
    start = np.sort(10_000_001 + rng.integers(0, 1_000_000 - 400, 800))
    return start, start + rng.integers(150, 400, 800)


def test_oracle_one_region_segment():
    start, end = _one_region_case()
    L = 248_956_422
    seg = {"chrom": np.zeros(3, np.int64), "start": np.array([1, 10_000_001, 11_000_001]), "end": np.array([10_000_000, 11_000_000, L]),
           "state": np.array([1, 2, 1])}
    plan = O.Plan(O.SEG_NOPROP, [L], 100_000, seg)
    assert plan.B == 100 + 10 + 2380
    y = O.Index(1, np.zeros(800, np.int64), start, end)
    for seed in range(1, 30):
        try:
            d = plan.draw_R(O.RRng(seed), 1)
        except ValueError:
            continue     # the region drew no block: the reference stops as well (R/bootRanges.R:303)
        rows = O.iteration_rows(plan, y, d[0][0], d[1][0], d[2][0])
        inside = (rows["end"] >= 10_000_001 - 100_000) & (rows["start"] <= 11_000_000 + 100_000)
        assert rows["src"].size > 0 and inside.mean() > 0.8   # the data are bootstrapped "just in this region"
        break
    else:
        pytest.fail("no seed produced a draw in the region")


@pytest.mark.gpu
def test_product_one_region_segment():
    from nullranges_b200 import GRanges, bootCounts, bootRanges, oneRegionSegment, set_seed
    start, end = _one_region_case()
    L = 248_956_422
    y = GRanges(["chr1"] * 800, start, end=end, seqlengths={"chr1": L})
    region = GRanges(["chr1"], [10_000_001], width=[1_000_000], seqlengths={"chr1": L})
    x = GRanges(["chr1"] * 10, 10_000_001 + np.arange(10) * 100_000, width=[10_000] * 10, seqlengths={"chr1": L})
    seg = oneRegionSegment(region, seqlength=L)
    assert seg.start.tolist() == [1, 10_000_001, 11_000_001] and seg.mcols["state"].tolist() == [1, 2, 1]
    bc = bootCounts(y, x, 100_000, R=200, seg=seg, proportionLength=False, rng="philox", seed=3)
    assert bc.feature_counts.shape == (200, 10)
    obs = np.array([np.sum((y.start <= e) & (y.end >= s)) for s, e in zip(x.start, x.end)])
    assert abs(bc.feature_counts.mean() / obs.mean() - 1.0) < 0.25     # the bootstrap keeps the region's density
    for seed in range(1, 30):
        set_seed(seed)
        try:
            boot = bootRanges(y, blockLength=100_000, R=1, seg=seg, proportionLength=False)
        except ValueError:
            continue
        assert ((boot.end >= 9_900_001) & (boot.start <= 11_100_000)).mean() > 0.8
        break
