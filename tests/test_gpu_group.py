"""Device groups (nrb200_config.n_devices > 1): a group handle shards the iterations of a run over its GPUs and must
deliver exactly what one GPU delivers (replicate() at R/bootRanges.R:90: independent iterations, one concatenated
result, :120-122).  On a one-GPU box the members share the GPU (NRB200_ALLOW_DUP_DEVICES): same code path -- worker
threads, shared window lists, sharded draws, direct writes into the caller's rows -- without the speed-up; with two or
more GPUs the same tests run across real devices as well."""
import os

import numpy as np
import pytest

from nullranges_b200 import Draws, Engine, GRanges, _lib, bootCounts, bootMoments
from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _device_lists():
    out = [[0, 0], [0, 0, 0]]
    n = _n_gpus()
    if n >= 2:
        out.append([0, 1])
    if n >= 4:
        out.append([0, 1, 2, 3])
    if n >= 8:
        out.append(list(range(8)))
    return out


@pytest.fixture(scope="module")
def groups():
    os.environ["NRB200_ALLOW_DUP_DEVICES"] = "1"
    made = {tuple(d): Engine(d) for d in _device_lists()}
    single = Engine(0)
    yield single, made
    for e in made.values():
        e.close()
    single.close()


def _draws(plan, kind, R, philox, seed):
    """(Draws for the engine, the oracle's keyword arguments for the same draws)"""
    if philox:
        return Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=seed, iter0=11), {"seed": seed, "iter0": 11}
    d = H.draw_R(plan, seed, R)
    return Draws(n_iter=R, rng_mode=_lib.RNG_HOST, bait_seqid=d[0], bait_start=d[1], dst_tile=d[2] if kind >= 4 else None), {"draws": d}


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("philox", [False, True])
@pytest.mark.parametrize("n_excl", [0, 20])
def test_group_counts_equal_single_gpu_and_oracle(groups, kind, philox, n_excl):
    single, made = groups
    case = H.random_case(seed=300 + kind, n=900, n_seq=4, n_query=70, n_excl=n_excl, n_seg=12, stranded=bool(n_excl))
    plan, y, q, x = H.oracle_objects(case, kind)
    R = 13  # not a multiple of any group size: ragged shards
    try:
        draws, how = _draws(plan, kind, R, philox, seed=5)
        ref = O.boot_counts(plan, y, q, R, exclude=x, **how)
    except ValueError as e:
        assert "drew no segments" in str(e)
        pytest.skip("proportionLength = FALSE: a state drew no segments (fails in the reference too)")
    H.load_engine(single, case, kind)
    one = single.run_counts(draws)
    for devs, eng in made.items():
        assert eng.n_devices == len(devs)
        assert H.load_engine(eng, case, kind) == plan.B
        got = eng.run_counts(draws)
        for k in ("iter_nranges", "iter_totals", "feature_counts"):
            assert np.array_equal(got[k], one[k]), (devs, k)
            assert np.array_equal(got[k], ref[k]), (devs, k)


def test_group_fewer_iterations_than_devices(groups):
    single, made = groups
    case = H.random_case(seed=77, n=400, n_seq=3, n_query=40)
    H.load_engine(single, case, 0)
    for R in (0, 1, 2):
        d = Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=3)
        one = single.run_counts(d)
        for devs, eng in made.items():
            H.load_engine(eng, case, 0)
            got = eng.run_counts(d)
            for k in ("iter_nranges", "iter_totals", "feature_counts"):
                assert np.array_equal(got[k], one[k]), (devs, R, k)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_group_moments_equal_single_gpu(groups, kind):
    single, made = groups
    case = H.random_case(seed=410 + kind, n=1200, n_seq=4, n_query=90, n_seg=9, points=kind == 1)
    d = Draws(n_iter=37, rng_mode=_lib.RNG_PHILOX, seed=17, iter0=2)
    H.load_engine(single, case, kind)
    obs = single.count_observed()[0]
    one = single.run_moments(d, observed=obs, batch_iter=8)
    for devs, eng in made.items():
        H.load_engine(eng, case, kind)
        assert np.array_equal(eng.count_observed()[0], obs)
        got = eng.run_moments(d, observed=obs, batch_iter=5)
        for k in ("iter_nranges", "iter_totals", "sum", "sumsq", "n_ge_observed"):
            assert np.array_equal(got[k], one[k]), (devs, k)


def test_group_weighted_sums(groups):
    single, made = groups
    case = H.random_case(seed=55, n=800, n_seq=3, n_query=60)
    w = np.random.default_rng(1).random(case["y"][0].size) * 10
    d = Draws(n_iter=9, rng_mode=_lib.RNG_PHILOX, seed=4)
    H.load_engine(single, case, 0)
    single.set_weights(w)
    one = single.run_weighted(d)
    one_m = single.run_weighted_moments(d, observed=one["feature_sums"][0])
    for devs, eng in made.items():
        H.load_engine(eng, case, 0)
        eng.set_weights(w)
        got = eng.run_weighted(d)
        # the same windows and running sums on every GPU: the per-iteration values are bit-identical; only the
        # moments add the members' partial sums in a different order (relative tolerance 1e-12)
        assert np.array_equal(got["feature_sums"], one["feature_sums"]) and np.array_equal(got["iter_sums"], one["iter_sums"])
        got_m = eng.run_weighted_moments(d, observed=one["feature_sums"][0])
        assert np.allclose(got_m["sum"], one_m["sum"], rtol=1e-12, atol=0) and np.allclose(got_m["sumsq"], one_m["sumsq"], rtol=1e-12, atol=0)
        assert np.array_equal(got_m["n_ge_observed"], one_m["n_ge_observed"])


@pytest.mark.parametrize("how", ["pageable", "pinned", "pinned_unsorted_stranded"])
def test_group_scattered_upload(groups, how, monkeypatch):
    """A large y crosses PCIe once for the whole group: every member uploads one slice and hands it to the others with peer copies
    (group_scatter_y; here forced for a small y with NRB200_SCATTER_MIN).  Same index, same counts as one GPU -- from a pageable y
    (goes through the group's pinned copy; needs columns of 256 KB or more), from pinned arrays, and for an unsorted, stranded y
    (sorted on every GPU after the exchange)."""
    import torch
    single, made = groups
    n = 70_000 if how == "pageable" else 3_001   # 3001: the slices' ragged ends
    case = H.random_case(seed=91, n=n, n_seq=5, max_len=400_000, n_query=300, L_b=20_000, stranded=how.endswith("stranded"),
                         sort=not how.endswith("unsorted_stranded"))
    if how != "pageable":
        case["y"] = tuple(None if a is None else torch.from_numpy(a).pin_memory().numpy() for a in case["y"])
    d = Draws(n_iter=7, rng_mode=_lib.RNG_PHILOX, seed=12)
    kind = 1 if how == "pinned_unsorted_stranded" else 0  # (the across-chromosome sampler insists on a sorted y)
    H.load_engine(single, case, kind)
    one = single.run_counts(d)
    for devs, eng in made.items():
        monkeypatch.setenv("NRB200_SCATTER_MIN", "0")  # (0: never)
        H.load_engine(eng, case, kind)
        plain = eng.run_counts(d)
        monkeypatch.setenv("NRB200_SCATTER_MIN", "1")
        H.load_engine(eng, case, kind)
        got = eng.run_counts(d)
        for k in ("iter_nranges", "iter_totals", "feature_counts"):
            assert np.array_equal(got[k], one[k]), (devs, k)
            assert np.array_equal(plain[k], one[k]), (devs, k)
        assert eng.ranges_info() == single.ranges_info()


@pytest.mark.parametrize("kind,n_excl", [(0, 0), (2, 40)])
def test_group_window_lists_built_by_all_members(groups, kind, n_excl):
    """From 4096 features on, the leader's two per-sequence passes of the window-list building run on the driver threads of
    all members (group_prepare / seq_runner): same lists, same counts as one GPU."""
    single, made = groups
    case = H.random_case(seed=640 + kind, n=30_000, n_seq=7, max_len=300_000, n_query=5_000, n_excl=n_excl, n_seg=21, L_b=9_000)
    d = Draws(n_iter=5, rng_mode=_lib.RNG_PHILOX, seed=21)
    H.load_engine(single, case, kind)
    one = single.run_counts(d)
    for devs, eng in made.items():
        H.load_engine(eng, case, kind)
        got = eng.run_counts(d)
        for k in ("iter_nranges", "iter_totals", "feature_counts"):
            assert np.array_equal(got[k], one[k]), (devs, k)


def test_group_wire_format_policy(groups):
    """A pinned [R x G] int32 destination: up to 2 GPUs the matrix crosses PCIe as uint16 and the host widens it (faster while the
    host has bandwidth to spare); from 3 members on every GPU copies its int32 rows by DMA (run.cuh, tune_wire16).  Same matrix."""
    import torch
    single, made = groups
    case = H.random_case(seed=19, n=2500, n_seq=3, n_query=130)
    R, G = 390, 130
    d = Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=6)
    H.load_engine(single, case, 0)
    one = single.run_counts(d)
    for devs, eng in made.items():
        keep = torch.empty((R, G), dtype=torch.int32, pin_memory=True)
        H.load_engine(eng, case, 0)
        got = eng.run_counts(d, out={"feature_counts": keep.numpy()})
        assert np.array_equal(got["feature_counts"], one["feature_counts"]), devs
        assert got["stats"]["d2h_bytes"] == 2 * R * 8 + R * G * (2 if len(devs) <= 2 else 4), devs


def test_group_errors_and_unsupported(groups):
    single, made = groups
    from nullranges_b200 import EngineError
    eng = next(iter(made.values()))
    case = H.random_case(seed=5, n=100, n_seq=2, n_query=10)
    H.load_engine(eng, case, 0)
    with pytest.raises(EngineError, match="per GPU"):
        eng.run_counts_resident(Draws(n_iter=2, rng_mode=_lib.RNG_PHILOX, seed=1))
    bad = np.array(case["y"][1])
    bad[3] = 0  # start < 1: every member rejects it, the group reports one of them
    with pytest.raises(EngineError, match="GPU"):
        eng.set_ranges(case["seqlen"], case["y"][0], bad, case["y"][2])
    with pytest.raises(EngineError, match="twice|device"):
        os.environ.pop("NRB200_ALLOW_DUP_DEVICES")
        try:
            Engine([0, 0])
        finally:
            os.environ["NRB200_ALLOW_DUP_DEVICES"] = "1"


def test_front_end_gpus_argument(groups):
    """bootCounts(..., gpus=) / bootMoments(..., gpus=): the product API reaches the group."""
    rng = np.random.default_rng(9)
    sl = {"chr1": 40000, "chr2": 30000}
    names = np.array(["chr1", "chr2"])[rng.integers(0, 2, 1500)]
    st = rng.integers(1, 29000, 1500)
    y = GRanges(names, st, width=rng.integers(1, 200, 1500), seqlengths=sl).sort()
    x = GRanges(np.array(["chr1", "chr2"])[rng.integers(0, 2, 80)], rng.integers(1, 28000, 80), width=rng.integers(50, 900, 80), seqlengths=sl)
    a = bootCounts(y, x, 2000, R=21, rng="philox", seed=8)
    devs = [0, 0, 0] if _n_gpus() < 2 else list(range(min(_n_gpus(), 4)))
    b = bootCounts(y, x, 2000, R=21, rng="philox", seed=8, gpus=devs)
    assert np.array_equal(a.feature_counts, b.feature_counts) and np.array_equal(a.iter_totals, b.iter_totals)
    assert np.array_equal(a.iter_nranges, b.iter_nranges)
    c = bootCounts(y, x, 2000, R=5, rng="R", seed=1)
    d = bootCounts(y, x, 2000, R=5, rng="R", seed=1, gpus=devs)
    assert np.array_equal(c.feature_counts, d.feature_counts)
    m1 = bootMoments(y, x, 2000, R=40, seed=3)
    m2 = bootMoments(y, x, 2000, R=40, seed=3, gpus=devs)
    assert np.array_equal(m1.sum, m2.sum) and np.array_equal(m1.sumsq, m2.sumsq) and np.array_equal(m1.p_ge, m2.p_ge)


@pytest.mark.skipif("_n_gpus() < 2")
def test_group_full_size_cfg2_across_real_gpus():
    """1M ranges x R = 1000 on all GPUs of the box == one GPU, whole [R x G] matrix (BASELINE configs[1])."""
    from nullranges_b200 import synth
    y, x = synth.atac_peaks(1_000_000, seed=3), synth.genes(20_000, seed=2)
    a = bootCounts(y, x, 500_000, R=1000, rng="philox", seed=2024)
    b = bootCounts(y, x, 500_000, R=1000, rng="philox", seed=2024, gpus=_n_gpus())
    assert np.array_equal(a.feature_counts, b.feature_counts)
    assert np.array_equal(a.iter_totals, b.iter_totals) and np.array_equal(a.iter_nranges, b.iter_nranges)


def test_front_end_boot_ranges_on_a_group(groups):
    """bootRanges(..., gpus=): the materialised table is produced by the group's first GPU -- same rows as one GPU,
    including blockStart for a device-drawn permutation (nrb200_fetch_dst_tiles through the group handle)."""
    from nullranges_b200 import bootRanges
    rng = np.random.default_rng(4)
    sl = {"chr1": 30000, "chr2": 20000}
    y = GRanges(np.array(["chr1", "chr2"])[rng.integers(0, 2, 900)], rng.integers(1, 19000, 900), width=rng.integers(1, 150, 900), seqlengths=sl).sort()
    devs = [0, 0] if _n_gpus() < 2 else [0, 1]
    for kw in (dict(type="bootstrap"), dict(type="permute", withinChrom=True, storeBlockStart=True)):
        a = bootRanges(y, 2000, R=6, rng="philox", seed=11, **kw)
        b = bootRanges(y, 2000, R=6, rng="philox", seed=11, gpus=devs, **kw)
        for k in ("seqid", "start", "end"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), k
        assert np.array_equal(a.mcols["iter"], b.mcols["iter"])
        if "storeBlockStart" in kw:
            assert np.array_equal(a.mcols["blockStart"], b.mcols["blockStart"])
            assert np.all((a.start >= a.mcols["blockStart"]) | (a.start == 1))  # a row lies on the tile its block landed on
