"""CPU: host-side logic of the product (no GPU, no compute calls) and the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from nullranges_b200 import GRanges, RRng, _lib, samplers, set_seed
from nullranges_b200 import bootranges as BR
from oracle import oracle as O
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "nrb200.h")).read()
    declared = set(re.findall(r"\b(nrb200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = ctypes.CDLL(_lib.so_path() if os.path.exists(_lib.so_path()) else _lib.load()._name)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/nrb200.h but not exported"
    assert declared == {n for n, _, _ in _lib.SYMBOLS}, "ctypes table and header disagree"
    # nothing else leaks out of the library (built with -fvisibility=hidden)
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.so_path()], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert {e for e in exported if not e.startswith("nrb200_")} <= {"_init", "_fini"}


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from nullranges_b200 import Engine, EngineError
    with pytest.raises(EngineError) as e:
        Engine(0)
    assert e.value.code == _lib.ERR_CUDA and "no CPU path" in str(e.value)
    y = GRanges(["chr1"] * 2, [1, 20], width=[5, 5], seqlengths={"chr1": 100})
    with pytest.raises(EngineError):
        BR.bootRanges(y, 10)  # the product never falls back to the CPU


def test_host_rng_matches_oracle_rng():
    a, b = RRng(99), O.RRng(99)
    assert np.array_equal(a.runif(50), b.runif(50))
    assert np.array_equal(a.sample(17), b.sample(17))
    p = np.arange(1, 301, dtype=np.float64)
    assert np.array_equal(a.sample(300, 500, True, p), b.sample(300, 500, True, p))   # Walker alias branch
    assert np.array_equal(a.sample(40, 500, True, p[:40]), b.sample(40, 500, True, p[:40]))  # inversion branch
    assert np.array_equal(a.sample(1000, 64, True), b.sample(1000, 64, True))
    assert np.array_equal(a.runif(6, [1, 2, 3], [1, 5, 3]), b.runif(6, [1.0, 2, 3], [1.0, 5, 3]))


@pytest.mark.parametrize("kind", range(6))
def test_python_samplers_follow_the_reference_draw_order(kind):
    """nullranges_b200/samplers.py (what the R wrapper does with R's own RNG) against the oracle's draws."""
    case = H.random_case(seed=500 + kind, n=10, n_seq=4, n_query=0, n_seg=12, n_states=2)
    plan = O.Plan(kind, case["seqlen"], case["L_b"], case["seg"] if kind in (2, 3) else None)
    L_s = case["seqlen"].astype(np.float64)
    seg = case["seg"]
    for seed in (1, 2, 3):
        try:
            want = plan.draw_R(O.RRng(seed), 3)
        except ValueError:
            continue
        rng = RRng(seed)
        for r in range(3):
            if kind == 0:
                d = samplers.unseg_bootstrap_across_chrom(rng, L_s, case["L_b"])
            elif kind == 1:
                d = samplers.unseg_bootstrap_within_chrom(rng, L_s, case["L_b"])
            elif kind == 2:
                d = samplers.seg_bootstrap_prop(rng, seg["chrom"], seg["start"], seg["end"], seg["state"], case["L_b"], L_s)
            elif kind == 3:
                d = samplers.seg_bootstrap_noprop(rng, seg["chrom"], seg["start"], seg["end"], seg["state"], case["L_b"])
            elif kind == 4:
                d = samplers.unseg_permute_within_chrom(rng, L_s, case["L_b"], plan.tile_start)
            else:
                d = samplers.unseg_permute_across_chrom(rng, plan.tile_chrom, plan.tile_start)
            assert np.array_equal(d[0], want[0][r]) and np.array_equal(d[1], want[1][r])
            if kind >= 4:
                assert np.array_equal(d[2], want[2][r])


def test_granges_sort_semantics():
    g = GRanges(["b", "a", "a", "a"], [5, 9, 3, 3], width=[2, 1, 4, 4], strand=["*", "-", "*", "+"], seqlevels=["a", "b"],
                seqlengths=[50, 50], score=[1, 2, 3, 4])
    s = g.sort()
    # seqlevel order, then strand + < - < *, then start
    assert s.seqnames.tolist() == ["a", "a", "a", "b"] and s.mcols["score"].tolist() == [4, 2, 3, 1]
    assert s.is_sorted() and not g.is_sorted()
    assert GRanges(["a"], [1], end=[3], seqlengths={"a": 9}).width.tolist() == [3]
    with pytest.raises(ValueError):
        GRanges(["zz"], [1], width=[1], seqlevels=["a"])


def test_front_end_argument_checks():
    y = GRanges(["chr1"] * 3, [1, 20, 40], width=[5, 5, 5], seqlengths={"chr1": 100})
    with pytest.raises(ValueError, match="is.na"):
        BR._setup(None, GRanges(["chr1"], [1], width=[2], seqlevels=["chr1"]), 10, None, None, "drop", True, "bootstrap", False)
    with pytest.raises(ValueError, match="should be one of"):
        BR._setup(None, y, 10, None, None, "keep", True, "bootstrap", False)
    with pytest.raises(ValueError, match="should be one of"):
        BR._setup(None, y, 10, None, None, "drop", True, "shuffle", False)
    with pytest.raises(ValueError, match="state"):
        BR._seg_arrays(GRanges(["chr1"], [1], end=[100], seqlengths={"chr1": 100}), y)
    assert BR._plan_kind(None, True, "bootstrap", False) == _lib.UNSEG_ACROSS
    assert BR._plan_kind(None, True, "permute", True) == _lib.PERMUTE_WITHIN
    assert BR._plan_kind(y, False, "permute", True) == _lib.SEG_NOPROP  # seg given: type / withinChrom ignored
    assert BR._match_arg(("drop", "trim"), ("drop", "trim"), "excludeOption") == "drop"


def test_set_seed_stream_is_shared_across_calls():
    set_seed(5)
    from nullranges_b200.rrng import global_rng
    a = global_rng().runif(2)
    b = global_rng().runif(2)
    assert np.array_equal(np.concatenate([a, b]), O.RRng(5).runif(4))


def test_r_shim_type_checks_against_the_header():
    """r-package/src/r_shim.c cannot be built here (no R); it is at least compiled (-fsyntax-only) against
    include/nrb200.h with declaration-only stand-ins of the R C API, so signature drift is caught."""
    import subprocess
    res = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type",
                          "-I", os.path.join(ROOT, "tests", "r_stubs"), "-I", os.path.join(ROOT, "include"),
                          os.path.join(ROOT, "r-package", "src", "r_shim.c")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    # every .Call used by the R wrapper is registered by the shim with the right number of arguments
    wrapper = open(os.path.join(ROOT, "r-package", "R", "bootRanges_b200.R")).read()
    shim = open(os.path.join(ROOT, "r-package", "src", "r_shim.c")).read()
    registered = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(nrb200_R_\w+)", \(DL_FUNC\)&\w+, (\d+)\}', shim)}
    for m in re.finditer(r'\.Call\("(nrb200_R_\w+)"((?:[^()]|\([^()]*(?:\([^()]*\)[^()]*)*\))*)\)', wrapper):
        name, args = m.group(1), m.group(2)
        depth, n = 0, 0
        for ch in args:
            depth += ch in "([" 
            depth -= ch in ")]"
            n += (ch == "," and depth == 0)
        assert name in registered, name
        assert registered[name] == n, f"{name}: wrapper passes {n} arguments, shim registers {registered[name]}"


def test_expand_runs_matches_numpy_repeat():
    """nrb200_expand_runs(): the dense iter column, rep(seq_len(R), lens) of R/bootRanges.R:122 (host utility, no GPU)."""
    import ctypes as C
    from nullranges_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for n_runs, scale in ((1, 5), (7, 3), (1000, 3000), (50, 100000), (3, 0)):
        lens = rng.integers(0, scale + 1, n_runs)
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64) + 17
        out = np.empty(int(offs[-1] - offs[0]), np.int32)
        lib.nrb200_expand_runs(_lib.ptr(offs, C.c_int64), n_runs, 1, _lib.ptr(out, C.c_int32))
        np.testing.assert_array_equal(out, np.repeat(np.arange(1, n_runs + 1, dtype=np.int32), lens))


def test_side_arrays_maps_seqlevels_and_drops_foreign_ranges():
    """Front-end helper: x is expressed in y's seqlevels; ranges on sequences y lacks are dropped (they can never be hit);
    identical seqlevels pass the arrays through untouched."""
    from nullranges_b200 import GRanges
    from nullranges_b200 import bootranges as BR
    y = GRanges(["chr1", "chr2"], [1, 5], width=[3, 3], seqlengths={"chr1": 100, "chr2": 50})
    same = GRanges(["chr2", "chr1"], [7, 9], width=[2, 2], seqlengths={"chr1": 100, "chr2": 50})
    sid, st, en, strand, keep = BR._side_arrays(same, y, "x")
    assert sid is same.seqid and st is same.start and en is same.end and strand is None and keep is None
    other = GRanges(["chr2", "chrX", "chr1"], [7, 1, 9], width=[2, 2, 2], strand=["+", "*", "-"],
                    seqlengths={"chr2": 50, "chrX": 10, "chr1": 100})
    sid, st, en, strand, keep = BR._side_arrays(other, y, "x")
    assert sid.tolist() == [1, 0] and st.tolist() == [7, 9] and en.tolist() == [8, 10]
    assert strand.tolist() == [1, 2] and keep.tolist() == [0, 2]
    reordered = GRanges(["chr2", "chr1"], [7, 9], width=[2, 2], seqlengths={"chr2": 50, "chr1": 100})
    sid, st, en, strand, keep = BR._side_arrays(reordered, y, "x")
    assert sid.tolist() == [1, 0] and keep is None and st is reordered.start


def test_zero_width_rule_is_one_constant_on_both_sides():
    """The engine and the oracle are built with the same zero-width rule (include/nrb200.h: nrb200_zero_width_rule);
    flipping it is one define on each side (NRB200_ZERO_WIDTH_RULE=1 at build time sets both)."""
    from nullranges_b200 import _lib
    from oracle import oracle as O
    lib = _lib.load()
    assert lib.nrb200_zero_width_rule() == O.zero_width_rule()
    assert lib.nrb200_zero_width_rule() == int(os.environ.get("NRB200_ZERO_WIDTH_RULE", "0"))
