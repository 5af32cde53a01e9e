"""Value-level pinning against the real reference: consumes tests/golden/r_reference_golden.json, the dump
tests/golden/make_golden.R writes where R + nullranges exist (every output row, rows per iteration and
countOverlaps(x, boots[iter == r]) of the reference's own fixtures under set.seed).  While that file is absent
(the build image has no R) the R-made tests skip, and the same consumer is run against an oracle-made file of the same
schema so that the plumbing cannot rot."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import r_golden as RG

HAVE_R_FILE = os.path.exists(RG.PATH)
need_r = pytest.mark.skipif(not HAVE_R_FILE, reason="no R-made golden file: run `Rscript tests/golden/make_golden.R` on a machine with R + nullranges")


def _load(path):
    d = json.load(open(path))
    return d, {c["name"]: c for c in d["cases"]}


def _check(got, c):
    want = RG.golden_rows(c)
    assert got["rows_per_iteration"] == list(np.atleast_1d(c["rows_per_iteration"])), "rows per iteration (lengths(br), R/bootRanges.R:120)"
    assert np.array_equal(RG.canonical(np.asarray(got["rows"])), RG.canonical(want)), "rows as a set (iter, seqname, start, end, strand, source range)"
    assert np.array_equal(np.asarray(got["counts"]), np.asarray(c["counts"]).reshape(np.asarray(got["counts"]).shape)), "countOverlaps(x, boots[iter == r])"


@pytest.fixture(scope="module")
def emulated(tmp_path_factory):
    return _load(RG.emulate(str(tmp_path_factory.mktemp("golden") / "emulated.json")))


NAMES = [c["name"] for c in RG._fixtures()]


@pytest.mark.parametrize("name", NAMES)
def test_consumer_on_oracle_made_file(emulated, name):
    """Plumbing only: the oracle against a file the oracle wrote (schema, draw order, row bookkeeping)."""
    _, cases = emulated
    _check(RG.oracle_case(cases[name]), cases[name])
    assert np.array_equal(RG.oracle_case(cases[name])["rows"], RG.golden_rows(cases[name]))


def test_emulated_file_reproduces_the_hand_derived_examples(emulated):
    """The emulated file's first cases are SURVEY 8c's worked examples (tests/golden/worked_examples.json)."""
    _, cases = emulated
    w = json.load(open(os.path.join(os.path.dirname(RG.PATH), "worked_examples.json")))
    assert cases["roxygen_example"]["rows"]["start"] == w["roxygen_example"]["rows"]["start"]
    assert cases["boot_across"]["rows_per_iteration"] == w["test_boot_fixture"]["rows_per_iteration"]


@need_r
@pytest.mark.parametrize("name", NAMES)
def test_oracle_equals_real_R(name):
    d, cases = _load(RG.PATH)
    assert d["provenance"].get("R"), "not an R-made file"
    _check(RG.oracle_case(cases[name]), cases[name])


@need_r
@pytest.mark.parametrize("name", NAMES)
def test_oracle_row_order_equals_real_R(name):
    """The contract order (iteration, block, y index) against the order the reference really emits
    (INTEGRATION.md R-check item 2)."""
    _, cases = _load(RG.PATH)
    assert np.array_equal(RG.oracle_case(cases[name])["rows"], RG.golden_rows(cases[name]))


@need_r
def test_zero_width_rule_matches_real_R():
    """INTEGRATION.md R-check item 1.  On failure: rebuild engine and oracle with NRB200_ZERO_WIDTH_RULE set to the value R reports."""
    d, _ = _load(RG.PATH)
    assert int(d["facts"]["zero_width_hits_range_past_end"]) == O.zero_width_rule()
    assert int(d["facts"]["zero_width_hits_range_inside"]) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_product_on_oracle_made_file(emulated, name):
    """The product's front end (bootRanges / bootCounts, rng = 'R') against the same schema: rows, lengths and counts."""
    _, cases = emulated
    _check(RG.product_case(cases[name]), cases[name])


@pytest.mark.gpu
@need_r
@pytest.mark.parametrize("name", NAMES)
def test_product_equals_real_R(name):
    _, cases = _load(RG.PATH)
    got = RG.product_case(cases[name])
    _check(got, cases[name])
    assert np.array_equal(got["rows"], RG.golden_rows(cases[name])), "row order as the reference emits it"
