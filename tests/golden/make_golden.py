"""Writes the fixtures under tests/golden/.  Run from the repo root: python tests/golden/make_golden.py

Provenance (the reference is pure R and R is not installed in the build image, so nothing here
was produced by running nullranges itself):
  r_known_answers.json   outputs of real R (>= 3.6, default RNGkind) for set.seed/runif/sample,
                         the canonical values found in R documentation and countless transcripts;
                         Philox4x32-10 known answers from Random123's kat_vectors.  Typed in by hand,
                         NOT produced by the oracle: they pin the oracle's RNG.
  worked_examples.json   the reference algorithm run BY HAND on the reference's own fixtures
                         (roxygen example R/bootRanges.R:65-69; tests/testthat/test_boot.R:7-16) with
                         the pinned RNG -- SURVEY.md section 8c.  Derived, not R-verified.
  regression_*.npz       oracle outputs on seeded random cases: regression vectors that keep the
                         oracle (and through it the CUDA path) from drifting.  Oracle-generated.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

R_KNOWN = {
    "runif": [
        {"seed": 1, "n": 5, "value": [0.2655087, 0.3721239, 0.5728534, 0.9082078, 0.2016819]},
        {"seed": 42, "n": 3, "value": [0.9148060, 0.9370754, 0.2861395]},
        {"seed": 123, "n": 3, "value": [0.2875775, 0.7883051, 0.4089769]},
        {"seed": 2, "n": 3, "value": [0.1848823, 0.7023740, 0.5733263]},
        {"seed": 10, "n": 3, "value": [0.5074782, 0.3067685, 0.4269077]},
    ],
    "sample": [
        {"seed": 1, "n": 10, "size": 10, "value": [9, 4, 7, 1, 2, 5, 3, 10, 6, 8]},
        {"seed": 42, "n": 10, "size": 10, "value": [1, 5, 10, 8, 2, 4, 6, 9, 7, 3]},
        {"seed": 123, "n": 5, "size": 5, "value": [3, 2, 5, 4, 1]},
        {"seed": 123, "n": 10, "size": 10, "value": [3, 10, 2, 8, 6, 9, 1, 7, 5, 4]},
        {"seed": 1, "n": 100, "size": 5, "value": [68, 39, 1, 34, 87]},
        {"seed": 1, "n": 5, "size": 5, "replace": True, "value": [1, 4, 1, 2, 5]},   # sample(1:5, 5, replace = TRUE)
    ],
    "philox4x32_10": [
        {"ctr": [0, 0, 0, 0], "key": [0, 0], "value": [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]},
        {"ctr": [0xffffffff] * 4, "key": [0xffffffff] * 2, "value": [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]},
        {"ctr": [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], "key": [0xa4093822, 0x299f31d0],
         "value": [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]},
    ],
}

WORKED = {
    "roxygen_example": {
        "source": "R/bootRanges.R:65-69: set.seed(1); gr on chr1 (seqlength 50); bootRanges(gr, blockLength=10)",
        "seqlengths": [50], "block_length": 10, "seed": 1,
        "y": {"chrom": [0] * 5, "start": [1, 11, 21, 31, 41], "end": [5, 15, 25, 35, 45]},
        "bait_start": [37, 39, 27, 26, 3],
        "rows": {"start": [5, 13, 25, 36, 39, 49], "end": [9, 17, 29, 40, 43, 50], "src_1based": [5, 5, 4, 4, 1, 2],
                 "block_1based": [1, 2, 3, 4, 5, 5]},
    },
    "test_boot_fixture": {
        "source": "tests/testthat/test_boot.R:7-16, set.seed(1), type='bootstrap', withinChrom=FALSE, blockLength=100, R=5",
        "seqlengths": [300, 450, 200], "block_length": 100, "seed": 1,
        "y": {"chrom": [0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2], "start": [1, 101, 121, 201, 101, 201, 216, 231, 401, 1, 101],
              "width": [20, 5, 5, 30, 20, 5, 5, 5, 30, 80, 40]},
        "iter1_bait_chrom": [1, 1, 0, 2, 1, 2, 2, 0, 0, 1],
        "iter1_bait_start": [83, 72, 138, 39, 309, 51, 73, 199, 77, 312],
        "iter1_rows": [[0, 19, 38, 5, 1], [0, 130, 149, 5, 2], [0, 264, 293, 4, 3], [1, 1, 42, 10, 4], [1, 63, 102, 11, 4],
                       [1, 193, 222, 9, 5], [1, 151, 230, 10, 6], [1, 251, 290, 11, 6], [1, 229, 308, 10, 7],
                       [1, 329, 368, 11, 7], [1, 403, 432, 4, 8], [2, 25, 29, 2, 9], [2, 45, 49, 3, 9], [2, 190, 200, 9, 10]],
        "rows_per_iteration": [14, 12, 17, 17, 10],
    },
}


def regression():
    from oracle import oracle as O
    from tests import helpers as H
    for kind in range(6):
        case = H.random_case(seed=900 + kind, n=300, n_seq=3, n_query=40, n_excl=10 if kind % 2 else 0, stranded=kind in (1, 4),
                             n_seg=9, n_states=2)
        plan, y, q, x = H.oracle_objects(case, kind)
        R = 4
        d = H.draw_R(plan, 11, R)
        ref = O.boot_counts(plan, y, q, R, exclude=x, draws=d)
        rows = H.oracle_rows(plan, y, x, d, R)
        np.savez_compressed(os.path.join(HERE, f"regression_kind{kind}.npz"), bait_chrom=d[0], bait_start=d[1], dst_tile=d[2],
                            iter_nranges=ref["iter_nranges"], iter_totals=ref["iter_totals"], feature_counts=ref["feature_counts"],
                            rows=rows)
    # Philox production-mode draws
    case = H.random_case(seed=950, n=300, n_seq=3, n_query=40)
    plan, y, q, x = H.oracle_objects(case, 0)
    ref = O.boot_counts(plan, y, q, 5, seed=2024, iter0=7)
    d = plan.draw_philox(2024, 7, 5)
    np.savez_compressed(os.path.join(HERE, "regression_philox.npz"), bait_chrom=d[0], bait_start=d[1],
                        iter_nranges=ref["iter_nranges"], iter_totals=ref["iter_totals"], feature_counts=ref["feature_counts"])


if __name__ == "__main__":
    json.dump(R_KNOWN, open(os.path.join(HERE, "r_known_answers.json"), "w"), indent=1)
    json.dump(WORKED, open(os.path.join(HERE, "worked_examples.json"), "w"), indent=1)
    regression()
    print("golden fixtures written to", HERE)
