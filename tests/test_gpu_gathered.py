"""Gathered result of a device group (nrb200_run_counts_gathered): every GPU ends up with the whole [R x G] matrix and
[R] vectors in its own memory, written there by the count kernels of all members over peer-mapped pointers (SURVEY
8e's all-gather of the [R/W x G] tiles, fused into the kernel).  Needs two GPUs with peer access."""
import numpy as np
import pytest

from nullranges_b200 import Draws, Engine, _lib
from nullranges_b200.sharding import DeviceView
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _fetch(eng, member, R, G):
    import torch
    dev, nr_p, tot_p, cnt_p = eng.gathered_pointers(member)
    with torch.cuda.device(dev):
        nr = torch.as_tensor(DeviceView(nr_p, (R,), "<i8"), device=f"cuda:{dev}").cpu().numpy()
        tot = torch.as_tensor(DeviceView(tot_p, (R,), "<i8"), device=f"cuda:{dev}").cpu().numpy()
        cnt = torch.as_tensor(DeviceView(cnt_p, (R, G), "<i4"), device=f"cuda:{dev}").cpu().numpy()
    return nr, tot, cnt


@pytest.mark.skipif("_n_gpus() < 2")
@pytest.mark.parametrize("kind", [0, 2, 4])
def test_gathered_equals_host_result_on_every_gpu(kind):
    n = min(_n_gpus(), 8)
    case = H.random_case(seed=900 + kind, n=3000, n_seq=4, n_query=150, n_excl=10 if kind == 2 else 0, n_seg=8)
    single, group = Engine(0), Engine(list(range(n)))
    try:
        H.load_engine(single, case, kind)
        H.load_engine(group, case, kind)
        R = 37
        d = Draws(n_iter=R, rng_mode=_lib.RNG_PHILOX, seed=5, iter0=3)
        want = single.run_counts(d)
        group.run_counts_gathered(d)
        for k in range(n):
            nr, tot, cnt = _fetch(group, k, R, 150)
            assert np.array_equal(cnt, want["feature_counts"]), k
            assert np.array_equal(tot, want["iter_totals"]) and np.array_equal(nr, want["iter_nranges"]), k
    finally:
        single.close()
        group.close()


def test_gathered_on_one_gpu_is_the_resident_result():
    case = H.random_case(seed=77, n=800, n_seq=3, n_query=40)
    eng = Engine(0)
    try:
        H.load_engine(eng, case, 0)
        d = Draws(n_iter=9, rng_mode=_lib.RNG_PHILOX, seed=2)
        want = eng.run_counts(d)
        eng.run_counts_gathered(d)
        nr, tot, cnt = _fetch(eng, 0, 9, 40)
        assert np.array_equal(cnt, want["feature_counts"]) and np.array_equal(tot, want["iter_totals"])
    finally:
        eng.close()
