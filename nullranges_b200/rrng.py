"""R's random number stream for the Python mirror (set.seed / runif / sample / round).

Thin wrapper over the host C++ in libnrb200 (csrc/rrng.cpp).  The R package does not use
this: there R itself draws, at the same call sites as the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class RRng:
    def __init__(self, seed: int):
        self._lib = _lib.load()
        self._h = self._lib.nrb200_rrng_create(C.c_uint32(int(seed) & 0xFFFFFFFF))

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.nrb200_rrng_destroy(self._h)
            self._h = None

    def unif_rand(self) -> float:
        return self._lib.nrb200_rrng_unif_rand(self._h)

    def runif(self, n: int, min=0.0, max=1.0) -> np.ndarray:
        lo = np.atleast_1d(np.asarray(min, dtype=np.float64))
        hi = np.atleast_1d(np.asarray(max, dtype=np.float64))
        out = np.empty(int(n), dtype=np.float64)
        rc = self._lib.nrb200_rrng_runif(self._h, int(n), _lib.ptr(lo, C.c_double), lo.size, _lib.ptr(hi, C.c_double), hi.size,
                                         _lib.ptr(out, C.c_double))
        if rc:
            raise ValueError("runif: invalid arguments")
        return out

    def sample(self, n: int, size: int | None = None, replace: bool = False, prob=None) -> np.ndarray:
        """sample.int(n, size, replace, prob); 1-based like R."""
        size = int(n if size is None else size)
        out = np.empty(size, dtype=np.int32)
        p = None if prob is None else np.ascontiguousarray(prob, dtype=np.float64)
        if p is not None and p.size != n:
            raise ValueError("incorrect number of probabilities")
        rc = self._lib.nrb200_rrng_sample(self._h, int(n), size, int(bool(replace)), _lib.ptr(p, C.c_double), _lib.ptr(out, C.c_int32))
        if rc:
            raise ValueError("sample: invalid arguments")
        return out


def r_round(x):
    """round(x): half to even, like R >= 4.0 at digits = 0."""
    return np.rint(x)


_global: RRng | None = None


def set_seed(seed: int) -> None:
    """set.seed(seed) for the stream bootRanges(rng="R") consumes."""
    global _global
    _global = RRng(seed)


def global_rng() -> RRng:
    global _global
    if _global is None:  # R seeds from time/pid when unseeded; any seed will do here
        import os
        import time
        _global = RRng((int(time.time()) ^ os.getpid()) & 0xFFFFFFFF)
    return _global
