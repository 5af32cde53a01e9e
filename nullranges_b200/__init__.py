"""nullranges_b200: B200-native engine for nullranges' bootRanges() block bootstrap.

Public surface mirrors the reference's: ``bootRanges(y, blockLength, R, seg, exclude,
excludeOption, proportionLength, type, withinChrom, storeBlockLength)`` returning a
``BootRanges`` with an ``iter`` column, plus the fused ``bootCounts()``.
The compute path is libnrb200.so (hand-written sm_100a CUDA behind include/nrb200.h);
there is no CPU fallback.
"""
from .bootranges import BootCounts, BootMoments, bootCounts, bootMoments, bootRanges
from .engine import Draws, Engine, EngineError, get_engine
from .granges import BootRanges, GRanges
from .rrng import RRng, set_seed
from .segment import countOverlaps, oneRegionSegment, segmentTileCounts, tileGenome

__all__ = ["GRanges", "BootRanges", "bootRanges", "bootCounts", "BootCounts", "bootMoments", "BootMoments", "set_seed", "RRng",
           "Engine", "EngineError", "Draws", "get_engine", "countOverlaps", "segmentTileCounts", "tileGenome", "oneRegionSegment"]
