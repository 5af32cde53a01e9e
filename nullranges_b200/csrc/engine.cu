// libnrb200: host side of the C ABI (include/nrb200.h) and the kernel launches.
//
// Memory layout in HBM (all int32 unless noted, canonical = sorted by (seqname, start, end)):
//   y        : start[N] end[N] pmax[N] (+ src[N] when the input was unsorted, strand[N] int8 when stranded),
//              seq_off[C+1], end_sorted[N] (ends ascending within each sequence; aliases start[] when all
//              ranges are equally wide), packed direct-address rank tables (start, sorted end; running-max
//              end on demand), ~4N entries each up to 8M; with several classes of y (strand, narrower than
//              minoverlap) a second, class-major copy of start / end / pmax with its own tables
//   query    : host mirror start[G] end[G] src[G] (+strand); rank path: signed window records (32 B) per
//              (feature, tile[, exclusion window], class) + pair_off[G+1]; streaming path: device copy +
//              per-tile window lo/hi [B]
//   exclude  : same shape as the query's streaming form (or gaps(exclude) for "trim") + per-tile windows
//   plan     : tile_seq[B] tile_start[B] bait_width[B], sampler tables (int64, a few KB)
//   draws    : validation mode only, bait_seq/bait_start(/dst_tile) [R x B]
//   results  : iter_nranges[R] iter_totals[R] (uint64), counts[R x G]
//
// One translation unit, in parts:
//   handle.h   the handle, device buffers, upload helpers
//   index.cuh  loading y / query / exclude, canonical order, rank tables, weight running sums
//   lists.cuh  per-tile slices, gaps(exclude), signed window lists of the rank path
//   run.cuh    dispatch, block draws, launches, core of the counting entry points
//   group.cuh  device groups: one handle sharding the iterations over several GPUs
//   api.cuh    extern "C" entry points (include/nrb200.h)
#include "handle.h"
#include "index.cuh"
#include "lists.cuh"
#include "run.cuh"
#include "group.cuh"
#include "api.cuh"
