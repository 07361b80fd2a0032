// K2: tile binning.  Replaces fill_tiledepth_aabb + cub::DeviceRadixSort + fill_start/end
// (aabb_culling.h:15-41, 70-103, 192-260).
//
// The reference sorts n_dub 64-bit keys (tile << 32 | depth bits) with 8 CUB passes.  Here the same
// LSD radix order is obtained with far less traffic by commuting the duplication past the low
// digits: the four depth-byte passes run on the N Gaussians (key = raw FP32 depth bits, payload =
// Gaussian id), duplicates are then emitted in depth order with a deterministic prefix scan (no
// global atomic counter), and only the tile digits (ceil(log2 n_tiles) bits -> 1-2 passes) are
// sorted over the n_dub duplicates.  Every pass is stable, so the final order is exactly the
// stable ascending sort of the reference's int64 keys with ties in ascending Gaussian id.
//
// One radix pass = per-block digit histogram -> digit-major exclusive scan -> stable scatter staged
// through shared memory so global writes are runs of consecutive addresses.  HBM-bound.
#include <stdlib.h>

#include "common.cuh"

namespace gs3d {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_IPT = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_IPT;  // 4096 items per block
constexpr int SORT_WARPS = SORT_THREADS / 32;

// ---------------------------------------------------------------- radix pass

template <int IPT = SORT_IPT>
__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const uint32_t *__restrict__ keys, uint32_t n, int shift, uint32_t mask,
                  uint32_t nblocks, uint32_t *__restrict__ table /*[RADIX][nblocks]*/,
                  uint32_t *__restrict__ totals /*[RADIX]*/, const uint32_t *__restrict__ n_dev) {
  constexpr int TILE_ITEMS = SORT_THREADS * IPT;  // must match the scatter kernel's tile
  __shared__ uint32_t h[RADIX];
  if (n_dev) n = min(n, *n_dev);  // capacity-sized launch: the real item count lives on the device
  h[threadIdx.x] = 0;
  __syncthreads();
  size_t base = (size_t)blockIdx.x * TILE_ITEMS;
  if (base + TILE_ITEMS <= n && (reinterpret_cast<uintptr_t>(keys) & 15) == 0) {
    // full tile: 16-byte loads (a histogram does not care which thread counts which key)
    const uint4 *k4 = reinterpret_cast<const uint4 *>(keys + base);
#pragma unroll
    for (int i = 0; i < IPT / 4; ++i) {
      const uint4 v = k4[i * SORT_THREADS + threadIdx.x];
      atomicAdd(&h[(v.x >> shift) & mask], 1u);
      atomicAdd(&h[(v.y >> shift) & mask], 1u);
      atomicAdd(&h[(v.z >> shift) & mask], 1u);
      atomicAdd(&h[(v.w >> shift) & mask], 1u);
    }
  } else {
#pragma unroll 4
    for (int i = 0; i < IPT; ++i) {
      size_t e = base + (size_t)i * SORT_THREADS + threadIdx.x;
      if (e < n) atomicAdd(&h[(keys[e] >> shift) & mask], 1u);
    }
  }
  __syncthreads();
  uint32_t c = h[threadIdx.x];
  table[(size_t)threadIdx.x * nblocks + blockIdx.x] = c;
  if (c) atomicAdd(&totals[threadIdx.x], c);
}

// One block per digit: base = sum of totals of lower digits, then exclusive scan along the blocks.  Every thread
// owns a run of consecutive blocks (two sequential sweeps around ONE block-wide scan: the table rows are a few
// thousand entries, a chunked scan with four barriers per 256 entries took 7-13 us per pass).
__global__ void __launch_bounds__(256)
radix_scan_kernel(uint32_t nblocks, uint32_t *__restrict__ table, const uint32_t *__restrict__ totals) {
  __shared__ uint32_t red[8];
  __shared__ uint32_t wsum[8];
  const int d = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t v = (threadIdx.x < d) ? totals[threadIdx.x] : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) red[warp] = v;
  uint32_t *row = table + (size_t)d * nblocks;
  const uint32_t per = (nblocks + 255u) / 256u;
  const uint32_t b0 = min(nblocks, threadIdx.x * per), b1 = min(nblocks, b0 + per);
  uint32_t s = 0;
  for (uint32_t b = b0; b < b1; ++b) s += row[b];
  uint32_t incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint32_t run = incl - s;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    run += red[w];                 // digits below this one
    if (w < warp) run += wsum[w];  // blocks of the warps before this one
  }
  for (uint32_t b = b0; b < b1; ++b) {
    const uint32_t x = row[b];
    row[b] = run;
    run += x;
  }
}

template <bool BALLOT_RANK>
__global__ void __launch_bounds__(SORT_THREADS, 3)
radix_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t n,
                     int shift, uint32_t mask, uint32_t nblocks,
                     const uint32_t *__restrict__ table, const uint32_t *__restrict__ n_dev) {
  __shared__ uint32_t s_keys[SORT_TILE];
  if (n_dev) n = min(n, *n_dev);
  if ((size_t)blockIdx.x * SORT_TILE >= n) return;  // (whole block: uniform)
  __shared__ uint32_t s_vals[SORT_TILE];
  __shared__ uint32_t warp_hist[SORT_WARPS][RADIX];  // per-warp running digit counts
  __shared__ uint32_t digit_base[RADIX];             // exclusive scan of block digit totals
  __shared__ uint32_t out_base[RADIX];               // global base - local base per digit
  __shared__ uint32_t warp_sums[SORT_WARPS];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&warp_hist[0][0])[i] = 0;
  __syncthreads();

  // warp-striped tile: warp w owns items [w*512, (w+1)*512), item (i, lane) = w*512 + i*32 + lane
  const size_t tile_base = (size_t)blockIdx.x * SORT_TILE;
  const size_t warp_base = tile_base + (size_t)warp * (32 * SORT_IPT);
  uint32_t k[SORT_IPT];
  uint16_t rank[SORT_IPT];
#pragma unroll
  for (int i = 0; i < SORT_IPT; ++i) {
    size_t e = warp_base + (size_t)i * 32 + lane;
    k[i] = (e < n) ? keys_in[e] : 0xffffffffu;
  }
  // Stable rank inside the warp chunk, in (i, lane) order.  All 16 MATCH.ANY are issued first
  // (independent, pipelined); the serial part is only the per-digit running counter update.
  uint32_t peers[SORT_IPT];
#pragma unroll
  for (int i = 0; i < SORT_IPT; ++i) {
    const size_t e = warp_base + (size_t)i * 32 + lane;
    const bool valid = e < n;
    const uint32_t active = __ballot_sync(0xffffffffu, valid);
    const uint32_t d = (k[i] >> shift) & mask;
    if (BALLOT_RANK) {
      // lanes with the same digit via one ballot per digit bit: MATCH.ANY serialises over the
      // distinct values in the warp (~28 for random bytes), eight VOTEs do not
      uint32_t m = active;
#pragma unroll
      for (int b = 0; b < RADIX_BITS; ++b) {
        const uint32_t vote = __ballot_sync(0xffffffffu, (d >> b) & 1u);
        m &= ((d >> b) & 1u) ? vote : ~vote;
      }
      peers[i] = valid ? m : 0u;
    } else {
      peers[i] = valid ? __match_any_sync(active, d) : 0u;
    }
  }
#pragma unroll
  for (int i = 0; i < SORT_IPT; ++i) {
    uint32_t r = 0;
    const uint32_t d = (k[i] >> shift) & mask;
    uint32_t before = 0;
    if (peers[i]) {
      before = warp_hist[warp][d];
      r = before + __popc(peers[i] & lt_mask);
    }
    __syncwarp();
    if (peers[i] && (peers[i] & lt_mask) == 0) warp_hist[warp][d] = before + __popc(peers[i]);
    __syncwarp();
    rank[i] = (uint16_t)r;
  }
  __syncthreads();
  // per digit: exclusive scan over warps (thread d owns digit d), block digit totals
  {
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
      uint32_t c = warp_hist[w][d];
      warp_hist[w][d] = run;
      run += c;
    }
    // exclusive scan of `run` over the 256 digits
    uint32_t incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t wb = 0;
    for (int w = 0; w < warp; ++w) wb += warp_sums[w];
    digit_base[d] = wb + incl - run;
    // global position of the block's first item of digit d, minus its block-local position:
    // the write-out loop then needs no dependent global load
    out_base[d] = table[(size_t)d * nblocks + blockIdx.x] - (wb + incl - run);
  }
  __syncthreads();
  // place items at their block-local sorted position
#pragma unroll
  for (int i = 0; i < SORT_IPT; ++i) {
    size_t e = warp_base + (size_t)i * 32 + lane;
    if (e < n) {
      uint32_t d = (k[i] >> shift) & mask;
      uint32_t lp = digit_base[d] + warp_hist[warp][d] + rank[i];
      s_keys[lp] = k[i];
      s_vals[lp] = vals_in[e];  // payload goes straight from global to its sorted slot
    }
  }
  __syncthreads();
  const uint32_t n_here = (uint32_t)min((size_t)SORT_TILE, (size_t)n - tile_base);
#pragma unroll 4
  for (int i = 0; i < SORT_IPT; ++i) {
    uint32_t lp = i * SORT_THREADS + threadIdx.x;
    if (lp < n_here) {
      uint32_t key = s_keys[lp];
      uint32_t d = (key >> shift) & mask;
      uint32_t pos = out_base[d] + lp;
      keys_out[pos] = key;
      vals_out[pos] = s_vals[lp];
    }
  }
}

// ---------------------------------------------------------------- specialised scatter
// The scatter of one radix pass (per-block digit counts and their digit-major scan come from radix_hist_kernel /
// radix_scan_kernel), specialised by the position of the pass in the binning:
//   FIRST_DEPTH  keys are read straight from the FP32 depth array, the payload (Gaussian index) is generated;
//   LAST_DEPTH   the keys are dropped and the tile rect of every Gaussian is gathered into depth order (packed
//                8 bytes), so the count and emit kernels stream instead of gathering by sorted id twice;
//   LAST_TILE    the tile ranges are extracted from the block's sorted tile and (unless the caller wants the
//                reference's int64 keys) the keys are dropped.
// BITS = width of the digit: one VOTE per digit bit, so the tile passes split their bits evenly (13 -> 7 + 6).
//
// Measured and rejected (cfg 2, tools/bench_binning.py --kernels, tools/experiments_onesweep_binning.patch): the same
// passes as single "onesweep" kernels (global histograms up front, decoupled look-back between tiles, tickets).
// With 733 (Gaussians) / 2478 (duplicates) tiles per pass the look-back chain costs what the histogram + scan
// kernels cost: 551 us of kernel time against 469 us for the three-kernel passes.

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t x, uint32_t *smem8, uint32_t &block_total);

#ifndef GS3D_SCATTER_MINB
#define GS3D_SCATTER_MINB 4  // blocks per SM the scatter is compiled for (64 registers, 43 KB shared memory each)
#endif
#ifndef GS3D_SCATTER_MINB_SMALL
#define GS3D_SCATTER_MINB_SMALL 6  // same for the 2048-item tiles of the Gaussian-level passes (27 KB each)
#endif
#ifndef GS3D_DEPTH_IPT
#define GS3D_DEPTH_IPT 8  // items per thread of the last depth pass (3 M Gaussians: 1465 tiles instead of 733)
#endif
enum { SC_PLAIN = 0, SC_FIRST_DEPTH = 1, SC_LAST_DEPTH = 2, SC_LAST_TILE = 3 };
constexpr int MAX_TILE_PASSES = 3;  // n_tiles < 2^24

struct ScArgs {
  const uint32_t *keys_in, *vals_in;
  uint32_t *keys_out, *vals_out;
  uint32_t n;
  const uint32_t *n_dev;  // optional: the real item count lives on the device (n = capacity)
  int shift;
  uint32_t nblocks;
  const uint32_t *table;  // [RADIX][nblocks] global base of (digit, block)
  // LAST_DEPTH
  const int32_t *tl, *br;
  uint2 *rects_out;
  // LAST_TILE
  int32_t *start, *end;
  uint32_t n_tiles;
};

// (tlx, tly, w, h) of a tile rect in 8 bytes; an empty / degenerate rect has w = h = 0
__device__ __forceinline__ uint2 pack_rect(int2 a, int2 b) {
  const int w = b.x - a.x + 1, h = b.y - a.y + 1;
  if (w <= 0 || h <= 0) return make_uint2(0u, 0u);
  return make_uint2(((uint32_t)a.x & 0xffffu) | ((uint32_t)a.y << 16), ((uint32_t)w & 0xffffu) | ((uint32_t)h << 16));
}

template <int MODE, int BITS, int IPT>
__global__ void __launch_bounds__(SORT_THREADS, (IPT <= 8 ? GS3D_SCATTER_MINB_SMALL : GS3D_SCATTER_MINB))
scatter_kernel(const ScArgs a) {
  constexpr uint32_t NB = 1u << BITS;  // bins
  constexpr int SORT_TILE = SORT_THREADS * IPT;  // (shadows the namespace constant: this kernel's own tile)
  __shared__ uint32_t s_keys[SORT_TILE];
  __shared__ uint32_t s_vals[SORT_TILE];
  __shared__ uint32_t warp_hist[SORT_WARPS][NB];
  __shared__ uint32_t digit_base[NB];
  __shared__ uint32_t out_base[NB];
  __shared__ uint32_t warp_sums[SORT_WARPS];

  uint32_t n = a.n;
  if (a.n_dev) n = min(n, *a.n_dev);
  const size_t tile_base = (size_t)blockIdx.x * SORT_TILE;
  if (tile_base >= n) return;  // (whole block: uniform)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t mask = NB - 1;
  const int shift = a.shift;
  for (int i = threadIdx.x; i < SORT_WARPS * (int)NB; i += SORT_THREADS) (&warp_hist[0][0])[i] = 0;
  __syncthreads();

  // warp-striped tile: warp w owns items [w*32*IPT, (w+1)*32*IPT), item (i, lane) = w*32*IPT + i*32 + lane
  const size_t warp_base = tile_base + (size_t)warp * (32 * IPT);
  uint32_t k[IPT];
  uint16_t rank[IPT];
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const size_t e = warp_base + (size_t)i * 32 + lane;
    k[i] = (e < n) ? a.keys_in[e] : 0xffffffffu;
  }
  // Stable rank inside the warp chunk, in (i, lane) order: lanes with the same digit through one ballot per
  // digit bit (MATCH.ANY serialises over the distinct values in the warp, the VOTEs do not); all of them are
  // issued first, the serial part is only the per-digit running counter update.
  uint32_t peers[IPT];
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const size_t e = warp_base + (size_t)i * 32 + lane;
    const bool valid = e < n;
    uint32_t m = __ballot_sync(0xffffffffu, valid);
    const uint32_t d = (k[i] >> shift) & mask;
#pragma unroll
    for (int b = 0; b < BITS; ++b) {
      const uint32_t vote = __ballot_sync(0xffffffffu, (d >> b) & 1u);
      m &= ((d >> b) & 1u) ? vote : ~vote;
    }
    peers[i] = valid ? m : 0u;
  }
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    uint32_t r = 0;
    const uint32_t d = (k[i] >> shift) & mask;
    uint32_t before = 0;
    if (peers[i]) {
      before = warp_hist[warp][d];
      r = before + __popc(peers[i] & lt_mask);
    }
    __syncwarp();
    if (peers[i] && (peers[i] & lt_mask) == 0) warp_hist[warp][d] = before + __popc(peers[i]);
    __syncwarp();
    rank[i] = (uint16_t)r;
  }
  __syncthreads();
  // per digit: exclusive scan over warps (thread d owns digit d), block digit totals
  {
    const int d = threadIdx.x;
    uint32_t run = 0;
    if (d < (int)NB) {
#pragma unroll
      for (int w = 0; w < SORT_WARPS; ++w) {
        uint32_t c = warp_hist[w][d];
        warp_hist[w][d] = run;
        run += c;
      }
    }
    uint32_t incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t wb = 0;
    for (int w = 0; w < warp; ++w) wb += warp_sums[w];
    if (d < (int)NB) {
      digit_base[d] = wb + incl - run;
      // global position of the block's first item of digit d, minus its block-local position
      out_base[d] = a.table[(size_t)d * a.nblocks + blockIdx.x] - (wb + incl - run);
    }
  }
  __syncthreads();
  // place items at their block-local sorted position
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const size_t e = warp_base + (size_t)i * 32 + lane;
    if (e < n) {
      const uint32_t d = (k[i] >> shift) & mask;
      const uint32_t lp = digit_base[d] + warp_hist[warp][d] + rank[i];
      s_keys[lp] = k[i];
      s_vals[lp] = (MODE == SC_FIRST_DEPTH) ? (uint32_t)e : a.vals_in[e];  // payload: global -> sorted slot
    }
  }
  __syncthreads();
  const uint32_t n_here = (uint32_t)min((size_t)SORT_TILE, (size_t)n - tile_base);
#pragma unroll 4
  for (int i = 0; i < IPT; ++i) {
    const uint32_t lp = i * SORT_THREADS + threadIdx.x;
    if (lp < n_here) {
      const uint32_t key = s_keys[lp];
      const uint32_t d = (key >> shift) & mask;
      const uint32_t pos = out_base[d] + lp;
      const uint32_t val = s_vals[lp];
      a.vals_out[pos] = val;
      if (MODE == SC_PLAIN || MODE == SC_FIRST_DEPTH) {
        a.keys_out[pos] = key;
      } else if (MODE == SC_LAST_DEPTH) {
        a.rects_out[pos] = pack_rect(reinterpret_cast<const int2 *>(a.tl)[val], reinterpret_cast<const int2 *>(a.br)[val]);
      } else {  // SC_LAST_TILE: `key` is the tile id and this is the final order
        if (a.keys_out) a.keys_out[pos] = key;
        if (key < a.n_tiles) {
          // Neighbours inside one digit run of this block are neighbours in the final order.  The first / last
          // item of a run may or may not open / close its tile's range (the neighbour lives in another block):
          // min / max over all candidates gives the true bounds.  start is all-ones (-1) when empty: unsigned
          // min; end is -1 when empty: signed max.
          const bool run_first = lp == 0 || ((s_keys[lp - 1] >> shift) & mask) != d;
          const bool run_last = lp + 1 == n_here || ((s_keys[lp + 1] >> shift) & mask) != d;
          if (run_first || s_keys[lp - 1] != key) atomicMin(reinterpret_cast<uint32_t *>(a.start) + key, pos);
          if (run_last || s_keys[lp + 1] != key) atomicMax(a.end + key, (int32_t)(pos + 1));
        }
      }
    }
  }
}

// duplicate counts in depth order from the packed rects (streaming) -> sums per block of EMIT_GPB Gaussians
constexpr uint32_t EMIT_GPB = 1024;  // Gaussians per count / emit block (four rounds of 256)
__global__ void __launch_bounds__(256)
count_rects_kernel(uint32_t N, const uint2 *__restrict__ rects_sorted, uint32_t *__restrict__ block_sums,
                   int32_t *__restrict__ start, int32_t *__restrict__ end, uint32_t n_tiles) {
  __shared__ uint32_t sm[8];
  for (uint32_t t = blockIdx.x * 256 + threadIdx.x; t < n_tiles; t += gridDim.x * 256) {  // aabb_culling.h:248-249
    start[t] = -1;
    end[t] = -1;
  }
  uint32_t c = 0;
#pragma unroll
  for (uint32_t r = 0; r < EMIT_GPB / 256; ++r) {
    const uint32_t j = blockIdx.x * EMIT_GPB + r * 256 + threadIdx.x;
    if (j < N) {
      const uint32_t wh = rects_sorted[j].y;
      c += (wh & 0xffffu) * (wh >> 16);
    }
  }
  uint32_t tot;
  block_exclusive_scan_256(c, sm, tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

// Emit (tile, id) pairs in depth order; x outer, y inner like aabb_culling.h:29-38.  Streams the depth-ordered
// ids and packed rects (no gather); offsets come from the scan, so the emission order is deterministic.  Every
// warp writes its pairs as ONE contiguous run: output slot q of the warp finds its source lane by a 5-step
// shuffle search over the lanes' inclusive counts, so stores are fully coalesced whatever the rect sizes.
__global__ void __launch_bounds__(256)
emit_rects_kernel(uint32_t N, uint32_t n_dub, uint32_t n_tiles_w, const uint32_t *__restrict__ ids_sorted,
                  const uint2 *__restrict__ rects_sorted, const uint32_t *__restrict__ block_offsets,
                  uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, const uint32_t *__restrict__ n_dev) {
  __shared__ uint32_t sm8[8];
  if (n_dev) n_dub = min(n_dub, *n_dev);
  const int lane = threadIdx.x & 31;
  uint32_t round_base = block_offsets[blockIdx.x];
  for (uint32_t r = 0; r < EMIT_GPB / 256; ++r) {
    const uint32_t j = blockIdx.x * EMIT_GPB + r * 256 + threadIdx.x;
    uint32_t g = 0, c = 0, txy = 0, hh = 1;
    if (j < N) {
      g = ids_sorted[j];
      const uint2 rc = rects_sorted[j];
      txy = rc.x;
      hh = rc.y >> 16;
      c = (rc.y & 0xffffu) * hh;
      if (hh == 0) hh = 1;
    }
    uint32_t tot;
    if (r) __syncthreads();  // sm8 of the previous round has been read by everyone
    const uint32_t excl = block_exclusive_scan_256(c, sm8, tot);
    const uint32_t excl0 = __shfl_sync(0xffffffffu, excl, 0);
    const uint32_t incl_w = excl + c - excl0;  // inclusive count within the warp
    const uint32_t T = __shfl_sync(0xffffffffu, incl_w, 31);
    const uint32_t wbase = round_base + excl0;
    const float rh = __frcp_rn((float)hh);
    for (uint32_t q0 = 0; q0 < T; q0 += 32) {
      const uint32_t q = q0 + lane;
      int s = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const uint32_t v = __shfl_sync(0xffffffffu, incl_w, s + step - 1);
        if (v <= q) s += step;
      }
      const uint32_t excl_s = __shfl_sync(0xffffffffu, incl_w - c, s);
      const uint32_t txy_s = __shfl_sync(0xffffffffu, txy, s);
      const uint32_t h_s = __shfl_sync(0xffffffffu, hh, s);
      const float rh_s = __shfl_sync(0xffffffffu, rh, s);
      const uint32_t g_s = __shfl_sync(0xffffffffu, g, s);
      const uint32_t dest = wbase + q;
      if (q < T && dest < n_dub) {  // never write past the caller's buffers
        const uint32_t ql = q - excl_s;
        // ql / h_s without the ~20-instruction integer division: float estimate + one correction step (exact
        // below 2^22; larger rects -- > 4 M tiles under one Gaussian -- take the integer path)
        uint32_t dx;
        if (ql < (1u << 22)) {
          dx = (uint32_t)((float)ql * rh_s);
          const int rem = (int)(ql - dx * h_s);
          dx += (rem >= (int)h_s) ? 1u : 0u;
          dx -= (rem < 0) ? 1u : 0u;
        } else {
          dx = ql / h_s;
        }
        const uint32_t dy = ql - dx * h_s;
        keys[dest] = ((txy_s >> 16) + dy) * n_tiles_w + (txy_s & 0xffffu) + dx;
        vals[dest] = g_s;
      }
    }
    round_base += tot;
  }
}

struct RadixBuffers {
  uint32_t *table;   // [RADIX][nblocks_max]  (three-kernel pass: per-block digit counts)
  uint32_t *totals;  // [RADIX]
};

// GS3D_SORT=classic selects the round-1 three-kernel passes (histogram, scan, scatter) for A/B runs.
static bool use_classic() {
  static const int v = [] {
    const char *e = getenv("GS3D_SORT");
    return (e && e[0] == 'c') ? 1 : 0;
  }();
  return v != 0;
}

static int radix_pass(const uint32_t *kin, const uint32_t *vin, uint32_t *kout, uint32_t *vout,
                      uint32_t n, int shift, int bits, const RadixBuffers &rb, cudaStream_t st,
                      const uint32_t *n_dev = nullptr) {
  uint32_t nblocks = div_up(n, (uint32_t)SORT_TILE);
  uint32_t mask = (1u << bits) - 1u;
  GS3D_CUDA(cudaMemsetAsync(rb.totals, 0, RADIX * sizeof(uint32_t), st));
  radix_hist_kernel<SORT_IPT><<<nblocks, SORT_THREADS, 0, st>>>(kin, n, shift, mask, nblocks, rb.table, rb.totals, n_dev);
  GS3D_LAUNCH_CHECK();
  radix_scan_kernel<<<RADIX, 256, 0, st>>>(nblocks, rb.table, rb.totals);
  GS3D_LAUNCH_CHECK();
  static const bool ballot = [] { const char *e = getenv("GS3D_RANK"); return !(e && e[0] == 'm'); }();
  if (ballot)
    radix_scatter_kernel<true><<<nblocks, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, mask,
                                                                 nblocks, rb.table, n_dev);
  else
    radix_scatter_kernel<false><<<nblocks, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, mask,
                                                                  nblocks, rb.table, n_dev);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

// ---------------------------------------------------------------- keys, counts, emit, ranges

__global__ void __launch_bounds__(256)
init_depth_keys_kernel(uint32_t N, const float *__restrict__ depth, uint32_t *__restrict__ keys,
                       uint32_t *__restrict__ vals) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  keys[i] = __float_as_uint(depth[i]);  // raw bits: low word of the reference key (quirk Q9)
  vals[i] = i;
}

__device__ __forceinline__ uint32_t rect_count(const int32_t *__restrict__ tl,
                                               const int32_t *__restrict__ br, uint32_t g, int &tlx,
                                               int &tly, int &h) {
  int2 a = reinterpret_cast<const int2 *>(tl)[g];
  int2 b = reinterpret_cast<const int2 *>(br)[g];
  tlx = a.x; tly = a.y;
  int w = b.x - a.x + 1;
  h = b.y - a.y + 1;
  return (w > 0 && h > 0) ? (uint32_t)w * (uint32_t)h : 0u;
}

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t x, uint32_t *smem8,
                                                             uint32_t &block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) smem8[warp] = incl;
  __syncthreads();
  uint32_t wb = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    uint32_t s = smem8[w];
    if (w < warp) wb += s;
    tot += s;
  }
  block_total = tot;
  return wb + incl - x;
}

__global__ void __launch_bounds__(256)
count_sorted_kernel(uint32_t N, const uint32_t *__restrict__ sorted_ids,
                    const int32_t *__restrict__ tl, const int32_t *__restrict__ br,
                    uint32_t *__restrict__ block_sums) {
  __shared__ uint32_t sm[8];
  uint32_t j = blockIdx.x * 256 + threadIdx.x;
  uint32_t c = 0;
  if (j < N) {
    int a, b, h;
    c = rect_count(tl, br, sorted_ids[j], a, b, h);
  }
  uint32_t tot;
  block_exclusive_scan_256(c, sm, tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

// single block: exclusive scan of block_sums in place, grand total to *total.  Every thread owns a run of
// consecutive entries (two sequential sweeps over its run around ONE block-wide scan).
__global__ void __launch_bounds__(1024)
scan_block_sums_kernel(uint32_t nb, uint32_t *__restrict__ block_sums, uint32_t *__restrict__ total,
                       uint32_t capacity, uint32_t *__restrict__ n_eff, int64_t *__restrict__ n_dub_out,
                       int32_t *__restrict__ overflow) {
  __shared__ uint32_t wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t per = (nb + 1023u) / 1024u;
  const uint32_t b0 = min(nb, threadIdx.x * per), b1 = min(nb, b0 + per);
  uint32_t s = 0;
  for (uint32_t b = b0; b < b1; ++b) s += block_sums[b];
  uint32_t incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint32_t wb = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 32; ++w) {
    const uint32_t x = wsum[w];
    if (w < warp) wb += x;
    tot += x;
  }
  uint32_t run = wb + incl - s;
  for (uint32_t b = b0; b < b1; ++b) {
    const uint32_t x = block_sums[b];
    block_sums[b] = run;
    run += x;
  }
  if (threadIdx.x == 0) {
    *total = tot;
    if (n_eff) *n_eff = min(tot, capacity);           // what the duplicate-level kernels process
    if (n_dub_out) *n_dub_out = (int64_t)tot;         // the true count (the reference's N_with_dub)
    if (overflow) *overflow = tot > capacity ? 1 : 0; // the caller's id buffer was too small: result truncated
  }
}

// Emit (tile, id) pairs in depth order; x outer, y inner like aabb_culling.h:29-38.  Offsets come
// from the scan, so the emission order is deterministic.  Small rects are written by their own
// thread, large ones cooperatively by the warp (coalesced, no long serial tails).
__global__ void __launch_bounds__(256)
emit_kernel(uint32_t N, uint32_t n_dub, uint32_t n_tiles_w, const uint32_t *__restrict__ sorted_ids,
            const int32_t *__restrict__ tl, const int32_t *__restrict__ br,
            const uint32_t *__restrict__ block_offsets, uint32_t *__restrict__ keys,
            uint32_t *__restrict__ vals, const uint32_t *__restrict__ n_dev) {
  __shared__ uint32_t sm[8];
  constexpr uint32_t SMALL = 8;
  if (n_dev) n_dub = min(n_dub, *n_dev);
  uint32_t j = blockIdx.x * 256 + threadIdx.x;
  uint32_t c = 0, g = 0;
  int tlx = 0, tly = 0, h = 1;
  if (j < N) {
    g = sorted_ids[j];
    c = rect_count(tl, br, g, tlx, tly, h);
  }
  uint32_t tot;
  uint32_t off = block_offsets[blockIdx.x] + block_exclusive_scan_256(c, sm, tot);
  if (off > n_dub) { off = n_dub; c = 0; }            // never write past the caller's buffers
  if ((uint64_t)off + c > n_dub) c = n_dub - off;
  if (c <= SMALL) {
    int x = tlx, y = tly;
    for (uint32_t q = 0; q < c; ++q) {
      keys[off + q] = (uint32_t)(y * (int)n_tiles_w + x);
      vals[off + q] = g;
      if (++y >= tly + h) { y = tly; ++x; }
    }
  }
  const int lane = threadIdx.x & 31;
  uint32_t todo = __ballot_sync(0xffffffffu, c > SMALL);
  while (todo) {
    int src = __ffs(todo) - 1;
    todo &= todo - 1;
    uint32_t c_s = __shfl_sync(0xffffffffu, c, src);
    uint32_t off_s = __shfl_sync(0xffffffffu, off, src);
    uint32_t g_s = __shfl_sync(0xffffffffu, g, src);
    int tlx_s = __shfl_sync(0xffffffffu, tlx, src);
    int tly_s = __shfl_sync(0xffffffffu, tly, src);
    int h_s = __shfl_sync(0xffffffffu, h, src);
    for (uint32_t q = lane; q < c_s; q += 32) {
      int x = tlx_s + (int)(q / (uint32_t)h_s);
      int y = tly_s + (int)(q % (uint32_t)h_s);
      keys[off_s + q] = (uint32_t)(y * (int)n_tiles_w + x);
      vals[off_s + q] = g_s;
    }
  }
}

__global__ void __launch_bounds__(256)
ranges_kernel(uint32_t n_dub, const uint32_t *__restrict__ tile_keys, int32_t *__restrict__ start,
              int32_t *__restrict__ end, uint32_t n_tiles, const uint32_t *__restrict__ n_dev) {
  if (n_dev) n_dub = min(n_dub, *n_dev);
  // four consecutive keys per thread (one 16-byte load) plus the two neighbours
  const size_t g0 = 4 * ((size_t)blockIdx.x * 256 + threadIdx.x);
  if (g0 >= n_dub) return;
  uint32_t k[6];
  const bool vec = g0 + 4 <= n_dub && (reinterpret_cast<uintptr_t>(tile_keys) & 15) == 0;
  if (vec) {
    const uint4 v = *reinterpret_cast<const uint4 *>(tile_keys + g0);
    k[1] = v.x; k[2] = v.y; k[3] = v.z; k[4] = v.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) k[1 + i] = g0 + i < n_dub ? tile_keys[g0 + i] : 0xffffffffu;
  }
  k[0] = g0 > 0 ? tile_keys[g0 - 1] : 0xffffffffu;
  k[5] = g0 + 4 < n_dub ? tile_keys[g0 + 4] : 0xffffffffu;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const size_t g = g0 + i;
    if (g >= n_dub) break;
    const uint32_t t = k[1 + i];
    if (t >= n_tiles) continue;  // cannot happen for rects clamped to the image
    if (g == 0 || k[i] != t) start[t] = (int32_t)g;
    if (g == n_dub - 1 || k[2 + i] != t) end[t] = (int32_t)(g + 1);
  }
}

__global__ void __launch_bounds__(256)
keys64_kernel(uint32_t n_dub, const uint32_t *__restrict__ tile_keys,
              const int32_t *__restrict__ ids, const float *__restrict__ depth,
              int64_t *__restrict__ keys64, const uint32_t *__restrict__ n_dev) {
  if (n_dev) n_dub = min(n_dub, *n_dev);
  uint32_t g = blockIdx.x * 256 + threadIdx.x;
  if (g >= n_dub) return;
  uint64_t k = ((uint64_t)tile_keys[g] << 32) | (uint64_t)__float_as_uint(depth[ids[g]]);
  keys64[g] = (int64_t)k;
}

__global__ void copy_u32_kernel(uint32_t n, const uint32_t *__restrict__ a, uint32_t *__restrict__ b) {
  uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) b[i] = a[i];
}

static uint32_t table_elems(uint32_t n) { return RADIX * div_up(n ? n : 1u, (uint32_t)(SORT_THREADS * GS3D_DEPTH_IPT)); }

constexpr uint32_t COUNTER_WORDS = 8 * RADIX + 64;    // digit totals [8 passes][RADIX], total / n_eff

template <int MODE, int IPT>
static int scatter_launch(const ScArgs &a, int bits, cudaStream_t st) {
  if (a.nblocks == 0) return GS3D_OK;
  switch (bits) {
    case 8: scatter_kernel<MODE, 8, IPT><<<a.nblocks, SORT_THREADS, 0, st>>>(a); break;
    case 7: scatter_kernel<MODE, 7, IPT><<<a.nblocks, SORT_THREADS, 0, st>>>(a); break;
    case 6: scatter_kernel<MODE, 6, IPT><<<a.nblocks, SORT_THREADS, 0, st>>>(a); break;
    case 5: scatter_kernel<MODE, 5, IPT><<<a.nblocks, SORT_THREADS, 0, st>>>(a); break;
    default: scatter_kernel<MODE, 4, IPT><<<a.nblocks, SORT_THREADS, 0, st>>>(a); break;
  }
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}
template <int MODE>
static int scatter_launch8(const ScArgs &a, cudaStream_t st) {  // depth passes: always 8 bits
  if (a.nblocks == 0) return GS3D_OK;
  scatter_kernel<MODE, 8, GS3D_DEPTH_IPT><<<a.nblocks, SORT_THREADS, 0, st>>>(a);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

// histogram + scan + specialised scatter of one pass; `totals` is this pass's own zeroed row
template <int MODE, bool DEPTH>
static int radix_pass2(ScArgs a, int bits, uint32_t *table, uint32_t *totals, cudaStream_t st) {
  // 2048-item tiles only pay off where the scatter is latency-bound on its gather (measured, cfg 2: last depth
  // pass 64 -> 48 us; the plain depth passes gain nothing and their histogram / scan get longer tables)
  constexpr int IPT = (DEPTH && MODE == SC_LAST_DEPTH) ? GS3D_DEPTH_IPT : SORT_IPT;
  a.nblocks = div_up(a.n, (uint32_t)(SORT_THREADS * IPT));
  a.table = table;
  if (a.nblocks == 0) return GS3D_OK;
  radix_hist_kernel<IPT><<<a.nblocks, SORT_THREADS, 0, st>>>(a.keys_in, a.n, a.shift, (1u << bits) - 1u, a.nblocks,
                                                             table, totals, a.n_dev);
  GS3D_LAUNCH_CHECK();
  radix_scan_kernel<<<RADIX, 256, 0, st>>>(a.nblocks, table, totals);
  GS3D_LAUNCH_CHECK();
  if (IPT != SORT_IPT) return scatter_launch8<MODE>(a, st);
  return scatter_launch<MODE, SORT_IPT>(a, bits, st);
}

}  // namespace gs3d

using namespace gs3d;

extern "C" {

size_t gs3d_binning_scratch_bytes(uint32_t N, uint32_t n_dub) {
  const uint32_t nmax = N > n_dub ? N : n_dub;
  size_t b = 0;
  b += 4 * align_up((size_t)N * 4);                                   // depth keys/vals ping-pong
  b += align_up((size_t)N * 8);                                       // packed rects in depth order
  b += align_up((size_t)div_up(N ? N : 1u, 256u) * 8);                // block sums / scan status (64-bit)
  b += 2 * align_up((size_t)table_elems(nmax) * 4);                   // digit table | look-back status A, B
  b += align_up(COUNTER_WORDS * 4);                                   // histograms + tickets + totals
  b += 3 * align_up((size_t)n_dub * 4);                               // dup keys x2, vals x1
  return b + 1024;
}

// n_dub = capacity of gaussian_ids.  device_count == false: the caller knows the exact count (reference
// contract).  device_count == true: the count is only known on the device -- duplicate-level kernels are launched
// for the capacity and bound themselves by the scanned total; nothing is read back.
static int binning_core(uint32_t N, uint32_t n_dub, uint32_t n_tiles_h, uint32_t n_tiles_w,
                        const int32_t *aabb_topleft, const int32_t *aabb_bottomright, const float *depth,
                        int32_t *gaussian_ids, int32_t *start, int32_t *end, int64_t *sorted_keys, int check_count,
                        bool device_count, int64_t *n_dub_out_dev, int32_t *overflow_dev, void *scratch,
                        size_t scratch_bytes, cudaStream_t st) {
  const uint32_t n_tiles = n_tiles_h * n_tiles_w;
  GS3D_REQUIRE(start && end && n_tiles > 0, GS3D_EINVAL, "tile_culling_aabb_start_end: bad tiles");
  if (N == 0) {
    // aabb_culling.h:248-249
    GS3D_CUDA(cudaMemsetAsync(start, 0xff, sizeof(int32_t) * n_tiles, st));
    GS3D_CUDA(cudaMemsetAsync(end, 0xff, sizeof(int32_t) * n_tiles, st));
    GS3D_REQUIRE(!check_count || n_dub == 0, GS3D_ECOUNT, "n_dub = %u but N = 0", n_dub);
    if (n_dub_out_dev) GS3D_CUDA(cudaMemsetAsync(n_dub_out_dev, 0, sizeof(int64_t), st));
    if (overflow_dev) GS3D_CUDA(cudaMemsetAsync(overflow_dev, 0, sizeof(int32_t), st));
    return GS3D_OK;
  }
  GS3D_REQUIRE(aabb_topleft && aabb_bottomright && depth && scratch, GS3D_EINVAL,
               "tile_culling_aabb_start_end: null argument");
  GS3D_REQUIRE(n_dub == 0 || gaussian_ids, GS3D_EINVAL, "tile_culling_aabb_start_end: null ids");
  GS3D_REQUIRE(n_dub < (1u << 30) && N < (1u << 30), GS3D_EUNSUPPORTED,
               "tile_culling_aabb_start_end: at most 2^30 - 1 Gaussians / duplicates (got %u / %u)", N, n_dub);
  GS3D_REQUIRE(scratch_bytes >= gs3d_binning_scratch_bytes(N, n_dub), GS3D_EINVAL,
               "tile_culling_aabb_start_end: scratch too small (%zu < %zu)", scratch_bytes,
               gs3d_binning_scratch_bytes(N, n_dub));
  const uint32_t nmax = N > n_dub ? N : n_dub;
  Scratch sc(scratch, scratch_bytes);
  uint32_t *kA = sc.take<uint32_t>(N), *kB = sc.take<uint32_t>(N);
  uint32_t *vA = sc.take<uint32_t>(N), *vB = sc.take<uint32_t>(N);
  uint2 *rects = sc.take<uint2>(N);
  const uint32_t nb256 = div_up(N, 256u);
  unsigned long long *scan_status = sc.take<unsigned long long>(nb256);
  uint32_t *statusA = sc.take<uint32_t>(table_elems(nmax)), *statusB = sc.take<uint32_t>(table_elems(nmax));
  uint32_t *counters = sc.take<uint32_t>(COUNTER_WORDS);
  uint32_t *dK0 = sc.take<uint32_t>(n_dub), *dK1 = sc.take<uint32_t>(n_dub);
  uint32_t *dV0 = sc.take<uint32_t>(n_dub);
  GS3D_REQUIRE(kA && kB && vA && vB && rects && scan_status && statusA && statusB && counters &&
                   (n_dub == 0 || (dK0 && dK1 && dV0)),
               GS3D_EINVAL, "tile_culling_aabb_start_end: scratch exhausted");
  uint32_t *total = counters + 8 * RADIX;     // total, n_eff (rows 0..7 of `counters`: digit totals per pass)
  uint32_t *n_eff = device_count ? total + 1 : nullptr;  // min(total, capacity), written by the scan
  int tile_bits = 0;
  while ((1u << tile_bits) < n_tiles) ++tile_bits;
  int n_pass = (tile_bits + RADIX_BITS - 1) / RADIX_BITS;
  if (n_pass < 1) n_pass = 1;
  uint32_t *ids_u = reinterpret_cast<uint32_t *>(gaussian_ids);

  if (use_classic()) {
    // ---------------- round-1 pipeline: three kernels per pass, gathers in count / emit
    RadixBuffers rb;
    rb.table = statusA;
    rb.totals = counters;  // (re-zeroed by every pass)
    uint32_t *block_sums = reinterpret_cast<uint32_t *>(scan_status);
    GS3D_CUDA(cudaMemsetAsync(start, 0xff, sizeof(int32_t) * n_tiles, st));
    GS3D_CUDA(cudaMemsetAsync(end, 0xff, sizeof(int32_t) * n_tiles, st));
    init_depth_keys_kernel<<<nb256, 256, 0, st>>>(N, depth, kA, vA);
    GS3D_LAUNCH_CHECK();
    for (int p = 0; p < 4; ++p) {
      int rc = (p & 1) ? radix_pass(kB, vB, kA, vA, N, 8 * p, 8, rb, st)
                       : radix_pass(kA, vA, kB, vB, N, 8 * p, 8, rb, st);
      if (rc) return rc;
    }
    count_sorted_kernel<<<nb256, 256, 0, st>>>(N, vA, aabb_topleft, aabb_bottomright, block_sums);
    GS3D_LAUNCH_CHECK();
    scan_block_sums_kernel<<<1, 1024, 0, st>>>(nb256, block_sums, total, n_dub, n_eff, n_dub_out_dev, overflow_dev);
    GS3D_LAUNCH_CHECK();
    if (check_count && !device_count) {
      int64_t *box = pinned_mailbox();
      GS3D_REQUIRE(box != nullptr, GS3D_ECUDA, "pinned mailbox unavailable");
      *box = 0;
      GS3D_CUDA(cudaMemcpyAsync(box, total, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      GS3D_CUDA(cudaStreamSynchronize(st));
      GS3D_REQUIRE((uint32_t)*box == n_dub, GS3D_ECOUNT,
                   "tile rects add up to %u duplicates but gaussian_ids has %u (aabb_culling.h:228)",
                   (uint32_t)*box, n_dub);
    }
    if (n_dub == 0) return GS3D_OK;
    uint32_t *vcur = (n_pass & 1) ? dV0 : ids_u;
    uint32_t *vnext = (n_pass & 1) ? ids_u : dV0;
    uint32_t *kcur = dK0, *knext = dK1;
    emit_kernel<<<nb256, 256, 0, st>>>(N, n_dub, n_tiles_w, vA, aabb_topleft, aabb_bottomright,
                                       block_sums, kcur, vcur, n_eff);
    GS3D_LAUNCH_CHECK();
    for (int p = 0; p < n_pass; ++p) {
      int bits = tile_bits - p * RADIX_BITS;
      if (bits > RADIX_BITS) bits = RADIX_BITS;
      if (bits < 1) bits = 1;
      int rc = radix_pass(kcur, vcur, knext, vnext, n_dub, p * RADIX_BITS, bits, rb, st, n_eff);
      if (rc) return rc;
      uint32_t *t = kcur; kcur = knext; knext = t;
      t = vcur; vcur = vnext; vnext = t;
    }
    ranges_kernel<<<div_up(n_dub, 1024u), 256, 0, st>>>(n_dub, kcur, start, end, n_tiles, n_eff);
    GS3D_LAUNCH_CHECK();
    if (sorted_keys) {
      keys64_kernel<<<div_up(n_dub, 256u), 256, 0, st>>>(n_dub, kcur, gaussian_ids, depth, sorted_keys, n_eff);
      GS3D_LAUNCH_CHECK();
    }
    return GS3D_OK;
  }

  // ---------------- specialised passes: one memset, 4 depth passes, count + scan, emit, tile passes
  GS3D_REQUIRE(n_pass <= MAX_TILE_PASSES, GS3D_EUNSUPPORTED, "more than 2^24 tiles (%u)", n_tiles);
  GS3D_REQUIRE(n_tiles_w <= 0xffffu && n_tiles_h <= 0xffffu, GS3D_EUNSUPPORTED,
               "tile grid %u x %u exceeds the packed rect range", n_tiles_h, n_tiles_w);
  GS3D_CUDA(cudaMemsetAsync(counters, 0, COUNTER_WORDS * sizeof(uint32_t), st));
  uint32_t *table = statusA;
  uint32_t *block_sums = reinterpret_cast<uint32_t *>(scan_status);
  {
    ScArgs a = {};
    a.n = N;
    a.tl = aabb_topleft; a.br = aabb_bottomright; a.rects_out = rects;
    // pass 0: depth -> (kA, vA);  1: -> (kB, vB);  2: -> (kA, vA);  3: -> ids in depth order (vB) + packed rects
    a.keys_in = reinterpret_cast<const uint32_t *>(depth); a.vals_in = nullptr; a.keys_out = kA; a.vals_out = vA;
    a.shift = 0;
    int rc = radix_pass2<SC_FIRST_DEPTH, true>(a, 8, table, counters + 0 * RADIX, st);
    if (rc) return rc;
    a.keys_in = kA; a.vals_in = vA; a.keys_out = kB; a.vals_out = vB; a.shift = 8;
    rc = radix_pass2<SC_PLAIN, true>(a, 8, table, counters + 1 * RADIX, st);
    if (rc) return rc;
    a.keys_in = kB; a.vals_in = vB; a.keys_out = kA; a.vals_out = vA; a.shift = 16;
    rc = radix_pass2<SC_PLAIN, true>(a, 8, table, counters + 2 * RADIX, st);
    if (rc) return rc;
    a.keys_in = kA; a.vals_in = vA; a.keys_out = nullptr; a.vals_out = vB; a.shift = 24;
    rc = radix_pass2<SC_LAST_DEPTH, true>(a, 8, table, counters + 3 * RADIX, st);
    if (rc) return rc;
  }
  // deterministic offsets: scan of the duplicate counts in depth order (+ start / end = -1)
  const uint32_t nb_emit = div_up(N, EMIT_GPB);
  count_rects_kernel<<<nb_emit, 256, 0, st>>>(N, rects, block_sums, start, end, n_tiles);
  GS3D_LAUNCH_CHECK();
  scan_block_sums_kernel<<<1, 1024, 0, st>>>(nb_emit, block_sums, total, n_dub, n_eff, n_dub_out_dev, overflow_dev);
  GS3D_LAUNCH_CHECK();
  if (check_count && !device_count) {
    int64_t *box = pinned_mailbox();
    GS3D_REQUIRE(box != nullptr, GS3D_ECUDA, "pinned mailbox unavailable");
    *box = 0;
    GS3D_CUDA(cudaMemcpyAsync(box, total, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    GS3D_CUDA(cudaStreamSynchronize(st));
    GS3D_REQUIRE((uint32_t)*box == n_dub, GS3D_ECOUNT,
                 "tile rects add up to %u duplicates but gaussian_ids has %u (aabb_culling.h:228)",
                 (uint32_t)*box, n_dub);
  }
  if (n_dub == 0) return GS3D_OK;
  uint32_t *vcur = (n_pass & 1) ? dV0 : ids_u;  // ping-pong so that the last pass lands in the caller's gaussian_ids
  uint32_t *vnext = (n_pass & 1) ? ids_u : dV0;
  uint32_t *kcur = dK0, *knext = dK1;
  emit_rects_kernel<<<nb_emit, 256, 0, st>>>(N, n_dub, n_tiles_w, vB, rects, block_sums, kcur, vcur, n_eff);
  GS3D_LAUNCH_CHECK();
  // tile-digit passes over the duplicates (high 32 bits of the reference key), bits split evenly
  int bits_pp = (tile_bits + n_pass - 1) / n_pass;
  if (bits_pp < 4) bits_pp = 4;
  for (int p = 0; p < n_pass; ++p) {
    ScArgs a = {};
    a.n = n_dub; a.n_dev = n_eff;
    a.keys_in = kcur; a.vals_in = vcur; a.vals_out = vnext;
    a.shift = p * bits_pp;
    int rc;
    if (p + 1 < n_pass) {
      a.keys_out = knext;
      rc = radix_pass2<SC_PLAIN, false>(a, bits_pp, table, counters + (4 + p) * RADIX, st);
    } else {
      a.keys_out = sorted_keys ? knext : nullptr;
      a.start = start; a.end = end; a.n_tiles = n_tiles;
      rc = radix_pass2<SC_LAST_TILE, false>(a, bits_pp, table, counters + (4 + p) * RADIX, st);
    }
    if (rc) return rc;
    uint32_t *t = kcur; kcur = knext; knext = t;
    t = vcur; vcur = vnext; vnext = t;
  }
  if (sorted_keys) {
    keys64_kernel<<<div_up(n_dub, 256u), 256, 0, st>>>(n_dub, kcur, gaussian_ids, depth, sorted_keys, n_eff);
    GS3D_LAUNCH_CHECK();
  }
  return GS3D_OK;
}

int gs3d_tile_culling_aabb_start_end(uint32_t N, uint32_t n_dub, uint32_t n_tiles_h,
                                     uint32_t n_tiles_w, const int32_t *aabb_topleft,
                                     const int32_t *aabb_bottomright, const float *depth,
                                     int32_t *gaussian_ids, int32_t *start, int32_t *end,
                                     int64_t *sorted_keys, int check_count, void *scratch,
                                     size_t scratch_bytes, void *stream) {
  return binning_core(N, n_dub, n_tiles_h, n_tiles_w, aabb_topleft, aabb_bottomright, depth, gaussian_ids, start,
                      end, sorted_keys, check_count, false, nullptr, nullptr, scratch, scratch_bytes,
                      as_stream(stream));
}

int gs3d_tile_culling_aabb_start_end_capacity(uint32_t N, uint32_t capacity, uint32_t n_tiles_h,
                                              uint32_t n_tiles_w, const int32_t *aabb_topleft,
                                              const int32_t *aabb_bottomright, const float *depth,
                                              int32_t *gaussian_ids, int32_t *start, int32_t *end,
                                              int64_t *n_dub_dev, int32_t *overflow_dev, void *scratch,
                                              size_t scratch_bytes, void *stream) {
  GS3D_REQUIRE(n_dub_dev && overflow_dev, GS3D_EINVAL, "tile_culling_aabb_start_end_capacity: null count / flag");
  return binning_core(N, capacity, n_tiles_h, n_tiles_w, aabb_topleft, aabb_bottomright, depth, gaussian_ids, start,
                      end, nullptr, 0, true, n_dub_dev, overflow_dev, scratch, scratch_bytes, as_stream(stream));
}

}  // extern "C"
