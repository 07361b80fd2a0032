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

__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const uint32_t *__restrict__ keys, uint32_t n, int shift, uint32_t mask,
                  uint32_t nblocks, uint32_t *__restrict__ table /*[RADIX][nblocks]*/,
                  uint32_t *__restrict__ totals /*[RADIX]*/, const uint32_t *__restrict__ n_dev) {
  __shared__ uint32_t h[RADIX];
  if (n_dev) n = min(n, *n_dev);  // capacity-sized launch: the real item count lives on the device
  h[threadIdx.x] = 0;
  __syncthreads();
  size_t base = (size_t)blockIdx.x * SORT_TILE;
  if (base + SORT_TILE <= n && (reinterpret_cast<uintptr_t>(keys) & 15) == 0) {
    // full tile: 16-byte loads (a histogram does not care which thread counts which key)
    const uint4 *k4 = reinterpret_cast<const uint4 *>(keys + base);
#pragma unroll
    for (int i = 0; i < SORT_IPT / 4; ++i) {
      const uint4 v = k4[i * SORT_THREADS + threadIdx.x];
      atomicAdd(&h[(v.x >> shift) & mask], 1u);
      atomicAdd(&h[(v.y >> shift) & mask], 1u);
      atomicAdd(&h[(v.z >> shift) & mask], 1u);
      atomicAdd(&h[(v.w >> shift) & mask], 1u);
    }
  } else {
#pragma unroll 4
    for (int i = 0; i < SORT_IPT; ++i) {
      size_t e = base + (size_t)i * SORT_THREADS + threadIdx.x;
      if (e < n) atomicAdd(&h[(keys[e] >> shift) & mask], 1u);
    }
  }
  __syncthreads();
  uint32_t c = h[threadIdx.x];
  table[(size_t)threadIdx.x * nblocks + blockIdx.x] = c;
  if (c) atomicAdd(&totals[threadIdx.x], c);
}

// One block per digit: base = sum of totals of lower digits, then exclusive scan along the blocks.
__global__ void __launch_bounds__(256)
radix_scan_kernel(uint32_t nblocks, uint32_t *__restrict__ table, const uint32_t *__restrict__ totals) {
  __shared__ uint32_t red[8];
  __shared__ uint32_t carry_s;
  const int d = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t v = (threadIdx.x < d) ? totals[threadIdx.x] : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0;
    for (int w = 0; w < 8; ++w) s += red[w];
    carry_s = s;
  }
  __syncthreads();
  uint32_t *row = table + (size_t)d * nblocks;
  for (uint32_t b0 = 0; b0 < nblocks; b0 += 256) {
    uint32_t b = b0 + threadIdx.x;
    uint32_t x = (b < nblocks) ? row[b] : 0u;
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    __syncthreads();  // red[] reuse + carry_s read ordering
    if (lane == 31) red[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += red[w];
    uint32_t carry = carry_s;
    if (b < nblocks) row[b] = carry + wbase + incl - x;
    __syncthreads();
    if (threadIdx.x == 255) carry_s = carry + wbase + incl;
    __syncthreads();
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

// ---------------------------------------------------------------- onesweep pass
// Single-kernel radix pass (Adinets & Merrill, "Onesweep", 2022): the global digit histograms of all
// passes are computed up front in one read of the keys; each block then ranks its tile, publishes its
// per-digit counts and obtains its global offsets by decoupled look-back over the preceding tiles.
// Tiles are handed out by an atomic ticket, so every tile a block waits on is already running.

constexpr uint32_t OS_FLAG_AGG = 1u << 30;     // tile's own digit count is published
constexpr uint32_t OS_FLAG_PREFIX = 2u << 30;  // inclusive prefix (all tiles up to this one) is published
constexpr uint32_t OS_VALUE_MASK = (1u << 30) - 1u;

__global__ void __launch_bounds__(SORT_THREADS)
multi_hist_kernel(const uint32_t *__restrict__ keys, uint32_t n, int shift0, int n_pass,
                  uint32_t *__restrict__ hist /*[n_pass][RADIX]*/) {
  __shared__ uint32_t h[4][RADIX];
  for (int p = 0; p < n_pass; ++p) h[p][threadIdx.x] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * SORT_TILE;
#pragma unroll 4
  for (int i = 0; i < SORT_IPT; ++i) {
    const size_t e = base + (size_t)i * SORT_THREADS + threadIdx.x;
    if (e < n) {
      const uint32_t k = keys[e] >> shift0;
      for (int p = 0; p < n_pass; ++p) atomicAdd(&h[p][(k >> (RADIX_BITS * p)) & (RADIX - 1)], 1u);
    }
  }
  __syncthreads();
  for (int p = 0; p < n_pass; ++p) {
    const uint32_t c = h[p][threadIdx.x];
    if (c) atomicAdd(&hist[p * RADIX + threadIdx.x], c);
  }
}

__device__ __forceinline__ uint32_t block_exclusive_scan_digits(uint32_t x, uint32_t *warp_sums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  __syncthreads();  // warp_sums may still be read from a previous call
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  uint32_t wb = 0;
  for (int w = 0; w < warp; ++w) wb += warp_sums[w];
  return wb + incl - x;
}

__global__ void __launch_bounds__(SORT_THREADS, 3)
radix_onesweep_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                      uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t n,
                      int shift, const uint32_t *__restrict__ ghist /*[RADIX] digit totals*/,
                      volatile uint32_t *status /*[tiles][RADIX], zeroed*/, uint32_t *ticket) {
  __shared__ uint32_t s_keys[SORT_TILE];
  __shared__ uint32_t s_vals[SORT_TILE];
  __shared__ uint32_t warp_hist[SORT_WARPS][RADIX];
  __shared__ uint32_t digit_base[RADIX];
  __shared__ uint32_t out_base[RADIX];
  __shared__ uint32_t warp_sums[SORT_WARPS];
  __shared__ uint32_t s_tile;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t mask = RADIX - 1;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&warp_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;

  const size_t tile_base = (size_t)tile * SORT_TILE;
  const size_t warp_base = tile_base + (size_t)warp * (32 * SORT_IPT);
  uint32_t k[SORT_IPT];
  uint16_t rank[SORT_IPT];
#pragma unroll
  for (int i = 0; i < SORT_IPT; ++i) {
    size_t e = warp_base + (size_t)i * 32 + lane;
    k[i] = (e < n) ? keys_in[e] : 0xffffffffu;
  }
  uint32_t peers[SORT_IPT];
#pragma unroll
  for (int i = 0; i < SORT_IPT; ++i) {
    const size_t e = warp_base + (size_t)i * 32 + lane;
    const bool valid = e < n;
    const uint32_t active = __ballot_sync(0xffffffffu, valid);
    peers[i] = valid ? __match_any_sync(active, (k[i] >> shift) & mask) : 0u;
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
  {
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
      uint32_t c = warp_hist[w][d];
      warp_hist[w][d] = run;
      run += c;
    }
    // publish this tile's digit count as early as possible
    volatile uint32_t *mine = status + (size_t)tile * RADIX + d;
    *mine = (tile == 0 ? OS_FLAG_PREFIX : OS_FLAG_AGG) | run;
    const uint32_t local_base = block_exclusive_scan_digits(run, warp_sums);
    const uint32_t gstart = block_exclusive_scan_digits(ghist[d], warp_sums);
    // decoupled look-back over the preceding tiles
    uint32_t excl = 0;
    if (tile != 0) {
      int tt = (int)tile - 1;
      while (true) {
        uint32_t v;
        do {
          v = status[(size_t)tt * RADIX + d];
        } while ((v >> 30) == 0);
        excl += v & OS_VALUE_MASK;
        if ((v >> 30) == 2 || tt == 0) break;
        --tt;
      }
      *mine = OS_FLAG_PREFIX | (excl + run);
    }
    digit_base[d] = local_base;
    out_base[d] = gstart + excl - local_base;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < SORT_IPT; ++i) {
    size_t e = warp_base + (size_t)i * 32 + lane;
    if (e < n) {
      uint32_t d = (k[i] >> shift) & mask;
      uint32_t lp = digit_base[d] + warp_hist[warp][d] + rank[i];
      s_keys[lp] = k[i];
      s_vals[lp] = vals_in[e];
    }
  }
  __syncthreads();
  const uint32_t n_here = (uint32_t)min((size_t)SORT_TILE, (size_t)n - tile_base);
#pragma unroll 4
  for (int i = 0; i < SORT_IPT; ++i) {
    uint32_t lp = i * SORT_THREADS + threadIdx.x;
    if (lp < n_here) {
      uint32_t key = s_keys[lp];
      uint32_t pos = out_base[(key >> shift) & mask] + lp;
      keys_out[pos] = key;
      vals_out[pos] = s_vals[lp];
    }
  }
}

struct RadixBuffers {
  uint32_t *table;   // [RADIX][nblocks_max]  (classic: per-block digit counts; onesweep: tile status)
  uint32_t *totals;  // [8][RADIX]  (classic uses row 0; onesweep: one histogram per pass) + tickets
};

// Both passes sort identically and measure the same on cfg 2 (0.64 vs 0.65 ms for the whole binning):
// the three-kernel pass stays the default because it has no inter-block spin-wait.
static bool use_onesweep() {
  static const int v = [] {
    const char *e = getenv("GS3D_SORT");
    return (e && e[0] == 'o') ? 1 : 0;  // GS3D_SORT=onesweep selects the single-kernel look-back pass
  }();
  return v != 0;
}

// histograms of `n_pass` consecutive 8-bit digits starting at bit `shift0`, rows [row0, row0+n_pass)
static int onesweep_histograms(const uint32_t *keys, uint32_t n, int shift0, int n_pass, int row0,
                               const RadixBuffers &rb, cudaStream_t st) {
  GS3D_CUDA(cudaMemsetAsync(rb.totals + row0 * RADIX, 0, (size_t)n_pass * RADIX * sizeof(uint32_t), st));
  multi_hist_kernel<<<div_up(n, (uint32_t)SORT_TILE), SORT_THREADS, 0, st>>>(keys, n, shift0, n_pass,
                                                                            rb.totals + row0 * RADIX);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

static int onesweep_pass(const uint32_t *kin, const uint32_t *vin, uint32_t *kout, uint32_t *vout,
                         uint32_t n, int shift, int hist_row, const RadixBuffers &rb, cudaStream_t st) {
  const uint32_t nblocks = div_up(n, (uint32_t)SORT_TILE);
  uint32_t *ticket = rb.totals + 8 * RADIX;
  GS3D_CUDA(cudaMemsetAsync(rb.table, 0, (size_t)nblocks * RADIX * sizeof(uint32_t), st));
  GS3D_CUDA(cudaMemsetAsync(ticket, 0, sizeof(uint32_t), st));
  radix_onesweep_kernel<<<nblocks, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift,
                                                          rb.totals + hist_row * RADIX, rb.table, ticket);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

static int radix_pass(const uint32_t *kin, const uint32_t *vin, uint32_t *kout, uint32_t *vout,
                      uint32_t n, int shift, int bits, const RadixBuffers &rb, cudaStream_t st,
                      const uint32_t *n_dev = nullptr) {
  uint32_t nblocks = div_up(n, (uint32_t)SORT_TILE);
  uint32_t mask = (1u << bits) - 1u;
  GS3D_CUDA(cudaMemsetAsync(rb.totals, 0, RADIX * sizeof(uint32_t), st));
  radix_hist_kernel<<<nblocks, SORT_THREADS, 0, st>>>(kin, n, shift, mask, nblocks, rb.table, rb.totals, n_dev);
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

// single block: exclusive scan of block_sums in place, grand total to *total
__global__ void __launch_bounds__(1024)
scan_block_sums_kernel(uint32_t nb, uint32_t *__restrict__ block_sums, uint32_t *__restrict__ total,
                       uint32_t capacity, uint32_t *__restrict__ n_eff, int64_t *__restrict__ n_dub_out,
                       int32_t *__restrict__ overflow) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < nb; b0 += 1024) {
    uint32_t b = b0 + threadIdx.x;
    uint32_t x = b < nb ? block_sums[b] : 0u;
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t wb = 0;
    for (int w = 0; w < warp; ++w) wb += wsum[w];
    uint32_t carry = carry_s;
    if (b < nb) block_sums[b] = carry + wb + incl - x;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wb + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const uint32_t t = carry_s;
    *total = t;
    if (n_eff) *n_eff = min(t, capacity);           // what the duplicate-level kernels process
    if (n_dub_out) *n_dub_out = (int64_t)t;         // the true count (the reference's N_with_dub)
    if (overflow) *overflow = t > capacity ? 1 : 0; // the caller's id buffer was too small: result truncated
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

static uint32_t table_elems(uint32_t n) { return RADIX * div_up(n ? n : 1u, (uint32_t)SORT_TILE); }

}  // namespace gs3d

using namespace gs3d;

extern "C" {

size_t gs3d_binning_scratch_bytes(uint32_t N, uint32_t n_dub) {
  size_t b = 0;
  b += 4 * align_up((size_t)N * 4);                                   // depth keys/vals ping-pong
  b += align_up((size_t)div_up(N ? N : 1u, 256u) * 4);                // block sums
  b += align_up((size_t)table_elems(N > n_dub ? N : n_dub) * 4);      // digit table
  b += align_up((8 * RADIX + 64) * 4) + 256;                          // histograms/totals + ticket + total
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
  // aabb_culling.h:248-249
  GS3D_CUDA(cudaMemsetAsync(start, 0xff, sizeof(int32_t) * n_tiles, st));
  GS3D_CUDA(cudaMemsetAsync(end, 0xff, sizeof(int32_t) * n_tiles, st));
  if (N == 0) {
    GS3D_REQUIRE(!check_count || n_dub == 0, GS3D_ECOUNT, "n_dub = %u but N = 0", n_dub);
    if (n_dub_out_dev) GS3D_CUDA(cudaMemsetAsync(n_dub_out_dev, 0, sizeof(int64_t), st));
    if (overflow_dev) GS3D_CUDA(cudaMemsetAsync(overflow_dev, 0, sizeof(int32_t), st));
    return GS3D_OK;
  }
  GS3D_REQUIRE(aabb_topleft && aabb_bottomright && depth && scratch, GS3D_EINVAL,
               "tile_culling_aabb_start_end: null argument");
  GS3D_REQUIRE(n_dub == 0 || gaussian_ids, GS3D_EINVAL, "tile_culling_aabb_start_end: null ids");
  GS3D_REQUIRE(scratch_bytes >= gs3d_binning_scratch_bytes(N, n_dub), GS3D_EINVAL,
               "tile_culling_aabb_start_end: scratch too small (%zu < %zu)", scratch_bytes,
               gs3d_binning_scratch_bytes(N, n_dub));
  Scratch sc(scratch, scratch_bytes);
  uint32_t *kA = sc.take<uint32_t>(N), *kB = sc.take<uint32_t>(N);
  uint32_t *vA = sc.take<uint32_t>(N), *vB = sc.take<uint32_t>(N);
  const uint32_t nb256 = div_up(N, 256u);
  uint32_t *block_sums = sc.take<uint32_t>(nb256);
  RadixBuffers rb;
  rb.table = sc.take<uint32_t>(table_elems(N > n_dub ? N : n_dub));
  rb.totals = sc.take<uint32_t>(8 * RADIX + 64);
  uint32_t *total = sc.take<uint32_t>(64);
  uint32_t *dK0 = sc.take<uint32_t>(n_dub), *dK1 = sc.take<uint32_t>(n_dub);
  uint32_t *dV0 = sc.take<uint32_t>(n_dub);
  GS3D_REQUIRE(kA && kB && vA && vB && block_sums && rb.table && rb.totals && total &&
                   (n_dub == 0 || (dK0 && dK1 && dV0)),
               GS3D_EINVAL, "tile_culling_aabb_start_end: scratch exhausted");
  uint32_t *n_eff = device_count ? total + 1 : nullptr;  // min(total, capacity), written by the scan kernel

  // 1. depth-byte passes over the Gaussians (low 32 bits of the reference key)
  init_depth_keys_kernel<<<nb256, 256, 0, st>>>(N, depth, kA, vA);
  GS3D_LAUNCH_CHECK();
  const bool onesweep = use_onesweep() && !device_count;
  if (onesweep) {
    int rc = onesweep_histograms(kA, N, 0, 4, 0, rb, st);
    if (rc) return rc;
  }
  for (int p = 0; p < 4; ++p) {
    int rc;
    if (onesweep)
      rc = (p & 1) ? onesweep_pass(kB, vB, kA, vA, N, 8 * p, p, rb, st)
                   : onesweep_pass(kA, vA, kB, vB, N, 8 * p, p, rb, st);
    else
      rc = (p & 1) ? radix_pass(kB, vB, kA, vA, N, 8 * p, 8, rb, st)
                   : radix_pass(kA, vA, kB, vB, N, 8 * p, 8, rb, st);
    if (rc) return rc;
  }
  // 2. deterministic offsets: scan of per-Gaussian duplicate counts in depth order
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
  // 3. tile-digit passes over the duplicates (high 32 bits of the reference key)
  int tile_bits = 0;
  while ((1u << tile_bits) < n_tiles) ++tile_bits;
  const int n_pass = (tile_bits + RADIX_BITS - 1) / RADIX_BITS;
  uint32_t *ids_u = reinterpret_cast<uint32_t *>(gaussian_ids);
  // ping-pong so that the last pass lands in the caller's gaussian_ids
  uint32_t *vcur = (n_pass & 1) ? dV0 : ids_u;
  uint32_t *vnext = (n_pass & 1) ? ids_u : dV0;
  uint32_t *kcur = dK0, *knext = dK1;
  emit_kernel<<<nb256, 256, 0, st>>>(N, n_dub, n_tiles_w, vA, aabb_topleft, aabb_bottomright,
                                     block_sums, kcur, vcur, n_eff);
  GS3D_LAUNCH_CHECK();
  if (onesweep && n_pass > 0) {
    int rc = onesweep_histograms(kcur, n_dub, 0, n_pass, 4, rb, st);
    if (rc) return rc;
  }
  for (int p = 0; p < n_pass; ++p) {
    int bits = tile_bits - p * RADIX_BITS;
    if (bits > RADIX_BITS) bits = RADIX_BITS;
    int rc = onesweep ? onesweep_pass(kcur, vcur, knext, vnext, n_dub, p * RADIX_BITS, 4 + p, rb, st)
                      : radix_pass(kcur, vcur, knext, vnext, n_dub, p * RADIX_BITS, bits, rb, st, n_eff);
    if (rc) return rc;
    uint32_t *t = kcur; kcur = knext; knext = t;
    t = vcur; vcur = vnext; vnext = t;
  }
  // 4. tile ranges (+ optional reconstruction of the reference's sorted int64 keys)
  const uint32_t nbd = div_up(n_dub, 256u);
  ranges_kernel<<<div_up(n_dub, 1024u), 256, 0, st>>>(n_dub, kcur, start, end, n_tiles, n_eff);
  GS3D_LAUNCH_CHECK();
  if (sorted_keys) {
    keys64_kernel<<<nbd, 256, 0, st>>>(n_dub, kcur, gaussian_ids, depth, sorted_keys, n_eff);
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
