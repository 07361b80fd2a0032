// Tile-row bands for the tile-sharded render (cfg 5: tiles are independent after
// projection, so rank r bins and composites only the tile rows of its band).
//
//   gs3d_row_duplicate_counts : duplicates per tile row (balances the bands) -- a per-block difference
//                               array in shared memory, one integer atomic per touched row and block
//   gs3d_clip_rects_to_rows   : restrict every rect to rows [row_begin, row_end) and COMPACT the
//                               Gaussians that still cover a tile (deterministic order: block counts ->
//                               scan -> write), so that the band's depth sort and emission run over
//                               ~N/world Gaussians instead of N
// Integer work, HBM-bound streaming (32 B read per Gaussian).
#include "common.cuh"

namespace gs3d {

constexpr int MAX_ROWS_SMEM = 4096;

__global__ void __launch_bounds__(256)
row_counts_kernel(uint32_t N, const int32_t *__restrict__ tl, const int32_t *__restrict__ br, int n_rows,
                  unsigned long long *__restrict__ diff /*[n_rows + 1], zeroed*/) {
  extern __shared__ int s_diff[];  // [n_rows + 1]
  for (int i = threadIdx.x; i <= n_rows; i += blockDim.x) s_diff[i] = 0;
  __syncthreads();
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < N; g += gridDim.x * blockDim.x) {
    const int2 a = reinterpret_cast<const int2 *>(tl)[g];
    const int2 b = reinterpret_cast<const int2 *>(br)[g];
    const int w = b.x - a.x + 1;
    if (w > 0 && b.y >= a.y) {
      const int y0 = min(max(a.y, 0), n_rows), y1 = min(max(b.y + 1, 0), n_rows);
      atomicAdd(&s_diff[y0], w);
      atomicAdd(&s_diff[y1], -w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= n_rows; i += blockDim.x) {
    const int v = s_diff[i];
    if (v) atomicAdd(&diff[i], (unsigned long long)(long long)v);  // two's complement: signed sum
  }
}

__global__ void row_counts_scan_kernel(int n_rows, const unsigned long long *__restrict__ diff,
                                       int64_t *__restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long run = 0;
    for (int i = 0; i < n_rows; ++i) {
      run += (long long)diff[i];
      out[i] = run;
    }
  }
}

__device__ __forceinline__ bool clip_one(const int32_t *tl, const int32_t *br, uint32_t g, int r0, int r1,
                                         int2 &a, int2 &b, uint32_t &cnt) {
  a = reinterpret_cast<const int2 *>(tl)[g];
  b = reinterpret_cast<const int2 *>(br)[g];
  a.y = max(a.y, r0);
  b.y = min(b.y, r1 - 1);
  const int w = b.x - a.x + 1, h = b.y - a.y + 1;
  cnt = (w > 0 && h > 0) ? (uint32_t)w * (uint32_t)h : 0u;
  return cnt != 0;
}

// pass 1: per block of 1024 Gaussians, (kept Gaussians, kept duplicates)
__global__ void __launch_bounds__(256)
clip_count_kernel(uint32_t N, const int32_t *__restrict__ tl, const int32_t *__restrict__ br, int r0, int r1,
                  uint32_t *__restrict__ blk_kept, unsigned long long *__restrict__ blk_dups) {
  __shared__ uint32_t s_k[8];
  __shared__ unsigned long long s_d[8];
  uint32_t kept = 0;
  unsigned long long dups = 0;
  const uint32_t base = blockIdx.x * 1024;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t g = base + i * 256 + threadIdx.x;
    if (g < N) {
      int2 a, b;
      uint32_t c;
      if (clip_one(tl, br, g, r0, r1, a, b, c)) { ++kept; dups += c; }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    kept += __shfl_xor_sync(0xffffffffu, kept, o);
    dups += __shfl_xor_sync(0xffffffffu, dups, o);
  }
  if ((threadIdx.x & 31) == 0) { s_k[threadIdx.x >> 5] = kept; s_d[threadIdx.x >> 5] = dups; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t k = 0; unsigned long long d = 0;
    for (int w = 0; w < 8; ++w) { k += s_k[w]; d += s_d[w]; }
    blk_kept[blockIdx.x] = k;
    blk_dups[blockIdx.x] = d;
  }
}

// pass 2 (one block): exclusive scan of the per-block kept counts; totals to totals[0..1]
__global__ void __launch_bounds__(1024)
clip_scan_kernel(uint32_t nb, uint32_t *__restrict__ blk_kept, const unsigned long long *__restrict__ blk_dups,
                 unsigned long long *__restrict__ totals) {
  __shared__ uint32_t wsum[32];
  __shared__ unsigned long long dsum[32];
  __shared__ uint32_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  unsigned long long dtot = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < nb; b0 += 1024) {
    const uint32_t b = b0 + threadIdx.x;
    const uint32_t x = b < nb ? blk_kept[b] : 0u;
    if (b < nb) dtot += blk_dups[b];
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t wb = 0;
    for (int w = 0; w < warp; ++w) wb += wsum[w];
    const uint32_t carry = carry_s;
    if (b < nb) blk_kept[b] = carry + wb + incl - x;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wb + incl;
    __syncthreads();
  }
  for (int o = 16; o > 0; o >>= 1) dtot += __shfl_xor_sync(0xffffffffu, dtot, o);
  if (lane == 0) dsum[warp] = dtot;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long d = 0;
    for (int w = 0; w < 32; ++w) d += dsum[w];
    totals[0] = carry_s;
    totals[1] = d;
  }
}

// pass 3: write the kept Gaussians in index order
__global__ void __launch_bounds__(256)
clip_write_kernel(uint32_t N, const int32_t *__restrict__ tl, const int32_t *__restrict__ br,
                  const float *__restrict__ depth, int r0, int r1, const uint32_t *__restrict__ blk_off,
                  int32_t *__restrict__ tl_out, int32_t *__restrict__ br_out, float *__restrict__ depth_out,
                  int32_t *__restrict__ index_out) {
  __shared__ uint32_t s_w[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t run = blk_off[blockIdx.x];
  const uint32_t base = blockIdx.x * 1024;
  for (int i = 0; i < 4; ++i) {
    const uint32_t g = base + i * 256 + threadIdx.x;
    int2 a = make_int2(0, 0), b = make_int2(0, 0);
    uint32_t c = 0;
    const bool keep = g < N && clip_one(tl, br, g, r0, r1, a, b, c);
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    uint32_t wb = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const uint32_t s = s_w[w];
      if (w < warp) wb += s;
      tot += s;
    }
    if (keep) {
      const uint32_t o = run + wb + __popc(m & ((1u << lane) - 1u));
      reinterpret_cast<int2 *>(tl_out)[o] = a;
      reinterpret_cast<int2 *>(br_out)[o] = b;
      depth_out[o] = depth[g];
      index_out[o] = (int32_t)g;
    }
    run += tot;
    __syncthreads();
  }
}

}  // namespace gs3d

using namespace gs3d;

extern "C" {

int gs3d_row_duplicate_counts(uint32_t N, const int32_t *aabb_topleft, const int32_t *aabb_bottomright,
                              uint32_t n_tiles_h, int64_t *row_counts, void *scratch, size_t scratch_bytes,
                              void *stream) {
  GS3D_REQUIRE(n_tiles_h > 0 && n_tiles_h < (uint32_t)MAX_ROWS_SMEM, GS3D_EINVAL,
               "row_duplicate_counts: n_tiles_h must be 1..%d (got %u)", MAX_ROWS_SMEM - 1, n_tiles_h);
  GS3D_REQUIRE(row_counts && scratch && scratch_bytes >= (n_tiles_h + 1) * sizeof(unsigned long long), GS3D_EINVAL,
               "row_duplicate_counts: bad output / scratch");
  cudaStream_t st = as_stream(stream);
  unsigned long long *diff = static_cast<unsigned long long *>(scratch);
  GS3D_CUDA(cudaMemsetAsync(diff, 0, (n_tiles_h + 1) * sizeof(unsigned long long), st));
  if (N) {
    GS3D_REQUIRE(aabb_topleft && aabb_bottomright, GS3D_EINVAL, "row_duplicate_counts: null rects");
    const uint32_t blocks = min(div_up(N, 256u), 148u * 8u);
    row_counts_kernel<<<blocks, 256, (n_tiles_h + 1) * sizeof(int), st>>>(N, aabb_topleft, aabb_bottomright,
                                                                          (int)n_tiles_h, diff);
    GS3D_LAUNCH_CHECK();
  }
  row_counts_scan_kernel<<<1, 32, 0, st>>>((int)n_tiles_h, diff, row_counts);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

size_t gs3d_clip_scratch_bytes(uint32_t N) {
  const size_t nb = div_up(N ? N : 1u, 1024u);
  return align_up(nb * sizeof(uint32_t)) + align_up(nb * sizeof(unsigned long long)) + 256;
}

int gs3d_clip_rects_to_rows(uint32_t N, const int32_t *aabb_topleft, const int32_t *aabb_bottomright,
                            const float *depth, int row_begin, int row_end, int32_t *tl_out,
                            int32_t *br_out, float *depth_out, int32_t *index_out, int64_t *counts_host,
                            void *scratch, size_t scratch_bytes, void *stream) {
  GS3D_REQUIRE(counts_host, GS3D_EINVAL, "clip_rects_to_rows: counts_host is null");
  counts_host[0] = counts_host[1] = 0;
  if (N == 0) return GS3D_OK;
  GS3D_REQUIRE(aabb_topleft && aabb_bottomright && depth && tl_out && br_out && depth_out && index_out && scratch,
               GS3D_EINVAL, "clip_rects_to_rows: null argument");
  GS3D_REQUIRE(scratch_bytes >= gs3d_clip_scratch_bytes(N), GS3D_EINVAL, "clip_rects_to_rows: scratch too small");
  cudaStream_t st = as_stream(stream);
  const uint32_t nb = div_up(N, 1024u);
  Scratch sc(scratch, scratch_bytes);
  uint32_t *blk_kept = sc.take<uint32_t>(nb);
  unsigned long long *blk_dups = sc.take<unsigned long long>(nb);
  unsigned long long *totals = sc.take<unsigned long long>(2);
  GS3D_REQUIRE(blk_kept && blk_dups && totals, GS3D_EINVAL, "clip_rects_to_rows: scratch exhausted");
  clip_count_kernel<<<nb, 256, 0, st>>>(N, aabb_topleft, aabb_bottomright, row_begin, row_end, blk_kept, blk_dups);
  GS3D_LAUNCH_CHECK();
  clip_scan_kernel<<<1, 1024, 0, st>>>(nb, blk_kept, blk_dups, totals);
  GS3D_LAUNCH_CHECK();
  clip_write_kernel<<<nb, 256, 0, st>>>(N, aabb_topleft, aabb_bottomright, depth, row_begin, row_end, blk_kept,
                                        tl_out, br_out, depth_out, index_out);
  GS3D_LAUNCH_CHECK();
  int64_t *box = pinned_mailbox();
  GS3D_REQUIRE(box != nullptr, GS3D_ECUDA, "pinned mailbox unavailable");
  GS3D_CUDA(cudaMemcpyAsync(box, totals, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));  // 64-byte mailbox
  GS3D_CUDA(cudaStreamSynchronize(st));
  counts_host[0] = box[0];
  counts_host[1] = box[1];
  return GS3D_OK;
}

}  // extern "C"
