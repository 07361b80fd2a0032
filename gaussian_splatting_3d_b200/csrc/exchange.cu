// Sparse gradient exchange for view-sharded training (new capability).
//
// One view touches only the Gaussians that reach a pixel before it saturates (cfg 2: ~8 % of the
// duplicates are ever composited), so the leaf gradients of a step -- 236 B/Gaussian, 708 MB dense at
// 3 M Gaussians, C = 4 -- are zero for most rows on every rank.  The compositing backward marks the
// rows it writes (`touched`); the ranks OR their marks, and only the union of touched rows is
// all-reduced: these two kernels pack those rows of the [sh | mean | qvec | svec | alpha] blocks into
// one dense [U, W] matrix and unpack the sum.  HBM-bound, one warp per row.
#include "common.cuh"

namespace gs3d {

constexpr int MAX_SEG = 8;
struct RowSegs {
  float *base[MAX_SEG];   // block s is [N, width[s]] row-major
  uint32_t width[MAX_SEG];
  uint32_t col[MAX_SEG];  // first column of block s in the packed row
  int n;
};

template <bool GATHER>
__global__ void __launch_bounds__(256)
rows_move_kernel(RowSegs segs, const int32_t *__restrict__ idx, uint32_t U, float *__restrict__ packed,
                 uint32_t W) {
  const uint32_t row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= U) return;
  const int lane = threadIdx.x & 31;
  const size_t g = (size_t)idx[row];
  float *dst = packed + (size_t)row * W;
  for (int s = 0; s < segs.n; ++s) {
    float *src = segs.base[s] + g * segs.width[s];
    for (uint32_t k = lane; k < segs.width[s]; k += 32) {
      if (GATHER) dst[segs.col[s] + k] = src[k];
      else src[k] = dst[segs.col[s] + k];
    }
  }
  if (GATHER) {  // zero the padding columns so that the all-reduce never sums garbage
    const uint32_t used = segs.col[segs.n - 1] + segs.width[segs.n - 1];
    for (uint32_t k = used + lane; k < W; k += 32) dst[k] = 0.0f;
  }
}

static int rows_move(bool gather, int n_seg, const uint64_t *bases, const uint32_t *widths,
                     const int32_t *idx, uint32_t U, float *packed, uint32_t W, void *stream) {
  GS3D_REQUIRE(n_seg >= 1 && n_seg <= MAX_SEG, GS3D_EINVAL, "rows: n_seg must be 1..%d (got %d)", MAX_SEG, n_seg);
  if (U == 0) return GS3D_OK;
  GS3D_REQUIRE(bases && widths && idx && packed, GS3D_EINVAL, "rows: null argument");
  RowSegs segs;
  segs.n = n_seg;
  uint32_t col = 0;
  for (int s = 0; s < n_seg; ++s) {
    segs.base[s] = reinterpret_cast<float *>(static_cast<uintptr_t>(bases[s]));
    segs.width[s] = widths[s];
    segs.col[s] = col;
    col += widths[s];
    GS3D_REQUIRE(segs.base[s] && widths[s] > 0, GS3D_EINVAL, "rows: bad block %d", s);
  }
  GS3D_REQUIRE(W >= col, GS3D_EINVAL, "rows: packed width %u < sum of block widths %u", W, col);
  cudaStream_t st = as_stream(stream);
  if (gather) rows_move_kernel<true><<<div_up(U, 8u), 256, 0, st>>>(segs, idx, U, packed, W);
  else rows_move_kernel<false><<<div_up(U, 8u), 256, 0, st>>>(segs, idx, U, packed, W);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

// ---------------------------------------------------------------- push exchange over NVLink
//
// Symmetric-memory form of the same idea, without NCCL and without a host round trip: every rank
// keeps its step's gradients in a PRIVATE buffer (the backward kernels' target) and, when its backward
// is done, pushes the rows it touched -- once each, already summed over its tiles and views -- into
// the RESULT buffer of every rank: one multimem.red.add on the NVSwitch multicast address of the
// result buffers (the switch replicates it), or one red.global.add per peer over NVLink.  It also
// sets the row's byte in every rank's union marks, so each rank knows which result rows to clear
// before the next step.  Per rank and step: U_own rows x 240 B leave the GPU (cfg 2: ~60 k rows,
// 15 MB), against 708 MB for a dense all-reduce.

struct PushArgs {
  const uint8_t *marks;
  uint32_t N;
  int n_seg, n_units, n_peers;
  const float *src[MAX_SEG];
  uint32_t width[MAX_SEG];
  uint64_t off[MAX_SEG];      // float offset of block s inside the result buffer
  uint32_t unit0[MAX_SEG];    // first unit of block s
  uint32_t vec[MAX_SEG];      // 1: units are float4, 0: scalars
  float *dst_mc;              // multicast address of the result buffers (or null)
  float *dst_peer[8];         // result buffer of every rank (this one included)
  uint8_t *union_peer[8];     // union marks of every rank
};

__device__ __forceinline__ void red_add_f32_sys(float *addr, float v) {
  asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_v4_sys(float *addr, float4 v) {
  asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mm_red_add_f32(float *addr, float v) {
  asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void mm_red_add_v4(float *addr, float4 v) {
  asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Warp-cooperative walk over the marked Gaussians of a 512-mark chunk: each lane fetches 16 marks with
// one 16-byte load (the common case -- nothing marked -- costs just that), then the warp visits every
// non-zero byte together.  f(g) is called by all 32 lanes.  clear != 0 zeroes the visited words.
template <typename F>
__device__ __forceinline__ void for_each_marked(uint8_t *marks, uint32_t N, int clear, F f,
                                                uint32_t chunk_stride = 1, uint32_t chunk_offset = 0) {
  const uint32_t warp = (blockIdx.x * 8 + (threadIdx.x >> 5)) * chunk_stride + chunk_offset;
  if ((size_t)warp * 512 >= N) return;
  const int lane = threadIdx.x & 31;
  const size_t c0 = (size_t)warp * 512 + (size_t)lane * 16;  // first mark of this lane
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  const bool full = c0 + 16 <= N && ((reinterpret_cast<uintptr_t>(marks) & 15) == 0);
  if (full) {
    const uint4 v = *reinterpret_cast<const uint4 *>(marks + c0);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  } else {
    for (int b = 0; b < 16; ++b)
      if (c0 + b < N && marks[c0 + b]) w[b >> 2] |= 0xffu << (8 * (b & 3));
  }
  const bool any = (w[0] | w[1] | w[2] | w[3]) != 0u;
  uint32_t todo = __ballot_sync(0xffffffffu, any);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const size_t base = (size_t)warp * 512 + (size_t)src * 16;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t word = __shfl_sync(0xffffffffu, w[q], src);
      while (word) {
        const int bit = __ffs(word) - 1;
        word &= ~(0xffu << (bit & ~7));
        f(base + 4 * q + (bit >> 3));
      }
    }
  }
  if (any && clear) {
    if (full) *reinterpret_cast<uint4 *>(marks + c0) = make_uint4(0u, 0u, 0u, 0u);
    else for (int b = 0; b < 16; ++b) if (c0 + b < N) marks[c0 + b] = 0;
  }
}

__global__ void __launch_bounds__(256)
rows_push_kernel(const PushArgs a) {
  const int lane = threadIdx.x & 31;
  for_each_marked(const_cast<uint8_t *>(a.marks), a.N, 0, [&](size_t g) {
    for (int u = lane; u < a.n_units; u += 32) {
      int s = 0;
      while (s + 1 < a.n_seg && (uint32_t)u >= a.unit0[s + 1]) ++s;
      const uint32_t k = u - a.unit0[s];
      if (a.vec[s]) {
        const size_t e = g * a.width[s] + 4 * k;
        const float4 v = *reinterpret_cast<const float4 *>(a.src[s] + e);
        if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
        if (a.dst_mc) mm_red_add_v4(a.dst_mc + a.off[s] + e, v);
        else for (int r = 0; r < a.n_peers; ++r) red_add_v4_sys(a.dst_peer[r] + a.off[s] + e, v);
      } else {
        const size_t e = g * a.width[s] + k;
        const float v = a.src[s][e];
        if (v == 0.f) continue;
        if (a.dst_mc) mm_red_add_f32(a.dst_mc + a.off[s] + e, v);
        else for (int r = 0; r < a.n_peers; ++r) red_add_f32_sys(a.dst_peer[r] + a.off[s] + e, v);
      }
    }
    if (lane < a.n_peers && a.union_peer[lane]) a.union_peer[lane][g] = 1;
  });
}

// ---------------------------------------------------------------- pull exchange (sparse NVLS all-reduce)
//
// At 8 ranks the push form makes every GPU ingest 8 x U rows as 16-byte reductions and the NVLink
// packet rate, not the bandwidth, bounds it.  The pull form lets the switch do the sum: the ranks
// first OR their marks into every rank's union marks (one byte store per marked row and peer), then
// rank r takes every n_peers-th 512-row chunk and, for each union-marked row in it, reads the SUM of
// all ranks' private rows with multimem.ld_reduce.add (reduced inside the NVSwitch) and broadcasts it
// into all result buffers with multimem.st.  Each GPU then ingests ~(1 + 1/n) x U rows instead of n x U.

struct PullArgs {
  uint8_t *union_marks;
  uint32_t N;
  int n_seg, n_units, n_peers, rank;
  uint32_t width[MAX_SEG];
  uint64_t off[MAX_SEG];      // float offset of block s inside the private AND result buffers (same layout)
  uint32_t unit0[MAX_SEG];
  uint32_t vec[MAX_SEG];
  const float *src_mc;        // multicast address of the private buffers (or null)
  float *dst_mc;              // multicast address of the result buffers (or null)
  const float *src_peer[8];
  float *dst_peer[8];
};

// One warp per 512-row chunk of the union marks.  The marked rows of the chunk are first compacted into a
// list in shared memory; the warp then walks the (row, 16-byte unit) pairs with FOUR independent
// multimem.ld_reduce in flight per lane before the matching stores, so the NVSwitch round trip (~3 us) is
// overlapped ~128-fold per warp instead of being paid once per row (the first version did one row at a time:
// 125-150 GB/s per GPU; profiles/r2_exchange_pull.txt).
__device__ __forceinline__ float4 pull_load_v4(const PullArgs &a, size_t e) {
  float4 v;
  if (a.src_mc) {
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a.src_mc + e));
  } else {
    v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < a.n_peers; ++r) {
      float4 t;
      asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(a.src_peer[r] + e));
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
  }
  return v;
}
__device__ __forceinline__ void pull_store_v4(const PullArgs &a, size_t e, float4 v) {
  if (a.dst_mc) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.dst_mc + e), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  } else {
    for (int r = 0; r < a.n_peers; ++r) *reinterpret_cast<float4 *>(a.dst_peer[r] + e) = v;
  }
}
__device__ __forceinline__ float pull_load_f32(const PullArgs &a, size_t e) {
  float v;
  if (a.src_mc) {
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(a.src_mc + e));
  } else {
    v = 0.f;
    for (int r = 0; r < a.n_peers; ++r) {
      float t;
      asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(t) : "l"(a.src_peer[r] + e));
      v += t;
    }
  }
  return v;
}
__device__ __forceinline__ void pull_store_f32(const PullArgs &a, size_t e, float v) {
  if (a.dst_mc) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(a.dst_mc + e), "f"(v) : "memory");
  } else {
    for (int r = 0; r < a.n_peers; ++r) a.dst_peer[r][e] = v;
  }
}

constexpr int PULL_MLP = 4;   // independent loads per lane
constexpr int PULL_MAX_UNITS = 64;

__global__ void __launch_bounds__(256)
rows_pull_kernel(const PullArgs a) {
  __shared__ uint16_t s_rows[8][512];
  __shared__ uint32_t s_unit[PULL_MAX_UNITS];  // per unit: segment | offset-in-row << 8 | vec << 31
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int u = threadIdx.x; u < a.n_units; u += blockDim.x) {
    int sg = 0;
    while (sg + 1 < a.n_seg && (uint32_t)u >= a.unit0[sg + 1]) ++sg;
    const uint32_t k = u - a.unit0[sg];
    s_unit[u] = (uint32_t)sg | ((a.vec[sg] ? 4 * k : k) << 8) | (a.vec[sg] ? 0x80000000u : 0u);
  }
  __syncthreads();
  const uint32_t chunk = (blockIdx.x * 8 + wib) * (uint32_t)a.n_peers + (uint32_t)a.rank;
  if ((size_t)chunk * 512 >= a.N) return;
  // ---- compact the marked rows of this chunk (lane owns 16 consecutive marks)
  const size_t c0 = (size_t)chunk * 512 + (size_t)lane * 16;
  uint32_t bits = 0;  // bit b: mark c0 + b is set
  if (c0 + 16 <= a.N && ((reinterpret_cast<uintptr_t>(a.union_marks) & 15) == 0)) {
    const uint4 v = *reinterpret_cast<const uint4 *>(a.union_marks + c0);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if ((w[q] >> (8 * b)) & 0xffu) bits |= 1u << (4 * q + b);
  } else {
    for (int b = 0; b < 16; ++b)
      if (c0 + b < a.N && a.union_marks[c0 + b]) bits |= 1u << b;
  }
  uint32_t cnt = __popc(bits), incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  const uint32_t n_rows = __shfl_sync(0xffffffffu, incl, 31);
  uint32_t pos = incl - cnt;
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    s_rows[wib][pos++] = (uint16_t)(lane * 16 + b);
  }
  __syncwarp();
  if (n_rows == 0) return;
  const size_t row0 = (size_t)chunk * 512;
  const uint32_t total = n_rows * (uint32_t)a.n_units;
  for (uint32_t i0 = 0; i0 < total; i0 += 32 * PULL_MLP) {
    float4 v[PULL_MLP];
    size_t e[PULL_MLP];
    uint32_t kind[PULL_MLP];  // 0 = none, 1 = scalar, 2 = vector
#pragma unroll
    for (int m = 0; m < PULL_MLP; ++m) {
      const uint32_t i = i0 + 32 * m + lane;
      kind[m] = 0;
      if (i < total) {
        const uint32_t r = i / (uint32_t)a.n_units, u = i - r * (uint32_t)a.n_units;
        const uint32_t info = s_unit[u];
        const uint32_t sg = info & 0xffu, k = (info >> 8) & 0x7fffffu;
        const size_t g = row0 + s_rows[wib][r];
        e[m] = a.off[sg] + g * a.width[sg] + k;
        if (info >> 31) { kind[m] = 2; v[m] = pull_load_v4(a, e[m]); }
        else { kind[m] = 1; v[m].x = pull_load_f32(a, e[m]); }
      }
    }
#pragma unroll
    for (int m = 0; m < PULL_MLP; ++m) {
      if (kind[m] == 2) pull_store_v4(a, e[m], v[m]);
      else if (kind[m] == 1) pull_store_f32(a, e[m], v[m].x);
    }
  }
}

struct MarkArgs {
  const uint8_t *marks;
  uint32_t N;
  int n_peers;
  uint8_t *union_peer[8];
  uint8_t *union_mc;  // NVSwitch multicast address of the union marks (or null)
};

// Multicast form: one thread per 16 marks; every non-zero 4-mark word is OR-ed into ALL ranks' union
// marks by the switch with a single multimem.red.or.b32 -- 1/8 of the packets of the per-peer byte
// stores at 8 ranks, and 4 marks per packet.
__global__ void __launch_bounds__(256)
marks_broadcast_mc_kernel(const MarkArgs a) {
  const size_t c0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (c0 + 16 > a.N) {  // tail (< 16 marks): per-peer byte stores
    for (size_t g = c0; g < a.N; ++g)
      if (a.marks[g])
        for (int r = 0; r < a.n_peers; ++r) a.union_peer[r][g] = 1;
    return;
  }
  const uint4 v = *reinterpret_cast<const uint4 *>(a.marks + c0);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (w[q])
      asm volatile("multimem.red.relaxed.sys.global.or.b32 [%0], %1;" ::"l"(a.union_mc + c0 + 4 * q), "r"(w[q])
                   : "memory");
}

__global__ void __launch_bounds__(256)
marks_broadcast_kernel(const MarkArgs a) {
  const int lane = threadIdx.x & 31;
  for_each_marked(const_cast<uint8_t *>(a.marks), a.N, 0, [&](size_t g) {
    if (lane < a.n_peers) a.union_peer[lane][g] = 1;
  });
}

struct ZeroArgs {
  uint8_t *marks;
  uint32_t N;
  int n_seg, clear, n_units;
  float *base[MAX_SEG];
  uint32_t width[MAX_SEG];
  uint32_t unit0[MAX_SEG];  // first store unit of the block: a unit is one float4 (vec) or one float
  int vec[MAX_SEG];
};

// zero the rows of every marked Gaussian in all blocks, then (clear != 0) the marks themselves
__global__ void __launch_bounds__(256)
rows_zero_marked_kernel(const ZeroArgs a) {
  // one store unit per lane: a row of all blocks (cfg 2: 48 + 3 + 4 + 3 + 1 floats, + 2 + 4 + 1 of the 2-D gradient
  // scratch) is 24 units -- one store instruction per marked row instead of one per block and 32 floats
  const int lane = threadIdx.x & 31;
  for_each_marked(a.marks, a.N, a.clear, [&](size_t g) {
    for (int u = lane; u < a.n_units; u += 32) {
      int s = 0;
      while (s + 1 < a.n_seg && (uint32_t)u >= a.unit0[s + 1]) ++s;
      const uint32_t k = u - a.unit0[s];
      if (a.vec[s]) *reinterpret_cast<float4 *>(a.base[s] + g * a.width[s] + 4 * k) = make_float4(0.f, 0.f, 0.f, 0.f);
      else a.base[s][g * a.width[s] + k] = 0.0f;
    }
  });
}

}  // namespace gs3d

extern "C" {

int gs3d_rows_push_marked(const uint8_t *marks, uint32_t N, int n_blocks, const uint64_t *src_ptrs_host,
                          const uint32_t *block_widths_host, const uint64_t *dst_offsets_host,
                          const uint64_t *peer_result_ptrs_host, const uint64_t *peer_union_ptrs_host,
                          int n_peers, void *multicast_result, void *stream) {
  using namespace gs3d;
  GS3D_REQUIRE(n_blocks >= 1 && n_blocks <= MAX_SEG, GS3D_EINVAL, "rows_push: n_blocks must be 1..%d", MAX_SEG);
  GS3D_REQUIRE(n_peers >= 1 && n_peers <= 8, GS3D_EINVAL, "rows_push: n_peers must be 1..8 (got %d)", n_peers);
  if (N == 0) return GS3D_OK;
  GS3D_REQUIRE(marks && src_ptrs_host && block_widths_host && dst_offsets_host && peer_result_ptrs_host,
               GS3D_EINVAL, "rows_push: null argument");
  PushArgs a = {};
  a.marks = marks; a.N = N; a.n_seg = n_blocks; a.n_peers = n_peers;
  a.dst_mc = static_cast<float *>(multicast_result);
  bool base_aligned = (reinterpret_cast<uintptr_t>(multicast_result) & 15) == 0;
  for (int r = 0; r < n_peers; ++r) {
    a.dst_peer[r] = reinterpret_cast<float *>(static_cast<uintptr_t>(peer_result_ptrs_host[r]));
    a.union_peer[r] = peer_union_ptrs_host
                          ? reinterpret_cast<uint8_t *>(static_cast<uintptr_t>(peer_union_ptrs_host[r])) : nullptr;
    GS3D_REQUIRE(a.dst_peer[r], GS3D_EINVAL, "rows_push: null peer pointer %d", r);
    base_aligned = base_aligned && (peer_result_ptrs_host[r] & 15) == 0;
  }
  uint32_t units = 0;
  for (int s = 0; s < n_blocks; ++s) {
    a.src[s] = reinterpret_cast<const float *>(static_cast<uintptr_t>(src_ptrs_host[s]));
    a.width[s] = block_widths_host[s];
    a.off[s] = dst_offsets_host[s];
    GS3D_REQUIRE(a.src[s] && a.width[s] > 0, GS3D_EINVAL, "rows_push: bad block %d", s);
    a.vec[s] = (base_aligned && a.width[s] % 4 == 0 && a.off[s] % 4 == 0 && (src_ptrs_host[s] & 15) == 0) ? 1 : 0;
    a.unit0[s] = units;
    units += a.vec[s] ? a.width[s] / 4 : a.width[s];
  }
  a.n_units = (int)units;
  rows_push_kernel<<<div_up(N, 4096u), 256, 0, as_stream(stream)>>>(a);  // 512 marks per warp
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

int gs3d_marks_broadcast(const uint8_t *marks, uint32_t N, const uint64_t *peer_union_ptrs_host, int n_peers,
                         void *multicast_union, void *stream) {
  using namespace gs3d;
  GS3D_REQUIRE(n_peers >= 1 && n_peers <= 8, GS3D_EINVAL, "marks_broadcast: n_peers must be 1..8 (got %d)", n_peers);
  if (N == 0) return GS3D_OK;
  GS3D_REQUIRE(marks && peer_union_ptrs_host, GS3D_EINVAL, "marks_broadcast: null argument");
  MarkArgs a = {};
  a.marks = marks; a.N = N; a.n_peers = n_peers;
  for (int r = 0; r < n_peers; ++r) {
    a.union_peer[r] = reinterpret_cast<uint8_t *>(static_cast<uintptr_t>(peer_union_ptrs_host[r]));
    GS3D_REQUIRE(a.union_peer[r], GS3D_EINVAL, "marks_broadcast: null peer pointer %d", r);
  }
  a.union_mc = static_cast<uint8_t *>(multicast_union);
  if (a.union_mc && (reinterpret_cast<uintptr_t>(marks) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.union_mc) & 15) == 0)
    marks_broadcast_mc_kernel<<<div_up(div_up(N, 16u), 256u), 256, 0, as_stream(stream)>>>(a);
  else
    marks_broadcast_kernel<<<div_up(N, 4096u), 256, 0, as_stream(stream)>>>(a);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

int gs3d_rows_pull_marked(uint8_t *union_marks, uint32_t N, int n_blocks, const uint32_t *block_widths_host,
                          const uint64_t *block_offsets_host, const uint64_t *peer_private_ptrs_host,
                          const uint64_t *peer_result_ptrs_host, int n_peers, int rank,
                          const void *multicast_private, void *multicast_result, void *stream) {
  using namespace gs3d;
  GS3D_REQUIRE(n_blocks >= 1 && n_blocks <= MAX_SEG, GS3D_EINVAL, "rows_pull: n_blocks must be 1..%d", MAX_SEG);
  GS3D_REQUIRE(n_peers >= 1 && n_peers <= 8 && rank >= 0 && rank < n_peers, GS3D_EINVAL,
               "rows_pull: bad n_peers / rank (%d / %d)", n_peers, rank);
  if (N == 0) return GS3D_OK;
  GS3D_REQUIRE(union_marks && block_widths_host && block_offsets_host && peer_private_ptrs_host &&
                   peer_result_ptrs_host, GS3D_EINVAL, "rows_pull: null argument");
  PullArgs a = {};
  a.union_marks = union_marks; a.N = N; a.n_seg = n_blocks; a.n_peers = n_peers; a.rank = rank;
  a.src_mc = static_cast<const float *>(multicast_private);
  a.dst_mc = static_cast<float *>(multicast_result);
  bool aligned = ((reinterpret_cast<uintptr_t>(multicast_private) | reinterpret_cast<uintptr_t>(multicast_result)) & 15) == 0;
  for (int r = 0; r < n_peers; ++r) {
    a.src_peer[r] = reinterpret_cast<const float *>(static_cast<uintptr_t>(peer_private_ptrs_host[r]));
    a.dst_peer[r] = reinterpret_cast<float *>(static_cast<uintptr_t>(peer_result_ptrs_host[r]));
    GS3D_REQUIRE(a.src_peer[r] && a.dst_peer[r], GS3D_EINVAL, "rows_pull: null peer pointer %d", r);
    aligned = aligned && ((peer_private_ptrs_host[r] | peer_result_ptrs_host[r]) & 15) == 0;
  }
  uint32_t units = 0;
  for (int s = 0; s < n_blocks; ++s) {
    a.width[s] = block_widths_host[s];
    a.off[s] = block_offsets_host[s];
    GS3D_REQUIRE(a.width[s] > 0, GS3D_EINVAL, "rows_pull: bad block %d", s);
    a.vec[s] = (aligned && a.width[s] % 4 == 0 && a.off[s] % 4 == 0) ? 1 : 0;
    a.unit0[s] = units;
    units += a.vec[s] ? a.width[s] / 4 : a.width[s];
  }
  a.n_units = (int)units;
  GS3D_REQUIRE(units <= (uint32_t)PULL_MAX_UNITS, GS3D_EUNSUPPORTED, "rows_pull: %u units per row (max %d)", units,
               PULL_MAX_UNITS);
  const uint32_t chunks = div_up(N, 512u);
  rows_pull_kernel<<<div_up(div_up(chunks, (uint32_t)n_peers), 8u), 256, 0, as_stream(stream)>>>(a);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

int gs3d_rows_zero_marked(uint8_t *marks, uint32_t N, int n_blocks, const uint64_t *block_ptrs_host,
                          const uint32_t *block_widths_host, int clear_marks, void *stream) {
  using namespace gs3d;
  GS3D_REQUIRE(n_blocks >= 0 && n_blocks <= MAX_SEG, GS3D_EINVAL, "rows_zero: n_blocks must be 0..%d", MAX_SEG);
  if (N == 0) return GS3D_OK;
  GS3D_REQUIRE(marks && (n_blocks == 0 || (block_ptrs_host && block_widths_host)), GS3D_EINVAL,
               "rows_zero: null argument");
  ZeroArgs a = {};
  a.marks = marks; a.N = N; a.n_seg = n_blocks; a.clear = clear_marks;
  for (int s = 0; s < n_blocks; ++s) {
    a.base[s] = reinterpret_cast<float *>(static_cast<uintptr_t>(block_ptrs_host[s]));
    a.width[s] = block_widths_host[s];
    GS3D_REQUIRE(a.base[s] && a.width[s] > 0, GS3D_EINVAL, "rows_zero: bad block %d", s);
    a.vec[s] = (a.width[s] % 4 == 0 && (block_ptrs_host[s] & 15) == 0) ? 1 : 0;
    a.unit0[s] = (uint32_t)a.n_units;
    a.n_units += (int)(a.vec[s] ? a.width[s] / 4 : a.width[s]);
  }
  rows_zero_marked_kernel<<<div_up(N, 4096u), 256, 0, as_stream(stream)>>>(a);  // 512 marks per warp
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

int gs3d_rows_gather(int n_blocks, const uint64_t *block_ptrs_host, const uint32_t *block_widths_host,
                     const int32_t *row_idx, uint32_t U, float *packed, uint32_t W, void *stream) {
  return gs3d::rows_move(true, n_blocks, block_ptrs_host, block_widths_host, row_idx, U, packed, W, stream);
}

int gs3d_rows_scatter(int n_blocks, const uint64_t *block_ptrs_host, const uint32_t *block_widths_host,
                      const int32_t *row_idx, uint32_t U, const float *packed, uint32_t W, void *stream) {
  return gs3d::rows_move(false, n_blocks, block_ptrs_host, block_widths_host, row_idx, U,
                         const_cast<float *>(packed), W, stream);
}

}  // extern "C"
