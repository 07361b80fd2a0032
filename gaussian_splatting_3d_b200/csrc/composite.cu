// K3 / K4a: per-tile front-to-back alpha compositing with per-pixel SH colour, forward and
// backward.  Replaces tile_based_vol_rendering_sh_entry / ..._backward_sh_entry
// (vol_render_sh.h:97-248, 268-455) and their _with_bg variants (vol_render_bg.h).
//
// One CTA (256 threads = 16x16 pixels, warps own 8x4 pixel blocks) per tile.  Gaussians of the
// tile's list are staged in batches into shared memory with cp.async (LDGSTS), double-buffered:
// a 48-byte staging record per Gaussian (3 x 16 B) plus its 3*C*C SH coefficients, gathered by id.
// Per (pixel, Gaussian) pair the fast path costs ~10 FP32 instructions: a pre-scaled conic form
// gives log2(G) and the 1/255 skip test is a compare against a per-Gaussian threshold, so no
// exponential is evaluated for skipped pairs.  Decisions that fall within a small margin of the
// threshold are re-evaluated in the reference's exact FP32 operation order (kernels.h:172-193 as
// compiled: FMA contraction pattern taken from the reference's SASS) so that skip decisions --
// which move a pixel by up to 1/255 -- match the reference bit for bit.
// Warp-ballot early termination: a warp stops evaluating when all its pixels have T < thresh, the
// CTA stops staging when all warps are done.
//
// Backward: forward recompute in list order (the reference's own scheme, `final - prefix` suffix).
// Per contributing (warp, Gaussian) the 3*C*C + 6 partial sums over the warp's 32 pixels are
// reduced through a shared-memory transpose (w[c][pixel] x Y[pixel][k] as a tiny GEMV per lane)
// into warp-private accumulators, summed over the 8 warps per batch and flushed with one vector
// red.global.add.v4.f32 per 4 coefficients: ~14 global reductions per (tile, Gaussian) instead of
// the reference's 55 shared atomics per (pixel, Gaussian) + 55 global atomics per (tile, Gaussian).
#include <stdlib.h>

#include "common.cuh"

namespace gs3d {

constexpr int TILE = 16;
constexpr int NTHREADS = 256;
constexpr int NWARPS = 8;
constexpr float MIN_RENDER_ALPHA = 1 / 255.0f;  // common.h:90
constexpr float DECISION_MARGIN = 0.004f;       // log2 units around the skip threshold
constexpr float DEAD = 1e30f;                   // bias that makes a finished pixel fail the skip test
#ifndef GS3D_UNROLL_F
#define GS3D_UNROLL_F 4
#endif
#ifndef GS3D_UNROLL_B
#define GS3D_UNROLL_B 2
#endif
constexpr int UF = GS3D_UNROLL_F, UB = GS3D_UNROLL_B;  // inner-loop unroll (forward, backward); divide 4

struct CompositeParams {
  const float4 *records;
  const float *sh;
  uint32_t sh_sg, sh_sc;
  const int32_t *start, *end, *ids;
  float *out;
  const float *topleft, *c2w;
  uint32_t ntw, nth;
  float psx, psy;
  uint32_t H, W;
  float thresh;
  const float *bg;
  float *final_T;
  int32_t *n_contrib;
  int exact;
  // backward only
  const float *out_saved, *grad_out;
  float *g_mean, *g_cov, *g_sh, *g_alpha;
  uint32_t gsh_sg, gsh_sc;
  int sh_vec;   // SH rows are 16-byte addressable (staging)
  int gsh_vec;  // grad SH rows are 16-byte addressable (vector reductions)
  // fused gradient exchange (view-sharded training): the SH gradient rows are reduced straight into
  // every rank's buffer -- through one multimem.red on an NVSwitch multicast address, or peer by peer
  // over NVLink -- instead of locally followed by a dense all-reduce.  n_peers == 0: local only.
  float *g_sh_peer[8];
  int n_peers;
  float *g_sh_mc;
  uint8_t *touched;  // optional: 1 for every Gaussian whose gradient rows this launch wrote
  unsigned long long *stats;  // optional (gs3d_set_stage_counters): [0] += staged duplicates (fwd), [1] (bwd)
};

// ---------------------------------------------------------------- small device helpers

__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_4(void *smem, const void *gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void red_add_v4(float *addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void multimem_red_add_v4(float *addr, float4 v) {
  asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid(s) = 1 / (1 + 2^(-s log2 e)); two MUFU ops, ~2 ulp (tolerance: images 1e-4)
__device__ __forceinline__ float fast_sigmoid(float s) {
  return rcp_approx(1.0f + ex2_approx(s * -1.4426950408889634f));
}
// Packed FP32 pairs (Blackwell FFMA2: one issue slot, two FMAs).  The compositing kernels are
// issue-bound (ncu: 70-80 % issue-slot utilisation, FMA pipe < 50 %), so the SH dot products and the
// backward's per-warp GEMV run on fma.rn.f32x2 with operands that come out of LDS.128 as register pairs.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float sum2(f32x2 v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ void lds128_2(uint32_t a, f32x2 &lo, f32x2 &hi) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(a));
}
// 32-bit shared-window addresses kept in registers: the hot loops address staging buffers with
// plain adds instead of re-deriving generic pointers every iteration.
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}

// shencoder.h:13-55, per-pixel basis (quirk Q2)
template <int C>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float *o) {
  o[0] = 0.28209479177387814f;
  if constexpr (C > 1) {
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
  }
  if constexpr (C > 2) {
    float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    if constexpr (C > 3) {
      o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
      o[10] = 2.8906114426405538f * xy * z;
      o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
      o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
      o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
      o[14] = 1.4453057213202769f * z * (x2 - y2);
      o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
    }
  }
}

// vol_render_sh.h:48-65 + :213-217: direction from the first nine floats of the [3,4] c2w (Q1).
template <int C>
__device__ __forceinline__ void pixel_basis(const float *c2w, float px, float py, float *Y) {
  float d0 = c2w[0] * px + c2w[1] * py + c2w[2] * 1.0f;
  float d1 = c2w[3] * px + c2w[4] * py + c2w[5] * 1.0f;
  float d2 = c2w[6] * px + c2w[7] * py + c2w[8] * 1.0f;
  float len = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
  sh_basis<C>(d0 / len, d1 / len, d2 / len, Y);
}

// The reference's Gaussian in its compiled operation order (kernels.h:172-193; SASS of
// tile_based_vol_rendering_sh_entry<4>, sm_100a, nvcc 12.9):
//   det  = fma(c0, c3, -(c2*c1));  tmpy = fma(c0, y, -(x*c1));  tmpx = fma(x, c3, -(c2*y));
//   num  = fma(x, tmpx, y*tmpy);   radial = num / det (IEEE);  arg = radial >= 0 ? -0.5*radial : -500
__device__ __forceinline__ float gaussian_exact(float x, float y, float4 cv) {
  float det = __fmaf_rn(cv.x, cv.w, -__fmul_rn(cv.z, cv.y));
  float tmpy = __fmaf_rn(cv.x, y, -__fmul_rn(x, cv.y));
  float tmpx = __fmaf_rn(x, cv.w, -__fmul_rn(cv.z, y));
  float num = __fmaf_rn(x, tmpx, __fmul_rn(y, tmpy));
  float radial = __fdiv_rn(num, det);
  float arg = (radial < 0.0f) ? -500.0f : __fmul_rn(radial, -0.5f);
  return expf(arg);
}

// The legacy RGB path evaluates the Gaussian in FP64 (kernel_gaussian_2d, kernels.h:195-214): every operand
// is widened first, so the only roundings are the double ones and the final cast of exp().
__device__ __forceinline__ float gaussian_exact_f64(float x, float y, float4 cv) {
  const double c0 = cv.x, c1 = cv.y, c2 = cv.z, c3 = cv.w;
  const double det = c0 * c3 - c1 * c2;
  const double dx = (double)x, dy = (double)y;  // query - mean is an FP32 subtraction promoted afterwards
  const double tmpx = dx * c3 - dy * c2;
  const double tmpy = -dx * c1 + dy * c0;
  double radial = (tmpx * dx + tmpy * dy) / det;
  if (radial < 0.0) radial = 1000.0;
  return (float)exp(-0.5 * radial);
}

// Skip decision for one (pixel, Gaussian) pair (pair_test + pair_decide).  r0 = {m.x, m.y, alpha_, log2 threshold},
// r1 = {qa, qb, qc, depth}.  Returns true when the pair contributes (alpha_*G >= 1/255) and then
// G is valid.  Far from the threshold the pre-scaled conic decides alone (no exponential for
// skipped pairs); within DECISION_MARGIN (or for a non-negative exponent, where the reference's
// `radial < 0 -> 1000` rule matters) the reference's exact arithmetic decides.
// Part 1: log2 G and its distance to the per-Gaussian threshold.
__device__ __forceinline__ float4 pair_test(float px, float py, uint32_t rec_addr, float &pw, float &diff) {
  const float4 r0 = lds128(rec_addr), r1 = lds128(rec_addr + 16);
  const float dx = px - r0.x, dy = py - r0.y;
  const float u = fmaf(r1.y, dy, r1.x * dx);
  pw = fmaf(r1.z * dy, dy, dx * u);  // log2 G
  diff = pw - r0.w;
  return r0;
}
// Part 2, only reached when part 1 did not clearly reject: decide, and produce G.
template <bool EXACT, bool RGB = false>
__device__ __forceinline__ bool pair_decide(float px, float py, float dead, float pw, float diff,
                                            const float4 r0, uint32_t cov_addr, float &G) {
  if (dead != 0.0f) return false;                // finished pixel that slipped through (non-finite record)
  if (diff >= DECISION_MARGIN && pw <= -1e-5f) {
    G = ex2_approx(pw);
    return true;
  }
  if (EXACT) {
    const float val = RGB ? gaussian_exact_f64(px - r0.x, py - r0.y, lds128(cov_addr))
                          : gaussian_exact(px - r0.x, py - r0.y, lds128(cov_addr));
    G = val;
    return !(r0.z * val < MIN_RENDER_ALPHA);
  }
  if (!(diff >= 0.0f) || pw > 0.0f) return false;
  G = ex2_approx(pw);
  return true;
}

// Stage `nb` Gaussians (ids already in s_ids) into shared memory with cp.async.
template <int CC, int B, int NT = NTHREADS>
__device__ __forceinline__ void stage_batch(const CompositeParams &p, const int *s_ids, int nb,
                                            float4 *s_rec, float *s_sh) {
  constexpr int SHF = 3 * CC;
  for (int e = threadIdx.x; e < nb * 3; e += NT) {
    int j = e / 3, r = e - 3 * j;
    cp_async_16(s_rec + e, p.records + 3 * (size_t)s_ids[j] + r);
  }
  // the compositing loops are unrolled by up to 4: pad with null records (threshold +inf: never
  // contributes) so they need no remainder handling
  if (threadIdx.x >= NT - 9) {
    const int e = nb * 3 + (NT - 1 - threadIdx.x);
    if (e < ((nb + 3) & ~3) * 3)
      s_rec[e] = (e % 3 == 0) ? make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (p.sh_vec) {
    constexpr int V = SHF / 4 > 0 ? SHF / 4 : 1;
    for (int e = threadIdx.x; e < nb * V; e += NT) {
      int j = e / V, r = (e - V * j) * 4;
      int c = r / CC, k = r - c * CC;
      cp_async_16(s_sh + j * SHF + r, p.sh + (size_t)s_ids[j] * p.sh_sg + c * p.sh_sc + k);
    }
  } else {
    for (int e = threadIdx.x; e < nb * SHF; e += NT) {
      int j = e / SHF, r = e - SHF * j;
      int c = r / CC, k = r - c * CC;
      cp_async_4(s_sh + e, p.sh + (size_t)s_ids[j] * p.sh_sg + c * p.sh_sc + k);
    }
  }
}

// Per-pixel SH basis held in registers: packed pairs when C*C is a multiple of 4 (C = 2, 4),
// scalars otherwise (C = 1, 3).
template <int CC>
struct Basis {
  static constexpr bool PACKED = (CC % 4 == 0);
  static constexpr int NP = PACKED ? CC / 2 : 1;
  static constexpr int NS = PACKED ? 1 : CC;
  f32x2 p[NP];
  float s[NS];
  __device__ __forceinline__ void set(const float *Y) {
    if constexpr (PACKED) {
#pragma unroll
      for (int k = 0; k < CC / 2; ++k) p[k] = pack2(Y[2 * k], Y[2 * k + 1]);
    } else {
#pragma unroll
      for (int k = 0; k < CC; ++k) s[k] = Y[k];
    }
  }
};

#ifndef GS3D_ABLATE
#define GS3D_ABLATE 0  // experiment switches (tools/ablate.sh); 0 in every shipped build
#endif
template <int CC, bool RGB = false>
__device__ __forceinline__ void sh_colour(uint32_t h_addr, const Basis<CC> &Y, float *y) {
  if constexpr (RGB) {  // legacy path: the staged row IS the colour (vol_render.h:150-152)
    y[0] = lds32(h_addr); y[1] = lds32(h_addr + 4); y[2] = lds32(h_addr + 8);
  } else {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float s;
    if constexpr (Basis<CC>::PACKED) {
      f32x2 acc = 0ull;
#if GS3D_ABLATE & 2   // no SH row loads: one LDS.128 per channel instead of CC/4
      f32x2 lo, hi;
      lds128_2(h_addr + 4 * (c * CC), lo, hi);
#endif
#pragma unroll
      for (int k = 0; k < CC / 4; ++k) {
#if !(GS3D_ABLATE & 2)
        f32x2 lo, hi;
        lds128_2(h_addr + 4 * (c * CC + 4 * k), lo, hi);
#endif
#if GS3D_ABLATE & 4   // a quarter of the FMAs
        if (k > 0) continue;
#endif
        acc = fma2(lo, Y.p[2 * k], acc);
        acc = fma2(hi, Y.p[2 * k + 1], acc);
      }
      s = sum2(acc);
    } else {
      s = 0.0f;
#pragma unroll
      for (int k = 0; k < CC; ++k) s = fmaf(lds32(h_addr + 4 * (c * CC + k)), Y.s[k], s);
    }
#if GS3D_ABLATE & 1     // no MUFU in the sigmoid
    float v = fminf(fmaxf(fmaf(s, 0.05f, 0.5f), 0.0f), 1.0f);
#else
    float v = fast_sigmoid(s);
#endif
    // vol_render_sh.h:151-159 zeroes a colour whose product with the (finite, NaN-guarded) weight is
    // NaN; the sigmoid is in [0,1] or NaN, so that is exactly "v is NaN"
    if (isnan(v)) v = 0.0f;
    y[c] = v;
  }
  }
}

// ---------------------------------------------------------------- forward
//
// Pipeline (one barrier per batch): the ids of batch b+2 are fetched into a register and published
// to a three-deep id ring while batch b+1 is in flight (cp.async) and batch b is composited.  The
// single __syncthreads_and per batch (i) makes batch b's staged rows visible, (ii) releases the
// buffer batch b-1 used, (iii) publishes the id ring slot and (iv) carries the "every pixel of the
// tile is saturated" vote that ends the tile early.

#ifndef GS3D_FWD_MINB
#define GS3D_FWD_MINB 4  // CTAs per SM the forward is compiled for (64 registers)
#endif
template <int C, int B, bool EXACT, bool RGB = false, bool STATS = false>
__global__ void __launch_bounds__(NTHREADS, GS3D_FWD_MINB)
composite_fwd_kernel(const CompositeParams p) {
  constexpr int CC = C * C;
  constexpr int SHF = 3 * CC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4 *s_rec = reinterpret_cast<float4 *>(smem_raw);        // [2][B][3]
  float *s_sh = reinterpret_cast<float *>(s_rec + 2 * B * 3);  // [2][B][SHF]
  int *s_ids = reinterpret_cast<int *>(s_sh + 2 * B * SHF);    // [3][B]

  const int tile_id = blockIdx.x;
  const int tile_y = tile_id / p.ntw, tile_x = tile_id - tile_y * p.ntw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lx = 8 * (warp & 1) + (lane & 7), ly = 4 * (warp >> 1) + (lane >> 3);
  const uint32_t gx = tile_x * TILE + lx, gy = tile_y * TILE + ly;
  const bool inside = gx < p.W && gy < p.H;
  const size_t pix = (size_t)gy * p.W + gx;

  const int first = p.start[tile_id];
  const int n_this = first == -1 ? 0 : p.end[tile_id] - first;
  if (first == -1) {  // vol_render_sh.h:185-188 / vol_render_bg.h:28-37
    if (inside) {
      if (p.bg) {
        p.out[3 * pix + 0] = p.bg[0];
        p.out[3 * pix + 1] = p.bg[1];
        p.out[3 * pix + 2] = p.bg[2];
      }
      if (p.final_T) p.final_T[pix] = 1.0f;
      if (p.n_contrib) p.n_contrib[pix] = 0;
    }
    return;
  }
  if (n_this <= 0) return;

  // pixel corner in camera-plane units (quirk Q4), same expression as vol_render_sh.h:213-214
  const float px = p.topleft[0] + gx * p.psx, py = p.topleft[1] + gy * p.psy;
  Basis<CC> Y;
  if constexpr (!RGB) {
    float Yf[CC];
    pixel_basis<C>(p.c2w, px, py, Yf);
    Y.set(Yf);
  }

  float T = 1.0f, o0 = 0.0f, o1 = 0.0f, o2 = 0.0f;
  int last = 0;
  [[maybe_unused]] unsigned int n_pairs = 0;  // STATS: contributing (pixel, Gaussian) pairs of this lane
  const float thresh = p.thresh;
  // the reference tests T (initially 1) before each Gaussian
  float dead = (inside && !(1.0f < thresh)) ? 0.0f : DEAD;
  const int32_t *ids = p.ids + first;
  const int n_batches = (n_this + B - 1) / B;
  const bool id_lane = threadIdx.x < B;

  // prologue: ids of batch 0 -> ring slot 0, batch 0 in flight, ids of batch 1 -> ring slot 1
  if (id_lane) s_ids[threadIdx.x] = threadIdx.x < n_this ? ids[threadIdx.x] : 0;
  int my_id = (id_lane && B + threadIdx.x < n_this) ? ids[B + threadIdx.x] : 0;
  __syncthreads();
  stage_batch<CC, B>(p, s_ids, min(B, n_this), s_rec, s_sh);
  cp_async_commit();
  if (id_lane) s_ids[B + threadIdx.x] = my_id;
  my_id = (id_lane && 2 * B + threadIdx.x < n_this) ? ids[2 * B + threadIdx.x] : 0;

  int cb = 0;
  for (; cb < n_batches; ++cb) {
    cp_async_wait<0>();
    if (__syncthreads_and(dead != 0.0f)) break;  // all pixels of the tile finished -> stop staging
    const int buf = cb & 1;
    if (cb + 1 < n_batches) {
      const int nbuf = buf ^ 1;
      stage_batch<CC, B>(p, s_ids + ((cb + 1) % 3) * B, min(B, n_this - (cb + 1) * B), s_rec + nbuf * B * 3,
                         s_sh + nbuf * B * SHF);
      cp_async_commit();
      if (id_lane) s_ids[((cb + 2) % 3) * B + threadIdx.x] = my_id;
      const int nxt = (cb + 3) * B + threadIdx.x;
      my_id = (id_lane && nxt < n_this) ? ids[nxt] : 0;
    }
    const int nb = min(B, n_this - cb * B);
    uint32_t rec_a = smem_u32(s_rec + buf * B * 3);
    uint32_t sh_a = smem_u32(s_sh + buf * B * SHF);
    for (int j = 0; j < nb; j += UF, rec_a += 48 * UF, sh_a += 4 * SHF * UF) {
      if (!__any_sync(0xffffffffu, dead == 0.0f)) break;
#pragma unroll
      for (int u = 0; u < UF; ++u) {
        // (testing the whole unroll group up front for ILP measured slower: 0.502 vs 0.492 ms)
        float pw, df;
        const float4 r0 = pair_test(px, py, rec_a + 48 * u, pw, df);
        // `dead` is 0 for a live pixel and 1e30 for a finished one (T < thresh, or outside the
        // image): finished pixels fail the common-case test without a branch of their own
        if (df - dead < -DECISION_MARGIN) continue;  // the common case: clearly below 1/255
        float G;
        if (pair_decide<EXACT, RGB>(px, py, dead, pw, df, r0, rec_a + 48 * u + 32, G)) {
          const float a = r0.z;
          float coeff = (a * T) * G;
          if (!RGB && isnan(coeff)) coeff = 0.0f;  // vol_render_sh.h:147-150 (the RGB path has no guard)
          float y[3];
          sh_colour<CC, RGB>(sh_a + 4 * SHF * u, Y, y);
          o0 = fmaf(coeff, y[0], o0);
          o1 = fmaf(coeff, y[1], o1);
          o2 = fmaf(coeff, y[2], o2);
          T *= (1 - a * G);
          if constexpr (STATS) ++n_pairs;
          last = cb * B + j + u + 1;
          // vol_render_sh.h:121-123 tests T before each Gaussian; T only changes here
          if (T < thresh) dead = DEAD;
        }
      }
    }
  }
  cp_async_wait<0>();
  // batches 0..cb have been staged when the loop ends at cb (all of them when it ran to completion)
  if (p.stats && threadIdx.x == 0) atomicAdd(p.stats, (unsigned long long)min(n_this, (cb + 1) * B));
  if constexpr (STATS) {
    const unsigned int wsum = __reduce_add_sync(0xffffffffu, n_pairs);
    if (p.stats && lane == 0 && wsum) atomicAdd(p.stats + 2, (unsigned long long)wsum);
  }
  if (!inside) return;
  if (p.bg && T > p.thresh) {  // vol_render_bg.h:90-94
    o0 = o0 * T + p.bg[0] * (1.0f - T);
    o1 = o1 * T + p.bg[1] * (1.0f - T);
    o2 = o2 * T + p.bg[2] * (1.0f - T);
  }
  p.out[3 * pix + 0] = o0;
  p.out[3 * pix + 1] = o1;
  p.out[3 * pix + 2] = o2;
  if (p.final_T) p.final_T[pix] = T;
  if (p.n_contrib) p.n_contrib[pix] = last;
}

// ---------------------------------------------------------------- backward

constexpr int WROW = 36;  // floats per row of the per-warp exchange tile (32 pixels + 4 pad: the
                          // scalar-sum LDS.128 of rows v and v+1 then fall on different banks)

// One float4 (quad q of a Gaussian's [3*CC + 6]-float gradient row, padded to a multiple of 4) -> global memory.
template <int CC>
__device__ __forceinline__ void emit_quad(const CompositeParams &p, size_t g, int q, float4 s) {
  constexpr int SHF = 3 * CC;
  const float sv4[4] = {s.x, s.y, s.z, s.w};
  const int r0 = 4 * q;
  if (CC % 4 == 0 && p.gsh_vec && r0 + 3 < SHF) {
    const int c = r0 / CC, k = r0 - c * CC;
    const size_t off = g * p.gsh_sg + c * p.gsh_sc + k;
    if (p.g_sh_mc) {
      multimem_red_add_v4(p.g_sh_mc + off, s);  // one instruction, the switch adds it on every GPU
    } else if (p.n_peers > 0) {
      for (int r = 0; r < p.n_peers; ++r) red_add_v4(p.g_sh_peer[r] + off, s);
    } else {
      red_add_v4(p.g_sh + off, s);
    }
    return;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = r0 + u;
    const float val = sv4[u];
    if (val == 0.f) continue;
    if (r < SHF) {
      const int c = r / CC, k = r - c * CC;
      const size_t off = g * p.gsh_sg + c * p.gsh_sc + k;
      if (p.n_peers > 0) {
        for (int q2 = 0; q2 < p.n_peers; ++q2) atomicAdd(p.g_sh_peer[q2] + off, val);
      } else {
        atomicAdd(p.g_sh + off, val);
      }
    } else {
      const int v = r - SHF;
      if (v == 0) atomicAdd(p.g_mean + 2 * g, val);
      else if (v == 1) atomicAdd(p.g_mean + 2 * g + 1, val);
      else if (v == 2) atomicAdd(p.g_cov + 4 * g, val);
      else if (v == 3) { atomicAdd(p.g_cov + 4 * g + 1, val); atomicAdd(p.g_cov + 4 * g + 2, val); }
      else if (v == 4) atomicAdd(p.g_cov + 4 * g + 3, val);
      else if (v == 5) atomicAdd(p.g_alpha + g, val);
    }
  }
}

// ---------------------------------------------------------------- backward kernel: mbarrier ring
//
// History (profiles/r2_bwd_variants.jsonl): the first version synchronised the CTA once per batch of 16 Gaussians and
// branched per lane; ncu showed ~22 % of all warp samples waiting at that barrier.  Work per
// (warp, Gaussian) is bimodal (~220 instructions when any of the warp's 32 pixels contributes, ~25 otherwise),
// so at every batch of 16 the CTA waits for its unluckiest warp.  Here nothing is CTA-synchronous:
//
//   * a ring of RS stages of RG Gaussians; each stage holds the staged rows (record + SH) AND the warp-private
//     accumulator rows of its Gaussians;
//   * warp NW (the 9th) is producer and flusher: it stages a batch with cp.async.bulk (one 48-byte record copy
//     and one SH-row copy per Gaussian, completion counted in bytes on the stage's `full` mbarrier), and when
//     all NW pixel warps have arrived on the stage's `done` mbarrier it sums their accumulator rows, issues the
//     global vector reductions and re-arms the stage with the batch RS stages ahead;
//   * a pixel warp only ever waits for `full` of its next stage, i.e. it can run up to RS-1 stages ahead of the
//     slowest warp of its tile; rows it did not write are skipped through a per-(stage, warp) bit mask, so
//     nothing is zero-filled.
// Early termination: every warp publishes "some pixel still alive" with its arrival; when no warp is alive the
// producer arms the next stage as a sentinel (0 Gaussians) and the warps leave.

#ifndef GS3D_RING_STAGES
#define GS3D_RING_STAGES 4
#endif
#ifndef GS3D_RING_GAUSSIANS
#define GS3D_RING_GAUSSIANS 8
#endif
constexpr int RS = GS3D_RING_STAGES;     // ring stages (power of two)
constexpr int RG = GS3D_RING_GAUSSIANS;  // Gaussians per stage (<= 8: one mask byte per warp and stage)

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes)
               : "memory");
}
// Blocking wait for the phase with the given parity.  One try_wait (the common case: the data is already there),
// then a short sleep between probes.  History (profiles/): a bare try_wait loop spent 17 % of all issued
// instructions probing; try_wait with a suspend-time hint compiles to TRYWAIT + NANOSLEEP.SYNCS, which wakes on
// every mbarrier event of the SM.  A warp that has to wait is waiting for a slower warp of its tile, i.e. for
// microseconds; A/B of 100..1000 ns sleeps and of the hint form gave the same kernel time (the probes of a blocked
// warp only use issue slots nobody else wants), so the simplest form stays.
#ifndef GS3D_WAIT_NS
#define GS3D_WAIT_NS 200u
#endif
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t sleep_ns = GS3D_WAIT_NS) {
  while (!mbar_try(bar, parity)) __nanosleep(sleep_ns);
}
// Arrival on `bar` that fires when all cp.async copies this thread has issued so far have landed (the pending count is
// NOT incremented: the barrier must have been initialised with this arrival counted in)
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine, UBLKCP); bytes is a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// Per-lane state of the backward walk and one step of it (one Gaussian for the warp's 32 pixels).
template <int C, bool EXACT, bool RGB, bool STATS = false>
struct BwdWalk {
  static constexpr int CC = C * C;
  static constexpr int SHF = 3 * CC;
  static constexpr int KL = CC < 16 ? CC : 16;
  static constexpr uint32_t FULL = 0xffffffffu;
  static constexpr float INV_K = 1.0f / (-0.5f * 1.4426950408889634f);  // undo the conic pre-scale
  float px, py, g0, g1, g2, f0, f1, f2, T, dead, thresh;
  Basis<CC> Y;
  f32x2 Yt[8];
  uint32_t w_st, w_sh, w_sc;
  int lane, sv_row, sv_q;
  unsigned int n_pairs;  // STATS: contributing (pixel, Gaussian) pairs of this lane

  // rec / sh: shared addresses of the staged record and SH row; row: shared address of this warp's accumulator
  // row for the Gaussian.  Returns (warp-uniform) whether the row was written.
  __device__ __forceinline__ bool step(uint32_t rec, uint32_t sh, uint32_t row) {
    float pw, df;
    const float4 r1 = lds128(rec + 16);
    const float4 r0 = pair_test(px, py, rec, pw, df);
    const bool cand = !(df - dead < -DECISION_MARGIN) && dead == 0.0f;
    bool contrib;
    float G = ex2_approx(pw);
    if constexpr (EXACT) {
      contrib = cand && df >= DECISION_MARGIN && pw <= -1e-5f;  // far from the threshold: decided
      const bool near_thr = cand && !contrib;                   // rare: the reference's arithmetic decides
      if (__any_sync(FULL, near_thr)) {
        if (near_thr) {
          const float val = RGB ? gaussian_exact_f64(px - r0.x, py - r0.y, lds128(rec + 32))
                                : gaussian_exact(px - r0.x, py - r0.y, lds128(rec + 32));
          G = val;
          contrib = !(r0.z * val < MIN_RENDER_ALPHA);
        }
      }
    } else {
      contrib = cand && df >= 0.0f && !(pw > 0.0f);
    }
    if (!__any_sync(FULL, contrib)) return false;
    G = contrib ? G : 0.0f;
    if constexpr (STATS) n_pairs += contrib ? 1u : 0u;
    // ---- straight-line for the whole warp (G = 0 makes a lane's contribution exactly zero)
    const float a = r0.z;
    const float aG = a * G;
    float coeff = (a * T) * G;
    if (!RGB && isnan(coeff)) coeff = 0.0f;
    float y[3];
    sh_colour<CC, RGB>(sh, Y, y);
    f0 = fmaf(-coeff, y[0], f0);
    f1 = fmaf(-coeff, y[1], f1);
    f2 = fmaf(-coeff, y[2], f2);
    float w0, w1, w2;
    if constexpr (RGB) {  // vol_render.h:305-307: grad_color += a T G * grad_out
      w0 = coeff * g0; w1 = coeff * g1; w2 = coeff * g2;
    } else {              // vol_render_sh.h:328-333
      w0 = coeff * (y[0] * (1.0f - y[0])) * g0;
      w1 = coeff * (y[1] * (1.0f - y[1])) * g1;
      w2 = coeff * (y[2] * (1.0f - y[2])) * g2;
    }
    // vol_render_sh.h:336-342
    const float one_m = 1.0f - aG;
    const float inv1m = -rcp_approx(one_m);
    float P = g0 * fmaf(y[0], T, f0 * inv1m);
    P = fmaf(g1, fmaf(y[1], T, f1 * inv1m), P);
    P = fmaf(g2, fmaf(y[2], T, f2 * inv1m), P);
    // kernels.h:394-418 with the inverse covariance recovered from the conic
    const float dx = px - r0.x, dy = py - r0.y;
    const float i00 = r1.x * INV_K, i11 = r1.z * INV_K, i01 = (-0.5f * INV_K) * r1.y;
    const float vx = fmaf(dx, i00, -dy * i01), vy = fmaf(dy, i11, -dx * i01);
    const float gam = P * aG;
    const float gmx = gam * vx, gmy = gam * vy;
    const float hg = 0.5f * gam;
    const float g00 = hg * vx * vx, g01 = hg * vx * vy, g11 = hg * vy * vy;
    const float ga = P * G;
    T *= one_m;
    if (T < thresh) dead = DEAD;  // vol_render_sh.h:296-298 (T only changes here)
    // ---- warp reduction over the 32 pixels through shared memory
    sts32(w_st + 4 * 0 * WROW, w0);
    sts32(w_st + 4 * 1 * WROW, w1);
    sts32(w_st + 4 * 2 * WROW, w2);
    sts32(w_st + 4 * 3 * WROW, gmx);
    sts32(w_st + 4 * 4 * WROW, gmy);
    sts32(w_st + 4 * 5 * WROW, g00);
    sts32(w_st + 4 * 6 * WROW, g01);
    sts32(w_st + 4 * 7 * WROW, g11);
    sts32(w_st + 4 * 8 * WROW, ga);
    __syncwarp();
    f32x2 A0 = 0ull, A1 = 0ull, A2 = 0ull;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      f32x2 lo, hi;
      lds128_2(w_sh + 16 * q, lo, hi);
      A0 = fma2(lo, Yt[2 * q], A0);
      A0 = fma2(hi, Yt[2 * q + 1], A0);
      lds128_2(w_sh + 4 * WROW + 16 * q, lo, hi);
      A1 = fma2(lo, Yt[2 * q], A1);
      A1 = fma2(hi, Yt[2 * q + 1], A1);
      lds128_2(w_sh + 8 * WROW + 16 * q, lo, hi);
      A2 = fma2(lo, Yt[2 * q], A2);
      A2 = fma2(hi, Yt[2 * q + 1], A2);
    }
    float a0 = sum2(A0), a1 = sum2(A1), a2 = sum2(A2);
    // six scalar sums: lane = 4*v + qd sums pixels 8*qd .. 8*qd+7 of row 3+v
    float sv;
    {
      const float4 x0 = lds128(w_sc), x1 = lds128(w_sc + 16);
      sv = ((x0.x + x0.y) + (x0.z + x0.w)) + ((x1.x + x1.y) + (x1.z + x1.w));
    }
    __syncwarp();
    a0 += __shfl_xor_sync(FULL, a0, 16);
    a1 += __shfl_xor_sync(FULL, a1, 16);
    a2 += __shfl_xor_sync(FULL, a2, 16);
    sv += __shfl_xor_sync(FULL, sv, 1);
    sv += __shfl_xor_sync(FULL, sv, 2);
    if (lane < KL) {  // a (warp, Gaussian) row is written at most once per use of its stage: plain stores
      sts32(row + 4 * lane, a0);
      sts32(row + 4 * (CC + lane), a1);
      sts32(row + 4 * (2 * CC + lane), a2);
    }
    if (sv_q == 0 && sv_row < 6) sts32(row + 4 * (SHF + sv_row), sv);
    return true;
  }
};

template <int C, bool EXACT, bool RGB, bool STATS>
__global__ void __launch_bounds__(NTHREADS + 32, 3)
composite_bwd3_kernel(const CompositeParams p) {
  constexpr int CC = C * C;
  constexpr int SHF = 3 * CC;
  constexpr int ROW = SHF + 6;
  constexpr int ROWP = (ROW + 3) & ~3;
  constexpr int NQ = ROWP / 4;
  constexpr int NW = NWARPS;
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4 *s_rec = reinterpret_cast<float4 *>(smem_raw);            // [RS][RG][3]
  float *s_sh = reinterpret_cast<float *>(s_rec + RS * RG * 3);    // [RS][RG][SHF]  (SHF*4 bytes per row)
  float *s_w = s_sh + RS * RG * SHF;                               // [NW][9][WROW]
  float *s_acc = s_w + NW * 9 * WROW;                              // [RS][NW][RG][ROWP]
  int *s_idr = reinterpret_cast<int *>(s_acc + RS * NW * RG * ROWP);  // [2*RS][RG] id ring
  int *s_nvalid = s_idr + 2 * RS * RG;                             // [RS]
  unsigned char *s_mask = reinterpret_cast<unsigned char *>(s_nvalid + RS);  // [RS][NW] rows written
  unsigned char *s_alive = s_mask + RS * NW;                       // [RS][NW]
  unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(s_alive + RS * NW);  // full[RS], done[RS]
  static_assert((RS * NW) % 4 == 0 && NW == 8, "mask rows are read as 64-bit words");

  const int tile_id = blockIdx.x;
  const int tile_y = tile_id / p.ntw, tile_x = tile_id - tile_y * p.ntw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int first = p.start[tile_id];
  if (first == -1) return;
  const int n_this = p.end[tile_id] - first;
  if (n_this <= 0) return;
  const int32_t *ids = p.ids + first;
  const int n_batches = (n_this + RG - 1) / RG;
  const uint32_t bar_full = smem_u32(s_bar), bar_done = smem_u32(s_bar + RS);

  // bulk staging (16-byte addressable SH rows): ONE arrival (+ the copies' byte count) completes `full`; element-wise
  // staging: every producer lane arrives, asynchronously, when its own cp.async copies have landed
  const bool bulk = RGB ? false : (p.sh_vec != 0);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s2 = 0; s2 < RS; ++s2) {
      mbar_init(bar_full + 8 * s2, bulk ? 1 : 32);
      mbar_init(bar_done + 8 * s2, NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < RS * RG) s_idr[threadIdx.x] = threadIdx.x < n_this ? ids[threadIdx.x] : 0;
  __syncthreads();

  if (warp == NW) {
    // ================================================================= producer / flusher warp
    unsigned long long staged = 0;
    int end_b = n_batches;  // first batch index armed as the sentinel
    auto arm = [&](int t) {
      const int s2 = t & (RS - 1);
      if (t > end_b) return;
      if (t == end_b) {  // sentinel: no Gaussians
        if (lane == 0) s_nvalid[s2] = 0;
        __syncwarp();
        if (bulk ? lane == 0 : true) mbar_arrive(bar_full + 8 * s2);
        return;
      }
      asm volatile("cp.async.wait_all;" ::: "memory");  // this warp's own id prefetches have landed
      __syncwarp();
      const int nv = min(RG, n_this - RG * t);
      const int *idp = s_idr + (t & (2 * RS - 1)) * RG;
      // ids of batch t + RS -> the ring slot batch t - RS used (its flush is done)
      {
        const int nxt = RG * (t + RS) + lane;
        if (lane < RG) {
          int *dst = s_idr + ((t + RS) & (2 * RS - 1)) * RG + lane;
          if (nxt < n_this) cp_async_4(dst, ids + nxt);
          else *dst = 0;
        }
      }
      // null records behind the batch (the walk is unrolled by UB): threshold +inf never contributes
      if (lane < (RG - nv) * 3) {
        const int e = nv * 3 + lane;
        s_rec[s2 * RG * 3 + e] = (e % 3 == 0) ? make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (bulk) {
        const bool dense = p.sh_sc == (uint32_t)CC;
        const uint32_t sh_bytes = 4 * SHF;
        __syncwarp();  // the padding stores above are ordered before lane 0's (releasing) arrival
        if (lane == 0) {
          s_nvalid[s2] = nv;
          mbar_arrive_expect_tx(bar_full + 8 * s2, (uint32_t)nv * (48 + sh_bytes));
        }
        __syncwarp();
        if (lane < nv) {
          const size_t g = (size_t)idp[lane];
          bulk_g2s(smem_u32(s_rec + (s2 * RG + lane) * 3), p.records + 3 * g, 48, bar_full + 8 * s2);
          const uint32_t dst = smem_u32(s_sh + (s2 * RG + lane) * SHF);
          if (dense) {
            bulk_g2s(dst, p.sh + g * p.sh_sg, sh_bytes, bar_full + 8 * s2);
          } else {
#pragma unroll
            for (int c = 0; c < 3; ++c)
              bulk_g2s(dst + 4 * c * CC, p.sh + g * p.sh_sg + c * p.sh_sc, 4 * CC, bar_full + 8 * s2);
          }
        }
      } else {
        // element-wise staging (SH rows that are not 16-byte addressable, RGB colours): every lane arrives on `full`
        // when ITS copies have landed (cp.async.mbarrier.arrive), so the warp does not wait for them and goes
        // straight back to flushing; the plain stores above (padding, s_nvalid) are fenced before the arrivals
        if (lane == 0) s_nvalid[s2] = nv;
        for (int e = lane; e < nv * 3; e += 32)
          cp_async_16(s_rec + s2 * RG * 3 + e, p.records + 3 * (size_t)idp[e / 3] + (e % 3));
        for (int e = lane; e < nv * SHF; e += 32) {
          const int j = e / SHF, r = e - SHF * j;
          const int c = r / CC, k = r - c * CC;
          cp_async_4(s_sh + (s2 * RG + j) * SHF + r, p.sh + (size_t)idp[j] * p.sh_sg + c * p.sh_sc + k);
        }
        __threadfence_block();
        cp_async_arrive_noinc(bar_full + 8 * s2);
      }
      staged += (unsigned long long)nv;
    };
    for (int t = 0; t < RS; ++t) arm(t);
    for (int b = 0; b < end_b; ++b) {
      const int s2 = b & (RS - 1);
      mbar_wait(bar_done + 8 * s2, (b / RS) & 1);
      const unsigned long long masks = *reinterpret_cast<const unsigned long long *>(s_mask + s2 * NW);
      const unsigned long long alive = *reinterpret_cast<const unsigned long long *>(s_alive + s2 * NW);
      if (alive == 0ull) end_b = min(end_b, b + RS);
      if (masks != 0ull) {
        const int nv = min(RG, n_this - RG * b);
        const int *idp = s_idr + (b & (2 * RS - 1)) * RG;
        const float *acc = s_acc + s2 * NW * RG * ROWP;
        for (int e = lane; e < nv * NQ; e += 32) {
          const int j = e / NQ, q = e - NQ * j;
          float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int w = 0; w < NW; ++w) {
            if ((masks >> (8 * w + j)) & 1ull) {
              const float4 v = *(reinterpret_cast<const float4 *>(acc + (w * RG + j) * ROWP) + q);
              sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
            }
          }
          if (sum.x == 0.f && sum.y == 0.f && sum.z == 0.f && sum.w == 0.f) continue;
          const size_t g = (size_t)idp[j];
          if (p.touched) p.touched[g] = 1;
          emit_quad<CC>(p, g, q, sum);
        }
      }
      __syncwarp();
      arm(b + RS);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (p.stats && lane == 0) atomicAdd(p.stats + 1, staged);
    return;
  }

  // ===================================================================== pixel warps
  const int lx = 8 * (warp & 1) + (lane & 7), ly = 4 * (warp >> 1) + (lane >> 3);
  const uint32_t gx = tile_x * TILE + lx, gy = tile_y * TILE + ly;
  const bool inside = gx < p.W && gy < p.H;
  const size_t pix = (size_t)gy * p.W + gx;
  BwdWalk<C, EXACT, RGB, STATS> wk;
  wk.n_pairs = 0;
  wk.px = p.topleft[0] + gx * p.psx;
  wk.py = p.topleft[1] + gy * p.psy;
  wk.lane = lane;
  const int kcol = lane & 15, half = lane >> 4;
  {
    float Yf[CC];
    if constexpr (RGB) Yf[0] = 1.0f;  // colour gradient = plain sum of the per-pixel weights
    else pixel_basis<C>(p.c2w, wk.px, wk.py, Yf);
    wk.Y.set(Yf);
    float *tr = s_w + warp * 9 * WROW;  // [32][9] floats (288 <= 9 * WROW)
#pragma unroll
    for (int i = 0; i < 8; ++i) wk.Yt[i] = 0ull;
#pragma unroll
    for (int rd = 0; rd < (CC + 7) / 8; ++rd) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (8 * rd + k < CC) tr[lane * 9 + k] = inside ? Yf[8 * rd + k] : 0.0f;
      __syncwarp();
      if ((kcol >> 3) == rd && kcol < CC) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          wk.Yt[i] = pack2(tr[(16 * half + 2 * i) * 9 + (kcol & 7)], tr[(16 * half + 2 * i + 1) * 9 + (kcol & 7)]);
      }
      __syncwarp();
    }
  }
  wk.g0 = wk.g1 = wk.g2 = wk.f0 = wk.f1 = wk.f2 = 0.f;
  if (inside) {
    wk.g0 = p.grad_out[3 * pix + 0]; wk.g1 = p.grad_out[3 * pix + 1]; wk.g2 = p.grad_out[3 * pix + 2];
    wk.f0 = p.out_saved[3 * pix + 0]; wk.f1 = p.out_saved[3 * pix + 1]; wk.f2 = p.out_saved[3 * pix + 2];
  }
  wk.T = 1.0f;
  wk.thresh = p.thresh;
  wk.dead = (inside && !(1.0f < p.thresh)) ? 0.0f : DEAD;
  {
    const uint32_t w_base = smem_u32(s_w + warp * 9 * WROW);
    wk.w_st = w_base + 4 * lane;
    wk.w_sh = w_base + 4 * (16 * half);
    wk.sv_row = lane >> 2;
    wk.sv_q = lane & 3;
    wk.w_sc = w_base + 4 * ((3 + (wk.sv_row < 6 ? wk.sv_row : 0)) * WROW + 8 * wk.sv_q);
  }
  const uint32_t rec0 = smem_u32(s_rec), sh0 = smem_u32(s_sh);
  const uint32_t acc0 = smem_u32(s_acc) + 4 * (warp * RG * ROWP);

  for (int b = 0;; ++b) {
    const int s2 = b & (RS - 1);
    mbar_wait(bar_full + 8 * s2, (b / RS) & 1);
    const int nv = s_nvalid[s2];
    if (nv == 0) break;
    uint32_t rec_a = rec0 + 48 * (s2 * RG);
    uint32_t sh_a = sh0 + 4 * SHF * (s2 * RG);
    uint32_t row_a = acc0 + 4 * (s2 * NW * RG * ROWP);
    uint32_t mask = 0;
    for (int j = 0; j < nv; j += UB, rec_a += 48 * UB, sh_a += 4 * SHF * UB, row_a += 4 * ROWP * UB) {
      if (!__any_sync(FULL, wk.dead == 0.0f)) break;
#pragma unroll
      for (int uu = 0; uu < UB; ++uu)
        if (wk.step(rec_a + 48 * uu, sh_a + 4 * SHF * uu, row_a + 4 * ROWP * uu)) mask |= 1u << (j + uu);
    }
    const bool alive = __any_sync(FULL, wk.dead == 0.0f);
    if (lane == 0) {
      s_mask[s2 * NW + warp] = (unsigned char)mask;
      s_alive[s2 * NW + warp] = alive ? 1 : 0;
    }
    __threadfence_block();  // this warp's accumulator rows / mask before its arrival
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_done + 8 * s2);
  }
  if constexpr (STATS) {
    const unsigned int wsum = __reduce_add_sync(FULL, wk.n_pairs);
    if (p.stats && lane == 0 && wsum) atomicAdd(p.stats + 3, (unsigned long long)wsum);
  }
}


// ---------------------------------------------------------------- host side

template <int C, int B>
static size_t fwd_smem() {
  return (size_t)2 * B * 3 * sizeof(float4) + (size_t)2 * B * 3 * C * C * sizeof(float) +
         (size_t)3 * B * sizeof(int);
}
template <int C>
static size_t bwd3_smem() {
  constexpr int ROWP = (3 * C * C + 6 + 3) & ~3;
  return (size_t)RS * RG * 3 * sizeof(float4) + (size_t)RS * RG * 3 * C * C * sizeof(float) +
         (size_t)NWARPS * 9 * WROW * sizeof(float) + (size_t)RS * NWARPS * RG * ROWP * sizeof(float) +
         (size_t)2 * RS * RG * sizeof(int) + (size_t)RS * sizeof(int) + (size_t)2 * RS * NWARPS +
         (size_t)2 * RS * sizeof(unsigned long long);
}

#ifndef GS3D_FWD_B
#define GS3D_FWD_B 64
#endif
constexpr int FWD_B = GS3D_FWD_B;  // (A/B, cfg 2: 32 / 64 / 128 Gaussians per batch x 3..6 CTAs per SM: 64 x 4 is fastest)

template <int C, int B, bool EXACT, bool RGB, bool STATS>
static int launch_fwd_t(const CompositeParams &p, uint32_t n_tiles, cudaStream_t st) {
  size_t sm = fwd_smem<C, B>();
  GS3D_CUDA(cudaFuncSetAttribute(composite_fwd_kernel<C, B, EXACT, RGB, STATS>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  composite_fwd_kernel<C, B, EXACT, RGB, STATS><<<n_tiles, NTHREADS, sm, st>>>(p);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}
template <int C, bool RGB = false>
static int launch_fwd(const CompositeParams &p, uint32_t n_tiles, cudaStream_t st) {
  if (p.stats)
    return p.exact ? launch_fwd_t<C, FWD_B, true, RGB, true>(p, n_tiles, st)
                   : launch_fwd_t<C, FWD_B, false, RGB, true>(p, n_tiles, st);
  return p.exact ? launch_fwd_t<C, FWD_B, true, RGB, false>(p, n_tiles, st)
                 : launch_fwd_t<C, FWD_B, false, RGB, false>(p, n_tiles, st);
}
template <int C, bool EXACT, bool RGB, bool STATS>
static int launch_bwd_t(const CompositeParams &p, uint32_t n_tiles, cudaStream_t st) {
  size_t sm = bwd3_smem<C>();
  GS3D_CUDA(cudaFuncSetAttribute(composite_bwd3_kernel<C, EXACT, RGB, STATS>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  composite_bwd3_kernel<C, EXACT, RGB, STATS><<<n_tiles, NTHREADS + 32, sm, st>>>(p);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}
template <int C, bool RGB = false>
static int launch_bwd(const CompositeParams &p, uint32_t n_tiles, cudaStream_t st) {
  if (p.stats)  // measurement launches (gs3d_set_stage_counters): same kernel + pair counting
    return p.exact ? launch_bwd_t<C, true, RGB, true>(p, n_tiles, st) : launch_bwd_t<C, false, RGB, true>(p, n_tiles, st);
  return p.exact ? launch_bwd_t<C, true, RGB, false>(p, n_tiles, st) : launch_bwd_t<C, false, RGB, false>(p, n_tiles, st);
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static unsigned long long *g_stage_counters = nullptr;  // gs3d_set_stage_counters

}  // namespace gs3d

using namespace gs3d;

extern "C" {

int gs3d_set_stage_counters(uint64_t *counters) {
  g_stage_counters = reinterpret_cast<unsigned long long *>(counters);
  return GS3D_OK;
}

int gs3d_composite_sh_forward(uint32_t M, const float *records, const float *sh_coeffs,
                              uint32_t sh_stride_g, uint32_t sh_stride_c, const int32_t *start,
                              const int32_t *end, const int32_t *gaussian_ids, float *out,
                              const float *topleft, const float *c2w, uint32_t tile_size,
                              uint32_t n_tiles_h, uint32_t n_tiles_w, float pixel_size_x,
                              float pixel_size_y, uint32_t H, uint32_t W, uint32_t C, float thresh,
                              const float *bg_rgb, float *final_T, int32_t *n_contrib,
                              int exact_decisions, void *stream) {
  (void)M;
  GS3D_REQUIRE(tile_size == TILE, GS3D_EUNSUPPORTED,
               "compositing kernels support tile_size 16 only (got %u)", tile_size);
  GS3D_REQUIRE(C >= 1 && C <= 4, GS3D_EINVAL, "SH order C must be 1..4 (got %u)", C);
  const uint32_t n_tiles = n_tiles_h * n_tiles_w;
  if (n_tiles == 0 || H == 0 || W == 0) return GS3D_OK;
  GS3D_REQUIRE(start && end && out && topleft && c2w, GS3D_EINVAL, "composite_sh_forward: null argument");
  GS3D_REQUIRE(aligned16(records), GS3D_EINVAL, "records must be 16-byte aligned");
  CompositeParams p = {};
  p.records = reinterpret_cast<const float4 *>(records);
  p.sh = sh_coeffs; p.sh_sg = sh_stride_g; p.sh_sc = sh_stride_c;
  p.start = start; p.end = end; p.ids = gaussian_ids;
  p.out = out; p.topleft = topleft; p.c2w = c2w;
  p.ntw = n_tiles_w; p.nth = n_tiles_h; p.psx = pixel_size_x; p.psy = pixel_size_y;
  p.H = H; p.W = W; p.thresh = thresh; p.bg = bg_rgb; p.final_T = final_T; p.n_contrib = n_contrib;
  p.exact = exact_decisions;
  p.stats = g_stage_counters;
  const uint32_t CC = C * C;
  p.sh_vec = (CC % 4 == 0) && (sh_stride_g % 4 == 0) && (sh_stride_c % 4 == 0) && aligned16(sh_coeffs);
  cudaStream_t st = as_stream(stream);
  switch (C) {
#ifndef GS3D_ONLY_C4  // (experimental builds compile the C = 4 kernels only)
    case 1: return launch_fwd<1>(p, n_tiles, st);
    case 2: return launch_fwd<2>(p, n_tiles, st);
    case 3: return launch_fwd<3>(p, n_tiles, st);
#endif
    default: return launch_fwd<4>(p, n_tiles, st);
  }
}

int gs3d_composite_sh_backward_peers(uint32_t M, const float *records, const float *sh_coeffs,
                               uint32_t sh_stride_g, uint32_t sh_stride_c, const int32_t *start,
                               const int32_t *end, const int32_t *gaussian_ids, const float *out,
                               const float *grad_out, float *grad_mean2d, float *grad_cov2d,
                               float *grad_sh, uint32_t gsh_stride_g, uint32_t gsh_stride_c,
                               float *grad_alpha, const float *topleft, const float *c2w,
                               uint32_t tile_size, uint32_t n_tiles_h, uint32_t n_tiles_w,
                               float pixel_size_x, float pixel_size_y, uint32_t H, uint32_t W,
                               uint32_t C, float thresh, int exact_decisions,
                               const uint64_t *peer_grad_sh_host, int n_peers,
                               void *multicast_grad_sh, uint8_t *touched, void *stream) {
  (void)M;
  GS3D_REQUIRE(tile_size == TILE, GS3D_EUNSUPPORTED,
               "compositing kernels support tile_size 16 only (got %u)", tile_size);
  GS3D_REQUIRE(C >= 1 && C <= 4, GS3D_EINVAL, "SH order C must be 1..4 (got %u)", C);
  const uint32_t n_tiles = n_tiles_h * n_tiles_w;
  if (n_tiles == 0 || H == 0 || W == 0) return GS3D_OK;
  GS3D_REQUIRE(start && end && out && grad_out && grad_mean2d && grad_cov2d && grad_sh && grad_alpha &&
                   topleft && c2w,
               GS3D_EINVAL, "composite_sh_backward: null argument");
  GS3D_REQUIRE(aligned16(records), GS3D_EINVAL, "records must be 16-byte aligned");
  CompositeParams p = {};
  p.records = reinterpret_cast<const float4 *>(records);
  p.sh = sh_coeffs; p.sh_sg = sh_stride_g; p.sh_sc = sh_stride_c;
  p.start = start; p.end = end; p.ids = gaussian_ids;
  p.topleft = topleft; p.c2w = c2w;
  p.ntw = n_tiles_w; p.nth = n_tiles_h; p.psx = pixel_size_x; p.psy = pixel_size_y;
  p.H = H; p.W = W; p.thresh = thresh; p.exact = exact_decisions;
  p.out_saved = out; p.grad_out = grad_out;
  p.g_mean = grad_mean2d; p.g_cov = grad_cov2d; p.g_sh = grad_sh; p.g_alpha = grad_alpha;
  p.gsh_sg = gsh_stride_g; p.gsh_sc = gsh_stride_c;
  p.touched = touched;
  p.stats = g_stage_counters;
  const uint32_t CC = C * C;
  p.sh_vec = (CC % 4 == 0) && (sh_stride_g % 4 == 0) && (sh_stride_c % 4 == 0) && aligned16(sh_coeffs);
  p.gsh_vec = (CC % 4 == 0) && (gsh_stride_g % 4 == 0) && (gsh_stride_c % 4 == 0) && aligned16(grad_sh);
  GS3D_REQUIRE(n_peers >= 0 && n_peers <= 8, GS3D_EINVAL, "n_peers must be 0..8 (got %d)", n_peers);
  GS3D_REQUIRE(n_peers == 0 || peer_grad_sh_host, GS3D_EINVAL, "peer pointer array is null");
  p.n_peers = n_peers;
  for (int r = 0; r < n_peers; ++r) {
    p.g_sh_peer[r] = reinterpret_cast<float *>(static_cast<uintptr_t>(peer_grad_sh_host[r]));
    GS3D_REQUIRE(p.g_sh_peer[r] && aligned16(p.g_sh_peer[r]), GS3D_EINVAL, "bad peer pointer %d", r);
  }
  // the multicast path needs vector reductions; otherwise fall back to the per-peer loop
  p.g_sh_mc = (multicast_grad_sh && p.gsh_vec && n_peers > 0) ? static_cast<float *>(multicast_grad_sh) : nullptr;
  cudaStream_t st = as_stream(stream);
  switch (C) {
#ifndef GS3D_ONLY_C4
    case 1: return launch_bwd<1>(p, n_tiles, st);
    case 2: return launch_bwd<2>(p, n_tiles, st);
    case 3: return launch_bwd<3>(p, n_tiles, st);
#endif
    default: return launch_bwd<4>(p, n_tiles, st);
  }
}

int gs3d_composite_sh_backward(uint32_t M, const float *records, const float *sh_coeffs,
                               uint32_t sh_stride_g, uint32_t sh_stride_c, const int32_t *start,
                               const int32_t *end, const int32_t *gaussian_ids, const float *out,
                               const float *grad_out, float *grad_mean2d, float *grad_cov2d,
                               float *grad_sh, uint32_t gsh_stride_g, uint32_t gsh_stride_c,
                               float *grad_alpha, const float *topleft, const float *c2w,
                               uint32_t tile_size, uint32_t n_tiles_h, uint32_t n_tiles_w,
                               float pixel_size_x, float pixel_size_y, uint32_t H, uint32_t W,
                               uint32_t C, float thresh, int exact_decisions, void *stream) {
  return gs3d_composite_sh_backward_peers(M, records, sh_coeffs, sh_stride_g, sh_stride_c, start, end,
                                          gaussian_ids, out, grad_out, grad_mean2d, grad_cov2d, grad_sh,
                                          gsh_stride_g, gsh_stride_c, grad_alpha, topleft, c2w, tile_size,
                                          n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C,
                                          thresh, exact_decisions, nullptr, 0, nullptr, nullptr, stream);
}

// ---- legacy RGB path: tile_based_vol_rendering_start_end{,_backward}, bindings.cpp:29-33 /
// render.cu / vol_render.h:716-923.  Same kernels, RGB mode: the staged row is the colour itself, the
// near-threshold decisions use the reference's FP64 Gaussian.

static int rgb_params(CompositeParams &p, const float *records, const float *color, const int32_t *start,
                      const int32_t *end, const int32_t *gaussian_ids, const float *topleft, uint32_t tile_size,
                      uint32_t n_tiles_h, uint32_t n_tiles_w, float psx, float psy, uint32_t H, uint32_t W,
                      float thresh, int exact) {
  GS3D_REQUIRE(tile_size == TILE, GS3D_EUNSUPPORTED, "compositing kernels support tile_size 16 only (got %u)",
               tile_size);
  GS3D_REQUIRE(start && end && topleft && color, GS3D_EINVAL, "composite_rgb: null argument");
  GS3D_REQUIRE(aligned16(records), GS3D_EINVAL, "records must be 16-byte aligned");
  p.records = reinterpret_cast<const float4 *>(records);
  p.sh = color; p.sh_sg = 3; p.sh_sc = 1; p.sh_vec = 0;
  p.start = start; p.end = end; p.ids = gaussian_ids; p.topleft = topleft; p.c2w = nullptr;
  p.ntw = n_tiles_w; p.nth = n_tiles_h; p.psx = psx; p.psy = psy; p.H = H; p.W = W; p.thresh = thresh;
  p.exact = exact;
  p.stats = g_stage_counters;
  return GS3D_OK;
}

int gs3d_composite_rgb_forward(uint32_t M, const float *records, const float *color, const int32_t *start,
                               const int32_t *end, const int32_t *gaussian_ids, float *out, const float *topleft,
                               uint32_t tile_size, uint32_t n_tiles_h, uint32_t n_tiles_w, float pixel_size_x,
                               float pixel_size_y, uint32_t H, uint32_t W, float thresh, int exact_decisions,
                               void *stream) {
  (void)M;
  const uint32_t n_tiles = n_tiles_h * n_tiles_w;
  if (n_tiles == 0 || H == 0 || W == 0) return GS3D_OK;
  GS3D_REQUIRE(out, GS3D_EINVAL, "composite_rgb_forward: out is null");
  CompositeParams p = {};
  int rc = rgb_params(p, records, color, start, end, gaussian_ids, topleft, tile_size, n_tiles_h, n_tiles_w,
                      pixel_size_x, pixel_size_y, H, W, thresh, exact_decisions);
  if (rc) return rc;
  p.out = out;
  cudaStream_t st = as_stream(stream);
  return launch_fwd<1, true>(p, n_tiles, st);
}

int gs3d_composite_rgb_backward(uint32_t M, const float *records, const float *color, const int32_t *start,
                                const int32_t *end, const int32_t *gaussian_ids, const float *out,
                                const float *grad_out, float *grad_mean2d, float *grad_cov2d, float *grad_color,
                                float *grad_alpha, const float *topleft, uint32_t tile_size, uint32_t n_tiles_h,
                                uint32_t n_tiles_w, float pixel_size_x, float pixel_size_y, uint32_t H, uint32_t W,
                                float thresh, int exact_decisions, void *stream) {
  (void)M;
  const uint32_t n_tiles = n_tiles_h * n_tiles_w;
  if (n_tiles == 0 || H == 0 || W == 0) return GS3D_OK;
  GS3D_REQUIRE(out && grad_out && grad_mean2d && grad_cov2d && grad_color && grad_alpha, GS3D_EINVAL,
               "composite_rgb_backward: null argument");
  CompositeParams p = {};
  int rc = rgb_params(p, records, color, start, end, gaussian_ids, topleft, tile_size, n_tiles_h, n_tiles_w,
                      pixel_size_x, pixel_size_y, H, W, thresh, exact_decisions);
  if (rc) return rc;
  p.out_saved = out; p.grad_out = grad_out;
  p.g_mean = grad_mean2d; p.g_cov = grad_cov2d; p.g_sh = grad_color; p.g_alpha = grad_alpha;
  p.gsh_sg = 3; p.gsh_sc = 1; p.gsh_vec = 0;
  cudaStream_t st = as_stream(stream);
  return launch_bwd<1, true>(p, n_tiles, st);
}

}  // extern "C"
