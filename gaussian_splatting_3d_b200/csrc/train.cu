// The two steps either side of the rasteriser in the training loop:
//
//   gs3d_adam_step      : `opt.step()` of main_sh.py:193 on the torch.optim.Adam that
//                         SHRenderer.get_optimizer builds (sh_renderer.py:720-729: five groups with their
//                         own lr, betas (0.9, 0.99), eps 1e-8, no weight decay / amsgrad).  ONE launch for
//                         all groups; HBM-bound streaming.  The reference re-creates the optimiser after
//                         every step (main_sh.py:238), so every step it takes is a FIRST Adam step: the
//                         moments are zero on entry.  state_mode 1 exploits that (moments written, never
//                         read: 20 B/element) and state_mode 2 drops them altogether (12 B/element) --
//                         against 28 B/element for the general step and ~10 full passes for torch's
//                         foreach implementation.
//   gs3d_adc_classify / gs3d_adc_plan / gs3d_adc_apply :
//                         split_gaussians / remove_low_alpha_gaussians / select_masked_gaussians
//                         (sh_renderer.py:426-600,731-741): ~60 boolean-mask gathers, cats and temporaries
//                         become classify -> deterministic block scan -> one fused row mover that writes the
//                         new parameter tensors in the reference's order
//                         [kept + to-be-cloned originals | clones | split samples (first copies, second copies)].
#include "common.cuh"

namespace gs3d {

// ------------------------------------------------------------------------------------------ Adam

constexpr int ADAM_MAX_SEGMENTS = 8;

struct AdamSeg {
  float *p;
  const float *g;
  float *m, *v;
  unsigned long long n;
  float neg_step_size;  // -(lr / (1 - beta1^t)), rounded to FP32 like torch's scalar
};

struct AdamParams {
  AdamSeg seg[ADAM_MAX_SEGMENTS];
  float w1;            // 1 - beta1 (lerp weight)
  float beta2, w2;     // beta2, 1 - beta2
  float bc2_sqrt, eps; // sqrt(1 - beta2^t), eps
};

// torch.optim.Adam single-tensor arithmetic (torch/optim/adam.py `_single_tensor_adam`):
//   m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, value=1-b2);
//   denom = v.sqrt() / bias_correction2_sqrt + eps;  p.addcdiv_(m, denom, value=-lr/bias_correction1)
template <int MODE>
__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, const AdamParams &P, float nss) {
  if (MODE == 0) {
    m = P.w1 < 0.5f ? fmaf(P.w1, g - m, m) : g - (g - m) * (1.0f - P.w1);  // ATen lerp
    v = __fmul_rn(v, P.beta2);
    v = v + P.w2 * (g * g);
  } else {  // zero moments on entry
    m = P.w1 < 0.5f ? P.w1 * g : g - g * (1.0f - P.w1);
    v = P.w2 * (g * g);
  }
  const float denom = sqrtf(v) / P.bc2_sqrt + P.eps;
  p = p + nss * (m / denom);
}

template <int MODE>
__global__ void __launch_bounds__(256) adam_kernel(const AdamParams P) {
  const AdamSeg s = P.seg[blockIdx.y];
  const float nss = s.neg_step_size;
  const unsigned long long n4 = s.n >> 2;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(s.p) | reinterpret_cast<uintptr_t>(s.g) |
                     reinterpret_cast<uintptr_t>(s.m) | reinterpret_cast<uintptr_t>(s.v)) & 15) == 0;
  unsigned long long done = 0;
  if (vec) {
    float4 *p4 = reinterpret_cast<float4 *>(s.p);
    const float4 *g4 = reinterpret_cast<const float4 *>(s.g);
    float4 *m4 = reinterpret_cast<float4 *>(s.m);
    float4 *v4 = reinterpret_cast<float4 *>(s.v);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 p = p4[i];
      const float4 g = __ldcs(g4 + i);  // gradients are read once
      float4 m = make_float4(0.f, 0.f, 0.f, 0.f), v = m;
      if (MODE == 0) { m = m4[i]; v = v4[i]; }
      adam_one<MODE>(p.x, g.x, m.x, v.x, P, nss);
      adam_one<MODE>(p.y, g.y, m.y, v.y, P, nss);
      adam_one<MODE>(p.z, g.z, m.z, v.z, P, nss);
      adam_one<MODE>(p.w, g.w, m.w, v.w, P, nss);
      p4[i] = p;
      if (MODE != 2) { m4[i] = m; v4[i] = v; }
    }
    done = n4 << 2;
  }
  for (unsigned long long i = done + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += stride) {
    float p = s.p[i], m = 0.f, v = 0.f;
    if (MODE == 0) { m = s.m[i]; v = s.v[i]; }
    adam_one<MODE>(p, s.g[i], m, v, P, nss);
    s.p[i] = p;
    if (MODE != 2) { s.m[i] = m; s.v[i] = v; }
  }
}

// ------------------------------------------------------------------------------------------ ADC

constexpr int ADC_BLOCK = 256;  // Gaussians per block of the plan / apply kernels

__device__ __forceinline__ uint8_t adc_class_of(const uint8_t *cls, uint32_t g, int is_keep_mask) {
  const uint8_t c = cls[g];
  return is_keep_mask ? (c ? GS3D_ADC_KEEP : GS3D_ADC_DROP) : c;
}

// sh_renderer.py:433-456: hot = grad_mean [/ (cnt + 1e-5)] > pos_grad_thresh; big = any(svec > split_scale_thresh)
__global__ void __launch_bounds__(256)
adc_classify_kernel(uint32_t N, const float *__restrict__ acc, const int32_t *__restrict__ cnt, int reduction,
                    float pos_thresh, const float *__restrict__ svec_param, int svec_act, float scale_thresh,
                    uint8_t *__restrict__ cls) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  float a = acc[g];
  if (reduction == 2) a = a / (static_cast<float>(cnt[g]) + 1e-5f);  // int32 + python float -> float32
  uint8_t c = GS3D_ADC_KEEP;
  if (a > pos_thresh) {
    bool big = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float s = svec_param[3 * (size_t)g + k];
      if (svec_act) s = expf(s);
      big |= s > scale_thresh;
    }
    c = big ? GS3D_ADC_SPLIT : GS3D_ADC_CLONE;
  }
  cls[g] = c;
}

// remove_low_alpha_gaussians (sh_renderer.py:542-560): keep iff act(alpha_param) >= thresh
__global__ void __launch_bounds__(256)
adc_classify_alpha_kernel(uint32_t N, const float *__restrict__ alpha_param, int alpha_act, float thresh,
                          uint8_t *__restrict__ cls) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  float a = alpha_param[g];
  if (alpha_act) a = 1.0f / (1.0f + expf(-a));
  cls[g] = a >= thresh ? GS3D_ADC_KEEP : GS3D_ADC_DROP;
}

// pass 1: per block, #rows that stay (KEEP or CLONE), #CLONE, #SPLIT
__global__ void __launch_bounds__(ADC_BLOCK)
adc_count_kernel(uint32_t N, const uint8_t *__restrict__ cls, int is_keep_mask, uint32_t *__restrict__ blk /*[nb][3]*/) {
  __shared__ uint32_t s_c[3][ADC_BLOCK / 32];
  const uint32_t g = blockIdx.x * ADC_BLOCK + threadIdx.x;
  const uint8_t c = g < N ? adc_class_of(cls, g, is_keep_mask) : (uint8_t)GS3D_ADC_DROP;
  const uint32_t b0 = __ballot_sync(0xffffffffu, c == GS3D_ADC_KEEP || c == GS3D_ADC_CLONE);
  const uint32_t b1 = __ballot_sync(0xffffffffu, c == GS3D_ADC_CLONE);
  const uint32_t b2 = __ballot_sync(0xffffffffu, c == GS3D_ADC_SPLIT);
  if ((threadIdx.x & 31) == 0) {
    s_c[0][threadIdx.x >> 5] = __popc(b0);
    s_c[1][threadIdx.x >> 5] = __popc(b1);
    s_c[2][threadIdx.x >> 5] = __popc(b2);
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    uint32_t t = 0;
    for (int w = 0; w < ADC_BLOCK / 32; ++w) t += s_c[threadIdx.x][w];
    blk[3 * (size_t)blockIdx.x + threadIdx.x] = t;
  }
}

// pass 2 (one block): exclusive scan of the three per-block counters; totals[0..2]
__global__ void __launch_bounds__(1024)
adc_scan_kernel(uint32_t nb, uint32_t *__restrict__ blk, unsigned long long *__restrict__ totals) {
  __shared__ uint32_t wsum[3][32];
  __shared__ uint32_t carry_s[3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 3) carry_s[threadIdx.x] = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < nb; b0 += 1024) {
    const uint32_t b = b0 + threadIdx.x;
    uint32_t x[3], incl[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      x[k] = b < nb ? blk[3 * (size_t)b + k] : 0u;
      incl[k] = x[k];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl[k], o);
        if (lane >= o) incl[k] += y;
      }
      if (lane == 31) wsum[k][warp] = incl[k];
    }
    __syncthreads();
    uint32_t out[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      uint32_t wb = 0;
      for (int w = 0; w < warp; ++w) wb += wsum[k][w];
      out[k] = carry_s[k] + wb + incl[k];  // inclusive
      if (b < nb) blk[3 * (size_t)b + k] = out[k] - x[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) {
#pragma unroll
      for (int k = 0; k < 3; ++k) carry_s[k] = out[k];
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) totals[threadIdx.x] = carry_s[threadIdx.x];
}

struct AdcApply {
  uint32_t N, sh_width;
  const uint8_t *cls;
  int is_keep_mask;
  const uint32_t *blk;  // exclusive block offsets [nb][3]
  uint32_t n_stay, n_clone, n_split;
  const float *mean, *qvec, *svec, *sh, *alpha;
  float *mean_o, *qvec_o, *svec_o, *sh_o, *alpha_o;
  const float *noise;  // [2 * n_split, 3]
  int svec_act;
  float shrink;
};

// pass 3: 16 lanes move one Gaussian row (59 floats at maxC = 4) to its one or two destination rows.
__global__ void __launch_bounds__(ADC_BLOCK) adc_apply_kernel(const AdcApply A) {
  __shared__ uint32_t s_w[3][ADC_BLOCK / 32];
  __shared__ uint32_t s_d0[ADC_BLOCK], s_d1[ADC_BLOCK];  // destination rows (0xffffffff = none)
  __shared__ uint8_t s_c[ADC_BLOCK];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = blockIdx.x * ADC_BLOCK;
  {
    const uint32_t g = base + threadIdx.x;
    const uint8_t c = g < A.N ? adc_class_of(A.cls, g, A.is_keep_mask) : (uint8_t)GS3D_ADC_DROP;
    const bool stay = c == GS3D_ADC_KEEP || c == GS3D_ADC_CLONE;
    const uint32_t b0 = __ballot_sync(0xffffffffu, stay);
    const uint32_t b1 = __ballot_sync(0xffffffffu, c == GS3D_ADC_CLONE);
    const uint32_t b2 = __ballot_sync(0xffffffffu, c == GS3D_ADC_SPLIT);
    if (lane == 0) {
      s_w[0][warp] = __popc(b0);
      s_w[1][warp] = __popc(b1);
      s_w[2][warp] = __popc(b2);
    }
    __syncthreads();
    uint32_t r[3];
    const uint32_t below = (1u << lane) - 1u;
    r[0] = __popc(b0 & below); r[1] = __popc(b1 & below); r[2] = __popc(b2 & below);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      for (int w = 0; w < warp; ++w) r[k] += s_w[k][w];
      r[k] += A.blk[3 * (size_t)blockIdx.x + k];
    }
    uint32_t d0 = 0xffffffffu, d1 = 0xffffffffu;
    if (stay) d0 = r[0];
    if (c == GS3D_ADC_CLONE) d1 = A.n_stay + r[1];
    if (c == GS3D_ADC_SPLIT) {  // .repeat(2, 1): all first copies, then all second copies
      d0 = A.n_stay + A.n_clone + r[2];
      d1 = d0 + A.n_split;
    }
    s_d0[threadIdx.x] = d0;
    s_d1[threadIdx.x] = d1;
    s_c[threadIdx.x] = c;
  }
  __syncthreads();
  const uint32_t W = A.sh_width;
  const uint32_t pts = A.n_stay + A.n_clone;
  for (uint32_t w = threadIdx.x; w < ADC_BLOCK * 16; w += ADC_BLOCK) {
    const uint32_t j = w >> 4, part = w & 15;
    const uint32_t g = base + j;
    if (g >= A.N) break;
    const uint8_t c = s_c[j];
    if (c == GS3D_ADC_DROP) continue;
    const uint32_t dst[2] = {s_d0[j], s_d1[j]};
    // SH row and opacity: plain copies
    for (uint32_t k = part; k < W; k += 16) {
      const float x = A.sh[(size_t)g * W + k];
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (dst[t] != 0xffffffffu) A.sh_o[(size_t)dst[t] * W + k] = x;
    }
    if (part == 10) {
      const float x = A.alpha[g];
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (dst[t] != 0xffffffffu) A.alpha_o[dst[t]] = x;
    } else if (part >= 3 && part < 7) {
      const float x = A.qvec[4 * (size_t)g + (part - 3)];
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (dst[t] != 0xffffffffu) A.qvec_o[4 * (size_t)dst[t] + (part - 3)] = x;
    } else if (part >= 7 && part < 10) {
      const uint32_t k = part - 7;
      float x = A.svec[3 * (size_t)g + k];
      if (c == GS3D_ADC_SPLIT) {  // svec_inv_act(svec / scale_shrink_factor), sh_renderer.py:521-523
        if (A.svec_act) x = logf(expf(x) / A.shrink);
        else x = x / A.shrink;
      }
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (dst[t] != 0xffffffffu) A.svec_o[3 * (size_t)dst[t] + k] = x;
    } else if (part < 3) {
      const float x = A.mean[3 * (size_t)g + part];
      if (c != GS3D_ADC_SPLIT) {
#pragma unroll
        for (int t = 0; t < 2; ++t)
          if (dst[t] != 0xffffffffu) A.mean_o[3 * (size_t)dst[t] + part] = x;
      } else {
        // mean + R(q)^T (randn * svec): sh_renderer.py:470-476 (einsum "bij,bj->bi" on the transposed matrix)
        float R[9];
        quat_to_rotmat(A.qvec[4 * (size_t)g], A.qvec[4 * (size_t)g + 1], A.qvec[4 * (size_t)g + 2],
                       A.qvec[4 * (size_t)g + 3], R, nullptr, nullptr);
        float s[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          s[k] = A.svec[3 * (size_t)g + k];
          if (A.svec_act) s[k] = expf(s[k]);
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const float *nz = A.noise + 3 * (size_t)(dst[t] - pts);
          const float g0 = nz[0] * s[0], g1 = nz[1] * s[1], g2 = nz[2] * s[2];
          const float o = (R[0 + part] * g0 + R[3 + part] * g1) + R[6 + part] * g2;
          A.mean_o[3 * (size_t)dst[t] + part] = x + o;
        }
      }
    }
  }
}

}  // namespace gs3d

using namespace gs3d;

extern "C" {

int gs3d_adam_step(int n_segments, const gs3d_adam_segment *segments_host, double beta1, double beta2, double eps,
                   uint32_t step, int state_mode, void *stream) {
  GS3D_REQUIRE(n_segments >= 0 && n_segments <= ADAM_MAX_SEGMENTS, GS3D_EINVAL,
               "adam_step: 0..%d segments per call (got %d)", ADAM_MAX_SEGMENTS, n_segments);
  GS3D_REQUIRE(state_mode >= 0 && state_mode <= 2, GS3D_EINVAL, "adam_step: state_mode must be 0, 1 or 2");
  GS3D_REQUIRE(step >= 1, GS3D_EINVAL, "adam_step: step counts from 1");
  GS3D_REQUIRE(state_mode == 0 || step == 1, GS3D_EINVAL,
               "adam_step: state_mode %d assumes zero moments, i.e. step == 1 (got %u)", state_mode, step);
  GS3D_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, GS3D_EINVAL,
               "adam_step: bad betas / eps");
  if (n_segments == 0) return GS3D_OK;
  GS3D_REQUIRE(segments_host, GS3D_EINVAL, "adam_step: segments_host is null");
  AdamParams P;
  double bc1 = 1.0, bc2 = 1.0, p1 = 1.0, p2 = 1.0;
  for (uint32_t i = 0; i < step && (p1 > 0.0 || p2 > 0.0); ++i) { p1 *= beta1; p2 *= beta2; }
  bc1 = 1.0 - p1;
  bc2 = 1.0 - p2;
  int n = 0;
  unsigned long long n_max = 0;
  for (int i = 0; i < n_segments; ++i) {
    const gs3d_adam_segment &s = segments_host[i];
    if (s.n == 0) continue;
    GS3D_REQUIRE(s.param && s.grad, GS3D_EINVAL, "adam_step: segment %d has a null param / grad", i);
    GS3D_REQUIRE(state_mode == 2 || (s.exp_avg && s.exp_avg_sq), GS3D_EINVAL,
                 "adam_step: segment %d has no moment buffers (state_mode %d)", i, state_mode);
    AdamSeg &d = P.seg[n++];
    d.p = s.param; d.g = s.grad;
    d.m = state_mode == 2 ? nullptr : s.exp_avg;
    d.v = state_mode == 2 ? nullptr : s.exp_avg_sq;
    d.n = s.n;
    d.neg_step_size = static_cast<float>(-(s.lr / bc1));
    n_max = s.n > n_max ? s.n : n_max;
  }
  if (n == 0) return GS3D_OK;
  P.w1 = static_cast<float>(1.0 - beta1);
  P.beta2 = static_cast<float>(beta2);
  P.w2 = static_cast<float>(1.0 - beta2);
  P.bc2_sqrt = static_cast<float>(sqrt(bc2));
  P.eps = static_cast<float>(eps);
  const unsigned long long want = div_up<unsigned long long>(div_up<unsigned long long>(n_max, 4ull), 256ull);
  const dim3 grid((unsigned)(want < 148ull * 8ull ? (want ? want : 1ull) : 148ull * 8ull), (unsigned)n);
  cudaStream_t st = as_stream(stream);
  if (state_mode == 0) adam_kernel<0><<<grid, 256, 0, st>>>(P);
  else if (state_mode == 1) adam_kernel<1><<<grid, 256, 0, st>>>(P);
  else adam_kernel<2><<<grid, 256, 0, st>>>(P);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

int gs3d_adc_classify(uint32_t N, const float *grad_mean_acc, const int32_t *cnt, int reduction,
                      float pos_grad_thresh, const float *svec_param, int svec_act, float split_scale_thresh,
                      uint8_t *cls, void *stream) {
  GS3D_REQUIRE(reduction == 1 || reduction == 2, GS3D_EINVAL, "adc_classify: reduction 1 (max) or 2 (mean)");
  if (N == 0) return GS3D_OK;
  GS3D_REQUIRE(grad_mean_acc && svec_param && cls && (reduction == 1 || cnt), GS3D_EINVAL,
               "adc_classify: null argument");
  adc_classify_kernel<<<div_up(N, 256u), 256, 0, as_stream(stream)>>>(N, grad_mean_acc, cnt, reduction,
                                                                      pos_grad_thresh, svec_param, svec_act,
                                                                      split_scale_thresh, cls);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

int gs3d_adc_classify_alpha(uint32_t N, const float *alpha_param, int alpha_act, float alpha_thresh, uint8_t *cls,
                            void *stream) {
  if (N == 0) return GS3D_OK;
  GS3D_REQUIRE(alpha_param && cls, GS3D_EINVAL, "adc_classify_alpha: null argument");
  adc_classify_alpha_kernel<<<div_up(N, 256u), 256, 0, as_stream(stream)>>>(N, alpha_param, alpha_act,
                                                                            alpha_thresh, cls);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

size_t gs3d_adc_scratch_bytes(uint32_t N) {
  const size_t nb = div_up(N ? N : 1u, (uint32_t)ADC_BLOCK);
  return align_up(3 * nb * sizeof(uint32_t)) + 256;
}

int gs3d_adc_plan(uint32_t N, const uint8_t *cls, int cls_is_keep_mask, int64_t *counts_host, void *scratch,
                  size_t scratch_bytes, void *stream) {
  GS3D_REQUIRE(counts_host, GS3D_EINVAL, "adc_plan: counts_host is null");
  counts_host[0] = counts_host[1] = counts_host[2] = 0;
  if (N == 0) return GS3D_OK;
  GS3D_REQUIRE(cls && scratch && scratch_bytes >= gs3d_adc_scratch_bytes(N), GS3D_EINVAL,
               "adc_plan: null argument / scratch too small");
  cudaStream_t st = as_stream(stream);
  const uint32_t nb = div_up(N, (uint32_t)ADC_BLOCK);
  Scratch sc(scratch, scratch_bytes);
  uint32_t *blk = sc.take<uint32_t>(3 * (size_t)nb);
  unsigned long long *totals = sc.take<unsigned long long>(3);
  GS3D_REQUIRE(blk && totals, GS3D_EINVAL, "adc_plan: scratch exhausted");
  adc_count_kernel<<<nb, ADC_BLOCK, 0, st>>>(N, cls, cls_is_keep_mask, blk);
  GS3D_LAUNCH_CHECK();
  adc_scan_kernel<<<1, 1024, 0, st>>>(nb, blk, totals);
  GS3D_LAUNCH_CHECK();
  int64_t *box = pinned_mailbox();
  GS3D_REQUIRE(box != nullptr, GS3D_ECUDA, "pinned mailbox unavailable");
  GS3D_CUDA(cudaMemcpyAsync(box, totals, 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  GS3D_CUDA(cudaStreamSynchronize(st));  // the reference's `.sum().item()` (sh_renderer.py:458-459)
  counts_host[0] = box[0];
  counts_host[1] = box[1];
  counts_host[2] = box[2];
  return GS3D_OK;
}

int gs3d_adc_apply(uint32_t N, const uint8_t *cls, int cls_is_keep_mask, const void *plan_scratch,
                   const int64_t *counts_host, const float *mean, const float *qvec, const float *svec_param,
                   const float *sh_coeffs, const float *alpha_param, uint32_t sh_width, int svec_act,
                   float scale_shrink_factor, const float *noise, float *mean_out, float *qvec_out,
                   float *svec_param_out, float *sh_coeffs_out, float *alpha_param_out, void *stream) {
  if (N == 0) return GS3D_OK;
  GS3D_REQUIRE(cls && plan_scratch && counts_host, GS3D_EINVAL, "adc_apply: null plan");
  GS3D_REQUIRE(mean && qvec && svec_param && sh_coeffs && alpha_param, GS3D_EINVAL, "adc_apply: null input");
  const int64_t n_out = counts_host[0] + counts_host[1] + 2 * counts_host[2];
  GS3D_REQUIRE(counts_host[0] >= 0 && counts_host[1] >= 0 && counts_host[2] >= 0 && n_out < (1ll << 32), GS3D_EINVAL,
               "adc_apply: bad counts");
  if (n_out == 0) return GS3D_OK;
  GS3D_REQUIRE(mean_out && qvec_out && svec_param_out && sh_coeffs_out && alpha_param_out, GS3D_EINVAL,
               "adc_apply: null output");
  GS3D_REQUIRE(counts_host[2] == 0 || (noise && scale_shrink_factor > 0.0f), GS3D_EINVAL,
               "adc_apply: splitting needs noise [2*n_split,3] and a positive scale_shrink_factor");
  AdcApply A;
  A.N = N; A.sh_width = sh_width; A.cls = cls; A.is_keep_mask = cls_is_keep_mask;
  A.blk = static_cast<const uint32_t *>(plan_scratch);
  A.n_stay = (uint32_t)counts_host[0]; A.n_clone = (uint32_t)counts_host[1]; A.n_split = (uint32_t)counts_host[2];
  A.mean = mean; A.qvec = qvec; A.svec = svec_param; A.sh = sh_coeffs; A.alpha = alpha_param;
  A.mean_o = mean_out; A.qvec_o = qvec_out; A.svec_o = svec_param_out; A.sh_o = sh_coeffs_out;
  A.alpha_o = alpha_param_out; A.noise = noise; A.svec_act = svec_act; A.shrink = scale_shrink_factor;
  adc_apply_kernel<<<div_up(N, (uint32_t)ADC_BLOCK), ADC_BLOCK, 0, as_stream(stream)>>>(A);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

}  // extern "C"
