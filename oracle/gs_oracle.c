/*
 * gs_oracle.c -- CPU restatement of the reference rasteriser kernels.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (gaussian_splatting_3d_b200/) never links or calls it and
 * fails loudly when its CUDA library is missing.
 *
 * Parity pinning: the primitives below are checked in tests/test_oracle.py against the
 * known-answer vectors produced by the reference's own host-compilable code (SURVEY.md 8c:
 * kernel_gaussian_2d_float = 0.951229393 for test/gaussian_test.py's vector; SIGMOID(0.3);
 * spherical_harmonic((1,2,3)/sqrt14, C=4)), the closed-form gradients of test/gaussian_test.py,
 * finite differences, and -- on the GPU box -- against the real reference extension built by
 * oracle/build_ref.py into oracle/_ref/.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/gs/src/include unless noted).  Plain FP32 arithmetic, compiled with
 * -ffp-contract=off so the operation order is the source order of the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GSO_MIN_RENDER_ALPHA (1 / 255.0f) /* common.h:90 */

/* ---------------------------------------------------------------- scalar primitives */

/* shencoder.h:4 */
float gso_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
/* shencoder.h:6 */
float gso_sigmoid_dsigmoid(float s) { return s * (1.0f - s); }

/* kernels.h:172-193 kernel_gaussian_2d_float (covariance form, c1 and c2 kept apart, radial<0 -> 1000,
 * double-literal -0.5 multiply folded to float before expf). */
float gso_gaussian_2d(const float *mean, const float *cov, const float *query) {
  float c0 = cov[0], c1 = cov[1], c2 = cov[2], c3 = cov[3];
  float det = c0 * c3 - c1 * c2;
  float x = query[0] - mean[0];
  float y = query[1] - mean[1];
  float tmpx = x * c3 - y * c2;
  float tmpy = -x * c1 + y * c0;
  float radial = tmpx * x + tmpy * y;
  radial /= det;
  if (radial < 0.0) radial = 1000.0f;
  return expf((float)(-0.5 * (double)radial));
}

/* kernels.h:394-418 kernel_gaussian_2d_backward: FP64 inside; outputs are the per-call addends
 * (grad already carries the Gaussian value). out[0..1] = d mean, out[2..5] = d cov. */
void gso_gaussian_2d_backward(const float *mean, const float *cov, const float *query, float grad,
                              double *out) {
  double d_grad = (double)grad;
  double c0 = cov[0], c1 = cov[1], c2 = cov[2], c3 = cov[3];
  double det = c0 * c3 - c1 * c2;
  double x = (double)(float)(query[0] - mean[0]); /* FP32 subtraction, then promoted */
  double y = (double)(float)(query[1] - mean[1]);
  double tmpx = (x * c3 - y * c2) / det;
  double tmpy = (-x * c1 + y * c0) / det;
  out[0] = (double)(float)(d_grad * tmpx);
  out[1] = (double)(float)(d_grad * tmpy);
  out[2] = (double)(float)(0.5 * (float)(d_grad * tmpx * tmpx));
  out[3] = (double)(float)(0.5 * (float)(d_grad * tmpx * tmpy));
  out[4] = (double)(float)(0.5 * (float)(d_grad * tmpy * tmpx));
  out[5] = (double)(float)(0.5 * (float)(d_grad * tmpy * tmpy));
}

/* shencoder.h:13-55 spherical_harmonic, degrees C = 1..4 (the bindings dispatch C in 1..4 only,
 * render.cu:506-543). */
void gso_spherical_harmonic(const float *direction, float *outputs, uint32_t C) {
  float x = direction[0], y = direction[1], z = direction[2];
  float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  outputs[0] = 0.28209479177387814f;
  if (C <= 1) return;
  outputs[1] = -0.48860251190291987f * y;
  outputs[2] = 0.48860251190291987f * z;
  outputs[3] = -0.48860251190291987f * x;
  if (C <= 2) return;
  outputs[4] = 1.0925484305920792f * xy;
  outputs[5] = -1.0925484305920792f * yz;
  outputs[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  outputs[7] = -1.0925484305920792f * xz;
  outputs[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  if (C <= 3) return;
  outputs[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
  outputs[10] = 2.8906114426405538f * xy * z;
  outputs[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
  outputs[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
  outputs[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
  outputs[14] = 1.4453057213202769f * z * (x2 - y2);
  outputs[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

/* vol_render_sh.h:48-65 calc_direction: the [3,4] row-major c2w is read as three consecutive
 * float3 (flat elements 0-2, 3-5, 6-8) -- quirk Q1 -- then normalised. pos = (px, py, 1). */
void gso_calc_direction(float *direction, const float *pos, const float *c2w) {
  for (int k = 0; k < 3; ++k)
    direction[k] = c2w[3 * k + 0] * pos[0] + c2w[3 * k + 1] * pos[1] + c2w[3 * k + 2] * pos[2];
  float length = sqrtf(direction[0] * direction[0] + direction[1] * direction[1] +
                       direction[2] * direction[2]);
  direction[0] /= length;
  direction[1] /= length;
  direction[2] /= length;
}

/* ---------------------------------------------------------------- frustum cull */

/* culling.h:11-20 + kernels.h:156-170: keep iff for all 6 planes dot(mean - pt_k, n_k) > -r,
 * r = max(svec) * thresh.  qvec is passed by the reference but unused.  helper_math dot order. */
void gso_culling_gaussian_bsphere(uint32_t N, const float *mean, const float *svec,
                                  const float *normal, const float *pts, uint8_t *mask,
                                  float thresh) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)N; ++i) {
    const float *m = mean + 3 * i, *s = svec + 3 * i;
    float r = fmaxf(fmaxf(s[0], s[1]), s[2]) * thresh;
    uint8_t keep = 1;
    for (int k = 0; k < 6; ++k) {
      float dx = m[0] - pts[3 * k], dy = m[1] - pts[3 * k + 1], dz = m[2] - pts[3 * k + 2];
      float d = dx * normal[3 * k] + dy * normal[3 * k + 1] + dz * normal[3 * k + 2];
      if (!(d > -r)) {
        keep = 0;
        break;
      }
    }
    mask[i] = keep;
  }
}

/* ---------------------------------------------------------------- binning */

/* Stable LSD radix sort of signed 64-bit keys with a 32-bit payload; the semantics of
 * cub::DeviceRadixSort::SortPairs<int64,int> over all 64 bits (aabb_culling.h:235-241). */
static void gso_sort_pairs(uint64_t n, int64_t *keys, int32_t *vals, int64_t *keys_tmp,
                           int32_t *vals_tmp) {
  int64_t *ksrc = keys, *kdst = keys_tmp;
  int32_t *vsrc = vals, *vdst = vals_tmp;
  for (int pass = 0; pass < 8; ++pass) {
    uint64_t hist[256];
    memset(hist, 0, sizeof hist);
    int shift = 8 * pass;
    for (uint64_t i = 0; i < n; ++i) {
      uint64_t u = (uint64_t)ksrc[i] ^ 0x8000000000000000ull;
      hist[(u >> shift) & 255]++;
    }
    uint64_t sum = 0;
    for (int b = 0; b < 256; ++b) {
      uint64_t c = hist[b];
      hist[b] = sum;
      sum += c;
    }
    for (uint64_t i = 0; i < n; ++i) {
      uint64_t u = (uint64_t)ksrc[i] ^ 0x8000000000000000ull;
      uint64_t p = hist[(u >> shift) & 255]++;
      kdst[p] = ksrc[i];
      vdst[p] = vsrc[i];
    }
    int64_t *kt = ksrc; ksrc = kdst; kdst = kt;
    int32_t *vt = vsrc; vsrc = vdst; vdst = vt;
  }
  /* 8 passes: data ends in the original buffers */
}

/* aabb_culling.h:15-41 (emit; x outer, y inner; key = tile<<32 | depth bits, quirk Q9),
 * :235-241 (stable sort), :70-103 (start/end, -1 for empty tiles), :192-260 (host sequence).
 * The reference emits in atomic order; this oracle emits in ascending Gaussian id, which is one of
 * the orders the reference can produce.  Returns the number of emitted pairs (must equal n_dub,
 * aabb_culling.h:228) or -1 on mismatch.  sorted_keys may be NULL. */
int64_t gso_tile_culling_aabb_start_end(uint32_t N, uint32_t n_dub, uint32_t n_tiles_h,
                                        uint32_t n_tiles_w, const int32_t *aabb_topleft,
                                        const int32_t *aabb_bottomright, const float *depth,
                                        int32_t *gaussian_ids, int32_t *start, int32_t *end,
                                        int64_t *sorted_keys) {
  uint64_t cap = n_dub ? n_dub : 1;
  int64_t *keys = (int64_t *)malloc(sizeof(int64_t) * cap);
  int64_t *keys_tmp = (int64_t *)malloc(sizeof(int64_t) * cap);
  int32_t *vals_tmp = (int32_t *)malloc(sizeof(int32_t) * cap);
  uint64_t pos = 0;
  int overflow = 0;
  for (uint32_t i = 0; i < N && !overflow; ++i) {
    int sx = aabb_topleft[2 * i], sy = aabb_topleft[2 * i + 1];
    int ex = aabb_bottomright[2 * i], ey = aabb_bottomright[2 * i + 1];
    uint32_t dbits;
    memcpy(&dbits, depth + i, 4);
    for (int x = sx; x <= ex && !overflow; ++x)
      for (int y = sy; y <= ey; ++y) {
        if (pos >= n_dub) { overflow = 1; break; }
        int32_t tile = y * (int)n_tiles_w + x;
        keys[pos] = (int64_t)(((uint64_t)(uint32_t)tile << 32) | dbits);
        gaussian_ids[pos] = (int32_t)i;
        ++pos;
      }
  }
  int64_t ret = (overflow || pos != n_dub) ? -1 : (int64_t)pos;
  if (ret >= 0) {
    gso_sort_pairs(pos, keys, gaussian_ids, keys_tmp, vals_tmp);
    uint32_t n_tiles = n_tiles_h * n_tiles_w;
    for (uint32_t t = 0; t < n_tiles; ++t) start[t] = end[t] = -1;
    for (uint64_t g = 0; g < pos; ++g) {
      int32_t tile = (int32_t)(keys[g] >> 32);
      if (g == 0 || (int32_t)(keys[g - 1] >> 32) != tile) start[tile] = (int32_t)g;
      if (g == pos - 1 || (int32_t)(keys[g + 1] >> 32) != tile) end[tile] = (int32_t)(g + 1);
    }
    if (sorted_keys) memcpy(sorted_keys, keys, sizeof(int64_t) * pos);
  }
  free(keys); free(keys_tmp); free(vals_tmp);
  return ret;
}

/* ---------------------------------------------------------------- SH compositing, forward */

/* vol_render_sh.h:171-248 (entry) + :97-169 (batch loop); vol_render_bg.h:12-100 when with_bg.
 * Caller pre-zeroes out (renderer.py:693).  Optional per-pixel diagnostics (may be NULL):
 *   final_T      transmittance after the walk,
 *   n_contrib    1 + list index of the last Gaussian that passed the 1/255 test (0 = none),
 *   margin       min over evaluated pairs of |alpha*G*255 - 1|  (how close the pixel came to
 *                flipping a skip decision; used by tests to separate FP-fragile pixels).
 * Staging batches of the reference do not change per-thread results and are not restated. */
void gso_tile_based_vol_rendering_sh(
    const float *mean, const float *cov, const float *sh_coeffs, const float *alpha,
    const int32_t *start, const int32_t *end, const int32_t *gaussian_ids, float *out_rgb,
    const float *topleft, const float *c2w, uint32_t tile_size, uint32_t n_tiles_h,
    uint32_t n_tiles_w, float pixel_size_x, float pixel_size_y, uint32_t H, uint32_t W, uint32_t C,
    float thresh, int with_bg, const float *bg_rgb, float *final_T, int32_t *n_contrib,
    float *margin) {
  const uint32_t CC = C * C;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t tile_id = 0; tile_id < (int64_t)n_tiles_h * n_tiles_w; ++tile_id) {
    uint32_t by = (uint32_t)(tile_id / n_tiles_w), bx = (uint32_t)(tile_id % n_tiles_w);
    int empty = (start[tile_id] == -1);
    int n_this = empty ? 0 : end[tile_id] - start[tile_id];
    const int32_t *ids = gaussian_ids + (empty ? 0 : start[tile_id]);
    for (uint32_t ly = 0; ly < tile_size; ++ly)
      for (uint32_t lx = 0; lx < tile_size; ++lx) {
        uint32_t gy = by * tile_size + ly, gx = bx * tile_size + lx;
        if (gy >= H || gx >= W) continue;
        uint64_t pix = (uint64_t)gy * W + gx;
        if (final_T) final_T[pix] = 1.0f;
        if (n_contrib) n_contrib[pix] = 0;
        if (margin) margin[pix] = 1.0f;
        if (empty) {
          if (with_bg) { /* vol_render_bg.h:28-37, intent of the undeclared-`out` lines */
            out_rgb[3 * pix + 0] = bg_rgb[0];
            out_rgb[3 * pix + 1] = bg_rgb[1];
            out_rgb[3 * pix + 2] = bg_rgb[2];
          }
          continue;
        }
        if (n_this == 0) continue;
        float pos[3] = {topleft[0] + gx * pixel_size_x, topleft[1] + gy * pixel_size_y, 1.0f};
        float direction[3], sh_consts[16];
        gso_calc_direction(direction, pos, c2w);
        gso_spherical_harmonic(direction, sh_consts, C);
        float out[3] = {0.0f, 0.0f, 0.0f};
        float cum_alpha = 1.0f;
        float mrg = 1.0f;
        int last = 0;
        for (int i = 0; i < n_this; ++i) {
          if (cum_alpha < thresh) break;
          int32_t g = ids[i];
          float alpha_ = fminf(alpha[g], 0.99f);
          float coeff = alpha_ * cum_alpha;
          float val = gso_gaussian_2d(mean + 2 * (int64_t)g, cov + 4 * (int64_t)g, pos);
          coeff *= val;
          float m = fabsf(alpha_ * val * 255.0f - 1.0f);
          if (m < mrg) mrg = m;
          if (alpha_ * val < GSO_MIN_RENDER_ALPHA) continue;
          if (isnan(coeff)) coeff = 0.0f;
          float y[3];
          for (int c = 0; c < 3; ++c) {
            const float *h = sh_coeffs + ((int64_t)g * 3 + c) * CC;
            float s = 0.0f;
            for (uint32_t k = 0; k < CC; ++k) s += h[k] * sh_consts[k];
            y[c] = gso_sigmoid(s);
            if (isnan(y[c] * coeff)) y[c] = 0.0f;
          }
          out[0] += coeff * y[0];
          out[1] += coeff * y[1];
          out[2] += coeff * y[2];
          cum_alpha *= (1 - alpha_ * val);
          last = i + 1;
        }
        if (with_bg && cum_alpha > thresh) { /* vol_render_bg.h:90-94 */
          out[0] = out[0] * cum_alpha + bg_rgb[0] * (1.0f - cum_alpha);
          out[1] = out[1] * cum_alpha + bg_rgb[1] * (1.0f - cum_alpha);
          out[2] = out[2] * cum_alpha + bg_rgb[2] * (1.0f - cum_alpha);
        }
        out_rgb[3 * pix + 0] = out[0];
        out_rgb[3 * pix + 1] = out[1];
        out_rgb[3 * pix + 2] = out[2];
        if (final_T) final_T[pix] = cum_alpha;
        if (n_contrib) n_contrib[pix] = last;
        if (margin) margin[pix] = mrg;
      }
  }
}

/* ---------------------------------------------------------------- SH compositing, backward */

static inline void gso_atomic_add(double *p, double v) {
#pragma omp atomic
  *p += v;
}

/* vol_render_sh.h:353-455 (entry) + :268-351 (batch loop) + :28-36 (backward_C) +
 * kernels.h:394-418.  The reference accumulates FP32 atomics in unspecified order; the oracle
 * accumulates the same FP32 addends in FP64 and rounds once (the "order-free" value the
 * 1e-3 gradient tolerance is measured against).  out_rgb is the saved forward image (`final`).
 * The bg variant (vol_render_bg.h:121-234) runs the same math on the blended image.
 * Gradients are ADDED to the caller's (pre-zeroed) buffers like the reference. */
void gso_tile_based_vol_rendering_backward_sh(
    uint32_t N, const float *mean, const float *cov, const float *sh_coeffs, const float *alpha,
    const int32_t *start, const int32_t *end, const int32_t *gaussian_ids, const float *out_rgb,
    float *grad_mean, float *grad_cov, float *grad_sh_coeffs, float *grad_alpha,
    const float *grad_out_rgb, const float *topleft, const float *c2w, uint32_t tile_size,
    uint32_t n_tiles_h, uint32_t n_tiles_w, float pixel_size_x, float pixel_size_y, uint32_t H,
    uint32_t W, uint32_t C, float thresh) {
  const uint32_t CC = C * C;
  const uint32_t row = 7 + 3 * CC;
  double *acc = (double *)calloc((size_t)N * row, sizeof(double));
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t tile_id = 0; tile_id < (int64_t)n_tiles_h * n_tiles_w; ++tile_id) {
    uint32_t by = (uint32_t)(tile_id / n_tiles_w), bx = (uint32_t)(tile_id % n_tiles_w);
    if (start[tile_id] == -1) continue;
    int n_this = end[tile_id] - start[tile_id];
    if (n_this == 0) continue;
    const int32_t *ids = gaussian_ids + start[tile_id];
    for (uint32_t ly = 0; ly < tile_size; ++ly)
      for (uint32_t lx = 0; lx < tile_size; ++lx) {
        uint32_t gy = by * tile_size + ly, gx = bx * tile_size + lx;
        if (gy >= H || gx >= W) continue;
        uint64_t pix = (uint64_t)gy * W + gx;
        float pos[3] = {topleft[0] + gx * pixel_size_x, topleft[1] + gy * pixel_size_y, 1.0f};
        float direction[3], sh_consts[16];
        gso_calc_direction(direction, pos, c2w);
        gso_spherical_harmonic(direction, sh_consts, C);
        float g_out[3], final[3];
        for (int c = 0; c < 3; ++c) {
          g_out[c] = grad_out_rgb[3 * pix + c];
          final[c] = out_rgb[3 * pix + c];
        }
        float out[3] = {0.0f, 0.0f, 0.0f};
        float cum_alpha = 1.0f;
        for (int i = 0; i < n_this; ++i) {
          if (cum_alpha < thresh) break;
          int32_t g = ids[i];
          float alpha_ = fminf(alpha[g], 0.99f);
          float G = gso_gaussian_2d(mean + 2 * (int64_t)g, cov + 4 * (int64_t)g, pos);
          if (alpha_ * G < GSO_MIN_RENDER_ALPHA) continue;
          float coeff = alpha_ * cum_alpha * G;
          if (isnan(coeff)) coeff = 0.0f;
          float y[3];
          for (int c = 0; c < 3; ++c) {
            const float *h = sh_coeffs + ((int64_t)g * 3 + c) * CC;
            float s = 0.0f;
            for (uint32_t k = 0; k < CC; ++k) s += h[k] * sh_consts[k];
            y[c] = gso_sigmoid(s);
            if (isnan(y[c] * coeff)) y[c] = 0.0f;
          }
          out[0] += coeff * y[0];
          out[1] += coeff * y[1];
          out[2] += coeff * y[2];
          double *a = acc + (size_t)g * row;
          for (int c = 0; c < 3; ++c) {
            float gc = coeff * gso_sigmoid_dsigmoid(y[c]) * g_out[c];
            for (uint32_t k = 0; k < CC; ++k)
              gso_atomic_add(a + 7 + c * CC + k, (double)(float)(gc * sh_consts[k]));
          }
          float partial_aG = 0.0f;
          for (int c = 0; c < 3; ++c)
            partial_aG += g_out[c] * (y[c] * cum_alpha - (final[c] - out[c]) / (1 - alpha_ * G));
          double gg[6];
          gso_gaussian_2d_backward(mean + 2 * (int64_t)g, cov + 4 * (int64_t)g, pos,
                                   partial_aG * alpha_ * G, gg);
          for (int k = 0; k < 6; ++k) gso_atomic_add(a + k, gg[k]);
          gso_atomic_add(a + 6, (double)(float)(partial_aG * G));
          cum_alpha *= (1 - alpha_ * G);
        }
      }
  }
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < (int64_t)N; ++g) {
    const double *a = acc + (size_t)g * row;
    grad_mean[2 * g + 0] += (float)a[0];
    grad_mean[2 * g + 1] += (float)a[1];
    for (int k = 0; k < 4; ++k) grad_cov[4 * g + k] += (float)a[2 + k];
    grad_alpha[g] += (float)a[6];
    for (uint32_t k = 0; k < 3 * CC; ++k) grad_sh_coeffs[(size_t)g * 3 * CC + k] += (float)a[7 + k];
  }
  free(acc);
}


/* ---------------------------------------------------------------- legacy RGB compositing (8f rank 2) */

/* kernels.h:195-214 kernel_gaussian_2d: every operand widened to FP64 first (the query - mean
 * subtraction is FP32, then promoted), exp in double, one final cast. */
float gso_gaussian_2d_f64(const float *mean, const float *cov, const float *query) {
  double c0 = cov[0], c1 = cov[1], c2 = cov[2], c3 = cov[3];
  double det = c0 * c3 - c1 * c2;
  double x = (double)(float)(query[0] - mean[0]);
  double y = (double)(float)(query[1] - mean[1]);
  double tmpx = x * c3 - y * c2;
  double tmpy = -x * c1 + y * c0;
  double radial = tmpx * x + tmpy * y;
  radial /= det;
  if (radial < 0.0) radial = 1000.0;
  return (float)exp(-0.5 * radial);
}

/* vol_render.h:716-798 (entry) + :169-250 (batch loop, same arithmetic as :87-167).  Caller pre-zeroes
 * out (renderer.py:558).  margin: as in the SH oracle. */
void gso_tile_based_vol_rendering_start_end(
    const float *mean, const float *cov, const float *color, const float *alpha, const int32_t *start,
    const int32_t *end, const int32_t *gaussian_ids, float *out_rgb, const float *topleft,
    uint32_t tile_size, uint32_t n_tiles_h, uint32_t n_tiles_w, float pixel_size_x, float pixel_size_y,
    uint32_t H, uint32_t W, float thresh, float *margin) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t tile_id = 0; tile_id < (int64_t)n_tiles_h * n_tiles_w; ++tile_id) {
    uint32_t by = (uint32_t)(tile_id / n_tiles_w), bx = (uint32_t)(tile_id % n_tiles_w);
    if (start[tile_id] == -1) continue;
    int n_this = end[tile_id] - start[tile_id];
    if (n_this == 0) continue;
    const int32_t *ids = gaussian_ids + start[tile_id];
    for (uint32_t ly = 0; ly < tile_size; ++ly)
      for (uint32_t lx = 0; lx < tile_size; ++lx) {
        uint32_t gy = by * tile_size + ly, gx = bx * tile_size + lx;
        if (gy >= H || gx >= W) continue;
        uint64_t pix = (uint64_t)gy * W + gx;
        float pos[2] = {topleft[0] + gx * pixel_size_x, topleft[1] + gy * pixel_size_y};
        float out[3] = {0.0f, 0.0f, 0.0f};
        float cum_alpha = 1.0f, mrg = 1.0f;
        for (int i = 0; i < n_this; ++i) {
          if (cum_alpha < thresh) break;
          int32_t g = ids[i];
          float alpha_ = fminf(alpha[g], 0.99f);
          float coeff = alpha_ * cum_alpha;
          float val = gso_gaussian_2d_f64(mean + 2 * (int64_t)g, cov + 4 * (int64_t)g, pos);
          coeff *= val;
          float m = fabsf(alpha_ * val * 255.0f - 1.0f);
          if (m < mrg) mrg = m;
          if (alpha_ * val < GSO_MIN_RENDER_ALPHA) continue;
          out[0] += color[3 * (int64_t)g + 0] * coeff;
          out[1] += color[3 * (int64_t)g + 1] * coeff;
          out[2] += color[3 * (int64_t)g + 2] * coeff;
          cum_alpha *= (1 - alpha_ * val);
        }
        out_rgb[3 * pix + 0] = out[0];
        out_rgb[3 * pix + 1] = out[1];
        out_rgb[3 * pix + 2] = out[2];
        if (margin) margin[pix] = mrg;
      }
  }
}

/* vol_render.h:800-923 (entry) + :252-352 (batch loop): grad_color += a*T*G*grad_out; partial_aG is a
 * DOUBLE accumulator there (:309-325); kernel_gaussian_2d_backward and grad_alpha take its float cast.
 * FP32 addends summed in FP64 and rounded once, like the SH oracle. */
void gso_tile_based_vol_rendering_backward_start_end(
    uint32_t N, const float *mean, const float *cov, const float *color, const float *alpha,
    const int32_t *start, const int32_t *end, const int32_t *gaussian_ids, const float *out_rgb,
    float *grad_mean, float *grad_cov, float *grad_color, float *grad_alpha, const float *grad_out_rgb,
    const float *topleft, uint32_t tile_size, uint32_t n_tiles_h, uint32_t n_tiles_w, float pixel_size_x,
    float pixel_size_y, uint32_t H, uint32_t W, float thresh) {
  const uint32_t row = 10;
  double *acc = (double *)calloc((size_t)N * row, sizeof(double));
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t tile_id = 0; tile_id < (int64_t)n_tiles_h * n_tiles_w; ++tile_id) {
    uint32_t by = (uint32_t)(tile_id / n_tiles_w), bx = (uint32_t)(tile_id % n_tiles_w);
    if (start[tile_id] == -1) continue;
    int n_this = end[tile_id] - start[tile_id];
    if (n_this == 0) continue;
    const int32_t *ids = gaussian_ids + start[tile_id];
    for (uint32_t ly = 0; ly < tile_size; ++ly)
      for (uint32_t lx = 0; lx < tile_size; ++lx) {
        uint32_t gy = by * tile_size + ly, gx = bx * tile_size + lx;
        if (gy >= H || gx >= W) continue;
        uint64_t pix = (uint64_t)gy * W + gx;
        float pos[2] = {topleft[0] + gx * pixel_size_x, topleft[1] + gy * pixel_size_y};
        float g_out[3], final[3];
        for (int c = 0; c < 3; ++c) {
          g_out[c] = grad_out_rgb[3 * pix + c];
          final[c] = out_rgb[3 * pix + c];
        }
        float out[3] = {0.0f, 0.0f, 0.0f};
        float cum_alpha = 1.0f;
        for (int i = 0; i < n_this; ++i) {
          if (cum_alpha < thresh) break;
          int32_t g = ids[i];
          float alpha_ = fminf(alpha[g], 0.99f);
          float G = gso_gaussian_2d_f64(mean + 2 * (int64_t)g, cov + 4 * (int64_t)g, pos);
          if (alpha_ * G < GSO_MIN_RENDER_ALPHA) continue;
          float coeff = alpha_ * cum_alpha * G;
          double *a = acc + (size_t)g * row;
          double partial_aG = 0.0;
          for (int c = 0; c < 3; ++c) {
            float col = color[3 * (int64_t)g + c];
            out[c] += col * coeff;
            gso_atomic_add(a + 7 + c, (double)(float)(coeff * g_out[c]));
            partial_aG += (double)((col * cum_alpha - (final[c] - out[c]) / (1 - alpha_ * G)) * g_out[c]);
          }
          double gg[6];
          gso_gaussian_2d_backward(mean + 2 * (int64_t)g, cov + 4 * (int64_t)g, pos,
                                   (float)(partial_aG * alpha_ * G), gg);
          for (int k = 0; k < 6; ++k) gso_atomic_add(a + k, gg[k]);
          gso_atomic_add(a + 6, (double)(float)(partial_aG * G));
          cum_alpha *= (1 - alpha_ * G);
        }
      }
  }
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < (int64_t)N; ++g) {
    const double *a = acc + (size_t)g * row;
    grad_mean[2 * g + 0] += (float)a[0];
    grad_mean[2 * g + 1] += (float)a[1];
    for (int k = 0; k < 4; ++k) grad_cov[4 * g + k] += (float)a[2 + k];
    grad_alpha[g] += (float)a[6];
    for (int k = 0; k < 3; ++k) grad_color[3 * g + k] += (float)a[7 + k];
  }
  free(acc);
}
