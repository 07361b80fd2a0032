// The loss step between K3 and K4a: utils/loss.py:5-24
//
//     loss = ssim_loss_mult * ssim_loss(out, gt, win, reduction="mean") + (1 - ssim_loss_mult) * base(out, gt)
//
// with base = mse_loss | l1_loss and kornia's ssim_loss (kornia is un-vendored and unpinned,
// requirements.txt:5; 0.6.x semantics restated: separable Gaussian window (sigma 1.5) with REFLECT border,
// C1 = 0.01^2, C2 = 0.03^2, ssim = num / (den + 1e-12), loss map = clamp((1 - ssim) / 2, 0, 1), mean over
// [1,3,H,W]).  Forward value AND d loss / d out in three HBM-bound passes over the image instead of
// ~60 ATen launches (5 padded convolutions forward, their transposes backward, ~40 elementwise ops):
//
//   1  ssim_stats_kernel  : per 16x16 tile (5-pixel halo through shared memory) the five filtered moments,
//                           the ssim / base loss partial sums, and the three derivative maps
//                           dL/d filt(x), dL/d filt(x^2), dL/d filt(x*y)
//   2  adjoint_conv_kernel: zero-padded separable correlation of the three maps on the EXTENDED domain
//                           (H+2R) x (W+2R)  -- the transpose of "conv" ...
//   3  fold_grad_kernel   : ... and the transpose of "reflect pad": every pixel gathers its own value plus
//                           its mirror images, then grad = adj(gA) + 2x adj(gB) + y adj(gC) + base term
//   4  loss_reduce_kernel : fixed-order sum of the per-tile partials (deterministic), one float out
//
// Image layout is the renderer's: HWC float32, pixel (x, y) channel c at 3 (y W + x) + c.
#include <math.h>

#include "common.cuh"

namespace gs3d {

constexpr int LT = 16;           // tile edge
constexpr int LR = 5;            // max window radius (window 11)
constexpr int LH = LT + 2 * LR;  // tile edge with halo

struct LossParams {
  const float *x, *y;  // out, gt
  uint32_t H, W;
  int base;  // 1 = l1, 2 = l2
  float mult;
  int R;  // window radius
  float w[2 * LR + 1];
  float inv_n;  // 1 / (3 H W)
  float *gA, *gB, *gC;       // derivative maps [H, W, 3]
  float *eA, *eB, *eC;       // adjoint correlations on the extended domain [(H+2R), (W+2R), 3]
  double *partials;          // [n_tiles][2]: ssim-loss sum, base-loss sum
  float *grad;               // [H, W, 3] or null
  float *loss;               // 1 float
  uint32_t tiles_x, tiles_y;
};

__device__ __forceinline__ int reflect_idx(int j, int n) {  // torch "reflect": -1 -> 1, n -> n-2
  if (j < 0) j = -j;
  if (j >= n) j = 2 * (n - 1) - j;
  return j;
}

__global__ void __launch_bounds__(256) ssim_stats_kernel(const LossParams p) {
  __shared__ float s_x[LH][LH * 3 + 1];
  __shared__ float s_y[LH][LH * 3 + 1];
  __shared__ float s_h[5][LH][LT + 1];
  __shared__ double s_red[2][8];
  const int R = p.R, E = LT + 2 * R;
  const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
  // stage the tile + halo of both images, reflect border, all three channels (coalesced along x*3+c)
  for (int e = threadIdx.x; e < E * E * 3; e += 256) {
    const int row = e / (E * 3), col = e - row * (E * 3);
    const int px = col / 3, c = col - 3 * px;
    const int gy = reflect_idx(y0 + row - R, (int)p.H), gx = reflect_idx(x0 + px - R, (int)p.W);
    const bool ok = gy >= 0 && gy < (int)p.H && gx >= 0 && gx < (int)p.W;  // tiles past the image edge
    const size_t a = 3 * ((size_t)gy * p.W + gx) + c;
    s_x[row][col] = ok ? p.x[a] : 0.0f;
    s_y[row][col] = ok ? p.y[a] : 0.0f;
  }
  __syncthreads();
  const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
  const int gx = x0 + lx, gy = y0 + ly;
  const bool inside = gx < (int)p.W && gy < (int)p.H;
  double ssim_sum = 0.0, base_sum = 0.0;
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  for (int c = 0; c < 3; ++c) {
    // horizontal pass: E rows x LT columns, five moments
    for (int e = threadIdx.x; e < E * LT; e += 256) {
      const int row = e / LT, col = e - row * LT;
      float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
      for (int k = 0; k <= 2 * R; ++k) {
        const float w = p.w[k];
        const float xv = s_x[row][3 * (col + k) + c], yv = s_y[row][3 * (col + k) + c];
        a = fmaf(w, xv, a);
        b = fmaf(w, yv, b);
        aa = fmaf(w, xv * xv, aa);
        bb = fmaf(w, yv * yv, bb);
        ab = fmaf(w, xv * yv, ab);
      }
      s_h[0][row][col] = a; s_h[1][row][col] = b; s_h[2][row][col] = aa; s_h[3][row][col] = bb;
      s_h[4][row][col] = ab;
    }
    __syncthreads();
    if (inside) {
      float mu1 = 0.f, mu2 = 0.f, m11 = 0.f, m22 = 0.f, m12 = 0.f;
      for (int k = 0; k <= 2 * R; ++k) {
        const float w = p.w[k];
        mu1 = fmaf(w, s_h[0][ly + k][lx], mu1);
        mu2 = fmaf(w, s_h[1][ly + k][lx], mu2);
        m11 = fmaf(w, s_h[2][ly + k][lx], m11);
        m22 = fmaf(w, s_h[3][ly + k][lx], m22);
        m12 = fmaf(w, s_h[4][ly + k][lx], m12);
      }
      const float s1 = m11 - mu1 * mu1, s2 = m22 - mu2 * mu2, s12 = m12 - mu1 * mu2;
      const float n1 = 2.0f * mu1 * mu2 + C1, n2 = 2.0f * s12 + C2;
      const float d1 = mu1 * mu1 + mu2 * mu2 + C1, d2 = s1 + s2 + C2;
      const float num = n1 * n2, D = d1 * d2 + 1e-12f;
      const float ssim = num / D;
      const float l = 0.5f * (1.0f - ssim);
      ssim_sum += (double)fminf(fmaxf(l, 0.0f), 1.0f);
      // torch.clamp passes the gradient where min <= value <= max
      const float dl = (l >= 0.0f && l <= 1.0f) ? -0.5f * p.mult * p.inv_n : 0.0f;  // dL / d ssim
      const float invD = 1.0f / D;
      const float dA = ((2.0f * mu2 * (n2 - n1)) - ssim * (2.0f * mu1 * (d2 - d1))) * invD;
      const float dB = -ssim * d1 * invD;
      const float dC = 2.0f * n1 * invD;
      const size_t o = 3 * ((size_t)gy * p.W + gx) + c;
      p.gA[o] = dl * dA;
      p.gB[o] = dl * dB;
      p.gC[o] = dl * dC;
      const float diff = s_x[ly + R][3 * (lx + R) + c] - s_y[ly + R][3 * (lx + R) + c];
      base_sum += p.base == 2 ? (double)(diff * diff) : (double)fabsf(diff);
    }
    __syncthreads();
  }
  // block reduction (fixed order inside the block: shuffles, then warp 0 adds the eight warp sums)
  for (int o = 16; o > 0; o >>= 1) {
    ssim_sum += __shfl_xor_sync(0xffffffffu, ssim_sum, o);
    base_sum += __shfl_xor_sync(0xffffffffu, base_sum, o);
  }
  if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = ssim_sum; s_red[1][threadIdx.x >> 5] = base_sum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
    const size_t t = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    p.partials[2 * t] = a;
    p.partials[2 * t + 1] = b;
  }
}

// Zero-padded separable correlation of one derivative map, evaluated on the extended domain
// [-R, H+R) x [-R, W+R): out_ext[j] = sum_q w[q - j + R] d[q] over pixels q inside the image.
__global__ void __launch_bounds__(256) adjoint_conv_kernel(const LossParams p) {
  __shared__ float s_d[LH][LH * 3 + 1];
  __shared__ float s_h[LH][LT * 3 + 1];
  const int R = p.R, E = LT + 2 * R;
  const float *src = blockIdx.z == 0 ? p.gA : (blockIdx.z == 1 ? p.gB : p.gC);
  float *dst = blockIdx.z == 0 ? p.eA : (blockIdx.z == 1 ? p.eB : p.eC);
  const int EW = (int)p.W + 2 * R, EH = (int)p.H + 2 * R;
  // extended coordinates of this tile's outputs: ex = ex0 + lx, image coordinate = ex - R
  const int ex0 = blockIdx.x * LT, ey0 = blockIdx.y * LT;
  for (int e = threadIdx.x; e < E * E * 3; e += 256) {
    const int row = e / (E * 3), col = e - row * (E * 3);
    const int px = col / 3, c = col - 3 * px;
    const int iy = ey0 + row - 2 * R, ix = ex0 + px - 2 * R;  // image coords of the input needed
    const bool ok = iy >= 0 && iy < (int)p.H && ix >= 0 && ix < (int)p.W;
    s_d[row][col] = ok ? src[3 * ((size_t)iy * p.W + ix) + c] : 0.0f;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E * LT * 3; e += 256) {
    const int row = e / (LT * 3), col = e - row * (LT * 3);
    const int px = col / 3, c = col - 3 * px;
    float a = 0.f;
    for (int k = 0; k <= 2 * R; ++k) a = fmaf(p.w[k], s_d[row][3 * (px + k) + c], a);
    s_h[row][col] = a;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < LT * LT * 3; e += 256) {
    const int row = e / (LT * 3), col = e - row * (LT * 3);
    const int px = col / 3;
    const int ey = ey0 + row, ex = ex0 + px;
    if (ey >= EH || ex >= EW) continue;
    float a = 0.f;
    for (int k = 0; k <= 2 * R; ++k) a = fmaf(p.w[k], s_h[row + k][col], a);
    dst[3 * ((size_t)ey * EW + ex) + (col - 3 * px)] = a;
  }
}

// Transpose of the reflect padding: pixel i also receives what the padded positions that mirror onto it
// received (-i for 1 <= i <= R, 2(n-1)-i for n-1-R <= i <= n-2), in both dimensions.
__global__ void __launch_bounds__(256) fold_grad_kernel(const LossParams p) {
  const size_t n = (size_t)p.H * p.W * 3;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % 3);
  const size_t pix = i / 3;
  const int x = (int)(pix % p.W), y = (int)(pix / p.W);
  const float xv = p.x[i], yv = p.y[i];
  const float diff = xv - yv;
  float g = p.base == 2 ? 2.0f * diff : (diff > 0.f ? 1.0f : (diff < 0.f ? -1.0f : 0.0f));
  g *= (1.0f - p.mult) * p.inv_n;
  if (p.mult != 0.0f) {
    const int R = p.R, EW = (int)p.W + 2 * R;
    int xs[3], ys[3], nx = 0, ny = 0;
    xs[nx++] = x;
    if (x >= 1 && x <= R) xs[nx++] = -x;
    if (x >= (int)p.W - 1 - R && x <= (int)p.W - 2) xs[nx++] = 2 * ((int)p.W - 1) - x;
    ys[ny++] = y;
    if (y >= 1 && y <= R) ys[ny++] = -y;
    if (y >= (int)p.H - 1 - R && y <= (int)p.H - 2) ys[ny++] = 2 * ((int)p.H - 1) - y;
    float a = 0.f, b = 0.f, cc = 0.f;
    for (int j = 0; j < ny; ++j)
      for (int k = 0; k < nx; ++k) {
        const size_t o = 3 * ((size_t)(ys[j] + R) * EW + (xs[k] + R)) + c;
        a += p.eA[o];
        b += p.eB[o];
        cc += p.eC[o];
      }
    g += a + 2.0f * xv * b + yv * cc;
  }
  p.grad[i] = g;
}

__global__ void loss_reduce_kernel(const LossParams p) {
  __shared__ double s_a[256], s_b[256];
  const uint32_t nt = p.tiles_x * p.tiles_y;
  double a = 0.0, b = 0.0;
  for (uint32_t t = threadIdx.x; t < nt; t += 256) { a += p.partials[2 * t]; b += p.partials[2 * t + 1]; }
  s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { s_a[threadIdx.x] += s_a[threadIdx.x + o]; s_b[threadIdx.x] += s_b[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0)
    *p.loss = (float)((double)p.mult * s_a[0] * (double)p.inv_n + (double)(1.0f - p.mult) * s_b[0] * (double)p.inv_n);
}

static size_t loss_layout(uint32_t H, uint32_t W, int R, size_t *maps, size_t *ext, size_t *parts) {
  const size_t m = align_up((size_t)H * W * 3 * sizeof(float));
  const size_t e = align_up((size_t)(H + 2 * R) * (W + 2 * R) * 3 * sizeof(float));
  const size_t pt = align_up((size_t)div_up(H, (uint32_t)LT) * div_up(W, (uint32_t)LT) * 2 * sizeof(double));
  if (maps) *maps = m;
  if (ext) *ext = e;
  if (parts) *parts = pt;
  return 3 * m + 3 * e + pt + 256;
}

}  // namespace gs3d

using namespace gs3d;

extern "C" {

size_t gs3d_image_loss_scratch_bytes(uint32_t H, uint32_t W) { return loss_layout(H, W, LR, nullptr, nullptr, nullptr); }

int gs3d_image_loss(const float *out, const float *gt, uint32_t H, uint32_t W, int base_loss, float ssim_mult,
                    uint32_t window_size, float *loss, float *grad, void *scratch, size_t scratch_bytes,
                    void *stream) {
  GS3D_REQUIRE(out && gt && loss, GS3D_EINVAL, "image_loss: null argument");
  GS3D_REQUIRE(base_loss == 1 || base_loss == 2, GS3D_EINVAL, "image_loss: base_loss 1 (l1) or 2 (l2)");
  GS3D_REQUIRE(window_size % 2 == 1 && window_size >= 1 && window_size <= 2 * LR + 1, GS3D_EUNSUPPORTED,
               "image_loss: odd window sizes up to %d (got %u)", 2 * LR + 1, window_size);
  const int R = (int)window_size / 2;
  GS3D_REQUIRE(H > (uint32_t)R && W > (uint32_t)R, GS3D_EINVAL,
               "image_loss: reflect padding needs H, W > %d (got %ux%u)", R, H, W);
  GS3D_REQUIRE(scratch && scratch_bytes >= gs3d_image_loss_scratch_bytes(H, W), GS3D_EINVAL,
               "image_loss: scratch too small");
  LossParams p = {};
  p.x = out; p.y = gt; p.H = H; p.W = W; p.base = base_loss; p.mult = ssim_mult; p.R = R;
  {  // kornia gaussian(window_size, 1.5): exp(-x^2 / (2 sigma^2)) normalised, FP32
    float sum = 0.f;
    for (int k = 0; k <= 2 * R; ++k) {
      const float xk = (float)(k - R);
      p.w[k] = expf(-(xk * xk) / (2.0f * 1.5f * 1.5f));
      sum += p.w[k];
    }
    for (int k = 0; k <= 2 * R; ++k) p.w[k] /= sum;
  }
  p.inv_n = (float)(1.0 / ((double)H * W * 3));
  size_t m, e, pt;
  loss_layout(H, W, LR, &m, &e, &pt);
  char *base = static_cast<char *>(scratch);
  p.gA = reinterpret_cast<float *>(base); p.gB = reinterpret_cast<float *>(base + m);
  p.gC = reinterpret_cast<float *>(base + 2 * m);
  p.eA = reinterpret_cast<float *>(base + 3 * m); p.eB = reinterpret_cast<float *>(base + 3 * m + e);
  p.eC = reinterpret_cast<float *>(base + 3 * m + 2 * e);
  p.partials = reinterpret_cast<double *>(base + 3 * m + 3 * e);
  p.grad = grad; p.loss = loss;
  p.tiles_x = div_up(W, (uint32_t)LT); p.tiles_y = div_up(H, (uint32_t)LT);
  cudaStream_t st = as_stream(stream);
  ssim_stats_kernel<<<dim3(p.tiles_x, p.tiles_y), 256, 0, st>>>(p);
  GS3D_LAUNCH_CHECK();
  loss_reduce_kernel<<<1, 256, 0, st>>>(p);
  GS3D_LAUNCH_CHECK();
  if (grad) {
    if (ssim_mult != 0.0f) {
      adjoint_conv_kernel<<<dim3(div_up(W + 2 * R, (uint32_t)LT), div_up(H + 2 * R, (uint32_t)LT), 3), 256, 0, st>>>(p);
      GS3D_LAUNCH_CHECK();
    }
    fold_grad_kernel<<<(unsigned)div_up((size_t)H * W * 3, (size_t)256), 256, 0, st>>>(p);
    GS3D_LAUNCH_CHECK();
  }
  return GS3D_OK;
}

}  // extern "C"
