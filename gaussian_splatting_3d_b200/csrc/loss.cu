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
//   1  ssim_stats_kernel  : per 32x32 tile (5-pixel halo through shared memory) the five filtered moments,
//                           the ssim / base loss partial sums, and the three derivative maps
//                           dL/d filt(x), dL/d filt(x^2), dL/d filt(x*y); a thread owns one column of the
//                           HWC-flattened tile and walks down the rows with the horizontal moments in an
//                           11-deep REGISTER ring, so the vertical filter costs no shared-memory traffic
//   2  adjoint_conv_kernel: zero-padded separable correlation of the three maps on the EXTENDED domain
//                           (H+10) x (W+10)  -- the transpose of "conv" ... (same register-ring walk)
//   3  fold_grad_kernel   : ... and the transpose of "reflect pad": every pixel gathers its own value plus
//                           its mirror images, then grad = adj(gA) + 2x adj(gB) + y adj(gC) + base term
//   4  loss_reduce_kernel : fixed-order sum of the per-tile partials (deterministic), one float out
//
// Image layout is the renderer's: HWC float32, pixel (x, y) channel c at 3 (y W + x) + c.
#include <math.h>

#include "common.cuh"

namespace gs3d {

constexpr int LR = 5;                 // filter radius the kernels are built for (window 11; smaller windows
                                      // run as 11 taps with zero outer weights)
constexpr int NT = 2 * LR + 1;        // taps
constexpr int TW = 32, TH = 32;       // output tile: 32 pixels x 32 rows
constexpr int FW = 3 * TW;            // ... = 96 floats per row in the HWC-flattened image (x*3 + c)
constexpr int SW = FW + 6 * LR;       // staged row: halo of LR pixels = 3*LR floats each side
constexpr int SH_ROWS = TH + 2 * LR;  // staged rows
constexpr int HALF = TH / 2;          // rows produced by one thread (two thread groups per tile)
constexpr int WALK = HALF + 2 * LR;   // rows a thread walks
constexpr int LOSS_THREADS = 2 * FW;  // 192: thread = (flattened column, upper / lower half of the tile)

struct LossParams {
  const float *x, *y;  // out, gt
  uint32_t H, W;
  int base;  // 1 = l1, 2 = l2
  float mult;
  int R;          // true window radius (fold)
  float w[NT];    // 11 taps, zero outside the true window
  float inv_n;    // 1 / (3 H W)
  float *gA, *gB, *gC;       // derivative maps [H, W, 3]
  float *eA, *eB, *eC;       // adjoint correlations on the extended domain [(H+2*LR), (W+2*LR), 3]
  double *partials;          // [n_tiles][2]: ssim-loss sum, base-loss sum
  float *grad;               // [H, W, 3] or null
  float *loss;               // 1 float
  uint32_t tiles_x, tiles_y;
};

__device__ __forceinline__ int reflect_idx(int j, int n) {  // torch "reflect": -1 -> 1, n -> n-2
  if (j < 0) j = -j;
  if (j >= n) j = 2 * (n - 1) - j;
  return min(max(j, 0), n - 1);  // only reachable under zero weights (window radius < LR on a tiny image)
}

__device__ __forceinline__ int floordiv3(int v) { return v >= 0 ? v / 3 : -((-v + 2) / 3); }

// Pass 1.  Each thread owns one column of the HWC-flattened tile (x*3 + c) and walks down WALK rows: the
// horizontal 11-tap moments of a row are formed from shared memory (taps are 3 floats apart), kept in an
// 11-deep register ring, and the vertical filter reads the ring -- no second shared-memory pass.
__global__ void __launch_bounds__(LOSS_THREADS) ssim_stats_kernel(const LossParams p) {
  __shared__ float s_x[SH_ROWS][SW];
  __shared__ float s_y[SH_ROWS][SW];
  __shared__ double s_red[2][LOSS_THREADS / 32];
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int fx0 = 3 * x0 - 3 * LR;
  for (int e = threadIdx.x; e < SH_ROWS * SW; e += LOSS_THREADS) {
    const int row = e / SW, col = e - row * SW;
    const int fc = fx0 + col;
    const int px = floordiv3(fc), c = fc - 3 * px;
    const int gy = reflect_idx(y0 + row - LR, (int)p.H), gx = reflect_idx(px, (int)p.W);
    const size_t a = 3 * ((size_t)gy * p.W + gx) + c;
    s_x[row][col] = p.x[a];
    s_y[row][col] = p.y[a];
  }
  __syncthreads();
  const int col = threadIdx.x % FW, half = threadIdx.x / FW;
  const int px = col / 3, c = col - 3 * px;
  const int gx = x0 + px;
  const int r0 = half * HALF;  // first staged row this thread reads
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  float ring[5][NT];
  double ssim_sum = 0.0, base_sum = 0.0;
#pragma unroll
  for (int r = 0; r < WALK; ++r) {
    float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      const float w = p.w[k];
      const float xv = s_x[r0 + r][col + 3 * k], yv = s_y[r0 + r][col + 3 * k];
      const float wx = w * xv, wy = w * yv;
      a += wx;
      b += wy;
      aa = fmaf(wx, xv, aa);
      bb = fmaf(wy, yv, bb);
      ab = fmaf(wx, yv, ab);
    }
    ring[0][r % NT] = a; ring[1][r % NT] = b; ring[2][r % NT] = aa; ring[3][r % NT] = bb; ring[4][r % NT] = ab;
    if (r >= NT - 1) {
      const int orow = r0 + r - (NT - 1);  // output row inside the tile
      const int gy = y0 + orow;
      float mu1 = 0.f, mu2 = 0.f, m11 = 0.f, m22 = 0.f, m12 = 0.f;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float w = p.w[j];
        const int slot = (r - (NT - 1) + j) % NT;
        mu1 = fmaf(w, ring[0][slot], mu1);
        mu2 = fmaf(w, ring[1][slot], mu2);
        m11 = fmaf(w, ring[2][slot], m11);
        m22 = fmaf(w, ring[3][slot], m22);
        m12 = fmaf(w, ring[4][slot], m12);
      }
      if (gx < (int)p.W && gy < (int)p.H) {
        const float s1 = m11 - mu1 * mu1, s2 = m22 - mu2 * mu2, s12 = m12 - mu1 * mu2;
        const float n1 = 2.0f * mu1 * mu2 + C1, n2 = 2.0f * s12 + C2;
        const float d1 = mu1 * mu1 + mu2 * mu2 + C1, d2 = s1 + s2 + C2;
        const float D = d1 * d2 + 1e-12f;
        const float invD = 1.0f / D;
        const float ssim = (n1 * n2) * invD;
        const float l = 0.5f * (1.0f - ssim);
        ssim_sum += (double)fminf(fmaxf(l, 0.0f), 1.0f);
        // torch.clamp passes the gradient where min <= value <= max
        const float dl = (l >= 0.0f && l <= 1.0f) ? -0.5f * p.mult * p.inv_n : 0.0f;  // dL / d ssim
        const float dA = ((2.0f * mu2 * (n2 - n1)) - ssim * (2.0f * mu1 * (d2 - d1))) * invD;
        const float dB = -ssim * d1 * invD;
        const float dC = 2.0f * n1 * invD;
        const size_t o = 3 * ((size_t)gy * p.W + gx) + c;
        p.gA[o] = dl * dA;
        p.gB[o] = dl * dB;
        p.gC[o] = dl * dC;
        const float diff = s_x[orow + LR][col + 3 * LR] - s_y[orow + LR][col + 3 * LR];
        base_sum += p.base == 2 ? (double)(diff * diff) : (double)fabsf(diff);
      }
    }
  }
  // block reduction (fixed order: shuffles, then thread 0 adds the warp sums)
  for (int o = 16; o > 0; o >>= 1) {
    ssim_sum += __shfl_xor_sync(0xffffffffu, ssim_sum, o);
    base_sum += __shfl_xor_sync(0xffffffffu, base_sum, o);
  }
  if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = ssim_sum; s_red[1][threadIdx.x >> 5] = base_sum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
    const size_t t = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    p.partials[2 * t] = a;
    p.partials[2 * t + 1] = b;
  }
}

// Pass 2.  Zero-padded separable correlation of the three derivative maps, evaluated on the extended domain
// [-LR, H+LR) x [-LR, W+LR): ext[j] = sum_q w[j - q + LR] d[q] over pixels q inside the image (the transpose
// of the convolution; the transpose of the reflect padding is the fold in pass 3).  Same register-ring walk.
__global__ void __launch_bounds__(LOSS_THREADS) adjoint_conv_kernel(const LossParams p) {
  extern __shared__ __align__(16) float s_dyn[];  // 3 x SH_ROWS x SW floats = 62 KB (dynamic: > 48 KB)
  float(*s_d)[SH_ROWS][SW] = reinterpret_cast<float(*)[SH_ROWS][SW]>(s_dyn);
  const int EW = (int)p.W + 2 * LR, EH = (int)p.H + 2 * LR;
  const int ex0 = blockIdx.x * TW, ey0 = blockIdx.y * TH;  // extended coordinates of the tile's outputs
  const int fx0 = 3 * (ex0 - 2 * LR);                      // flattened image column of staged column 0
  for (int e = threadIdx.x; e < SH_ROWS * SW; e += LOSS_THREADS) {
    const int row = e / SW, col = e - row * SW;
    const int fc = fx0 + col;
    const int ix = floordiv3(fc), c = fc - 3 * ix;
    const int iy = ey0 + row - 2 * LR;
    const bool ok = iy >= 0 && iy < (int)p.H && ix >= 0 && ix < (int)p.W;
    const size_t a = ok ? 3 * ((size_t)iy * p.W + ix) + c : 0;
    s_d[0][row][col] = ok ? p.gA[a] : 0.0f;
    s_d[1][row][col] = ok ? p.gB[a] : 0.0f;
    s_d[2][row][col] = ok ? p.gC[a] : 0.0f;
  }
  __syncthreads();
  const int col = threadIdx.x % FW, half = threadIdx.x / FW;
  const int px = col / 3, c = col - 3 * px;
  const int ex = ex0 + px;
  const int r0 = half * HALF;
  float ring[3][NT];
#pragma unroll
  for (int r = 0; r < WALK; ++r) {
    float a = 0.f, b = 0.f, cc = 0.f;
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      const float w = p.w[k];
      a = fmaf(w, s_d[0][r0 + r][col + 3 * k], a);
      b = fmaf(w, s_d[1][r0 + r][col + 3 * k], b);
      cc = fmaf(w, s_d[2][r0 + r][col + 3 * k], cc);
    }
    ring[0][r % NT] = a; ring[1][r % NT] = b; ring[2][r % NT] = cc;
    if (r >= NT - 1) {
      const int ey = ey0 + r0 + r - (NT - 1);
      float oa = 0.f, ob = 0.f, oc = 0.f;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float w = p.w[j];
        const int slot = (r - (NT - 1) + j) % NT;
        oa = fmaf(w, ring[0][slot], oa);
        ob = fmaf(w, ring[1][slot], ob);
        oc = fmaf(w, ring[2][slot], oc);
      }
      if (ex < EW && ey < EH) {
        const size_t o = 3 * ((size_t)ey * EW + ex) + c;
        p.eA[o] = oa;
        p.eB[o] = ob;
        p.eC[o] = oc;
      }
    }
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
    const int R = p.R, EW = (int)p.W + 2 * LR;
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
        const size_t o = 3 * ((size_t)(ys[j] + LR) * EW + (xs[k] + LR)) + c;
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
  const size_t pt = align_up((size_t)div_up(H, (uint32_t)TH) * div_up(W, (uint32_t)TW) * 2 * sizeof(double));
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
    for (int k = 0; k < NT; ++k) {
      const float xk = (float)(k - LR);
      p.w[k] = (k >= LR - R && k <= LR + R) ? expf(-(xk * xk) / (2.0f * 1.5f * 1.5f)) : 0.0f;
      sum += p.w[k];
    }
    for (int k = 0; k < NT; ++k) p.w[k] /= sum;
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
  p.tiles_x = div_up(W, (uint32_t)TW); p.tiles_y = div_up(H, (uint32_t)TH);
  cudaStream_t st = as_stream(stream);
  ssim_stats_kernel<<<dim3(p.tiles_x, p.tiles_y), LOSS_THREADS, 0, st>>>(p);
  GS3D_LAUNCH_CHECK();
  loss_reduce_kernel<<<1, 256, 0, st>>>(p);
  GS3D_LAUNCH_CHECK();
  if (grad) {
    if (ssim_mult != 0.0f) {
      constexpr size_t ADJ_SMEM = (size_t)3 * SH_ROWS * SW * sizeof(float);
      GS3D_CUDA(cudaFuncSetAttribute(adjoint_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ADJ_SMEM));
      adjoint_conv_kernel<<<dim3(div_up(W + 2 * LR, (uint32_t)TW), div_up(H + 2 * LR, (uint32_t)TH)), LOSS_THREADS,
                            ADJ_SMEM, st>>>(p);
      GS3D_LAUNCH_CHECK();
    }
    fold_grad_kernel<<<(unsigned)div_up((size_t)H * W * 3, (size_t)256), 256, 0, st>>>(p);
    GS3D_LAUNCH_CHECK();
  }
  return GS3D_OK;
}

}  // extern "C"
