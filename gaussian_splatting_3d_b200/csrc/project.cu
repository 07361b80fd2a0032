// K1 / K4b: per-Gaussian kernels -- frustum planes, sphere cull, EWA projection, tile rects,
// duplicate counting, staging records, and the projection backward with ADC accumulation.
// HBM-bound streaming kernels: one thread per Gaussian, SoA inputs read with the widest loads the
// reference layouts allow (qvec as float4; mean/svec are [N,3] so three coalesced 4-byte loads).
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace gs3d {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

int64_t *pinned_mailbox() {
  static thread_local int64_t *box = nullptr;
  if (!box) {
    if (cudaMallocHost(&box, 64) != cudaSuccess) {
      box = nullptr;
      cudaGetLastError();
    }
  }
  return box;
}

// ---------------------------------------------------------------- camera constants

struct CamConst {
  float fx, fy, cx, cy;
  int w, h;
  float near_plane, far_plane;
  float half_vside, half_hside;  // far * tan(yfov/2) [* aspect], computed in double on the host
};

static CamConst make_cam(const gs3d_camera *c) {
  CamConst k;
  k.fx = (float)c->fx; k.fy = (float)c->fy; k.cx = (float)c->cx; k.cy = (float)c->cy;
  k.w = c->w; k.h = c->h;
  k.near_plane = (float)c->near_plane; k.far_plane = (float)c->far_plane;
  // utils/camera.py:225-226,255-256: yfov = 2*arctan(h / (2 fy)), aspect = w / h (Python doubles)
  double yfov = 2.0 * atan((double)c->h / (2.0 * (double)c->fy));
  double hv = (double)c->far_plane * tan(yfov * 0.5);
  double hh = hv * ((double)c->w / (double)c->h);
  k.half_vside = (float)hv;
  k.half_hside = (float)hh;
  return k;
}

__device__ __forceinline__ void cross3(const float *a, const float *b, float *o) {
  o[0] = __fsub_rn(__fmul_rn(a[1], b[2]), __fmul_rn(a[2], b[1]));
  o[1] = __fsub_rn(__fmul_rn(a[2], b[0]), __fmul_rn(a[0], b[2]));
  o[2] = __fsub_rn(__fmul_rn(a[0], b[1]), __fmul_rn(a[1], b[0]));
}

// utils/camera.py:249-283 in FP32, one op per torch op (no contraction).
__device__ void frustum_planes(const float *c2w, const CamConst cam, float *normals, float *pts) {
  float up[3], right[3], look[3], t[3];
  for (int r = 0; r < 3; ++r) {
    right[r] = c2w[4 * r + 0];
    up[r] = -c2w[4 * r + 1];
    look[r] = c2w[4 * r + 2];
    t[r] = c2w[4 * r + 3];
  }
  float nearp[3], farp[3], a[3], n[6][3];
  for (int r = 0; r < 3; ++r) {
    nearp[r] = __fmul_rn(cam.near_plane, look[r]);
    farp[r] = __fmul_rn(cam.far_plane, look[r]);
    n[0][r] = look[r];
    n[1][r] = -look[r];
  }
  for (int r = 0; r < 3; ++r) a[r] = __fsub_rn(farp[r], __fmul_rn(cam.half_hside, right[r]));
  cross3(a, up, n[2]);
  for (int r = 0; r < 3; ++r) a[r] = __fadd_rn(farp[r], __fmul_rn(cam.half_hside, right[r]));
  cross3(up, a, n[3]);
  for (int r = 0; r < 3; ++r) a[r] = __fadd_rn(farp[r], __fmul_rn(cam.half_vside, up[r]));
  cross3(a, right, n[4]);
  for (int r = 0; r < 3; ++r) a[r] = __fsub_rn(farp[r], __fmul_rn(cam.half_vside, up[r]));
  cross3(right, a, n[5]);
  for (int k = 0; k < 6; ++k) {
    float nn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(n[k][0], n[k][0]), __fmul_rn(n[k][1], n[k][1])),
                                    __fmul_rn(n[k][2], n[k][2])));
    float d = fmaxf(nn, 1e-12f);  // F.normalize eps
    for (int r = 0; r < 3; ++r) normals[3 * k + r] = __fdiv_rn(n[k][r], d);
  }
  for (int r = 0; r < 3; ++r) {
    pts[0 + r] = __fadd_rn(nearp[r], t[r]);
    pts[3 + r] = __fadd_rn(farp[r], t[r]);
    pts[6 + r] = t[r];
    pts[9 + r] = t[r];
    pts[12 + r] = t[r];
    pts[15 + r] = t[r];
  }
}

__global__ void frustum_kernel(const float *c2w, CamConst cam, float *normals, float *pts) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float c[12];
    for (int i = 0; i < 12; ++i) c[i] = c2w[i];
    frustum_planes(c, cam, normals, pts);
  }
}

// culling.h:18-19 + kernels.h:156-170 (helper_math dot: x*x + y*y + z*z, left to right).
__device__ __forceinline__ bool sphere_in_frustum(float mx, float my, float mz, float r,
                                                  const float *normal, const float *pts) {
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    float dx = mx - pts[3 * k], dy = my - pts[3 * k + 1], dz = mz - pts[3 * k + 2];
    float d = dx * normal[3 * k] + dy * normal[3 * k + 1] + dz * normal[3 * k + 2];
    if (!(d > -r)) return false;
  }
  return true;
}

__global__ void __launch_bounds__(256)
cull_bsphere_kernel(uint32_t N, const float *__restrict__ mean, const float *__restrict__ svec,
                    const float *__restrict__ normal, const float *__restrict__ pts,
                    uint8_t *__restrict__ mask, float thresh) {
  __shared__ float pl[36];
  if (threadIdx.x < 18) pl[threadIdx.x] = normal[threadIdx.x];
  else if (threadIdx.x < 36) pl[threadIdx.x] = pts[threadIdx.x - 18];
  __syncthreads();
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float sx = svec[3 * (size_t)i], sy = svec[3 * (size_t)i + 1], sz = svec[3 * (size_t)i + 2];
  float r = fmaxf(fmaxf(sx, sy), sz) * thresh;
  mask[i] = sphere_in_frustum(mean[3 * (size_t)i], mean[3 * (size_t)i + 1], mean[3 * (size_t)i + 2],
                              r, pl, pl + 18)
                ? 1 : 0;
}

// ---------------------------------------------------------------- projection (a4)

struct Projected {
  float u[3];    // camera-space mean
  float R[9];    // rotation from the normalised quaternion
  float qn[4];   // normalised quaternion
  float qinv;    // 1 / max(|q|, eps)
  float JW[9];
  float S[4];    // 2-D covariance, row-major, S01 / S10 separately
  float m2[2];   // mean2d
};

// ---- rounding-exact building blocks.  The reference computes the projection with ATen / cuBLAS kernels
// (gs/renderer.py:381-419).  tools/probe_torch_order.py measured, on a B200 with torch 2.11, where those
// kernels round (profiles/r2_probe_torch_order.json, 1 M Gaussians, identity and posed camera, 100 % bit match):
//   * every small matrix product (the einsum of project_pts and of JW, `rotmat @ rotmat^T`, both bmm of the
//     covariance) accumulates k = 0,1,2 as  fma(a2, b2, fma(a1, b1, a0 * b0));
//   * torch.norm(u, dim=-1) of the Jacobian is sqrt((u0^2 + u2^2) + u1^2), products rounded separately;
//   * F.normalize(q) (kornia's normalize_quaternion) is q / max(sqrt((q0^2 + q2^2) + (q1^2 + q3^2)), eps);
//   * everything else is one rounding per torch op.
// project_one reproduces exactly that with un-contracted intrinsics, so that the 2-D covariance -- and with
// it tile rects, duplicate counts and the 1/255 skip decisions -- is bit-identical to the reference flow.
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float dot3_gemm(float a0, float b0, float a1, float b1, float a2, float b2) {
  return __fmaf_rn(a2, b2, __fmaf_rn(a1, b1, __fmul_rn(a0, b0)));
}

// kornia 0.6.x quaternion_to_rotation_matrix(q, WXYZ) op for op (utils/transforms.py:31-45 -> kornia,
// un-vendored): normalize_quaternion, then the t* product form, one rounding per torch op.
__device__ __forceinline__ void quat_to_rotmat_exact(const float *q, float *R, float *qn, float *inv_norm) {
  const float n = __fsqrt_rn(xadd(xadd(xmul(q[0], q[0]), xmul(q[2], q[2])), xadd(xmul(q[1], q[1]), xmul(q[3], q[3]))));
  const float d = fmaxf(n, 1e-12f);
  const float w = xdiv(q[0], d), x = xdiv(q[1], d), y = xdiv(q[2], d), z = xdiv(q[3], d);
  const float tx = xmul(2.0f, x), ty = xmul(2.0f, y), tz = xmul(2.0f, z);
  const float twx = xmul(tx, w), twy = xmul(ty, w), twz = xmul(tz, w);
  const float txx = xmul(tx, x), txy = xmul(ty, x), txz = xmul(tz, x);
  const float tyy = xmul(ty, y), tyz = xmul(tz, y), tzz = xmul(tz, z);
  R[0] = xsub(1.0f, xadd(tyy, tzz));
  R[1] = xsub(txy, twz);
  R[2] = xadd(txz, twy);
  R[3] = xadd(txy, twz);
  R[4] = xsub(1.0f, xadd(txx, tzz));
  R[5] = xsub(tyz, twx);
  R[6] = xsub(txz, twy);
  R[7] = xadd(tyz, twx);
  R[8] = xsub(1.0f, xadd(txx, tyy));
  qn[0] = w; qn[1] = x; qn[2] = y; qn[3] = z;
  *inv_norm = 1.0f / d;  // (backward only; not on the bit-exact forward path)
}

// gs/renderer.py:381-419 restated per Gaussian, rounding where the reference's GPU kernels round (see above).
__device__ __forceinline__ void project_one(const float *p, const float *q, const float *s,
                                            const float *c2w, Projected &o) {
  // project_pts: W = c2w[:3,:3]^T, d = -t, u = einsum("ij,bj->bi", W, p + d)
  float pd[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) pd[j] = xadd(p[j], -c2w[4 * j + 3]);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    o.u[i] = dot3_gemm(c2w[4 * 0 + i], pd[0], c2w[4 * 1 + i], pd[1], c2w[4 * 2 + i], pd[2]);
  quat_to_rotmat_exact(q, o.R, o.qn, &o.qinv);
  // rotmat = svec.unsqueeze(-2) * R  ->  A[i][j] = s[j] * R[i][j];  sigma = A @ A^T
  float A[9], Sg[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A[3 * i + j] = xmul(s[j], o.R[3 * i + j]);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Sg[3 * i + j] = dot3_gemm(A[3 * i], A[3 * j], A[3 * i + 1], A[3 * j + 1], A[3 * i + 2], A[3 * j + 2]);
  // jacobian (renderer.py:366-377)
  const float ux = o.u[0], uy = o.u[1], uz = o.u[2];
  const float l = __fsqrt_rn(xadd(xadd(xmul(ux, ux), xmul(uz, uz)), xmul(uy, uy)));
  const float inv_z = xdiv(1.0f, uz);
  const float J[9] = {inv_z, 0.0f, xdiv(xdiv(-ux, uz), uz), 0.0f, inv_z, xdiv(xdiv(-uy, uz), uz),
                      xdiv(ux, l), xdiv(uy, l), xdiv(uz, l)};
  // JW = einsum("bij,jk->bik", J, W), W[j][k] = c2w[k][j]
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k)
      o.JW[3 * i + k] = dot3_gemm(J[3 * i], c2w[4 * k + 0], J[3 * i + 1], c2w[4 * k + 1], J[3 * i + 2], c2w[4 * k + 2]);
  // cov = bmm(bmm(JW, sigma), JW^T)[:2,:2]
  float X[6];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int k = 0; k < 3; ++k)
      X[3 * a + k] = dot3_gemm(o.JW[3 * a], Sg[k], o.JW[3 * a + 1], Sg[3 + k], o.JW[3 * a + 2], Sg[6 + k]);
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
      o.S[2 * a + b] = dot3_gemm(X[3 * a], o.JW[3 * b], X[3 * a + 1], o.JW[3 * b + 1], X[3 * a + 2], o.JW[3 * b + 2]);
  o.m2[0] = xdiv(ux, uz);
  o.m2[1] = xdiv(uy, uz);
}

__global__ void __launch_bounds__(256)
project_kernel(uint32_t N, const float *__restrict__ mean, const float *__restrict__ qvec,
               const float *__restrict__ svec, const float *__restrict__ c2w_g,
               float *__restrict__ mean2d, float *__restrict__ cov2d, float *__restrict__ JW,
               float *__restrict__ depth) {
  __shared__ float c2w[12];
  if (threadIdx.x < 12) c2w[threadIdx.x] = c2w_g[threadIdx.x];
  __syncthreads();
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float p[3] = {mean[3 * (size_t)i], mean[3 * (size_t)i + 1], mean[3 * (size_t)i + 2]};
  float4 q4 = reinterpret_cast<const float4 *>(qvec)[i];
  float q[4] = {q4.x, q4.y, q4.z, q4.w};
  float s[3] = {svec[3 * (size_t)i], svec[3 * (size_t)i + 1], svec[3 * (size_t)i + 2]};
  Projected o;
  project_one(p, q, s, c2w, o);
  reinterpret_cast<float2 *>(mean2d)[i] = make_float2(o.m2[0], o.m2[1]);
  reinterpret_cast<float4 *>(cov2d)[i] = make_float4(o.S[0], o.S[1], o.S[2], o.S[3]);
  depth[i] = o.u[2];
  if (JW) {
#pragma unroll
    for (int k = 0; k < 9; ++k) JW[9 * (size_t)i + k] = o.JW[k];
  }
}

// ---------------------------------------------------------------- tile rects (a5)

struct Rect {
  int tlx, tly, brx, bry;
};

// gs/culling.py:16-31 + utils/camera.py:290-303: every step is a separate FP32 torch op, so no
// contraction; `.to(int32)` truncates toward zero; clamp; floor-divide (operands >= 0).
__device__ __forceinline__ Rect tile_rect(float mx, float my, float s00, float s11, float D,
                                          const CamConst &cam, int tile) {
  float ax = __fsqrt_rn(__fmul_rn(D, s00));
  float ay = __fsqrt_rn(__fmul_rn(D, s11));
  float tlx = __fadd_rn(__fmul_rn(__fsub_rn(mx, ax), cam.fx), cam.cx);
  float tly = __fadd_rn(__fmul_rn(__fsub_rn(my, ay), cam.fy), cam.cy);
  float brx = __fadd_rn(__fmul_rn(__fadd_rn(mx, ax), cam.fx), cam.cx);
  float bry = __fadd_rn(__fmul_rn(__fadd_rn(my, ay), cam.fy), cam.cy);
  Rect r;
  int ix0 = __float2int_rz(tlx), iy0 = __float2int_rz(tly);
  int ix1 = __float2int_rz(brx), iy1 = __float2int_rz(bry);
  ix0 = min(max(ix0, 0), cam.w - 1);
  ix1 = min(max(ix1, 0), cam.w - 1);
  iy0 = min(max(iy0, 0), cam.h - 1);
  iy1 = min(max(iy1, 0), cam.h - 1);
  r.tlx = ix0 / tile; r.tly = iy0 / tile; r.brx = ix1 / tile; r.bry = iy1 / tile;
  return r;
}

__device__ __forceinline__ void block_count_add(unsigned long long local, unsigned long long *total) {
  // warp shuffle sum then one 64-bit atomic per warp (integer: order-independent, deterministic)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(total, local);
}

__global__ void __launch_bounds__(256)
rect_count_kernel(uint32_t N, const float *__restrict__ mean2d, const float *__restrict__ cov2d,
                  float D, CamConst cam, int tile, int32_t *__restrict__ tl,
                  int32_t *__restrict__ br, unsigned long long *__restrict__ total) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long cnt = 0;
  if (i < N) {
    float2 m = reinterpret_cast<const float2 *>(mean2d)[i];
    float4 c = reinterpret_cast<const float4 *>(cov2d)[i];
    Rect r = tile_rect(m.x, m.y, c.x, c.w, D, cam, tile);
    reinterpret_cast<int2 *>(tl)[i] = make_int2(r.tlx, r.tly);
    reinterpret_cast<int2 *>(br)[i] = make_int2(r.brx, r.bry);
    // torch.prod(br - tl + 1): may be <= 0 for NaN-degenerate inputs; keep the signed product
    long long w = (long long)(r.brx - r.tlx + 1), h = (long long)(r.bry - r.tly + 1);
    long long pr = w * h;
    cnt = pr > 0 ? (unsigned long long)pr : 0ull;
  }
  block_count_add(cnt, total);
}

// ---------------------------------------------------------------- staging records

__device__ __forceinline__ void make_record(float m2x, float m2y, float c0, float c1, float c2,
                                            float c3, float alpha, float depth, float4 *rec) {
  float a = fminf(alpha, 0.99f);  // vol_render_sh.h:124
  // skip test alpha_*G < 1/255  <=>  log2(G) < log2((1/255)/alpha_)
  float lthr = a > 0.0f ? log2f((1.0f / 255.0f) / a) : INFINITY;
  // fast conic: det in FP64 to keep the approximation error far inside the exact-path margin
  double det = (double)c0 * (double)c3 - (double)c1 * (double)c2;
  const double k = -0.5 * 1.4426950408889634;
  float qa = (float)(k * (double)c3 / det);
  float qb = (float)(-k * ((double)c1 + (double)c2) / det);
  float qc = (float)(k * (double)c0 / det);
  rec[0] = make_float4(m2x, m2y, a, lthr);
  rec[1] = make_float4(qa, qb, qc, depth);
  rec[2] = make_float4(c0, c1, c2, c3);
}

__global__ void __launch_bounds__(256)
pack_records_kernel(uint32_t N, const float *__restrict__ mean2d, const float *__restrict__ cov2d,
                    const float *__restrict__ alpha, const float *__restrict__ depth,
                    float *__restrict__ records) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float2 m = reinterpret_cast<const float2 *>(mean2d)[i];
  float4 c = reinterpret_cast<const float4 *>(cov2d)[i];
  make_record(m.x, m.y, c.x, c.y, c.z, c.w, alpha[i], depth ? depth[i] : 0.0f,
              reinterpret_cast<float4 *>(records) + 3 * (size_t)i);
}

// ---------------------------------------------------------------- fused K1

__device__ __forceinline__ float act_exp(float x, int on) { return on ? expf(x) : x; }
__device__ __forceinline__ float act_sigmoid(float x, int on) {
  return on ? 1.0f / (1.0f + expf(-x)) : x;
}

__global__ void __launch_bounds__(256)
project_cull_fused_kernel(uint32_t N, const float *__restrict__ mean,
                          const float *__restrict__ qvec, const float *__restrict__ svec_param,
                          const float *__restrict__ alpha_param, int svec_act, int alpha_act,
                          const float *__restrict__ c2w_g, const float *__restrict__ planes_g,
                          CamConst cam, float frustum_radius,
                          int skip_cull, float tile_D, int tile, uint8_t *__restrict__ mask,
                          float *__restrict__ mean2d, float *__restrict__ cov2d,
                          float *__restrict__ depth, int32_t *__restrict__ tl,
                          int32_t *__restrict__ br, float *__restrict__ records,
                          float *__restrict__ svec_out, float *__restrict__ alpha_out,
                          int32_t *__restrict__ cnt, unsigned long long *__restrict__ total) {
  __shared__ float c2w[12];
  __shared__ float planes[36];
  // frustum planes come from frustum_kernel (one thread, once per frame), not from every CTA
  if (threadIdx.x < 12) c2w[threadIdx.x] = c2w_g[threadIdx.x];
  else if (threadIdx.x >= 32 && threadIdx.x < 68) planes[threadIdx.x - 32] = planes_g[threadIdx.x - 32];
  __syncthreads();
  // staging records leave through shared memory: the 256 records of a block are 12 KB of CONTIGUOUS global
  // memory, written as full 16-byte-per-lane coalesced rows instead of three 48-byte-strided stores per thread
  __shared__ float4 s_rec[256 * 3];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long n_dup = 0;
  float4 rec[3] = {make_float4(0.f, 0.f, 0.f, INFINITY), make_float4(0.f, 0.f, 0.f, 0.f),
                   make_float4(1.f, 0.f, 0.f, 1.f)};  // culled: never contributes
  if (i < N) {
    // every load is issued before the first use (one memory round trip, the cull test included)
    const float p[3] = {mean[3 * (size_t)i], mean[3 * (size_t)i + 1], mean[3 * (size_t)i + 2]};
    const float sp[3] = {svec_param[3 * (size_t)i], svec_param[3 * (size_t)i + 1], svec_param[3 * (size_t)i + 2]};
    const float ap = alpha_param[i];
    const float4 q4 = reinterpret_cast<const float4 *>(qvec)[i];
    float s[3] = {act_exp(sp[0], svec_act), act_exp(sp[1], svec_act), act_exp(sp[2], svec_act)};
    float a = act_sigmoid(ap, alpha_act);
    if (svec_out) {
      svec_out[3 * (size_t)i] = s[0];
      svec_out[3 * (size_t)i + 1] = s[1];
      svec_out[3 * (size_t)i + 2] = s[2];
    }
    if (alpha_out) alpha_out[i] = a;
    bool keep = true;
    if (!skip_cull) {
      float r = fmaxf(fmaxf(s[0], s[1]), s[2]) * frustum_radius;
      keep = sphere_in_frustum(p[0], p[1], p[2], r, planes, planes + 18);
    }
    mask[i] = keep ? 1 : 0;
    Rect rc = {0, 0, -1, -1};
    float2 m2 = make_float2(0.f, 0.f);
    float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
    float dz = 0.f;
    if (keep) {
      float q[4] = {q4.x, q4.y, q4.z, q4.w};
      Projected o;
      project_one(p, q, s, c2w, o);
      m2 = make_float2(o.m2[0], o.m2[1]);
      cv = make_float4(o.S[0], o.S[1], o.S[2], o.S[3]);
      dz = o.u[2];
      rc = tile_rect(m2.x, m2.y, cv.x, cv.w, tile_D, cam, tile);
      long long w = (long long)(rc.brx - rc.tlx + 1), h = (long long)(rc.bry - rc.tly + 1);
      long long pr = w * h;
      n_dup = pr > 0 ? (unsigned long long)pr : 0ull;
      if (cnt) cnt[i] += 1;  // sh_renderer.py:215-216
      if (records) make_record(m2.x, m2.y, cv.x, cv.y, cv.z, cv.w, a, dz, rec);
    }
    // mean2d / cov2d are optional: the staging record already holds both (floats 0-1 and 8-11)
    if (mean2d) reinterpret_cast<float2 *>(mean2d)[i] = m2;
    if (cov2d) reinterpret_cast<float4 *>(cov2d)[i] = cv;
    depth[i] = dz;
    reinterpret_cast<int2 *>(tl)[i] = make_int2(rc.tlx, rc.tly);
    reinterpret_cast<int2 *>(br)[i] = make_int2(rc.brx, rc.bry);
  }
  if (records) {
    s_rec[3 * threadIdx.x + 0] = rec[0];
    s_rec[3 * threadIdx.x + 1] = rec[1];
    s_rec[3 * threadIdx.x + 2] = rec[2];
    __syncthreads();
    const size_t base4 = (size_t)blockIdx.x * (256 * 3);  // float4 index of the block's first record
    const size_t end4 = (size_t)N * 3;
    float4 *out4 = reinterpret_cast<float4 *>(records);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const size_t e = base4 + r * 256 + threadIdx.x;
      if (e < end4) out4[e] = s_rec[r * 256 + threadIdx.x];
    }
  }
  block_count_add(n_dup, total);
}

// ---------------------------------------------------------------- projection backward (a9)

struct LeafGrad {
  float gp[3], gq[4], gs[3];
};

// Chain rule through project_one. J (and JW) are constants (renderer.py:365 @torch.no_grad);
// depth is detached in mean2d = xy / depth unless detach_depth == 0 (renderer.py:408-417).
__device__ __forceinline__ void project_backward_one(const float *p, const float *q, const float *s,
                                                     const float *c2w, const float *gm2,
                                                     const float *gS, float gdepth, int detach,
                                                     LeafGrad &g) {
  Projected o;
  project_one(p, q, s, c2w, o);
  float uz = o.u[2];
  float gu[3] = {gm2[0] / uz, gm2[1] / uz, 0.0f};
  if (!detach) gu[2] = -(gm2[0] * o.u[0] + gm2[1] * o.u[1]) / (uz * uz) + gdepth;
  // u = W pd, W[i][j] = c2w[4j+i]  ->  gp_j = sum_i W[i][j] gu_i
#pragma unroll
  for (int j = 0; j < 3; ++j)
    g.gp[j] = c2w[4 * j + 0] * gu[0] + c2w[4 * j + 1] * gu[1] + c2w[4 * j + 2] * gu[2];
  // S[a][b] = sum_jk M[a][j] Sg[j][k] M[b][k], M = JW rows 0..1
  float gSg[9];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float acc = 0.0f;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) acc += gS[2 * a + b] * o.JW[3 * a + j] * o.JW[3 * b + k];
      gSg[3 * j + k] = acc;
    }
  // Sg = A A^T  ->  gA = (gSg + gSg^T) A ;  A[i][j] = R[i][j] s[j]
  float gR[9];
  g.gs[0] = g.gs[1] = g.gs[2] = 0.0f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float acc = 0.0f;
#pragma unroll
      for (int k = 0; k < 3; ++k) acc += (gSg[3 * i + k] + gSg[3 * k + i]) * (o.R[3 * k + j] * s[j]);
      g.gs[j] += acc * o.R[3 * i + j];
      gR[3 * i + j] = acc * s[j];
    }
  float w = o.qn[0], x = o.qn[1], y = o.qn[2], z = o.qn[3];
  float gqn[4];
  gqn[0] = 2.0f * (-z * gR[1] + y * gR[2] + z * gR[3] - x * gR[5] - y * gR[6] + x * gR[7]);
  gqn[1] = 2.0f * (y * gR[1] + z * gR[2] + y * gR[3] - 2.0f * x * gR[4] - w * gR[5] + z * gR[6] +
                   w * gR[7] - 2.0f * x * gR[8]);
  gqn[2] = 2.0f * (-2.0f * y * gR[0] + x * gR[1] + w * gR[2] + x * gR[3] + z * gR[5] - w * gR[6] +
                   z * gR[7] - 2.0f * y * gR[8]);
  gqn[3] = 2.0f * (-2.0f * z * gR[0] - w * gR[1] + x * gR[2] + w * gR[3] - 2.0f * z * gR[4] +
                   y * gR[5] + x * gR[6] + y * gR[7]);
  // q_hat = q * inv, inv = 1 / max(|q|, eps): for |q| > eps, gq = inv (g - q_hat (q_hat . g));
  // for |q| <= eps the denominator is the constant eps.
  float nrm2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  float dotv = (nrm2 > 1e-24f) ? (w * gqn[0] + x * gqn[1] + y * gqn[2] + z * gqn[3]) : 0.0f;
  g.gq[0] = o.qinv * (gqn[0] - w * dotv);
  g.gq[1] = o.qinv * (gqn[1] - x * dotv);
  g.gq[2] = o.qinv * (gqn[2] - y * dotv);
  g.gq[3] = o.qinv * (gqn[3] - z * dotv);
}

struct K4bArgs {
  uint32_t N;
  const uint8_t *mask;
  const float *mean, *qvec, *svec_param, *alpha_param;
  int svec_act, alpha_act;
  const float *c2w_g;
  int detach;
  const float *gm2d, *gcov, *gdepth, *galpha;
  float *gmean, *gqvec, *gsvec, *galpha_param, *adc_acc;
  int adc_mode, accumulate;
};

// One Gaussian of the projection backward.  keep == false: no gradient reaches this row (culled / unmarked).
__device__ __forceinline__ void k4b_row(const K4bArgs &a, const float *c2w, uint32_t i, bool keep) {
  LeafGrad g;
  float ga = 0.0f;
  float2 gm = make_float2(0.f, 0.f);
  float4 gc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (keep) {
    // A view touches a small fraction of the Gaussians (cfg 2: ~2 %): when every upstream gradient of
    // this Gaussian is zero the chain rule gives zeros -- skip the parameter loads and the arithmetic
    // (and, when accumulating, the read-modify-write).  NaN/Inf upstream values compare unequal to 0.
    gm = reinterpret_cast<const float2 *>(a.gm2d)[i];
    gc = reinterpret_cast<const float4 *>(a.gcov)[i];
    const float ga_up = a.galpha ? a.galpha[i] : 0.0f;
    const float gd_up = a.gdepth ? a.gdepth[i] : 0.0f;
    if (gm.x == 0.f && gm.y == 0.f && gc.x == 0.f && gc.y == 0.f && gc.z == 0.f && gc.w == 0.f &&
        ga_up == 0.f && gd_up == 0.f)
      keep = false;
  }
  if (keep) {
    float p[3] = {a.mean[3 * (size_t)i], a.mean[3 * (size_t)i + 1], a.mean[3 * (size_t)i + 2]};
    float4 q4 = reinterpret_cast<const float4 *>(a.qvec)[i];
    float q[4] = {q4.x, q4.y, q4.z, q4.w};
    float s[3] = {act_exp(a.svec_param[3 * (size_t)i], a.svec_act),
                  act_exp(a.svec_param[3 * (size_t)i + 1], a.svec_act),
                  act_exp(a.svec_param[3 * (size_t)i + 2], a.svec_act)};
    float gm2[2] = {gm.x, gm.y};
    float gS[4] = {gc.x, gc.y, gc.z, gc.w};
    project_backward_one(p, q, s, c2w, gm2, gS, a.gdepth ? a.gdepth[i] : 0.0f, a.detach, g);
    if (a.svec_act) {
      g.gs[0] *= s[0]; g.gs[1] *= s[1]; g.gs[2] *= s[2];
    }
    if (a.galpha) {
      ga = a.galpha[i];
      if (a.alpha_act) {
        float al = act_sigmoid(a.alpha_param[i], 1);
        ga *= al * (1.0f - al);
      }
    }
    if (a.adc_mode && a.adc_acc) {  // sh_renderer.py:612-623, split_type "2d_mean_grad"
      float nrm = sqrtf(gm.x * gm.x + gm.y * gm.y);
      if (a.adc_mode == 1) a.adc_acc[i] = fmaxf(a.adc_acc[i], nrm);
      else a.adc_acc[i] += nrm;
    }
  } else {
    g.gp[0] = g.gp[1] = g.gp[2] = 0.f;
    g.gq[0] = g.gq[1] = g.gq[2] = g.gq[3] = 0.f;
    g.gs[0] = g.gs[1] = g.gs[2] = 0.f;
  }
  float *gmean = a.gmean, *gqvec = a.gqvec, *gsvec = a.gsvec, *galpha_param = a.galpha_param;
  if (a.accumulate) {  // several views per step: sum into the caller's gradient buffers
    if (!keep) return;
    gmean[3 * (size_t)i] += g.gp[0]; gmean[3 * (size_t)i + 1] += g.gp[1]; gmean[3 * (size_t)i + 2] += g.gp[2];
    float4 q0 = reinterpret_cast<float4 *>(gqvec)[i];
    reinterpret_cast<float4 *>(gqvec)[i] = make_float4(q0.x + g.gq[0], q0.y + g.gq[1], q0.z + g.gq[2], q0.w + g.gq[3]);
    gsvec[3 * (size_t)i] += g.gs[0]; gsvec[3 * (size_t)i + 1] += g.gs[1]; gsvec[3 * (size_t)i + 2] += g.gs[2];
    if (galpha_param) galpha_param[i] += ga;
    return;
  }
  gmean[3 * (size_t)i] = g.gp[0]; gmean[3 * (size_t)i + 1] = g.gp[1]; gmean[3 * (size_t)i + 2] = g.gp[2];
  reinterpret_cast<float4 *>(gqvec)[i] = make_float4(g.gq[0], g.gq[1], g.gq[2], g.gq[3]);
  gsvec[3 * (size_t)i] = g.gs[0]; gsvec[3 * (size_t)i + 1] = g.gs[1]; gsvec[3 * (size_t)i + 2] = g.gs[2];
  if (galpha_param) galpha_param[i] = ga;
}

__global__ void __launch_bounds__(256)
project_backward_kernel(const K4bArgs a) {
  __shared__ float c2w[12];
  if (threadIdx.x < 12) c2w[threadIdx.x] = a.c2w_g[threadIdx.x];
  __syncthreads();
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  k4b_row(a, c2w, i, a.mask ? a.mask[i] != 0 : true);
}

// Sparse row filter (accumulate == 2: the mask is the compositing backward's `touched` marks, ~2 % of the rows):
// with one thread per Gaussian almost every warp would run the whole chain rule for a single live lane -- 67 k
// rows cost 57 us at cfg 2, latency-bound.  Here a warp reads 512 marks (16 bytes per lane), compacts the marked
// indices into shared memory and processes them 32 at a time with all lanes busy.  Accumulating only.
__global__ void __launch_bounds__(256)
project_backward_marked_kernel(const K4bArgs a) {
  __shared__ float c2w[12];
  __shared__ uint32_t s_idx[8][512];
  if (threadIdx.x < 12) c2w[threadIdx.x] = a.c2w_g[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t w0 = ((size_t)blockIdx.x * 8 + warp) * 512;  // first Gaussian of this warp
  if (w0 >= a.N) return;
  const size_t c0 = w0 + (size_t)lane * 16;
  uint32_t bits = 0;  // one bit per marked Gaussian of this lane's 16
  if (c0 + 16 <= a.N && (reinterpret_cast<uintptr_t>(a.mask) & 15) == 0) {
    const uint4 v = *reinterpret_cast<const uint4 *>(a.mask + c0);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if ((w[q] >> (8 * b)) & 0xffu) bits |= 1u << (4 * q + b);
  } else {
    for (int b = 0; b < 16; ++b)
      if (c0 + b < a.N && a.mask[c0 + b]) bits |= 1u << b;
  }
  const uint32_t cnt = __popc(bits);
  uint32_t incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  uint32_t pos = incl - cnt;
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    s_idx[warp][pos++] = (uint32_t)(c0 + b);
  }
  __syncwarp();
  for (uint32_t j = lane; j < total; j += 32) k4b_row(a, c2w, s_idx[warp][j], true);
}

static int read_count(unsigned long long *dev_total, int64_t *n_dub_host, cudaStream_t st) {
  int64_t *box = pinned_mailbox();
  GS3D_REQUIRE(box != nullptr, GS3D_ECUDA, "cudaMallocHost for the count mailbox failed");
  GS3D_CUDA(cudaMemcpyAsync(box, dev_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  GS3D_CUDA(cudaStreamSynchronize(st));
  *n_dub_host = *box;
  return GS3D_OK;
}

}  // namespace gs3d

using namespace gs3d;

extern "C" {

int gs3d_version(void) { return 100; }
const char *gs3d_last_error(void) { return g_err; }
uint64_t gs3d_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int gs3d_get_frustum(const float *c2w, const gs3d_camera *cam_host, float *normals, float *pts,
                     void *stream) {
  GS3D_REQUIRE(c2w && cam_host && normals && pts, GS3D_EINVAL, "gs3d_get_frustum: null argument");
  frustum_kernel<<<1, 32, 0, as_stream(stream)>>>(c2w, make_cam(cam_host), normals, pts);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

int gs3d_culling_gaussian_bsphere(uint32_t N, const float *mean, const float *qvec,
                                  const float *svec, const float *normal, const float *pts,
                                  uint8_t *mask, float thresh, void *stream) {
  (void)qvec;  // unused by the reference as well (culling.h:18-19)
  if (N == 0) return GS3D_OK;
  GS3D_REQUIRE(mean && svec && normal && pts && mask, GS3D_EINVAL,
               "gs3d_culling_gaussian_bsphere: null argument");
  cull_bsphere_kernel<<<div_up(N, 256u), 256, 0, as_stream(stream)>>>(N, mean, svec, normal, pts,
                                                                      mask, thresh);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

int gs3d_project_gaussians(uint32_t N, const float *mean, const float *qvec, const float *svec,
                           const float *c2w, float *mean2d, float *cov2d, float *JW, float *depth,
                           void *stream) {
  if (N == 0) return GS3D_OK;
  GS3D_REQUIRE(mean && qvec && svec && c2w && mean2d && cov2d && depth, GS3D_EINVAL,
               "gs3d_project_gaussians: null argument");
  project_kernel<<<div_up(N, 256u), 256, 0, as_stream(stream)>>>(N, mean, qvec, svec, c2w, mean2d,
                                                                 cov2d, JW, depth);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

int gs3d_project_gaussians_backward(uint32_t N, const float *mean, const float *qvec,
                                    const float *svec, const float *c2w, const float *grad_mean2d,
                                    const float *grad_cov2d, const float *grad_depth,
                                    int detach_depth, float *grad_mean, float *grad_qvec,
                                    float *grad_svec, void *stream) {
  if (N == 0) return GS3D_OK;
  GS3D_REQUIRE(mean && qvec && svec && c2w && grad_mean2d && grad_cov2d && grad_mean && grad_qvec &&
                   grad_svec,
               GS3D_EINVAL, "gs3d_project_gaussians_backward: null argument");
  K4bArgs a = {};
  a.N = N; a.mean = mean; a.qvec = qvec; a.svec_param = svec; a.c2w_g = c2w; a.detach = detach_depth;
  a.gm2d = grad_mean2d; a.gcov = grad_cov2d; a.gdepth = grad_depth;
  a.gmean = grad_mean; a.gqvec = grad_qvec; a.gsvec = grad_svec;
  project_backward_kernel<<<div_up(N, 256u), 256, 0, as_stream(stream)>>>(a);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

size_t gs3d_count_scratch_bytes(uint32_t N) {
  (void)N;
  return 256;
}

int gs3d_tile_culling_aabb_count(uint32_t N, const float *mean2d, const float *cov2d,
                                 uint32_t tile_size, const gs3d_camera *cam_host, float D,
                                 int32_t *aabb_topleft, int32_t *aabb_bottomright,
                                 int64_t *n_dub_host, void *scratch, size_t scratch_bytes,
                                 void *stream) {
  GS3D_REQUIRE(cam_host && n_dub_host && scratch && scratch_bytes >= 8 && tile_size > 0, GS3D_EINVAL,
               "gs3d_tile_culling_aabb_count: bad argument");
  cudaStream_t st = as_stream(stream);
  unsigned long long *total = static_cast<unsigned long long *>(scratch);
  GS3D_CUDA(cudaMemsetAsync(total, 0, 8, st));
  if (N) {
    GS3D_REQUIRE(mean2d && cov2d && aabb_topleft && aabb_bottomright, GS3D_EINVAL,
                 "gs3d_tile_culling_aabb_count: null argument");
    rect_count_kernel<<<div_up(N, 256u), 256, 0, st>>>(N, mean2d, cov2d, D, make_cam(cam_host),
                                                       (int)tile_size, aabb_topleft,
                                                       aabb_bottomright, total);
    GS3D_LAUNCH_CHECK();
  }
  return read_count(total, n_dub_host, st);
}

int gs3d_pack_records(uint32_t N, const float *mean2d, const float *cov2d, const float *alpha,
                      const float *depth, float *records, void *stream) {
  if (N == 0) return GS3D_OK;
  GS3D_REQUIRE(mean2d && cov2d && alpha && records, GS3D_EINVAL, "gs3d_pack_records: null argument");
  pack_records_kernel<<<div_up(N, 256u), 256, 0, as_stream(stream)>>>(N, mean2d, cov2d, alpha,
                                                                      depth, records);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

int gs3d_project_cull_fused(uint32_t N, const float *mean, const float *qvec,
                            const float *svec_param, const float *alpha_param, int svec_act,
                            int alpha_act, const float *c2w, const gs3d_camera *cam_host,
                            float frustum_radius, int skip_frustum_culling, float tile_D,
                            uint32_t tile_size, uint8_t *mask, float *mean2d, float *cov2d,
                            float *depth, int32_t *aabb_topleft, int32_t *aabb_bottomright,
                            float *records, float *svec_out, float *alpha_out, int32_t *cnt,
                            int64_t *n_dub_host, void *scratch, size_t scratch_bytes, void *stream) {
  GS3D_REQUIRE(cam_host && scratch && scratch_bytes >= 256 && tile_size > 0 && c2w,
               GS3D_EINVAL, "gs3d_project_cull_fused: bad argument (scratch must be >= 256 bytes)");
  cudaStream_t st = as_stream(stream);
  unsigned long long *total = static_cast<unsigned long long *>(scratch);
  GS3D_CUDA(cudaMemsetAsync(total, 0, 8, st));
  if (N) {
    GS3D_REQUIRE(mean && qvec && svec_param && alpha_param && mask && depth && aabb_topleft && aabb_bottomright,
                 GS3D_EINVAL, "gs3d_project_cull_fused: null argument");
    GS3D_REQUIRE((mean2d && cov2d) || records, GS3D_EINVAL,
                 "gs3d_project_cull_fused: mean2d / cov2d may only be NULL when records are written");
    float *planes = reinterpret_cast<float *>(static_cast<char *>(scratch) + 64);  // 36 floats
    frustum_kernel<<<1, 32, 0, st>>>(c2w, make_cam(cam_host), planes, planes + 18);
    GS3D_LAUNCH_CHECK();
    project_cull_fused_kernel<<<div_up(N, 256u), 256, 0, st>>>(
        N, mean, qvec, svec_param, alpha_param, svec_act, alpha_act, c2w, planes, make_cam(cam_host),
        frustum_radius, skip_frustum_culling, tile_D, (int)tile_size, mask, mean2d, cov2d, depth,
        aabb_topleft, aabb_bottomright, records, svec_out, alpha_out, cnt, total);
    GS3D_LAUNCH_CHECK();
  }
  if (!n_dub_host) return GS3D_OK;  // no read-back: the count stays in the first 8 bytes of `scratch`
  return read_count(total, n_dub_host, st);
}

int gs3d_project_backward_fused(uint32_t N, const uint8_t *mask, const float *mean,
                                const float *qvec, const float *svec_param,
                                const float *alpha_param, int svec_act, int alpha_act,
                                const float *c2w, int detach_depth, const float *grad_mean2d,
                                const float *grad_cov2d, const float *grad_alpha,
                                float *grad_mean, float *grad_qvec, float *grad_svec_param,
                                float *grad_alpha_param, float *grad_mean_acc, int adc_mode,
                                int accumulate, void *stream) {
  if (N == 0) return GS3D_OK;
  GS3D_REQUIRE(mean && qvec && svec_param && alpha_param && c2w && grad_mean2d && grad_cov2d &&
                   grad_alpha && grad_mean && grad_qvec && grad_svec_param && grad_alpha_param,
               GS3D_EINVAL, "gs3d_project_backward_fused: null argument");
  K4bArgs a = {};
  a.N = N; a.mask = mask; a.mean = mean; a.qvec = qvec; a.svec_param = svec_param; a.alpha_param = alpha_param;
  a.svec_act = svec_act; a.alpha_act = alpha_act; a.c2w_g = c2w; a.detach = detach_depth;
  a.gm2d = grad_mean2d; a.gcov = grad_cov2d; a.gdepth = nullptr; a.galpha = grad_alpha;
  a.gmean = grad_mean; a.gqvec = grad_qvec; a.gsvec = grad_svec_param; a.galpha_param = grad_alpha_param;
  a.adc_acc = grad_mean_acc; a.adc_mode = adc_mode; a.accumulate = accumulate ? 1 : 0;
  if (accumulate == 2 && mask)  // sparse row filter: compact the marked rows first
    project_backward_marked_kernel<<<div_up(N, 4096u), 256, 0, as_stream(stream)>>>(a);
  else
    project_backward_kernel<<<div_up(N, 256u), 256, 0, as_stream(stream)>>>(a);
  GS3D_LAUNCH_CHECK();
  return GS3D_OK;
}

}  // extern "C"
