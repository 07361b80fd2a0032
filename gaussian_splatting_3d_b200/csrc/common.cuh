// Shared helpers for libgs3d_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gs3d_b200.h"

#ifndef __CUDA_ARCH__
#define GS3D_HOST_ONLY
#endif

namespace gs3d {

// thread-local error slot (gs3d_last_error)
void set_error(const char *fmt, ...);
void count_launch();

#define GS3D_REQUIRE(cond, code, ...)      \
  do {                                     \
    if (!(cond)) {                         \
      gs3d::set_error(__VA_ARGS__);        \
      return (code);                       \
    }                                      \
  } while (0)

#define GS3D_CUDA(call)                                                                       \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      gs3d::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return GS3D_ECUDA;                                                                      \
    }                                                                                         \
  } while (0)

// every kernel launch goes through this macro: counts launches (gs3d_launch_count) and checks
#define GS3D_LAUNCH_CHECK()          \
  do {                               \
    gs3d::count_launch();            \
    GS3D_CUDA(cudaGetLastError());   \
  } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
static inline T div_up(T a, T b) {
  return (a + b - 1) / b;
}

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over caller-provided scratch.
struct Scratch {
  char *base;
  size_t cap, off;
  Scratch(void *p, size_t bytes) : base(static_cast<char *>(p)), cap(bytes), off(0) {}
  template <typename T>
  T *take(size_t n) {
    size_t bytes = align_up(n * sizeof(T));
    if (off + bytes > cap) return nullptr;
    T *r = reinterpret_cast<T *>(base + off);
    off += bytes;
    return r;
  }
};

// pinned 8-byte mailbox for the duplicate count (one per process, created lazily)
int64_t *pinned_mailbox();

}  // namespace gs3d

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
namespace gs3d {

struct Mat3 {
  float m[9];
};

// kornia 0.6.x quaternion_to_rotation_matrix(q, WXYZ): normalise (eps 1e-12), then the standard
// matrix built from t* = 2*q* products (utils/transforms.py:31-45 -> kornia, un-vendored).
__device__ __forceinline__ void quat_to_rotmat(float qw, float qx, float qy, float qz, float *R,
                                               float *qn /*[4] normalised*/, float *inv_norm) {
  float n = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
  float inv = 1.0f / fmaxf(n, 1e-12f);
  float w = qw * inv, x = qx * inv, y = qy * inv, z = qz * inv;
  float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
  float twx = tx * w, twy = ty * w, twz = tz * w;
  float txx = tx * x, txy = ty * x, txz = tz * x;
  float tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0f - (tyy + tzz);
  R[1] = txy - twz;
  R[2] = txz + twy;
  R[3] = txy + twz;
  R[4] = 1.0f - (txx + tzz);
  R[5] = tyz - twx;
  R[6] = txz - twy;
  R[7] = tyz + twx;
  R[8] = 1.0f - (txx + tyy);
  if (qn) {
    qn[0] = w; qn[1] = x; qn[2] = y; qn[3] = z;
  }
  if (inv_norm) *inv_norm = inv;
}

}  // namespace gs3d
#endif
