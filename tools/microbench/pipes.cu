// Pipe-throughput microbenchmarks for B200 (sm_100a): FFMA vs FFMA2 vs mixed, MUFU, LDS.128 broadcast.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ float ffma(float a, float b, float c) {
  float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float4 lds128(unsigned a) {
  float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
constexpr int ITERS = 4096;
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float a, float b) {
  __shared__ float4 sm[256];
  sm[threadIdx.x] = make_float4(a, b, a, b);
  __syncthreads();
  unsigned sa = (unsigned)__cvta_generic_to_shared(sm);
  float x[8]; f32x2 y[8];
  for (int i = 0; i < 8; ++i) { x[i] = a * i + threadIdx.x; y[i] = (f32x2)(threadIdx.x + i) * 0x0000000100000001ull; }
  f32x2 a2 = ((f32x2)__float_as_uint(a) << 32) | __float_as_uint(b);
  float acc = 0;
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {            // 8 independent FFMA chains
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = ffma(x[i], a, b);
    } else if (MODE == 1) {     // 8 independent FFMA2 chains
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = fma2(y[i], a2, a2);
    } else if (MODE == 2) {     // 4 FFMA2 + 4 FFMA
#pragma unroll
      for (int i = 0; i < 4; ++i) { y[i] = fma2(y[i], a2, a2); x[i] = ffma(x[i], a, b); }
    } else if (MODE == 3) {     // 8 MUFU.EX2
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = ex2(x[i]);
    } else if (MODE == 4) {     // 8 LDS.128 broadcast (all lanes same address)
#pragma unroll
      for (int i = 0; i < 8; ++i) { float4 v = lds128(sa + 16 * ((it + i) & 255)); x[i] += v.x; }
    } else if (MODE == 5) {     // 8 LDS.128 per-lane distinct (conflict-free 128-bit)
#pragma unroll
      for (int i = 0; i < 8; ++i) { float4 v = lds128(sa + 16 * ((threadIdx.x + it + i) & 255)); x[i] += v.x; }
    } else if (MODE == 6) {     // 8 LDS.128 broadcast, no dependent add (pure LSU)
#pragma unroll
      for (int i = 0; i < 8; ++i) { float4 v = lds128(sa + 16 * ((it + i) & 255)); acc = v.y; }
    } else if (MODE == 7) {     // 8 LDS.32 broadcast
#pragma unroll
      for (int i = 0; i < 8; ++i) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sa + 4 * ((it + i) & 255))); x[i] += v; }
    }
  }
  float s = acc;
  for (int i = 0; i < 8; ++i) s += x[i] + (float)(y[i] & 0xffff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name, int ctas_per_sm, double ops_per_inst) {
  float *out; cudaMalloc(&out, 148 * 8 * 256 * 4 * 4);
  int grid = 148 * ctas_per_sm;
  k<MODE><<<grid, 256>>>(out, 1.0001f, 0.5f); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<MODE><<<grid, 256>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  double winst = (double)grid * 8 * ITERS * 8;  // warp-instructions of the tested kind
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cyc = ms * 1e-3 * clk * 1e3;
  printf("%-44s ctas/SM %d: %.3f ms  %.3f warp-inst/clk/SM (at %d MHz nominal) -> %.2f inst/clk/SMSP, %.1f Gop/s\n", name,
         ctas_per_sm, ms, winst / cyc / 148, clk / 1000, winst / cyc / 148 / 4, winst * 32 * ops_per_inst / ms / 1e6);
  cudaFree(out);
}
int main() {
  for (int c : {2, 4, 8}) {
    if (c == 2) { run<0>("FFMA x8 indep", 2, 2); run<1>("FFMA2 x8 indep", 2, 4); run<2>("FFMA2 x4 + FFMA x4", 2, 3); run<3>("MUFU.EX2 x8", 2, 1);
                  run<4>("LDS.128 broadcast + FADD", 2, 16); run<5>("LDS.128 distinct + FADD", 2, 16); run<6>("LDS.128 broadcast only", 2, 16); run<7>("LDS.32 broadcast + FADD", 2, 4); }
    if (c == 4) { run<0>("FFMA x8 indep", 4, 2); run<1>("FFMA2 x8 indep", 4, 4); run<2>("FFMA2 x4 + FFMA x4", 4, 3); run<3>("MUFU.EX2 x8", 4, 1);
                  run<4>("LDS.128 broadcast + FADD", 4, 16); run<5>("LDS.128 distinct + FADD", 4, 16); run<6>("LDS.128 broadcast only", 4, 16); run<7>("LDS.32 broadcast + FADD", 4, 4); }
    if (c == 8) { run<0>("FFMA x8 indep", 8, 2); run<1>("FFMA2 x8 indep", 8, 4); run<2>("FFMA2 x4 + FFMA x4", 8, 3); run<3>("MUFU.EX2 x8", 8, 1);
                  run<4>("LDS.128 broadcast + FADD", 8, 16); run<5>("LDS.128 distinct + FADD", 8, 16); run<6>("LDS.128 broadcast only", 8, 16); run<7>("LDS.32 broadcast + FADD", 8, 4); }
  }
  return 0;
}
