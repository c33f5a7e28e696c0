// Issue-rate microbenchmark, second set (round 2): does the packed FP32 pipe (FFMA2 / FADD2, sm_100a) relieve the
// issue slots of the FFT butterflies, and what is the FP32 FMA-burn peak to divide the acquisition roofline by?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates2 pipe_rates2.cu && ./pipe_rates2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { return ((u64)__float_as_uint(b) << 32) | __float_as_uint(a); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ int iadd3(int a, int b, int c) { int d; asm volatile("{.reg .s32 t; add.s32 t, %1, %2; add.s32 %0, t, %3;}" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ float2 lds64(const float2* p) { float2 v; unsigned a = (unsigned)__cvta_generic_to_shared(p); asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ int lop(int a, int b) { int d; asm volatile("xor.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }

// OP: 0 FFMA reg,reg,reg   1 FFMA2 reg   2 FFMA2 with a scalar (broadcast) multiplier   3 FADD2
//     4 FFMA + LOP (1:1)   5 FFMA2 + LOP (1:1)   6 FFMA + LDS.64 (4:1)   7 FFMA2 + LDS.64 (4:1)   8 FFMA2 + LDS.64 (2:1)
//     9 FFMA2 + FFMA (1:1)  10 FFMA + LDS.64 (2:1)
template <int OP> __global__ void __launch_bounds__(512) k(int iters, float* io, float g) {
  __shared__ float2 sm[1024];
  sm[threadIdx.x] = make_float2(threadIdx.x, 1.f); sm[threadIdx.x + 512] = make_float2(2.f, threadIdx.x);
  __syncthreads();
  float x = io[threadIdx.x & 31], y = io[32 + (threadIdx.x & 31)];
  float f[16]; u64 p[8]; int a[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = x * (i + 1);
#pragma unroll
  for (int i = 0; i < 8; ++i) { p[i] = pk(f[i], f[i + 8]); a[i] = threadIdx.x * (i + 3); }
  const u64 xx = pk(x, y), yy = pk(y, x), gg = pk(g, g);
  int addr = threadIdx.x;
  float2 l0 = make_float2(0.f, 0.f), l1 = l0;
  for (int it = 0; it < iters; ++it) {
    if (OP == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = ffma(f[i], x, y);
    } else if (OP == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], xx, yy);
    } else if (OP == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], gg, yy);
    } else if (OP == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fadd2(p[i], xx);
    } else if (OP == 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { f[i] = ffma(f[i], x, y); a[i] = lop(a[i], addr); }
    } else if (OP == 5) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { p[i] = ffma2(p[i], xx, yy); a[i] = lop(a[i], addr); }
    } else if (OP == 6 || OP == 10) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        f[i] = ffma(f[i], x, y);
        if (i % (OP == 6 ? 4 : 2) == 0) { float2 v = lds64(&sm[(addr + i * 32) & 1023]); l0.x += v.x; }
      }
    } else if (OP == 7 || OP == 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        p[i] = ffma2(p[i], xx, yy);
        if (i % (OP == 7 ? 4 : 2) == 0) { float2 v = lds64(&sm[(addr + i * 32) & 1023]); l1.x += v.x; }
      }
    } else if (OP == 9) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { p[i] = ffma2(p[i], xx, yy); f[i] = ffma(f[i], x, y); }
    }
  }
  float r = l0.x + l1.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += f[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + a[i];
  if (r == 0.12345f) io[0] = r;
}
template <int OP> void run(const char* name, double inst_per_iter, double fma_per_iter, int threads = 512, int bps = 4) {
  float* io; cudaMalloc(&io, 4096); cudaMemset(io, 0, 4096);
  const int iters = 20000, blocks = 148 * bps;
  k<OP><<<blocks, threads>>>(100, io, 1.0001f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); k<OP><<<blocks, threads>>>(iters, io, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double thr = (double)blocks * threads * iters;
  const double wi = thr * inst_per_iter / 32.0 / (best * 1e-3) / 148.0 / (clk * 1e3);   // warp-instructions / clk / SM
  const double tf = thr * fma_per_iter * 2.0 / (best * 1e-3) / 1e12;
  printf("{\"op\": \"%s\", \"threads\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"warp_inst_per_clk_per_sm\": %.3f, \"fp32_tflops\": %.2f, \"clock_mhz\": %d}\n",
         name, threads, bps, best, wi, tf, clk / 1000);
  cudaFree(io);
}
int main() {
  run<0>("FFMA r,r,r", 16, 16);
  run<1>("FFMA2 r,r,r", 8, 16);
  run<2>("FFMA2 r,scalar,r", 8, 16);
  run<3>("FADD2", 8, 8);
  run<4>("FFMA + LOP 1:1", 16, 8);
  run<5>("FFMA2 + LOP 1:1", 16, 16);
  run<6>("FFMA + LDS.64 4:1", 20 + 4, 16);
  run<10>("FFMA + LDS.64 2:1", 24 + 8, 16);
  run<7>("FFMA2 + LDS.64 4:1", 10 + 2, 16);
  run<8>("FFMA2 + LDS.64 2:1", 12 + 4, 16);
  run<9>("FFMA2 + FFMA 1:1", 16, 24);
  run<0>("FFMA r,r,r 4 warps/SM", 16, 16, 128, 1);
  run<1>("FFMA2 r,r,r 4 warps/SM", 8, 16, 128, 1);
  run<0>("FFMA r,r,r 8 warps/SM", 16, 16, 256, 1);
  run<1>("FFMA2 r,r,r 8 warps/SM", 8, 16, 256, 1);
  return 0;
}
