// Issue-rate microbenchmark: how many FFMA / DP4A / IMAD.WIDE / PRMT / I2F per clock per SM?
// (decides which correlator formulation the tracking kernel should use)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(int iters, int* out, int seed) {
  int a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
  float f0 = a0, f1 = a1, f2 = a2, f3 = a3, f4 = a4, f5 = a5, f6 = a6, f7 = a7;
  long long w0 = a0, w1 = a1, w2 = a2, w3 = a3, w4 = a4, w5 = a5, w6 = a6, w7 = a7;
  double d0 = a0, d1 = a1, d2 = a2, d3 = a3, d4 = a4, d5 = a5, d6 = a6, d7 = a7;
  const double dg = 1.0000001;
  const int b = seed * 77 + 1;
  const float g = 1.0001f;
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) { f0 = fmaf(f0, g, 1.f); f1 = fmaf(f1, g, 1.f); f2 = fmaf(f2, g, 1.f); f3 = fmaf(f3, g, 1.f); f4 = fmaf(f4, g, 1.f); f5 = fmaf(f5, g, 1.f); f6 = fmaf(f6, g, 1.f); f7 = fmaf(f7, g, 1.f); }
    if (OP == 1) { a0 = __dp4a(a0, b, a0); a1 = __dp4a(a1, b, a1); a2 = __dp4a(a2, b, a2); a3 = __dp4a(a3, b, a3); a4 = __dp4a(a4, b, a4); a5 = __dp4a(a5, b, a5); a6 = __dp4a(a6, b, a6); a7 = __dp4a(a7, b, a7); }
    if (OP == 2) { w0 += (long long)a0 * b; w1 += (long long)a1 * b; w2 += (long long)a2 * b; w3 += (long long)a3 * b; w4 += (long long)a4 * b; w5 += (long long)a5 * b; w6 += (long long)a6 * b; w7 += (long long)a7 * b; }
    if (OP == 3) { a0 = __byte_perm(a0, b, 0x7540); a1 = __byte_perm(a1, b, 0x7541); a2 = __byte_perm(a2, b, 0x7542); a3 = __byte_perm(a3, b, 0x7543); a4 = __byte_perm(a4, b, 0x3210 ^ 0x1111); a5 = __byte_perm(a5, b, 0x6420); a6 = __byte_perm(a6, b, 0x7531); a7 = __byte_perm(a7, b, 0x0123); }
    if (OP == 4) { f0 += (float)a0; f1 += (float)a1; f2 += (float)a2; f3 += (float)a3; f4 += (float)a4; f5 += (float)a5; f6 += (float)a6; f7 += (float)a7; a0 += i; a1 += i; a2 += i; a3 += i; a4 += i; a5 += i; a6 += i; a7 += i; }
    if (OP == 5) { a0 = a0 * b + a0; a1 = a1 * b + a1; a2 = a2 * b + a2; a3 = a3 * b + a3; a4 = a4 * b + a4; a5 = a5 * b + a5; a6 = a6 * b + a6; a7 = a7 * b + a7; }
    if (OP == 6) { d0 = fma(d0, dg, 1.0); d1 = fma(d1, dg, 1.0); d2 = fma(d2, dg, 1.0); d3 = fma(d3, dg, 1.0); d4 = fma(d4, dg, 1.0); d5 = fma(d5, dg, 1.0); d6 = fma(d6, dg, 1.0); d7 = fma(d7, dg, 1.0); }
    if (OP == 7) { d0 += (double)a0; d1 += (double)a1; d2 += (double)a2; d3 += (double)a3; d4 += (double)a4; d5 += (double)a5; d6 += (double)a6; d7 += (double)a7; a0 += i; a1 += i; a2 += i; a3 += i; a4 += i; a5 += i; a6 += i; a7 += i; }
    if (OP == 8) { d0 = fma(d0, dg, 1.0); d0 = fma(d0, dg, 1.0); d0 = fma(d0, dg, 1.0); d0 = fma(d0, dg, 1.0); d0 = fma(d0, dg, 1.0); d0 = fma(d0, dg, 1.0); d0 = fma(d0, dg, 1.0); d0 = fma(d0, dg, 1.0); }
    if (OP == 9) { a0 = __dp4a(a0, b, a0); a0 = __dp4a(a0, b, a0); a0 = __dp4a(a0, b, a0); a0 = __dp4a(a0, b, a0); a0 = __dp4a(a0, b, a0); a0 = __dp4a(a0, b, a0); a0 = __dp4a(a0, b, a0); a0 = __dp4a(a0, b, a0); }
  }
  int r = (int)(d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7) ^ a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ (int)(f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7) ^ (int)(w0 + w1 + w2 + w3 + w4 + w5 + w6 + w7);
  if (r == 0x12345678) out[0] = r;
}
template <int OP> void run(const char* name, int ops_per_iter, int blocks = 148 * 4, int threads = 512) {
  int* out; cudaMalloc(&out, 4);
  const int iters = 20000;
  k<OP><<<blocks, threads>>>(100, out, 3);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<OP><<<blocks, threads>>>(iters, out, 3); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * threads * iters * ops_per_iter;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-10s %.1f Gop/s  = %.1f thread-ops/clk/SM at %d MHz (nominal)\n", name, ops / ms / 1e6, ops / (ms * 1e-3) / 148.0 / (clk * 1e3), clk / 1000);
}
int main() { run<0>("FFMA", 8); run<1>("DP4A", 8); run<2>("IMAD.WIDE", 8); run<3>("PRMT", 8); run<4>("I2F(+IADD)", 8); run<5>("IMAD", 8);
  run<6>("DFMA", 8); run<7>("I2F.F64(+DADD,IADD)", 8);
  // dependent chains, one warp per SM: ops/clk/SM * 32 lanes -> latency = 32 / value
  run<8>("DFMA chain (1 warp/SM: latency = 32/value clk)", 8, 148, 32); run<9>("DP4A chain (1 warp/SM)", 8, 148, 32);
  run<6>("DFMA 8 warps/SM", 8, 148, 256); run<6>("DFMA 16 warps/SM", 8, 148, 512); run<1>("DP4A 8 warps/SM", 8, 148, 256);
  return 0; }
