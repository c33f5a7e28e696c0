// Fine-frequency search (A10, acquisition.py:170-193) for nfft = 2^22 as a pruned, real-input two-step transform.
//
// The reference transforms 10 ms of code-stripped samples, nvalid = 381 920 real values zero-padded to
// nfft = 8 * 2^19 = 4 194 304 points, and takes the arg-max of |X[k]| over k in [4, nfft/2 - 4).  The first version
// (sgx_acq.cu, ProFine/EpiFine) evaluates five 2^19-point sub-transforms in three HBM passes each: 80 MB of traffic
// and 26 launches of small tiles per 25 detections (profiles/ncu_summary_r2_v2.md: the 64-point middle pass sits at
// its barriers, 38 % issue utilisation).  Here the 2^22-point transform is split 2048 x 2048:
//
//   n = n1 + 2048 n2,  k = k2 + 2048 k1:
//   X[k2 + 2048 k1] = sum_n1  w_N^(n1 k2) * [ sum_n2 x[n1 + 2048 n2] w_2048^(n2 k2) ]  *  w_2048^(n1 k1)
//
//   step 1 (fine_cols_kernel): 2048-point transforms over n2 for every column n1.  Only the first 187 rows are
//     non-zero (zero padding), and the input is real: two real columns travel as one complex column and are separated
//     afterwards, and only k2 = 0..1024 is kept (the other half is the mirror image).  The inter-step twiddle
//     w_N^(n1 k2) is applied on the way out.  Output Y[k2][n1], 16.8 MB per detection.
//   step 2 (fine_rows_kernel): 2048-point transforms over n1 for the 1025 rows k2; outputs k1 < 1024 are the bins
//     k2 + 2048 k1, outputs k1 >= 1024 the mirrored bins (2048 - k2) + 2048 (2047 - k1) (real input: |X[N-k]| = |X[k]|),
//     so every bin below nfft/2 appears exactly once.  |.|^2 and the slice-relative arg-max are taken from the
//     registers of the last butterfly; nothing is written but one key per (detection, tile).
//
// Both kernels keep a whole [2048][4] tile in shared memory (64 KB; step 1 adds its twiddle table: two CTAs per SM, step 2 three) and run the transform in place (decimation in
// frequency, radices 16 x 16 x 8, output in digit-reversed positions -- the consumer computes the index instead of
// permuting).  Traffic per detection: 33.6 MB instead of 80; two launches for all detections of a call.
#include "sgx_fine.cuh"

namespace sgx {
namespace fine {

constexpr int R = 2048;        // length of both steps
constexpr int C = 4;           // complex columns of a tile (64 KB of shared memory per tile)
constexpr int CP = C;          // row stride; bank conflicts are avoided by the column swizzle below, not by padding
constexpr int NT = 256;
constexpr int ROWS_KEPT = R / 2 + 1;   // k2 = 0..1024

// Element (row, column c) of a tile lives at (row ^ rs(row))*4 + (c ^ cs(row)), rs = (row >> 3) & 3 (changes only the
// two low bits of the row), cs = ((row >> 2) ^ (row >> 4)) & 3.  With 8-byte elements a half-warp's 16 accesses are
// conflict-free when they fall into 16 different 8-byte bank pairs, i.e. 16 different (row' & 3, c') pairs:
//   * butterfly stages of block sizes 2048 and 128, the pruned first stage and the output walk of step 1: 4 consecutive
//     rows x 4 columns per half-warp -- rs and cs are constant over them, any swizzle works;
//   * last stage (radix 8, block 8): rows 8t + u for 4 consecutive t x 4 columns -- the rows agree in their two low
//     bits, rs = t & 3 separates them (without it: 4-way conflicts on every access of that stage, ncu r2x: 8
//     wavefronts per request instead of 2, a third of all shared-memory traffic of step 1);
//   * transposed tile load of step 2: 16 consecutive rows of one column -- cs takes four values over the four groups
//     of 4 rows, the low bits the other four.
// Both swizzles together are one 4-bit mask on the linear address: at = (row*4 + c) ^ mask(row), and the mask is
// linear over GF(2) in the bits of the row.  Hence for a row offset d that shares no bit with row0 (every stage below:
// row0 = block * L + j, d a multiple of the butterfly distance, or the per-iteration step of a thread)
//   at(row0 + d, c) = at(row0, c) ^ (4 d ^ mask(d)) = (at(row0, c) ^ low4) + high      (at_off; d a compile-time constant)
// -- one logic operation per access at most, none when d is a multiple of 64.
__host__ __device__ constexpr int swz_mask(int row) { return (((row >> 3) & 3) << 2) | (((row >> 2) ^ (row >> 4)) & 3); }
__device__ __forceinline__ int at(int row, int c) { return (row * CP + c) ^ swz_mask(row); }
__device__ __forceinline__ int at_off(int a0, int d) {
  const int k = (d * CP) ^ swz_mask(d);
  return (a0 ^ (k & 15)) + (k & ~15);
}

__host__ __device__ constexpr int digit_rev(int k) {   // k = ka + 16 kb + 256 kc  ->  position ka*128 + kb*8 + kc
  return (k & 15) * 128 + ((k >> 4) & 15) * 8 + (k >> 8);
}

// One decimation-in-frequency stage, in place: blocks of L rows, radix r; twiddles w_L^(j u) = W[(TL/L) j u mod TL],
// W a table of the TL-th roots of unity.  The block-128 stage uses the compact 128-entry table: in the 2048-entry one
// its twiddles are 128 bytes apart, every lane of a half-warp in the same bank (ncu r2x: 8 wavefronts per request).
// WG: the twiddle table is read from global memory through L1 (step 2: three CTAs per SM), else from shared memory.
template <int r, int L, bool WG, int TL = R>
__device__ __forceinline__ void dif_stage(cpx* x, const cpx* __restrict__ W, int tid) {
  static_assert(TL % L == 0, "table too short for this stage");
  constexpr int q = L / r, TSTEP = NT / C, ITERS = (R / r) * C / NT;
  static_assert((R / r) * C % NT == 0 && (q > TSTEP ? q % TSTEP == 0 : TSTEP % q == 0), "whole iterations; constant row step");
  // item t = t0 + TSTEP*it of column c: j = t % q, row0 = (t / q) * L + j = row0(it = 0) + DSTEP * it, no bits in common
  constexpr int DSTEP = q > TSTEP ? TSTEP : (TSTEP / q) * L;
  const int c = tid % C, t0 = tid / C;
  const int j0 = t0 % q;
  const int a0 = at((t0 / q) * L + j0, c);
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int j = q > TSTEP ? j0 + TSTEP * it : j0;
    cpx v[r];
#pragma unroll
    for (int u = 0; u < r; ++u) v[u] = x[at_off(a0, DSTEP * it + u * q)];
    fft::Dft<r, false>::run(v);
    if (L > r) {
#pragma unroll
      for (int u = 1; u < r; ++u) {
        const int wi = ((TL / L) * j * u) & (TL - 1);
        v[u] = fft::cmulf(v[u], WG ? __ldg(W + wi) : W[wi]);
      }
    }
#pragma unroll
    for (int u = 0; u < r; ++u) x[at_off(a0, DSTEP * it + u * q)] = v[u];
  }
}

// v * w_16^u, w_16 = exp(-2 pi i / 16), u a compile-time constant after unrolling (the trivial powers cost no multiply)
__device__ __forceinline__ cpx mul_w16(cpx v, int u) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, H = 0.70710678118654752f;
  switch (u & 15) {
    case 0: return v;
    case 4: return make_float2(v.y, -v.x);
    case 8: return make_float2(-v.x, -v.y);
    case 12: return make_float2(-v.y, v.x);
    case 2: return make_float2((v.x + v.y) * H, (v.y - v.x) * H);
    case 6: return make_float2((v.y - v.x) * H, -(v.x + v.y) * H);
    case 10: return make_float2(-(v.x + v.y) * H, (v.x - v.y) * H);
    case 14: return make_float2((v.x - v.y) * H, (v.x + v.y) * H);
    default: {
      // cos / -sin of 2 pi u / 16 for the odd powers
      const float c = (u == 1 || u == 15) ? C1 : (u == 3 || u == 13) ? S1 : (u == 5 || u == 11) ? -S1 : -C1;
      const float sn = (u == 1 || u == 7) ? -S1 : (u == 3 || u == 5) ? -C1 : (u == 9 || u == 15) ? S1 : C1;
      return make_float2(v.x * c - v.y * sn, v.x * sn + v.y * c);
    }
  }
}

// (x - mean) * code for the nvalid samples of every detection, zero-padded to whole rows of 2048 (acquisition.py:177)
__global__ void strip_kernel(Args a, int padded) {
  const FineItem it = a.items[blockIdx.y];
  const float mean = (float)((double)a.sums[it.rec] / (double)a.n_samples);
  const int8_t* sig = a.sig + (long long)it.rec * a.rec_stride + it.codePhase;
  const int8_t* chips = a.chips + it.prn * 1023;
  float* out = a.stripped + (long long)blockIdx.y * a.strip_stride;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < padded; n += gridDim.x * blockDim.x)
    out[n] = n < a.nvalid ? ((float)sig[n] - mean) * (float)chips[a.idx[n]] : 0.f;
}

// step 1: persistent CTAs walk the (column tile of 2 C real columns, detection) pairs
__global__ void __launch_bounds__(NT, 2) fine_cols_kernel(Args a) {
  SGX_DYN_SMEM(smem);
  cpx* x = reinterpret_cast<cpx*>(smem);
  cpx* W = x + R * CP;
  cpx* W128 = W + R;
  float4* S = reinterpret_cast<float4*>(W128 + 128);   // [16][C] per-tile twiddle steps of the output walk
  const int tid = threadIdx.x;
  for (int i = tid; i < R; i += NT) W[i] = a.w2048[i];        // once per (persistent) CTA
  if (tid < 128) W128[tid] = a.w128[tid];
  const int n_rows = (a.nvalid + R - 1) / R;    // rows n2 that hold samples (187 of 2048: the rest is zero padding)
  constexpr int COL_TILES = R / (2 * C), PIT = 128 * C / NT;
  static_assert(128 * C % NT == 0, "whole iterations");
  // the two non-zero inputs of this thread's first-stage butterflies, fetched one work item ahead (they travel during
  // the output walk of the previous tile).  Code-stripped samples of the detection (strip_kernel), padded with zeros to
  // whole rows: two reals = one element
  cpx p0[PIT], p1[PIT];
  auto fetch = [&](int work) {
    const int tile = work % COL_TILES, item = work / COL_TILES;
    const cpx* xs = reinterpret_cast<const cpx*>(a.stripped + (long long)item * a.strip_stride) + tile * C;
#pragma unroll
    for (int it = 0; it < PIT; ++it) {
      const int c = tid % C, j = tid / C + (NT / C) * it;
      p0[it] = j < n_rows ? __ldg(xs + (long long)j * (R / 2) + c) : make_float2(0.f, 0.f);
      p1[it] = j + 128 < n_rows ? __ldg(xs + (long long)(j + 128) * (R / 2) + c) : make_float2(0.f, 0.f);
    }
  };
  if (blockIdx.x < COL_TILES * a.n_items) fetch(blockIdx.x);
#pragma unroll 1
  for (int work = blockIdx.x; work < COL_TILES * a.n_items; work += gridDim.x) {
  const int tile = work % COL_TILES, item = work / COL_TILES;
  if (tid < 16 * C) {   // S[it][c] = (w_N^(n1 it), w_N^((n1 + 1) it)), n1 = first real column of complex column c (read after 3 barriers)
    const int it = tid / C, n1 = tile * 2 * C + 2 * (tid % C);
    const int m = n1 * it;                                    // < 2^15
    const cpx sa = fft::cmulf(a.w2048[m >> 11], __ldg(a.wlo + (m & (R - 1))));
    const cpx sb = fft::cmulf(sa, __ldg(a.wlo + it));
    S[tid] = make_float4(sa.x, sa.y, sb.x, sb.y);
  }
  // Pruned first stage (radix 16 over rows j + 128 u): only u = 0 and u = 1 can be non-zero, so
  //   y[u'] = (v0 + w_16^u' v1) w_2048^(j u')
  // straight from the prefetched registers into the tile (requires n_rows <= 256; it is 187).
  __syncthreads();   // the twiddle table is complete
#pragma unroll
  for (int it = 0; it < PIT; ++it) {
    const int c = tid % C, j = tid / C + (NT / C) * it;          // rows j + 128 u: j < 128 shares no bit with 128 u
    const int a0 = at(tid / C, c);
    const cpx v0 = p0[it], v1 = p1[it];
    x[at_off(a0, (NT / C) * it)] = fft::cadd(v0, v1);
#pragma unroll
    for (int u = 1; u < 16; ++u)   // W[(128 + j) u] = w_16^u W[j u]: one table read per output
      x[at_off(a0, (NT / C) * it + 128 * u)] = fft::cmulf(fft::cadd(v0, mul_w16(v1, u)), W[(j * u) & (R - 1)]);
  }
  __syncthreads();
  dif_stage<16, 128, false, 128>(x, W128, tid);
  __syncthreads();
  dif_stage<8, 8, false>(x, W, tid);
  __syncthreads();
  if (work + (int)gridDim.x < COL_TILES * a.n_items) fetch(work + gridDim.x);
  // separate the two real columns of every complex column, apply w_N^(n1 k2), store Y[k2][n1].  Thread (c, kc, kb) walks
  // k2 = it + 16 kb + 256 kc, it = 0..15: the value sits at tile row it*128 + kb*8 + kc, its mirror R - k2 at
  // (16 - it)*128 + (15 - kb)*8 + 7 - kc (it >= 1) -- constant offsets from two per-thread addresses; a half-warp reads
  // 4 consecutive rows x 4 columns, conflict-free.  The twiddle w_N^(n1 k2) = w_N^(n1 k2(it = 0)) * S[it][c]: one
  // two-level table product per thread and tile instead of one per value (random gathers: up to 7 wavefronts each).
  // k2 = 1024 (position 4) is done by the first C threads.
  cpx* out = a.y + ((long long)item * ROWS_KEPT) * R + tile * 2 * C;
  {
    const int c = tid & (C - 1), kc = (tid >> 2) & 3, kb = tid >> 4;
    const int k20 = 16 * kb + 256 * kc;
    const int m0 = (tile * 2 * C + 2 * c) * k20;               // (n1 + 1) k2 < 2^21: no wrap modulo N = 2^22
    const cpx ba = fft::cmulf(W[m0 >> 11], __ldg(a.wlo + (m0 & (R - 1))));
    const cpx bb = fft::cmulf(ba, __ldg(a.wlo + k20));
    const int ap = at(kb * 8 + kc, c), am = at((15 - kb) * 8 + 7 - kc, c);
    const int a_first = at(digit_rev((R - k20) & (R - 1)), c);
    cpx* o = out + (long long)k20 * R + 2 * c;
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const cpx zp = x[at_off(ap, 128 * it)];
      const cpx zm = it == 0 ? x[a_first] : x[at_off(am, (16 - it) * 128)];
      const float4 s4 = S[it * C + c];
      cpx ya = make_float2(zp.x + zm.x, zp.y - zm.y);           // 2 X_a[k2]
      cpx yb = make_float2(zp.y + zm.y, zm.x - zp.x);           // 2 X_b[k2]
      ya = fft::cmulf(ya, fft::cmulf(ba, make_float2(s4.x, s4.y)));
      yb = fft::cmulf(yb, fft::cmulf(bb, make_float2(s4.z, s4.w)));
      *reinterpret_cast<float4*>(o + (long long)it * R) = make_float4(ya.x, ya.y, yb.x, yb.y);
    }
  }
  if (tid < C) {
    const int c = tid, k2 = R / 2;
    const cpx zp = x[at(digit_rev(k2), c)], zm = zp;             // R - k2 = k2
    cpx ya = make_float2(zp.x + zm.x, zp.y - zm.y);
    cpx yb = make_float2(zp.y + zm.y, zm.x - zp.x);
    const int ma = (tile * 2 * C + 2 * c) * k2;
    const cpx ta = fft::cmulf(W[ma >> 11], __ldg(a.wlo + (ma & (R - 1))));
    ya = fft::cmulf(ya, ta);
    yb = fft::cmulf(yb, fft::cmulf(ta, __ldg(a.wlo + k2)));
    *reinterpret_cast<float4*>(out + (long long)k2 * R + 2 * c) = make_float4(ya.x, ya.y, yb.x, yb.y);
  }
  __syncthreads();   // the tile is rewritten by the next work item
  }
}

// step 2: blockIdx.x = row tile (C values of k2), blockIdx.y = detection
__global__ void __launch_bounds__(NT, 3) fine_rows_kernel(Args a) {
  SGX_DYN_SMEM(smem);
  cpx* x = reinterpret_cast<cpx*>(smem);
  const cpx* W = a.w2048;
  __shared__ unsigned long long red[NT / 32];
  const int tid = threadIdx.x, tile = blockIdx.x;
  const cpx* in = a.y + ((long long)blockIdx.y * ROWS_KEPT + tile * C) * R;
  const int rows = min(C, ROWS_KEPT - tile * C);
  const int a_t = at(tid, 0);                // at(tid + 256 it, rr) = (a_t + 1024 it) ^ rr
#pragma unroll 2
  for (int it = 0; it < R / NT; ++it) {      // all rows of a column in flight before the first store
    const int n1 = tid + NT * it;
    cpx v[C];
#pragma unroll
    for (int rr = 0; rr < C; ++rr) v[rr] = rr < rows ? __ldcs(in + (long long)rr * R + n1) : make_float2(0.f, 0.f);
#pragma unroll
    for (int rr = 0; rr < C; ++rr) x[(a_t + NT * CP * it) ^ rr] = v[rr];
  }
  __syncthreads();
  dif_stage<16, 2048, true>(x, W, tid);
  __syncthreads();
  // compact table through L1 as well: one more kilobyte of shared memory would push three CTAs past the 196 KB
  // carve-out and take 32 KB of L1 away from the twiddles of the first stage (measured: 0.69 -> 0.82 ms per 85 items)
  dif_stage<16, 128, true, 128>(x, a.w128, tid);
  __syncthreads();
  // last stage (radix 8, no twiddles): outputs stay in registers -> |.|^2, bin index, arg-max
  unsigned long long best = 0ull;
  const int a_l = at((tid / C) * 8, tid % C);
  for (int it = 0; it < (R / 8) * C / NT; ++it) {
    const int c = tid % C, t = tid / C + (NT / C) * it;   // t = ka*16 + kb: the outputs of this butterfly are k1 = ka + 16 kb + 256 u
    cpx v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = x[at_off(a_l + (NT / C) * 8 * CP * it, u)];
    fft::Dft<8, false>::run(v);
    if (c < rows) {
      const int k2 = tile * C + c;
      const int k1lo = (t >> 4) + 16 * (t & 15);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k1 = k1lo + 256 * u;
        int k;
        if (k1 < R / 2) k = k2 + R * k1;
        else if (k2 > 0 && k2 < R / 2) k = (R - k2) + R * (R - 1 - k1);
        else continue;
        if (k < a.lo || k >= a.hi) continue;
        const unsigned long long key = fft::peak_key(fmaf(v[u].x, v[u].x, v[u].y * v[u].y), (unsigned)(k - a.lo));
        best = key > best ? key : best;
      }
    }
  }
#pragma unroll
  for (int msk = 16; msk > 0; msk >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, msk);
    best = o > best ? o : best;
  }
  if ((tid & 31) == 0) red[tid >> 5] = best;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < NT / 32; ++w) best = red[w] > best ? red[w] : best;
    a.partial[(long long)blockIdx.y * gridDim.x + tile] = best;
  }
}

__global__ void fine_argmax_kernel(const unsigned long long* partial, int ntiles, int* index) {
  __shared__ unsigned long long red[4];
  unsigned long long best = 0ull;
  for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
    const unsigned long long k = partial[(long long)blockIdx.x * ntiles + t];
    best = k > best ? k : best;
  }
#pragma unroll
  for (int msk = 16; msk > 0; msk >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, msk);
    best = o > best ? o : best;
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)blockDim.x / 32; ++w) best = red[w] > best ? red[w] : best;
    index[blockIdx.x] = (int)fft::key_index(best);
  }
}

__global__ void fine_tables_kernel(cpx* w2048, cpx* wlo, cpx* w128) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  double s, c;
  sincospi(-2.0 * (double)i / (double)R, &s, &c);
  w2048[i] = make_float2((float)c, (float)s);
  if (i % (R / 128) == 0) w128[i / (R / 128)] = w2048[i];
  sincospi(-2.0 * (double)i / (double)NFFT, &s, &c);
  wlo[i] = make_float2((float)c, (float)s);
}

struct Scratch {
  DevBuf w2048, wlo, w128, y, partial, stripped;
  bool tables = false;
};
static Scratch g_fine;

int run(Args a, int n_items, int* d_index, cudaStream_t s) {
  if (n_items <= 0) return SGX_OK;
  static_assert(C == 4, "index arithmetic of the output phase of step 1");
  if ((a.nvalid + R - 1) / R > 256) return fail(SGX_ERR_ARG, "fine search", "more than 256 non-zero rows");
  Scratch& g = g_fine;
  if (!g.tables) {
    if (g.w2048.reserve(sizeof(cpx) * R) || g.wlo.reserve(sizeof(cpx) * R) || g.w128.reserve(sizeof(cpx) * 128))
      return fail(SGX_ERR_CUDA, "cudaMalloc", "fine tables");
    SGX_COUNTED_LAUNCH(fine_tables_kernel, dim3(R / 256), dim3(256), 0, s, g.w2048.as<cpx>(), g.wlo.as<cpx>(), g.w128.as<cpx>());
    g.tables = true;
  }
  const int col_tiles = R / (2 * C), row_tiles = (ROWS_KEPT + C - 1) / C;
  long long chunk = (2LL << 30) / ((long long)sizeof(cpx) * ROWS_KEPT * R);      // detections per 2 GB of Y
  if (chunk < 1) chunk = 1;
  if (chunk > n_items) chunk = n_items;
  const int padded = (a.nvalid + R - 1) / R * R;
  if (g.stripped.reserve(sizeof(float) * (size_t)chunk * padded)) return fail(SGX_ERR_CUDA, "cudaMalloc", "fine search scratch");
  a.stripped = g.stripped.as<float>();
  a.strip_stride = padded;
  if (g.y.reserve(sizeof(cpx) * (size_t)chunk * ROWS_KEPT * R) ||
      g.partial.reserve(sizeof(unsigned long long) * (size_t)chunk * row_tiles))
    return fail(SGX_ERR_CUDA, "cudaMalloc", "fine search scratch");
  const size_t smem = sizeof(cpx) * (size_t)R * CP;            // step 2: tile only (twiddles through L1)
  const size_t smem_cols = smem + sizeof(cpx) * (R + 128) + sizeof(float4) * 16 * C;  // step 1: tile + twiddle tables (2048, 128) + per-tile steps
  SGX_CUDA(cudaFuncSetAttribute(fine_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
  SGX_CUDA(cudaFuncSetAttribute(fine_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  a.w2048 = g.w2048.as<cpx>();
  a.wlo = g.wlo.as<cpx>();
  a.w128 = g.w128.as<cpx>();
  a.y = g.y.as<cpx>();
  a.partial = g.partial.as<unsigned long long>();
  const FineItem* items = a.items;
  for (int i0 = 0; i0 < n_items; i0 += (int)chunk) {
    const int cnt = n_items - i0 < chunk ? n_items - i0 : (int)chunk;
    a.items = items + i0;
    SGX_COUNTED_LAUNCH(strip_kernel, dim3(64, cnt), dim3(256), 0, s, a, padded);
    a.n_items = cnt;
    {
      int dev = 0, n_sm = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
      long long grid = 2LL * (n_sm > 0 ? n_sm : 148);
      if (grid > (long long)col_tiles * cnt) grid = (long long)col_tiles * cnt;
      SGX_COUNTED_LAUNCH(fine_cols_kernel, dim3((unsigned)grid), dim3(NT), smem_cols, s, a);
    }
    SGX_COUNTED_LAUNCH(fine_rows_kernel, dim3(row_tiles, cnt), dim3(NT), smem, s, a);
    SGX_COUNTED_LAUNCH(fine_argmax_kernel, dim3(cnt), dim3(128), 0, s, a.partial, row_tiles, d_index + i0);
  }
  SGX_CUDA(cudaGetLastError());
  return SGX_OK;
}

}  // namespace fine
}  // namespace sgx
