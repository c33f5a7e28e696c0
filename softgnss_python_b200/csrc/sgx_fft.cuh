// Mixed-radix FFT engine for the acquisition path (sizes such as 38192 = 16*7*11*31 and 2^22).
//
// A transform of length N is a short sequence of "passes" (Stockham autosort, decimation in time).
// Pass p has a composite radix R (<= 512) and Ls = product of the earlier radices, m = N / R:
//
//     for j in [0, m):  k = j mod Ls
//        y[(j-k)*R + k + u*Ls] = sum_t  x[j + t*m] * w_{Ls*R}^{t*k} * w_R^{t*u}        u in [0, R)
//
// One CTA owns T consecutive j ("columns"): it gathers the T x R inputs (rows of T contiguous
// elements -> coalesced), applies the inter-pass twiddles, runs the R-point sub-FFT of all T
// columns in shared memory (again Stockham, radices from {2,3,4,5,7,8,11,16,31}, butterflies in
// registers with constant-bank twiddles) and scatters the result, or hands it to a fused epilogue
// (|.|^2 + arg-max) without ever writing it.  Input prologues are fused the same way (int8 samples
// x carrier NCO, spectrum x code spectrum, code-stripped real samples).  Output is in natural order
// after the last pass, so bin/code-phase indices need no permutation.
//
// Tensor cores are deliberately not used: the 31-point stage as a dense DFT GEMM costs 4x the FMAs
// of the symmetric butterfly below and would need a 3xTF32 split to hold the 1e-6 accuracy
// ("Why not tcgen05").
#pragma once
#include "sgx_common.cuh"
#include "sgx_fft_tables.h"

namespace sgx {
namespace fft {

typedef float2 cpx;
constexpr int MAX_SUB = 6;
constexpr int FFT_THREADS = 128;

struct Pass {
  int N, R, Ls, m, T, nsub, ntiles, inverse;
  int radix[MAX_SUB];
  const cpx* twg;  // [R][Ls]: w_{Ls*R}^{t*k} (forward sign); nullptr when Ls == 1
  const cpx* wr;   // [R]    : w_R^q           (forward sign)
};

__device__ __forceinline__ cpx cmulf(cpx a, cpx b) {
  return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return make_float2(a.x - b.x, a.y - b.y); }
template <bool INV> __device__ __forceinline__ cpx rot90(cpx a) {
  // forward: multiply by -i ; inverse: multiply by +i
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// ---- butterflies (in place, natural order in and out) ---------------------------------------
template <int N, bool INV> struct Pow2 {
  static __device__ __forceinline__ void run(cpx* v) {
    cpx e[N / 2], o[N / 2];
#pragma unroll
    for (int i = 0; i < N / 2; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    Pow2<N / 2, INV>::run(e);
    Pow2<N / 2, INV>::run(o);
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
      cpx t;
      if (k == 0) t = o[0];
      else if (4 * k == N) t = rot90<INV>(o[k]);
      else {
        const float c = COS16[k * (16 / N)], s = SIN16[k * (16 / N)];
        t = cmulf(o[k], make_float2(c, INV ? s : -s));
      }
      v[k] = cadd(e[k], t);
      v[k + N / 2] = csub(e[k], t);
    }
  }
};
template <bool INV> struct Pow2<1, INV> {
  static __device__ __forceinline__ void run(cpx*) {}
};

template <int P, bool INV>
__device__ __forceinline__ void dft_prime(cpx* v, const float* C, const float* S) {
  // conjugate-pair form: a_j = x_j + x_{P-j}, b_j = x_j - x_{P-j};  X_k = C_k -/+ i S_k
  constexpr int H = (P - 1) / 2;
  cpx a[H], b[H];
#pragma unroll
  for (int j = 1; j <= H; ++j) { a[j - 1] = cadd(v[j], v[P - j]); b[j - 1] = csub(v[j], v[P - j]); }
  const cpx x0 = v[0];
  cpx sum = x0;
#pragma unroll
  for (int j = 0; j < H; ++j) sum = cadd(sum, a[j]);
  v[0] = sum;
#pragma unroll
  for (int k = 1; k <= H; ++k) {
    float cr = x0.x, ci = x0.y, sr = 0.f, si = 0.f;
#pragma unroll
    for (int j = 1; j <= H; ++j) {
      const int q = (j * k) % P;
      cr = fmaf(a[j - 1].x, C[q], cr);
      ci = fmaf(a[j - 1].y, C[q], ci);
      sr = fmaf(b[j - 1].x, S[q], sr);
      si = fmaf(b[j - 1].y, S[q], si);
    }
    const cpx lo = make_float2(cr + si, ci - sr), hi = make_float2(cr - si, ci + sr);
    v[k] = INV ? hi : lo;
    v[P - k] = INV ? lo : hi;
  }
}

template <int R, bool INV> struct Dft;
template <bool INV> struct Dft<2, INV> { static __device__ __forceinline__ void run(cpx* v) { Pow2<2, INV>::run(v); } };
template <bool INV> struct Dft<4, INV> { static __device__ __forceinline__ void run(cpx* v) { Pow2<4, INV>::run(v); } };
template <bool INV> struct Dft<8, INV> { static __device__ __forceinline__ void run(cpx* v) { Pow2<8, INV>::run(v); } };
template <bool INV> struct Dft<16, INV> { static __device__ __forceinline__ void run(cpx* v) { Pow2<16, INV>::run(v); } };
template <bool INV> struct Dft<3, INV> { static __device__ __forceinline__ void run(cpx* v) { dft_prime<3, INV>(v, COS3, SIN3); } };
template <bool INV> struct Dft<5, INV> { static __device__ __forceinline__ void run(cpx* v) { dft_prime<5, INV>(v, COS5, SIN5); } };
template <bool INV> struct Dft<7, INV> { static __device__ __forceinline__ void run(cpx* v) { dft_prime<7, INV>(v, COS7, SIN7); } };
template <bool INV> struct Dft<11, INV> { static __device__ __forceinline__ void run(cpx* v) { dft_prime<11, INV>(v, COS11, SIN11); } };
template <bool INV> struct Dft<31, INV> { static __device__ __forceinline__ void run(cpx* v) { dft_prime<31, INV>(v, COS31, SIN31); } };

constexpr int TILE = 16;        // columns per CTA: 16 x 8 B = one 128-byte line per row
constexpr int TILE_P = TILE + 1;  // padded row stride in shared memory
constexpr int ROWS_PER_ITER = FFT_THREADS / TILE;

// One Stockham sub-pass of radix r over the R x TILE tile held in shared memory.
template <int r, bool INV>
__device__ __forceinline__ void subpass(const cpx* __restrict__ in, cpx* __restrict__ out,
                                        const cpx* __restrict__ W, int R, int ls) {
  const int mm = R / r;
  const int total = mm * TILE;
  const int wstride = R / (ls * r);
  for (int idx = threadIdx.x; idx < total; idx += FFT_THREADS) {
    const int jj = idx & (TILE - 1);
    const int b = idx / TILE;
    const int kk = ls > 1 ? b % ls : 0;
    cpx v[r];
#pragma unroll
    for (int u = 0; u < r; ++u) v[u] = in[(b + u * mm) * TILE_P + jj];
    if (ls > 1) {
#pragma unroll
      for (int u = 1; u < r; ++u) v[u] = cmulf(v[u], W[u * kk * wstride]);
    }
    Dft<r, INV>::run(v);
    const int base = (b - kk) * r + kk;
#pragma unroll
    for (int u = 0; u < r; ++u) out[(base + u * ls) * TILE_P + jj] = v[u];
  }
}

// Last sub-pass of a pass (ls == R / r): base + u*ls == b + u*mm, i.e. every butterfly writes the rows it read.
template <int r, bool INV>
__device__ __forceinline__ void subpass_inplace(cpx* buf, const cpx* __restrict__ W, int R, int ls) {
  const int mm = R / r;
  const int total = mm * TILE;
  const int wstride = R / (ls * r);
  for (int idx = threadIdx.x; idx < total; idx += FFT_THREADS) {
    const int jj = idx & (TILE - 1);
    const int b = idx / TILE;
    const int kk = ls > 1 ? b % ls : 0;
    cpx v[r];
#pragma unroll
    for (int u = 0; u < r; ++u) v[u] = buf[(b + u * mm) * TILE_P + jj];
    if (ls > 1) {
#pragma unroll
      for (int u = 1; u < r; ++u) v[u] = cmulf(v[u], W[u * kk * wstride]);
    }
    Dft<r, INV>::run(v);
    const int base = (b - kk) * r + kk;
#pragma unroll
    for (int u = 0; u < r; ++u) buf[(base + u * ls) * TILE_P + jj] = v[u];
  }
}

template <bool INV, bool BIG>
__device__ __forceinline__ void run_subpass(int r, const cpx* in, cpx* out, const cpx* W, int R, int ls) {
  switch (r) {
    case 2: subpass<2, INV>(in, out, W, R, ls); break;
    case 4: subpass<4, INV>(in, out, W, R, ls); break;
    case 8: subpass<8, INV>(in, out, W, R, ls); break;
    case 16: subpass<16, INV>(in, out, W, R, ls); break;
    default:
      if (BIG) {
        switch (r) {
          case 3: subpass<3, INV>(in, out, W, R, ls); break;
          case 5: subpass<5, INV>(in, out, W, R, ls); break;
          case 7: subpass<7, INV>(in, out, W, R, ls); break;
          case 11: subpass<11, INV>(in, out, W, R, ls); break;
          case 31: subpass<31, INV>(in, out, W, R, ls); break;
        }
      }
  }
}

// Pro:  void prepare(int batch);  cpx load(int n)            -- element n of transform `batch`
// Epi:  void begin(int batch);    void put(int n, cpx v);    void finish(int batch, int tile)
// Every thread owns one column (jj = tid % 16) for the whole kernel, so all index arithmetic that
// depends on the column or on the batch is done once.
template <class Pro, class Epi, bool INV, bool BIG, int R0 = 0, int R1 = 0>
__global__ void __launch_bounds__(FFT_THREADS) fft_pass_kernel(Pass P, Pro pro, Epi epi) {
  SGX_DYN_SMEM(smem);
  const int R = R0 ? R0 * R1 : P.R;   // R0, R1: compile-time radices of a two-sub-pass shape (see the persistent kernel)
  const int m = P.m, Ls = P.Ls;
  cpx* A = reinterpret_cast<cpx*>(smem);
  cpx* B = A + R * TILE_P;
  cpx* W = B + R * TILE_P;
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int batch = blockIdx.y;
  const int j0 = tile * TILE;
  const int jj = tid & (TILE - 1), t0 = tid / TILE;
  const int j = j0 + jj;
  const bool valid = j < m;
  const int k = Ls > 1 ? j % Ls : 0;

  pro.prepare(batch);
  for (int q = tid; q < R; q += FFT_THREADS) {
    cpx w = P.wr[q];
    if (INV) w.y = -w.y;
    W[q] = w;
  }
  if (Ls > 1) {
    const cpx* tw = P.twg + k;
#pragma unroll 4
    for (int t = t0; t < R; t += ROWS_PER_ITER) {
      cpx v = make_float2(0.f, 0.f);
      if (valid) {
        v = pro.load(j + t * m);
        cpx w = tw[(size_t)t * Ls];
        if (INV) w.y = -w.y;
        v = cmulf(v, w);
      }
      A[t * TILE_P + jj] = v;
    }
  } else {
#pragma unroll 4
    for (int t = t0; t < R; t += ROWS_PER_ITER) {
      cpx v = make_float2(0.f, 0.f);
      if (valid) v = pro.load(j + t * m);
      A[t * TILE_P + jj] = v;
    }
  }
  __syncthreads();
  cpx* src = A;
  cpx* dst = B;
  if (R0) {
    subpass<(R0 ? R0 : 2), INV>(A, B, W, R, 1);
    __syncthreads();
    subpass_inplace<(R1 ? R1 : 2), INV>(B, W, R, R0);
    __syncthreads();
    src = B;
  } else {
    int ls = 1;
    for (int s = 0; s < P.nsub; ++s) {
      const int r = P.radix[s];
      run_subpass<INV, BIG>(r, src, dst, W, R, ls);
      ls *= r;
      __syncthreads();
      cpx* tmp = src; src = dst; dst = tmp;
    }
  }
  epi.begin(batch);
  if (Ls == 1) {
    // y[j*R + u]: one column is R contiguous outputs
    const int ncol = min(TILE, m - j0);
    for (int c = 0; c < ncol; ++c) {
      const int base = (j0 + c) * R;
      for (int u = tid; u < R; u += FFT_THREADS) epi.put(base + u, src[u * TILE_P + c]);
    }
  } else if (valid) {
    const int base = (j - k) * R + k;
#pragma unroll 4
    for (int u = t0; u < R; u += ROWS_PER_ITER) epi.put(base + u * Ls, src[u * TILE_P + jj]);
  }
  epi.finish(batch, tile);
}

// ---- asynchronous, persistent variant of the pass kernel ----------------------------------------
// Same mathematics as fft_pass_kernel, restructured after the first ncu captures (profiles/): the
// synchronous version spent ~65 % of its warp samples waiting for the global loads of the tile.  Here
//  * a CTA owns one tile position (16 columns) and walks over `items_per_cta` consecutive transforms;
//  * the streamed operand of a transform is staged with cp.async (no register staging, all rows in
//    flight at once) into shared memory;
//  * the second operand -- the inter-pass twiddle tile, or the code-spectrum tile of the current PRN
//    -- stays in shared memory and is reloaded only when it changes, which also removes the L2 hot
//    spot of dozens of CTAs re-reading the same lines;
//  * the product of the two is formed while the first butterfly reads its inputs.
enum AuxKind { AUX_NONE = 0, AUX_SAME = 1, AUX_TWIDDLE = 2 };

// Src:  int locate(int batch, const cpx*& src, const cpx*& aux) const   -> AuxKind
template <int r, bool INV, int AUX>
__device__ __forceinline__ void subpass_first(const cpx* __restrict__ s1, const cpx* __restrict__ s2,
                                              cpx* __restrict__ out, int R) {
  const int mm = R / r;
  const int total = mm * TILE;
  for (int idx = threadIdx.x; idx < total; idx += FFT_THREADS) {
    const int jj = idx & (TILE - 1);
    const int b = idx / TILE;
    cpx v[r];
#pragma unroll
    for (int u = 0; u < r; ++u) {
      const int e = (b + u * mm) * TILE + jj;
      cpx x = s1[e];
      if (AUX != AUX_NONE) {
        cpx w = s2[e];
        if (AUX == AUX_TWIDDLE && INV) w.y = -w.y;
        x = cmulf(x, w);
      }
      v[u] = x;
    }
    Dft<r, INV>::run(v);
    const int base = b * r;   // ls == 1
#pragma unroll
    for (int u = 0; u < r; ++u) out[(base + u) * TILE_P + jj] = v[u];
  }
}

template <bool INV, bool BIG, int AUX>
__device__ __forceinline__ void run_subpass_first(int r, const cpx* s1, const cpx* s2, cpx* out, int R) {
  switch (r) {
    case 2: subpass_first<2, INV, AUX>(s1, s2, out, R); break;
    case 4: subpass_first<4, INV, AUX>(s1, s2, out, R); break;
    case 8: subpass_first<8, INV, AUX>(s1, s2, out, R); break;
    case 16: subpass_first<16, INV, AUX>(s1, s2, out, R); break;
    default:
      if (BIG) {
        switch (r) {
          case 3: subpass_first<3, INV, AUX>(s1, s2, out, R); break;
          case 5: subpass_first<5, INV, AUX>(s1, s2, out, R); break;
          case 7: subpass_first<7, INV, AUX>(s1, s2, out, R); break;
          case 11: subpass_first<11, INV, AUX>(s1, s2, out, R); break;
          case 31: subpass_first<31, INV, AUX>(s1, s2, out, R); break;
        }
      }
  }
}

// First sub-pass of the compile-time-radix kernel: the streamed operand goes from global memory straight into the
// butterfly's registers (r independent 8-byte loads per thread, the 16 threads of a row read one 128-byte line), is
// multiplied by the cached second operand and leaves as the sub-pass result in A.  No staging tile, no cp.async.
// AUXD: the second operand is read from global memory (L2) too; otherwise from the cached tile s2 [R][16].
__host__ __device__ constexpr bool aux_direct(int r0) { return r0 != 0 && r0 <= 16; }   // 2 x 31 operands do not fit the register file
template <int r, bool INV, int AUX, bool AUXD>
__device__ __forceinline__ void subpass_first_direct(const cpx* __restrict__ src, size_t row_stride, bool valid,
                                                     const cpx* __restrict__ aux, size_t aux_stride,
                                                     const cpx* __restrict__ s2, cpx* __restrict__ out, int R) {
  const int mm = R / r;
  const int total = mm * TILE;
  for (int idx = threadIdx.x; idx < total; idx += FFT_THREADS) {
    const int jj = idx & (TILE - 1);
    const int b = idx / TILE;
    cpx v[r];
#pragma unroll
    for (int u = 0; u < r; ++u) v[u] = valid ? __ldg(src + (size_t)(b + u * mm) * row_stride) : make_float2(0.f, 0.f);
    if (AUX != AUX_NONE) {
      // second operand (code spectrum / inter-pass twiddles) from L2 as well: it is shared by all CTAs of the launch
      cpx w[r];
#pragma unroll
      for (int u = 0; u < r; ++u) {
        if (AUXD) w[u] = valid ? __ldg(aux + (size_t)(b + u * mm) * aux_stride) : make_float2(0.f, 0.f);
        else w[u] = s2[(b + u * mm) * TILE + jj];
      }
#pragma unroll
      for (int u = 0; u < r; ++u) {
        if (AUX == AUX_TWIDDLE && INV) w[u].y = -w[u].y;
        v[u] = cmulf(v[u], w[u]);
      }
    }
    Dft<r, INV>::run(v);
    const int base = b * r;   // ls == 1
#pragma unroll
    for (int u = 0; u < r; ++u) out[(base + u) * TILE_P + jj] = v[u];
  }
}

// R0, R1 != 0: the pass is exactly two sub-passes of radices R0 and R1 known at compile time (the hot shapes
// 217 = 31 x 7, 176 = 16 x 11 of the 38192-point search and 256 = 16 x 16, 128 = 16 x 8 of the fine search), so
// every loop over rows has a constant trip count and the shared-memory index arithmetic folds into immediates --
// profiles/ncu_summary_r1_v4.md: the generic kernel spends ~40 % of its instructions on integer/branch work.
template <class Src, class Epi, bool INV, bool BIG, int AUX, int R0 = 0, int R1 = 0>
__global__ void __launch_bounds__(FFT_THREADS, (R0 == 16 ? 6 : 1)) fft_pass_async_kernel(Pass P, Src srcd, Epi epi, int n_batch,
                                                                      int items_per_cta) {
  SGX_DYN_SMEM(smem);
  const int R = R0 ? R0 * R1 : P.R;
  const int m = P.m, Ls = P.Ls;
  // generic: S1 staging / second work buffer, A, S2, W.  Compile-time radices: A, S2, W only (direct loads, the
  // last sub-pass runs in place) -- smem_direct() bytes, which lets a third (R = 217) / fourth (R = 176) CTA onto the SM (more once the second operand is read from L2 too).
  cpx* S1 = reinterpret_cast<cpx*>(smem);          // staged operand [R][16]; later a padded work buffer
  cpx* A = R0 ? S1 : S1 + R * TILE_P;              // padded work buffer [R][17]
  cpx* S2 = A + R * TILE_P;                        // second operand tile [R][16] (generic kernel only)
  constexpr bool AUXD = aux_direct(R0);
  cpx* W = AUXD ? S2 : S2 + R * TILE;
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int j0 = tile * TILE;
  const int ncol = min(TILE, m - j0);
  const int jj = tid & (TILE - 1), t0 = tid / TILE;
  const int j = j0 + jj;
  const bool valid = j < m;
  const int k = Ls > 1 ? j % Ls : 0;
  const int k0 = Ls > 1 ? j0 % Ls : 0;
  const int b_begin = blockIdx.y * items_per_cta;
  const int b_end = min(n_batch, b_begin + items_per_cta);

  for (int q = tid; q < R; q += FFT_THREADS) {
    cpx w = P.wr[q];
    if (INV) w.y = -w.y;
    W[q] = w;
  }
  const cpx* aux_cached = nullptr;
  for (int batch = b_begin; batch < b_end; ++batch) {
    const cpx* src;
    const cpx* aux;
    srcd.locate(batch, src, aux);
    if (!AUXD && AUX != AUX_NONE && aux != aux_cached) {
      if (AUX == AUX_SAME) {
        for (int t = t0; t < R; t += ROWS_PER_ITER)
          if (valid) cp_async8(&S2[t * TILE + jj], aux + (size_t)t * m + j);
      } else {
        for (int t = t0; t < R; t += ROWS_PER_ITER)
          if (valid) cp_async8(&S2[t * TILE + jj], aux + (size_t)t * Ls + k0 + jj);
      }
      aux_cached = aux;
    }
    if (!R0) {
      for (int t = t0; t < R; t += ROWS_PER_ITER)
        if (valid) cp_async8(&S1[t * TILE + jj], src + (size_t)t * m + j);
    }
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();   // second operand (and staged tile) visible; A free (previous epilogue finished)
    // first sub-pass: operands -> A
    cpx* cur = A;
    cpx* oth = S1;
    if (R0) {
      subpass_first_direct<(R0 ? R0 : 2), INV, AUX, AUXD>(src + j, (size_t)m, valid,
                                                          AUX == AUX_SAME ? aux + j : aux + k0 + jj,
                                                          AUX == AUX_SAME ? (size_t)m : (size_t)Ls, S2, A, R);
      __syncthreads();
      subpass_inplace<(R1 ? R1 : 2), INV>(A, W, R, R0);   // last sub-pass: ls == R / r, every butterfly writes where it read
      __syncthreads();
    } else {
      run_subpass_first<INV, BIG, AUX>(P.radix[0], S1, S2, A, R);
      __syncthreads();
      int ls = P.radix[0];
      for (int s = 1; s < P.nsub; ++s) {
        const int r = P.radix[s];
        run_subpass<INV, BIG>(r, cur, oth, W, R, ls);
        ls *= r;
        __syncthreads();
        cpx* tmp = cur; cur = oth; oth = tmp;
      }
    }
    epi.begin(batch);
    if (Ls == 1) {
      for (int c = 0; c < ncol; ++c) {
        const int base = (j0 + c) * R;
        for (int u = tid; u < R; u += FFT_THREADS) epi.put(base + u, cur[u * TILE_P + c]);
      }
    } else if (valid) {
      const int base = (j - k) * R + k;
#pragma unroll 4
      for (int u = t0; u < R; u += ROWS_PER_ITER) epi.put(base + u * Ls, cur[u * TILE_P + jj]);
    }
    epi.finish(batch, tile);
    if (!R0) __syncthreads();   // S1 is refilled by the next transform's cp.async (compile-time radices: the barrier
                                // at the top of the loop is the only one needed before A is rewritten)
  }
}

// ---- generic prologues / epilogues -----------------------------------------------------------
struct LoadCpx {  // plain complex input, transforms `stride` apart
  const cpx* in;
  long long stride;
  const cpx* cur;
  __device__ __forceinline__ void prepare(int batch) { cur = in + (long long)batch * stride; }
  __device__ __forceinline__ cpx load(int n) const { return cur[n]; }
};
struct StoreCpx {
  cpx* out;
  long long stride;
  float scale;   // applied to both parts
  int conj;      // store the conjugate
  cpx* cur;
  __device__ __forceinline__ void begin(int batch) { cur = out + (long long)batch * stride; }
  __device__ __forceinline__ void put(int n, cpx v) const {
    cur[n] = make_float2(v.x * scale, conj ? -v.y * scale : v.y * scale);
  }
  __device__ __forceinline__ void finish(int, int) {}
};
struct SrcPlain {  // async variant: plain complex input (+ the pass' twiddle table as second operand)
  const cpx* in;
  long long stride;
  const cpx* tw;
  __device__ __forceinline__ void locate(int batch, const cpx*& src, const cpx*& aux) const {
    src = in + (long long)batch * stride;
    aux = tw;
  }
};

// packed (value, index) key: larger value wins, then the smaller index (numpy arg-max tie rule)
__device__ __forceinline__ unsigned long long peak_key(float v, unsigned idx) {
  return ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ float key_value(unsigned long long k) { return __uint_as_float((unsigned)(k >> 32)); }
__device__ __forceinline__ unsigned key_index(unsigned long long k) { return 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFu); }

__device__ __forceinline__ unsigned long long block_max_key(unsigned long long key) {
  __shared__ unsigned long long wk[FFT_THREADS / 32];
#pragma unroll
  for (int msk = 16; msk > 0; msk >>= 1) {
    unsigned long long o = __shfl_xor_sync(0xffffffffu, key, msk);
    key = o > key ? o : key;
  }
  if ((threadIdx.x & 31) == 0) wk[threadIdx.x >> 5] = key;
  __syncthreads();
  unsigned long long best = wk[0];
#pragma unroll
  for (int w = 1; w < FFT_THREADS / 32; ++w) best = wk[w] > best ? wk[w] : best;
  return best;
}

// ---- twiddle / table setup --------------------------------------------------------------------
static __global__ void twiddle_kernel(cpx* out, long long count, int Ls, long long M) {
  // Ls > 0: out[t*Ls + k] = exp(-2*pi*i * (t*k mod M) / M);   Ls == 0: out[q] = exp(-2*pi*i*q/M)
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < count;
       e += (long long)gridDim.x * blockDim.x) {
    long long q = e % M;
    if (Ls > 0) {
      long long t = e / Ls, k = e % Ls;
      q = (t * k) % M;
    }
    double s, c;
    sincospi(-2.0 * (double)q / (double)M, &s, &c);
    out[e] = make_float2((float)c, (float)s);
  }
}

// ---- host-side plan ---------------------------------------------------------------------------
struct Plan {
  int N = 0, npass = 0;
  bool big = false;   // uses radices beyond {2,4,8,16}
  Pass pass[4];
  DevBuf tw[4], wr[4];
  size_t smem[4];
  size_t smem_async[4];
  size_t smem_direct[4];
  bool async_ok[4];
};

inline bool factor_small(int n, int* radices, int& cnt) {
  cnt = 0;
  int twos = 0;
  while (n % 2 == 0) { n /= 2; ++twos; }
  while (twos >= 4) { radices[cnt++] = 16; twos -= 4; }
  if (twos == 3) radices[cnt++] = 8;
  if (twos == 2) radices[cnt++] = 4;
  if (twos == 1) radices[cnt++] = 2;
  const int odd[5] = {3, 5, 7, 11, 31};
  for (int i = 0; i < 5; ++i)
    while (n % odd[i] == 0) {
      if (cnt >= 16) return false;
      radices[cnt++] = odd[i];
      n /= odd[i];
    }
  return n == 1;
}

// Split the radices of N into `np` passes with the most even products (exhaustive, tiny).
inline bool plan_passes(int N, int maxR, int groups[4][MAX_SUB], int gcnt[4], int& np) {
  int rad[16], cnt;
  if (N >= 2 && (N & (N - 1)) == 0) {
    // power of two: passes of (nearly) equal size, each one or two sub-passes (16 x 2^k) -- the radix list of
    // factor_small (16, 16, ..., 8) cannot be split evenly (2^19 came out as 256 x 128 x 16, and a 16-point pass
    // keeps 16 of the 128 threads busy)
    int e = 0, lim = 0;
    while ((1 << e) < N) ++e;
    while ((2 << lim) <= maxR) ++lim;                 // largest exponent of a pass
    if (lim > 8) lim = 8;                             // two sub-passes of radix <= 16
    np = (e + lim - 1) / lim;
    if (np > 4) return false;
    for (int g = 0; g < np; ++g) {
      const int x = e / np + (g < e % np ? 1 : 0);    // exponent of this pass
      gcnt[g] = 0;
      if (x > 4) { groups[g][gcnt[g]++] = 16; groups[g][gcnt[g]++] = 1 << (x - 4); }
      else groups[g][gcnt[g]++] = 1 << x;
      if (x > 8) return false;
    }
    return true;
  }
  if (!factor_small(N, rad, cnt)) return false;
  for (np = 1; np <= 4; ++np) {
    long long best = -1;
    int bestAssign[16];
    int assign[16] = {0};
    long long total = 1;
    for (int i = 0; i < cnt; ++i) total *= np;
    for (long long code = 0; code < total; ++code) {
      long long c = code;
      long long prod[4] = {1, 1, 1, 1};
      int gc[4] = {0, 0, 0, 0};
      bool ok = true;
      for (int i = 0; i < cnt; ++i) {
        assign[i] = (int)(c % np);
        c /= np;
        prod[assign[i]] *= rad[i];
        if (++gc[assign[i]] > MAX_SUB) ok = false;
      }
      long long mx = 0;
      for (int g = 0; g < np; ++g) { if (prod[g] > mx) mx = prod[g]; if (prod[g] == 1) ok = false; }
      if (!ok || mx > maxR) continue;
      if (best < 0 || mx < best) { best = mx; memcpy(bestAssign, assign, sizeof(assign)); }
    }
    if (best > 0) {
      for (int g = 0; g < np; ++g) gcnt[g] = 0;
      for (int i = 0; i < cnt; ++i) groups[bestAssign[i]][gcnt[bestAssign[i]]++] = rad[i];
      if (np >= 3) {
        // Pass order is free in a Stockham transform.  The persistent pass kernels need the twiddle rows of a tile not to
        // wrap (Ls % 16 == 0, async_ok below): put a pass whose radix is a multiple of 16 first, then every later Ls is one
        // too (381 920 = 62 x 77 x 80 ran its middle pass through the synchronous kernel; as 80 x 62 x 77 it does not).
        int first = -1;
        long long fprod = 0;
        for (int g = 0; g < np; ++g) {
          long long prod = 1;
          for (int i = 0; i < gcnt[g]; ++i) prod *= groups[g][i];
          if (prod % 16 == 0 && prod > fprod) { first = g; fprod = prod; }
        }
        if (first > 0) {
          int tmp[MAX_SUB];
          const int tc = gcnt[first];
          memcpy(tmp, groups[first], sizeof(tmp));
          for (int g = first; g > 0; --g) { memcpy(groups[g], groups[g - 1], sizeof(tmp)); gcnt[g] = gcnt[g - 1]; }
          memcpy(groups[0], tmp, sizeof(tmp));
          gcnt[0] = tc;
        }
      }
      return true;
    }
  }
  return false;
}

inline int build_plan(Plan& pl, int N, bool inverse, cudaStream_t s, int maxR = 512) {
  int groups[4][MAX_SUB], gcnt[4], np;
  if (!plan_passes(N, maxR, groups, gcnt, np))
    return fail(SGX_ERR_ARG, "fft plan", "length has a prime factor outside {2,3,5,7,11,31} or is too large");
  pl.N = N;
  pl.npass = np;
  pl.big = false;
  int Ls = 1;
  for (int p = 0; p < np; ++p) {
    Pass& P = pl.pass[p];
    P.N = N;
    P.R = 1;
    P.nsub = gcnt[p];
    for (int i = 0; i < gcnt[p]; ++i) {
      // big radices first: their butterflies then run without sub-pass twiddles (ls == 1)
      P.radix[i] = groups[p][i];
    }
    for (int i = 0; i < P.nsub; ++i)
      for (int j = i + 1; j < P.nsub; ++j)
        if (P.radix[j] > P.radix[i]) { int t = P.radix[i]; P.radix[i] = P.radix[j]; P.radix[j] = t; }
    for (int i = 0; i < P.nsub; ++i) {
      P.R *= P.radix[i];
      int r = P.radix[i];
      if (!(r == 2 || r == 4 || r == 8 || r == 16)) pl.big = true;
    }
    P.Ls = Ls;
    P.m = N / P.R;
    P.T = TILE;
    P.ntiles = (P.m + TILE - 1) / TILE;
    P.inverse = inverse ? 1 : 0;
    pl.smem[p] = sizeof(cpx) * ((size_t)2 * P.R * TILE_P + P.R);
    pl.smem_async[p] = sizeof(cpx) * ((size_t)2 * P.R * TILE_P + (size_t)P.R * TILE + P.R);
    // compile-time-radix kernel: work buffer (+ cached second operand when the first radix is 31) + W
    pl.smem_direct[p] = sizeof(cpx) * ((size_t)P.R * TILE_P + (aux_direct(P.radix[0]) ? 0 : (size_t)P.R * TILE) + P.R);
    pl.async_ok[p] = (Ls == 1) || (Ls % TILE == 0) || (P.m <= Ls);   // twiddle rows of a tile must not wrap
    if (pl.wr[p].reserve(sizeof(cpx) * P.R)) return fail(SGX_ERR_CUDA, "cudaMalloc", "fft tables");
    SGX_COUNTED_LAUNCH(twiddle_kernel, dim3(4), dim3(128), 0, s, pl.wr[p].as<cpx>(), (long long)P.R, 0, (long long)P.R);
    P.wr = pl.wr[p].as<cpx>();
    P.twg = nullptr;
    if (Ls > 1) {
      long long M = (long long)Ls * P.R;
      if (pl.tw[p].reserve(sizeof(cpx) * (size_t)M)) return fail(SGX_ERR_CUDA, "cudaMalloc", "fft twiddles");
      int blocks = (int)((M + 255) / 256 < 2048 ? (M + 255) / 256 : 2048);
      SGX_COUNTED_LAUNCH(twiddle_kernel, dim3(blocks), dim3(256), 0, s, pl.tw[p].as<cpx>(), M, Ls, M);
      P.twg = pl.tw[p].as<cpx>();
    }
    Ls *= P.R;
  }
  return SGX_OK;
}

}  // namespace fft
}  // namespace sgx
