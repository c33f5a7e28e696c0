// Prime-factor (Good-Thomas) search transform: the hot loop of the acquisition (A7), one CTA per transform.
//
// The parallel code-phase search needs  |IFFT(S * C)|^2  for every (recording, block, bin, PRN)
// (acquisition.py:115-126) -- 59 392 inverse transforms of N = 38 192 points per 32-recording batch.  The Stockham
// engine of sgx_fft.cuh computes one in two kernels (R = 217, R = 176) with the 305 KB intermediate of every
// transform going out to HBM and back, inter-pass and sub-pass twiddle multiplies, and 205 executed
// thread-instructions per point (profiles/ncu_summary_r1_v5.md).  This kernel uses the arithmetic structure of N:
//
//   N = 31 * 7 * 16 * 11, all four factors pairwise coprime  ->  Good-Thomas: with the input index taken by residues,
//   k -> (k mod 31, k mod 7, k mod 16, k mod 11), and the output index by
//   tau = (tau1 N/31 + tau2 N/7 + tau3 N/16 + tau4 N/11) mod N,  k tau / N = sum_i k_i tau_i / P_i (mod 1):
//   the transform is a plain 4-D DFT.  No twiddle factors anywhere (no table loads, no complex multiplies between the
//   stages), and because both the spectra and the code spectra are produced by this library, they are simply stored in
//   the residue order (StorePerm epilogue of the forward transforms); the output order only enters the arg-max key.
//
//   One CTA computes a whole transform: pass A (DFT over k1, k2) writes its result to a CTA-private scratch block in
//   global memory, pass B (DFT over k3, k4) reads it back; |.|^2 and the maximum are taken from the registers of the
//   last butterfly, and one key per transform is written.  The 36 GB/batch of separate-kernel intermediates become a
//   scratch that is reused for every transform (444 resident CTAs x 315 KB).
//
//   Within a pass the work is split into warp-private slices (Shape: 4 columns x 217 rows in pass A, 8 columns x 176
//   rows in pass B): every butterfly stage of a slice reads only what the same warp wrote, so the only block-wide
//   barriers are the two per transform.  Spectra, code spectra and the scratch are stored in the order they are
//   accessed (Shape::slot, Shape::scratch_index): a pass-A slice is one contiguous block (one TMA bulk copy for the code
//   slice), a pass-B round reads consecutive 256-byte blocks.
#pragma once
#include "sgx_acq_types.cuh"
#include "sgx_pfa_tables.h"

namespace sgx {
namespace pfa {
using fft::cpx;


struct SearchArgs {
  const cpx* spec;               // [rec][blk][bin][N], residue order
  const cpx* codeF;              // [32][N], residue order, conj and 1/N folded in
  cpx* scratch;                  // [gridDim.x][NB][SROW]
  unsigned long long* partial;   // one (value, index) key per transform
  const PeakSel* sel;            // MASKED: winning (bin, block) per (rec, prn)
  PeakSel* sel_out;              // MASKED: code phase and peak value are written here (may alias sel)
  SearchDims d;
  long long nitems;
  int chip;                      // samples per chip (second-peak exclusion, acquisition.py:145)
  int p2;                        // second factor of the transform length: 7 (N = 38 192) or 3 (N = 16 368)
};

template <int P1, int P2, int P3, int P4>
struct Shape {
  static constexpr int p1 = P1, p2 = P2, p3 = P3, p4 = P4;
  static constexpr int NA = P1 * P2, NB = P3 * P4, N = NA * NB;
  static constexpr int SROW = (NA + 31) / 32 * 32;            // scratch row: NA intermediate values, padded to whole lines
  static constexpr int CA = 4, CB = 8;                        // columns of a warp slice in pass A / pass B
  static constexpr int NSA = NB / CA, NSB = (NA + CB - 1) / CB;
  static_assert(NB % CA == 0, "pass-A slices must be whole");
  static_assert(P2 <= 32 / CA && P3 % (32 / CB) == 0, "lane mapping of the first stages");
  static_assert(CA == 32 / CB && P3 % CA == 0, "a pass-A slice holds the k3 values of one pass-B round");
  // complex values per warp: X + Y in pass A, one NB x CB tile in pass B
  static constexpr int WARP_TILE = 2 * NA * CA > NB * CB ? 2 * NA * CA : NB * CB;
  static constexpr size_t smem_per_warp = sizeof(cpx) * (size_t)WARP_TILE;
  static constexpr size_t scratch_per_cta = sizeof(cpx) * (size_t)NB * SROW;
  // Storage order of the spectra and code spectra ("residue order", slice-major): element (kA, kB), kA = k1*P2 + k2,
  // kB = k3*P4 + k4, lives in pass-A slice (k3 / CA) * P4 + k4 at row kA, column k3 % CA.  A slice (NA x CA values,
  // 6 944 bytes) is one contiguous block: the search kernel fetches a code slice with one bulk copy (the [kA][NB]
  // order took 16 shared-memory wavefronts per cp.async request instead of 4: every 32-byte row piece arrived on its
  // own) and reads a spectrum row group (P2 x CA values) as 224 contiguous bytes instead of 7 separate sectors.  The
  // four columns of a slice are the k3 values one pass-B warp round reads together (k3 = kb + 4 r for its lanes
  // kb = 0..3), so that the scratch can be laid out for 256-byte pass-B reads (see scratch_index).
  __host__ __device__ static constexpr int slice_of(int kB) { return ((kB / P4) / CA) * P4 + kB % P4; }
  __host__ __device__ static constexpr int slot(int kA, int kB) { return (slice_of(kB) * NA + kA) * CA + (kB / P4) % CA; }
  // Scratch order of the intermediate (pass A -> pass B): [block of CB values of kA][slice][column][position].  A pass-B
  // round (lane = column * CB + position, P4 slices of one k3 group) reads P4 consecutive 256-byte blocks.  The default
  // search kernel (DIRECT) numbers the kA values as its radix-P2 stage produces them -- block tau2 * 4 + tau1 / 8,
  // position tau1 % 8, one padding value per 32 -- so that every store instruction of that stage writes one whole block;
  // the copy-loop variant and this helper use block kA / CB, position kA % CB.
  __host__ __device__ static constexpr int scratch_index(int kA, int sa, int col) {
    return (((kA / CB) * NSA + sa) * CA + col) * CB + kA % CB;
  }
  static int storage_index(int k) {   // where frequency bin k of a natural-order spectrum is stored
    return slot((k % P1) * P2 + k % P2, (k % P3) * P4 + k % P4);
  }
};

// Forward-transform epilogue: natural-order output n goes to its residue-order position.
struct StorePerm {
  cpx* out;
  long long stride;
  float scale;
  int conj;
  const int* perm;
  cpx* cur;
  __device__ __forceinline__ void begin(int batch) { cur = out + (long long)batch * stride; }
  __device__ __forceinline__ void put(int n, cpx v) const {
    cur[perm[n]] = make_float2(v.x * scale, conj ? -v.y * scale : v.y * scale);
  }
  __device__ __forceinline__ void finish(int, int) {}
};


// ---- radix-31 inverse butterfly in three rolled groups of five output pairs --------------------------------------
// The fully unrolled conjugate-pair butterfly is ~1100 instructions (17 KB); sixteen unsynchronised warps streaming
// through it (and the rest of a 62 KB kernel) miss the 32 KB L1.5 instruction cache all the time
// (profiles/ncu_summary_r2_v1.md: `no_instruction` 1.5 stalled warps per issue).  With the pairs taken in the order of
// the powers of the primitive root 3, cos(2 pi j_n k_m / 31) = C[(n + m) mod 15]: the 15 x 15 cosine matrix is a
// circulant (the sine matrix a skew-circulant), so output group q = 0, 1, 2 is the same straight-line code applied to
// the pair arrays rotated by 5 q places -- 400 instructions executed three times, plus 2 x 60 register moves.
struct R31 {
  cpx A[15], B[15];   // a'_n = x[j_n] + x[31 - j_n],  b'_n = sg_n (x[j_n] - x[31 - j_n])
  cpx x0;
  // v: spectrum values, y: code-spectrum values (row stride ys); returns the DC output
  __device__ __forceinline__ cpx prepare(const cpx* v, const cpx* y, int ys) {
    constexpr int J[15] = SGX_R31_J;
    constexpr int SG[15] = SGX_R31_SG;
    x0 = fft::cmulf(v[0], y[0]);
    cpx dc = x0;
#pragma unroll
    for (int n = 0; n < 15; ++n) {
      const cpx p = fft::cmulf(v[J[n]], y[J[n] * ys]), q = fft::cmulf(v[31 - J[n]], y[(31 - J[n]) * ys]);
      A[n] = fft::cadd(p, q);
      B[n] = SG[n] > 0 ? fft::csub(p, q) : fft::csub(q, p);
      dc = fft::cadd(dc, A[n]);
    }
    return dc;
  }
  // output pair r of the current group: hi -> row KHI31[5 q + r], lo -> row 31 - KHI31[5 q + r]
  template <int r>
  __device__ __forceinline__ void pair(cpx& hi, cpx& lo) const {
    float cr = x0.x, ci = x0.y, sr = 0.f, si = 0.f;
#pragma unroll
    for (int n = 0; n < 15; ++n) {
      const float c = C31R[(n + r) % 15], s = S31R[n + r];
      cr = fmaf(A[n].x, c, cr);
      ci = fmaf(A[n].y, c, ci);
      sr = fmaf(B[n].x, s, sr);
      si = fmaf(B[n].y, s, si);
    }
    hi = make_float2(cr - si, ci + sr);
    lo = make_float2(cr + si, ci - sr);
  }
  // output pair m = 0..14 from the unrotated arrays (fully unrolled variant: no register moves, 3x the code)
  template <int m>
  __device__ __forceinline__ void pair_at(cpx& hi, cpx& lo) const {
    float cr = x0.x, ci = x0.y, sr = 0.f, si = 0.f;
#pragma unroll
    for (int n = 0; n < 15; ++n) {
      const float c = C31R[(n + m) % 15], s = S31R[n + m];
      cr = fmaf(A[n].x, c, cr);
      ci = fmaf(A[n].y, c, ci);
      sr = fmaf(B[n].x, s, sr);
      si = fmaf(B[n].y, s, si);
    }
    hi = make_float2(cr - si, ci + sr);
    lo = make_float2(cr + si, ci - sr);
  }
  __device__ __forceinline__ void rotate() {   // A_n <- A_(n-5 mod 15);  B_n <- B_(n-5), antiperiodic
    cpx tA[15], tB[15];
#pragma unroll
    for (int n = 0; n < 15; ++n) {
      tA[n] = A[(n + 10) % 15];
      tB[n] = n < 5 ? make_float2(-B[n + 10].x, -B[n + 10].y) : B[n - 5];
    }
#pragma unroll
    for (int n = 0; n < 15; ++n) { A[n] = tA[n]; B[n] = tB[n]; }
  }
};


// ---- forward transform into the residue order ---------------------------------------------------------------------
// X[k] = sum_n x[n] exp(-2 pi i n k / N) with the input taken in the Ruritanian order n = (n1 N/31 + n2 N/7 + n3 N/16 +
// n4 N/11) mod N: the 4-D output coordinates are then the residues (k mod 31, k mod 7, k mod 16, k mod 11), i.e. exactly
// the storage order the search kernel reads -- no permutation pass, no scattered stores.  Same two passes and slices as
// the search kernel; one CTA per transform, `Pro::load(n)` supplies sample n (int8 x carrier NCO, or a code table row).
struct ForwardArgs {
  cpx* scratch;      // [gridDim.x][NB][SROW]
  cpx* out;          // [batch][N], residue order
  long long out_stride;
  float scale;
  int conj;
  int nbatch;
};

// Shared memory of the forward kernel: the staged input block (Pro::STAGED) and one tile per warp.
template <class Pro, class S>
struct ForwardSmem {
  static constexpr int TILE = S::NA * S::CA > S::NB * S::CB ? S::NA * S::CA : S::NB * S::CB;   // complex values per warp
  static constexpr int STAGE = Pro::STAGED ? (S::N + 127) / 128 * 128 : 0;                      // bytes
  static constexpr size_t bytes(int warps) { return (size_t)STAGE + sizeof(cpx) * (size_t)TILE * warps; }
};

// Pro::STAGED: the prologue reads one int8 block of N samples per transform (ProNco).  The prime-factor input order
// makes every lane of a load touch its own 32-byte sector, and every pass-A slice touches nearly all sectors of the
// block: read from global memory that is 44 L2 reads of the block per transform.  The block is copied to shared memory
// once per transform instead (the copy for the next transform is issued between pass A and pass B) and gathered there.
template <class Pro, int P1, int P2, int P3, int P4, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) pfa_forward_kernel(ForwardArgs a, Pro pro0) {
  typedef Shape<P1, P2, P3, P4> S;
  typedef ForwardSmem<Pro, S> M;
  static_assert(P1 == 31, "stage 1 is the radix-31 butterfly");
  constexpr int NA = S::NA, NB = S::NB, N = S::N, SROW = S::SROW;
  constexpr int Q1 = N / P1, Q2 = N / P2, Q3 = N / P3, Q4 = N / P4;
  constexpr int CA = S::CA, CB = S::CB, KA = 32 / CA, KB = 32 / CB;
  SGX_DYN_SMEM(smem);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  int8_t* stage = reinterpret_cast<int8_t*>(smem);
  cpx* X = reinterpret_cast<cpx*>(smem + M::STAGE) + (size_t)w * M::TILE;
  cpx* scr = a.scratch + (size_t)blockIdx.x * NB * SROW;
  const int ja = lane & (CA - 1), ka = lane / CA;
  const int jb = lane & (CB - 1), kb = lane / CB;
  auto stage_in = [&](int item) {
    if constexpr (Pro::STAGED) {
      static_assert(N % 16 == 0, "16-byte copies");
      Pro q = pro0;
      q.prepare(item);
      const int8_t* g = q.block();
      if ((reinterpret_cast<unsigned long long>(g) & 15ull) == 0) {
        for (int i = tid; i < N / 16; i += WARPS * 32) reinterpret_cast<uint4*>(stage)[i] = __ldg(reinterpret_cast<const uint4*>(g) + i);
      } else {
        for (int i = tid; i < N; i += WARPS * 32) stage[i] = g[i];
      }
    }
  };
  if (Pro::STAGED && blockIdx.x < a.nbatch) {
    stage_in(blockIdx.x);
    __syncthreads();
  }
#pragma unroll 1
  for (int item = blockIdx.x; item < a.nbatch; item += gridDim.x) {
    Pro pro = pro0;
    pro.prepare(item);
    if constexpr (Pro::STAGED) pro.use(stage);
    // pass A: DFT over (n1, n2); slice = CA values of nB = n3*P4 + n4
#pragma unroll 1
    for (int sa = w; sa < S::NSA; sa += WARPS) {
      if (ka < P2) {
        const int nB = sa * CA + ja;
        int n = (ka * Q2 + (nB / P4) * Q3 + (nB % P4) * Q4) % N;
        cpx v[P1];
#pragma unroll
        for (int n1 = 0; n1 < P1; ++n1) {
          v[n1] = pro.load(n);
          n += Q1;
          if (n >= N) n -= N;
        }
        fft::Dft<P1, false>::run(v);
        cpx* x = X + ka * CA + ja;
#pragma unroll
        for (int k1 = 0; k1 < P1; ++k1) x[k1 * (P2 * CA)] = v[k1];
      }
      __syncwarp();
#pragma unroll 1
      for (int k1 = ka; k1 < P1; k1 += KA) {
        cpx u[P2];
        cpx* rp = X + (k1 * P2) * CA + ja;
#pragma unroll
        for (int n2 = 0; n2 < P2; ++n2) u[n2] = rp[n2 * CA];
        fft::Dft<P2, false>::run(u);
#pragma unroll
        for (int k2 = 0; k2 < P2; ++k2) rp[k2 * CA] = u[k2];
      }
      __syncwarp();
      {
        cpx* dst = scr + (size_t)(sa * CA + (lane & 1) * 2) * SROW;
#pragma unroll 1
        for (int c = lane >> 1; c < NA; c += 16) {
          const float4 x = *reinterpret_cast<const float4*>(X + c * CA + (lane & 1) * 2);
          __stcg(dst + c, make_float2(x.x, x.y));
          __stcg(dst + SROW + c, make_float2(x.z, x.w));
        }
      }
      __syncwarp();
    }
    __syncthreads();
    if (Pro::STAGED && item + (int)gridDim.x < a.nbatch) stage_in(item + gridDim.x);   // every warp is done with the block
    // pass B: DFT over (n4, n3); slice = CB values of kA = k1*P2 + k2; output X[kA][k3*P4 + k4]
    cpx* out = a.out + (long long)item * a.out_stride;
#pragma unroll 1
    for (int sb = w; sb < S::NSB; sb += WARPS) {
      const int kA = sb * CB + jb;
      const bool valid = kA < NA;
#pragma unroll 1
      for (int n3 = kb; n3 < P3; n3 += KB) {
        cpx u[P4];
        const cpx* p = scr + (size_t)(n3 * P4) * SROW + (valid ? kA : 0);
#pragma unroll
        for (int n4 = 0; n4 < P4; ++n4) u[n4] = __ldcg(p + (size_t)n4 * SROW);
        fft::Dft<P4, false>::run(u);
        cpx* rp = X + (n3 * P4) * CB + jb;
#pragma unroll
        for (int k4 = 0; k4 < P4; ++k4) rp[k4 * CB] = u[k4];
      }
      __syncwarp();
#pragma unroll 1
      for (int k4 = kb; k4 < P4; k4 += KB) {
        cpx u[P3];
        const cpx* rp = X + k4 * CB + jb;
#pragma unroll
        for (int n3 = 0; n3 < P3; ++n3) u[n3] = rp[(n3 * P4) * CB];
        fft::Dft<P3, false>::run(u);
        if (valid) {
#pragma unroll
          for (int k3 = 0; k3 < P3; ++k3)
            out[S::slot(kA, k3 * P4 + k4)] = make_float2(u[k3].x * a.scale, a.conj ? -u[k3].y * a.scale : u[k3].y * a.scale);
        }
      }
      __syncwarp();
    }
    __syncthreads();   // the scratch is rewritten by the next transform
  }
}

template <class Pro, int P2, int WARPS, int MINB>
inline int launch_forward_cfg(Pro pro, int batch, cpx* out, float scale, int conj, DevBuf& scratch, cudaStream_t s) {
  typedef Shape<31, P2, 16, 11> S;
  auto kfn = pfa_forward_kernel<Pro, 31, P2, 16, 11, WARPS, MINB>;
  const size_t smem = ForwardSmem<Pro, S>::bytes(WARPS);
  SGX_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, n_sm = 0, occ = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  if (n_sm <= 0) n_sm = 148;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, WARPS * 32, smem) != cudaSuccess || occ < 1) occ = 1;
  long long grid = (long long)n_sm * occ;
  if (grid > batch) grid = batch;
  if (scratch.reserve(S::scratch_per_cta * (size_t)grid)) return fail(SGX_ERR_CUDA, "cudaMalloc", "forward scratch");
  ForwardArgs a;
  a.scratch = scratch.as<cpx>(); a.out = out; a.out_stride = S::N; a.scale = scale; a.conj = conj; a.nbatch = batch;
  SGX_COUNTED_LAUNCH(kfn, dim3((unsigned)grid), dim3(WARPS * 32), smem, s, a, pro);
  SGX_CUDA(cudaGetLastError());
  return SGX_OK;
}

// `batch` forward transforms of N = 31 * P2 * 16 * 11 points into out[batch][N] (residue order).  Grows `scratch`.
template <class Pro, int P2 = 7>
inline int launch_forward(Pro pro, int batch, cpx* out, float scale, int conj, DevBuf& scratch, cudaStream_t s) {
  if (batch <= 0) return SGX_OK;
  if constexpr (Pro::STAGED) {
    // 38 KB of staged samples + 11 KB per warp: two CTAs per SM
    static const int cfg = getenv("SGX_PFA_FWD_CFG") ? atoi(getenv("SGX_PFA_FWD_CFG")) : 62;
    if constexpr (P2 == 7) {
      if (cfg == 42) return launch_forward_cfg<Pro, P2, 4, 2>(pro, batch, out, scale, conj, scratch, s);
      if (cfg == 81) return launch_forward_cfg<Pro, P2, 8, 1>(pro, batch, out, scale, conj, scratch, s);
    }
    return launch_forward_cfg<Pro, P2, 6, 2>(pro, batch, out, scale, conj, scratch, s);
  } else {
    return launch_forward_cfg<Pro, P2, 4, 4>(pro, batch, out, scale, conj, scratch, s);
  }
}

typedef Shape<31, 7, 16, 11> SearchShape;   // 38 192 = 31 * 7 * 16 * 11 (fs = 38.192 MHz, 1 ms)
typedef Shape<31, 3, 16, 11> SearchShape3;  // 16 368 = 31 * 3 * 16 * 11 (fs = 16.3676 MHz rounded, 1 ms): the radix-31 stage
                                            // runs on 12 of 32 lanes there, still ahead of the generic passes
// second factor of the prime-factor shape for transform length n, 0 if there is none
inline int pfa_shape_p2(long long n) { return n == SearchShape::N ? 7 : n == SearchShape3::N ? 3 : 0; }

// All search transforms of a call (masked: the one second-peak transform per (rec, prn)) in one launch of resident
// CTAs.  `scratch` is grown to gridDim.x x Shape::scratch_per_cta.  Defined in sgx_pfa.cu.
int launch_search(SearchArgs args, bool masked, DevBuf& scratch, cudaStream_t s);

}  // namespace pfa
}  // namespace sgx
