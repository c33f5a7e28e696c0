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
//   One CTA computes a whole transform: pass A (DFT over k1, k2; 11 tiles of 16 columns) writes its result to a
//   CTA-private scratch row block in global memory, pass B (DFT over k3, k4; 14 tiles) reads it back.  The scratch of
//   all resident CTAs (296 x 315 KB = 93 MB) is reused for every transform, so it lives in the 126 MB L2 and the
//   36 GB/batch HBM round trip of the two-kernel version disappears; |.|^2 and the arg-max are taken from the
//   registers of the last butterfly, and one key per transform is written.
//
//   A CTA is NG independent groups of 128 threads (named barriers), each working on its own tiles of the current
//   transform; two __syncthreads per transform (pass A -> pass B, and the key reduction).
#pragma once
#include "sgx_acq_types.cuh"

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
  static constexpr int WARP_TILE = 2 * NA * CA;   // complex values per warp: X + Y in pass A, one NB x CB tile in pass B
  static_assert(2 * NA * CA >= NB * CB, "pass-B tile must fit");
  static constexpr size_t smem_per_warp = sizeof(cpx) * (size_t)WARP_TILE;
  static constexpr size_t scratch_per_cta = sizeof(cpx) * (size_t)NB * SROW;
  static int storage_index(int k) {   // where frequency bin k of a natural-order spectrum is stored
    return ((k % P1) * P2 + k % P2) * NB + (k % P3) * P4 + k % P4;
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

typedef Shape<31, 7, 16, 11> SearchShape;   // 38 192 = 31 * 7 * 16 * 11 (fs = 38.192 MHz, 1 ms)

// All search transforms of a call (masked: the one second-peak transform per (rec, prn)) in one launch of resident
// CTAs.  `scratch` is grown to gridDim.x x Shape::scratch_per_cta.  Defined in sgx_pfa.cu.
int launch_search(SearchArgs args, bool masked, DevBuf& scratch, cudaStream_t s);

}  // namespace pfa
}  // namespace sgx
