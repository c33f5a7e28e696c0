// Types shared by the acquisition translation units (sgx_acq.cu, sgx_pfa.cu).
#pragma once
#include "sgx_fft.cuh"

namespace sgx {

struct SearchDims {
  int nprn, nbins, blocks, prn_first;
};

struct PeakSel {  // per (rec, prn): result of A8/A9 first half
  int bin, blk, codePhase;
  float peak;
};

struct FineItem {   // a detection handed to the fine-frequency search
  int rec, prn, codePhase, pad;
};

// acquisition.py:147-159: is code phase i a candidate for the second peak?
__device__ __forceinline__ bool second_peak_candidate(int i, int c, int w, int n) {
  const int lo = c - w, hi = c + w;
  if (lo <= 0) return i >= hi && i <= n + lo;  // (index n itself, the reference's IndexError, cannot occur)
  if (hi >= n - 1) {
    const int a = hi - n, b = lo - 1;           // a >= -1; -1 is numpy's "last element"
    return (i >= a && i <= b) || (a < 0 && i == n + a);
  }
  return i <= lo || i >= hi;
}

}  // namespace sgx
