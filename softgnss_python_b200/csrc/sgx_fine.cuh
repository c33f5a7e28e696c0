// Fine-frequency search for nfft = 2^22 as a pruned, real-input 2048 x 2048 transform (see sgx_fine.cu).
#pragma once
#include "sgx_acq_types.cuh"

namespace sgx {
namespace fine {
using fft::cpx;

constexpr int NFFT = 1 << 22;   // 8 * 2^ceil(log2(10 * 38192)), acquisition.py:179

struct Args {
  const int8_t* sig;             // [rec][rec_stride]
  long long rec_stride;
  const long long* sums;         // [rec] integer sum of the whole longSignal (DC removal, acquisition.py:59)
  long long n_samples;
  const int8_t* chips;           // [32][1023]
  const unsigned short* idx;     // [nvalid] chip index of acquisition.py:172-174
  const FineItem* items;         // detections: recording, PRN, code phase
  int nvalid;
  int n_items;                   // detections of this launch (persistent step-1 kernel)
  int lo, hi;                    // candidate bins lo <= k < hi; the key index is k - lo (slice-relative, :186-187)
  const cpx* w2048;              // [2048] exp(-2 pi i q / 2048)
  const cpx* wlo;                // [2048] exp(-2 pi i q / 2^22)
  const cpx* w128;               // [128] exp(-2 pi i q / 128)
  float* stripped;               // [items][strip_stride] (x - mean) * code, zero-padded to whole rows of 2048
  long long strip_stride;
  cpx* y;                        // [items][1025][2048] step-1 output
  unsigned long long* partial;   // [items][row tiles]
};

// arg-max index (slice-relative) of every detection -> d_index[n_items] (device)
int run(Args a, int n_items, int* d_index, cudaStream_t s);

}  // namespace fine
}  // namespace sgx
