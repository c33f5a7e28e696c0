// K0: synthetic int8 IF recordings generated on the device, bit-identical to synth.generate_cpu.
// Integer-only signal model (softgnss_python_b200/synth.py): 64-bit carrier and code NCOs,
// a 4096-entry cosine table in shared memory, counter-hash Irwin-Hall noise.  Needed because the
// batched tracking configuration (BASELINE.json config 4) is hundreds of GB of samples.
#include "sgx_common.cuh"

namespace sgx {

constexpr int SYN_THREADS = 256;
constexpr int SYN_PER_THREAD = 16;  // one 128-bit store per thread per tile

struct SynthArgs {
  int8_t* out;
  long long rec_stride, n_samples, start;
  const sgx_synth_spec* specs;  // device [R]
  const int8_t* bits;           // device [R][MAX_SATS][n_bits]
  const short* lut;             // device [4096]
  const int8_t* chips;          // device [32][1023]
  int tiles_per_rec;
};

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(SYN_THREADS) synth_kernel(SynthArgs a) {
  __shared__ short lutS[4096];
  __shared__ int8_t chipS[SGX_SYNTH_MAX_SATS][1024];
  __shared__ sgx_synth_spec sp;
  const int tid = threadIdx.x;
  const int r = blockIdx.y;
  if (tid == 0) sp = a.specs[r];
  for (int i = tid; i < 4096; i += SYN_THREADS) lutS[i] = a.lut[i];
  __syncthreads();
  const int ns = sp.n_sats;
  for (int i = tid; i < ns * 1023; i += SYN_THREADS) {
    int s = i / 1023, c = i - s * 1023;
    chipS[s][c] = a.chips[(sp.prn[s] - 1) * 1023 + c];
  }
  __syncthreads();
  const int8_t* bits = a.bits + (long long)r * SGX_SYNTH_MAX_SATS * sp.n_bits;
  int8_t* out = a.out + (long long)r * a.rec_stride;
  for (long long tile = blockIdx.x; tile < a.tiles_per_rec; tile += gridDim.x) {
    long long j0 = (tile * SYN_THREADS + tid) * SYN_PER_THREAD;  // index within this call
    if (j0 >= a.n_samples) continue;
    int acc[SYN_PER_THREAD];
#pragma unroll
    for (int j = 0; j < SYN_PER_THREAD; ++j) {
      unsigned long long n = (unsigned long long)(a.start + j0 + j);
      unsigned long long h = splitmix64(sp.seed + n * 0x9E3779B97F4A7C15ULL);
      unsigned lo = (unsigned)h, hi = (unsigned)(h >> 32);
      // sum of the 8 bytes
      unsigned sum = (lo & 0x00FF00FFu) + ((lo >> 8) & 0x00FF00FFu) + (hi & 0x00FF00FFu) + ((hi >> 8) & 0x00FF00FFu);
      int bsum = (int)((sum & 0xFFFFu) + (sum >> 16));
      acc[j] = sp.noise_k * (bsum - 1020);
    }
    for (int s = 0; s < ns; ++s) {
      const unsigned long long dphi = sp.dphi[s], dcp = sp.dcp[s];
      unsigned long long n0 = (unsigned long long)(a.start + j0);
      unsigned long long ph = sp.phi0[s] + n0 * dphi;
      unsigned long long cp = sp.cp0[s] + n0 * dcp;
      const int amp = sp.amp[s], per0 = sp.per0[s], nb = sp.n_bits;
      const int8_t* sb = bits + (long long)s * nb;
#pragma unroll
      for (int j = 0; j < SYN_PER_THREAD; ++j) {
        unsigned chips = (unsigned)(cp >> 32);
        unsigned period = chips / 1023u;
        unsigned chip = chips - period * 1023u;
        int bit = (int)(((period + (unsigned)per0) / 20u) % (unsigned)nb);
        int sign = (int)chipS[s][chip] * (int)sb[bit];
        acc[j] += amp * sign * (int)lutS[ph >> 52];
        ph += dphi;
        cp += dcp;
      }
    }
    unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < SYN_PER_THREAD; ++j) {
      int q = (acc[j] + (1 << 21)) >> 22;
      q = max(-128, min(127, q));
      w[j >> 2] |= ((unsigned)q & 0xFFu) << (8 * (j & 3));
    }
    if (j0 + SYN_PER_THREAD <= a.n_samples && ((a.rec_stride | (long long)(uintptr_t)a.out) & 15) == 0) {
      *reinterpret_cast<uint4*>(out + j0) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      for (int j = 0; j < SYN_PER_THREAD && j0 + j < a.n_samples; ++j)
        out[j0 + j] = (int8_t)((w[j >> 2] >> (8 * (j & 3))) & 0xFF);
    }
  }
}

struct SynthScratch {
  DevBuf specs, bits, lut, chips, out;
};
static SynthScratch g_syn;

}  // namespace sgx

using namespace sgx;

extern "C" int sgx_synth_generate(int8_t* out, int64_t rec_stride, int64_t n_samples, int64_t start,
                                  int32_t n_recordings, const sgx_synth_spec* specs, const int8_t* bits,
                                  const int16_t* lut, const int8_t* ca_chips, void* cuda_stream) {
  if (sgx_device_count() <= 0) return fail(SGX_ERR_NODEV, "sgx_synth_generate", "no CUDA device");
  SGX_API_GUARD();
  if (!out || !specs || !bits || !lut || !ca_chips || n_recordings <= 0 || n_samples <= 0 ||
      rec_stride < n_samples)
    return fail(SGX_ERR_ARG, "sgx_synth_generate", "bad argument");
  const int nb = specs[0].n_bits;
  for (int r = 0; r < n_recordings; ++r) {
    if (specs[r].n_bits != nb || specs[r].n_sats < 1 || specs[r].n_sats > SGX_SYNTH_MAX_SATS)
      return fail(SGX_ERR_ARG, "sgx_synth_generate", "n_bits must be uniform, 1..12 satellites");
    for (int s = 0; s < specs[r].n_sats; ++s)
      if (specs[r].prn[s] < 1 || specs[r].prn[s] > 32) return fail(SGX_ERR_ARG, "sgx_synth_generate", "PRN outside 1..32");
  }
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const size_t bits_bytes = (size_t)n_recordings * SGX_SYNTH_MAX_SATS * nb;
  if (g_syn.specs.reserve(sizeof(sgx_synth_spec) * n_recordings) || g_syn.bits.reserve(bits_bytes) ||
      g_syn.lut.reserve(4096 * 2) || g_syn.chips.reserve(32 * 1023))
    return fail(SGX_ERR_CUDA, "cudaMalloc", "synth scratch");
  SGX_CUDA(cudaMemcpyAsync(g_syn.specs.p, specs, sizeof(sgx_synth_spec) * n_recordings, cudaMemcpyHostToDevice, s));
  SGX_CUDA(cudaMemcpyAsync(g_syn.bits.p, bits, bits_bytes, cudaMemcpyHostToDevice, s));
  SGX_CUDA(cudaMemcpyAsync(g_syn.lut.p, lut, 4096 * 2, cudaMemcpyHostToDevice, s));
  SGX_CUDA(cudaMemcpyAsync(g_syn.chips.p, ca_chips, 32 * 1023, cudaMemcpyHostToDevice, s));
  const bool on_host = !is_device_ptr(out);
  int8_t* d_out = out;
  long long stride = rec_stride;
  if (on_host) {
    stride = (n_samples + 15) & ~15LL;
    if (g_syn.out.reserve((size_t)stride * n_recordings)) return fail(SGX_ERR_CUDA, "cudaMalloc", "synth output");
    d_out = g_syn.out.as<int8_t>();
  }
  SynthArgs a;
  a.out = d_out;
  a.rec_stride = stride;
  a.n_samples = n_samples;
  a.start = start;
  a.specs = g_syn.specs.as<sgx_synth_spec>();
  a.bits = g_syn.bits.as<int8_t>();
  a.lut = g_syn.lut.as<short>();
  a.chips = g_syn.chips.as<int8_t>();
  const long long per_tile = (long long)SYN_THREADS * SYN_PER_THREAD;
  a.tiles_per_rec = (int)((n_samples + per_tile - 1) / per_tile);
  int gx = a.tiles_per_rec < 148 * 8 ? a.tiles_per_rec : 148 * 8;
  SGX_COUNTED_LAUNCH(synth_kernel, dim3(gx, n_recordings), dim3(SYN_THREADS), 0, s, a);
  SGX_CUDA(cudaGetLastError());
  if (on_host)
    for (int r = 0; r < n_recordings; ++r)
      SGX_CUDA(cudaMemcpyAsync(out + (size_t)r * rec_stride, d_out + (size_t)r * stride, (size_t)n_samples,
                               cudaMemcpyDeviceToHost, s));
  SGX_CUDA(cudaStreamSynchronize(s));
  return SGX_OK;
}
