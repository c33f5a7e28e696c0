// Process-wide state of the C ABI (include/softgnss_b200.h): error text, launch counter, device probe.
#include "sgx_common.cuh"

namespace sgx {
char g_err[512] = "";
long long g_launches = 0;
int g_bound_device = -1;
std::recursive_mutex g_api_mutex;

// FP32 FMA burn: 16 independent chains per thread (the acquisition roofline divides by what this measures)
__global__ void __launch_bounds__(512) fma_burn_kernel(int iters, float* out, float x, float y) {
  float f[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = x * (float)(i + 1) + (float)threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = fmaf(f[i], x, y);
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += f[i];
  if (r == 0.12345f) out[0] = r;
}
}  // namespace sgx

extern "C" int sgx_abi_version(void) { return SGX_ABI_VERSION; }
extern "C" const char* sgx_last_error(void) { return sgx::g_err; }
extern "C" int64_t sgx_kernel_launch_count(void) { return sgx::g_launches; }

extern "C" int sgx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int sgx_set_device(int device) {
  if (sgx_device_count() <= 0) return sgx::fail(SGX_ERR_NODEV, "sgx_set_device", "no CUDA device");
  std::lock_guard<std::recursive_mutex> lock(sgx::g_api_mutex);
  if (sgx::g_bound_device >= 0 && device != sgx::g_bound_device)
    return sgx::fail(SGX_ERR_ARG, "sgx_set_device",
                     "plans and scratch buffers of this process already live on another device (one process per GPU)");
  SGX_CUDA(cudaSetDevice(device));
  sgx::g_bound_device = device;
  return SGX_OK;
}

extern "C" int sgx_fp32_peak(double* tflops, void* cuda_stream) {
  if (sgx_device_count() <= 0) return sgx::fail(SGX_ERR_NODEV, "sgx_fp32_peak", "no CUDA device");
  if (!tflops) return sgx::fail(SGX_ERR_ARG, "sgx_fp32_peak", "null pointer");
  SGX_API_GUARD();
  cudaStream_t s = (cudaStream_t)cuda_stream;
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  if (n_sm <= 0) n_sm = 148;
  static sgx::DevBuf out;
  if (out.reserve(64)) return sgx::fail(SGX_ERR_CUDA, "cudaMalloc", "fp32 peak");
  const int iters = 20000, blocks = n_sm * 4, threads = 512;
  cudaEvent_t e0, e1;
  SGX_CUDA(cudaEventCreate(&e0));
  SGX_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    SGX_CUDA(cudaEventRecord(e0, s));
    SGX_COUNTED_LAUNCH(sgx::fma_burn_kernel, dim3(blocks), dim3(threads), 0, s, rep == 0 ? 100 : iters, out.as<float>(),
                       1.0001f, 0.5f);
    SGX_CUDA(cudaEventRecord(e1, s));
    SGX_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    SGX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = (double)blocks * threads * iters * 16.0 * 2.0 / ((double)best * 1e-3) / 1e12;
  return SGX_OK;
}
