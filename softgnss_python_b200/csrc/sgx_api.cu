// Process-wide state of the C ABI (include/softgnss_b200.h): error text, launch counter, device probe.
#include "sgx_common.cuh"

namespace sgx {
char g_err[512] = "";
long long g_launches = 0;
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
  SGX_CUDA(cudaSetDevice(device));
  return SGX_OK;
}
