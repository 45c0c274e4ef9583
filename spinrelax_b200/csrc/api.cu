// Library-level entry points of the C ABI: versioning, error text, device probe.
#include <cstdarg>

#include "common.cuh"

static thread_local char g_err[512] = "";

void sr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int sr_abi_version(void) { return SR_ABI_VERSION; }

extern "C" const char* sr_last_error(void) { return g_err; }

extern "C" int sr_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* global_mem_bytes) {
  int dev = 0;
  SR_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  SR_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (global_mem_bytes) *global_mem_bytes = p.totalGlobalMem;
  return SR_OK;
}
