// Shared device/host helpers for libspinrelax_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/spinrelax_b200.h"

// ---- error plumbing (C ABI returns int status, message kept per thread) ------------------------
void sr_set_error(const char* fmt, ...);

#define SR_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      sr_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return SR_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define SR_REQUIRE(cond, ...)                                                                 \
  do {                                                                                        \
    if (!(cond)) { sr_set_error(__VA_ARGS__); return SR_ERR_ARG; }                            \
  } while (0)

static inline long long sr_round_up(long long x, long long m) { return (x + m - 1) / m * m; }

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk -> SASS UBLKCP) ------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t sr_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void sr_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sr_smem_u32(bar)), "r"(count) : "memory");
}
// make barrier init visible to the async (TMA) proxy
__device__ __forceinline__ void sr_fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void sr_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sr_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sr_tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          sr_smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(sr_smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void sr_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t addr = sr_smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void sr_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sr_smem_u32(bar)) : "memory");
}
// generic-proxy reads of a stage are done; order them before the next async-proxy write into it
__device__ __forceinline__ void sr_fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ double sr_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif
