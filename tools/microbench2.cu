// Issue-slot microbenchmarks for sm_100a: does FFMA2 (fma.rn.f32x2) leave issue slots free for ALU / LDS work?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench2 tools/microbench2.cu
// Every probe prints one JSON line: lane-FMAs per clock per SM (peak 128) at the effective SM clock.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

static double g_clk_mhz = 1965.0;
template <typename F>
static float time_ms(F launch, int reps = 5) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  return (uint64_t)__float_as_uint(lo) | ((uint64_t)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ unsigned alu1(unsigned a, unsigned b) {
  unsigned d;
  asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

// PACKED: 16 FFMA2 per iteration (32 lane-FMAs per thread), else 32 scalar FFMA.  NALU independent XORs and
// NLDS conflict-free LDS.64 are interleaved evenly.
template <bool PACKED, int NALU, int NLDS, int LDK = 0>
__global__ void __launch_bounds__(256) k_mix(float* out, int iters, float x0, float y0) {
  __shared__ __align__(16) float2 sm[256 * 8];
  for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = make_float2(i * 1e-3f, 1.f);
  __syncthreads();
  uint64_t acc2[16];
  float acc1[32];
  unsigned ia[16];
  unsigned lsum[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) { acc2[i] = pack2(threadIdx.x * 1e-3f + i, 1.f + i); ia[i] = threadIdx.x + i; }
#pragma unroll
  for (int i = 0; i < 32; ++i) acc1[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) lsum[i] = 0u;
  const uint64_t X = pack2(x0, x0 + 1e-6f), Y = pack2(y0, y0 - 1e-6f);
  const float2* p = sm + threadIdx.x;
#pragma unroll 4
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (PACKED) acc2[i] = ffma2(acc2[i], X, Y);
      else { acc1[2 * i] = ffma1(acc1[2 * i], x0, y0); acc1[2 * i + 1] = ffma1(acc1[2 * i + 1], x0, y0); }
      if (NALU > 0 && (i % (16 / (NALU > 16 ? 16 : NALU))) == 0) {
        ia[i] = alu1(ia[i], ia[(i + 1) & 15]);
        if (NALU > 16) ia[(i + 8) & 15] = alu1(ia[(i + 8) & 15], ia[(i + 3) & 15]);
      }
      if (NLDS > 0 && (i % (16 / NLDS)) == 0) {
        uint4 v = make_uint4(0, 0, 0, 0);
        const int sel = ((it + i) & 3);
        if (LDK == 0) asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(__cvta_generic_to_shared(p + sel * 256)) : "memory");
        if (LDK == 1) asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v.x) : "l"(__cvta_generic_to_shared((const float*)sm + threadIdx.x + sel * 256)) : "memory");
        if (LDK == 2) asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v.x) : "l"(__cvta_generic_to_shared((const float*)sm + (threadIdx.x >> 5) + sel * 256)) : "memory");
        if (LDK == 3) asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(__cvta_generic_to_shared((const float4*)sm + threadIdx.x + sel * 256)) : "memory");
        if (LDK == 4) asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(__cvta_generic_to_shared((const float4*)sm + (threadIdx.x >> 5) + sel * 256)) : "memory");
        lsum[(i / (16 / NLDS)) & 7] ^= v.x + v.y + v.z + v.w;
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += __uint_as_float((uint32_t)acc2[i]) + __uint_as_float((uint32_t)(acc2[i] >> 32)) + (float)ia[i];
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc1[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += (float)lsum[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// three distinct 64-bit sources per FFMA2 (no operand reuse): register-bandwidth probe
template <bool PACKED>
__global__ void __launch_bounds__(256) k_distinct(float* out, int iters, float x0, float y0) {
  uint64_t acc2[16], x2[4], y2[4];
  float acc1[32], x1[4], y1[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc2[i] = pack2(threadIdx.x * 1e-3f + i, 1.f + i);
#pragma unroll
  for (int i = 0; i < 32; ++i) acc1[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) { x2[i] = pack2(x0 + i * 1e-3f, x0 - i * 1e-3f); y2[i] = pack2(y0 + i * 1e-3f, y0 + threadIdx.x * 1e-6f); x1[i] = x0 + i * 1e-3f; }
#pragma unroll
  for (int i = 0; i < 8; ++i) y1[i] = y0 + i * 1e-3f + threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (PACKED) acc2[i] = ffma2(x2[i & 3], y2[(i >> 2) & 3], acc2[i]);
      else { acc1[2 * i] = ffma1(x1[i & 3], y1[(i >> 1) & 7], acc1[2 * i]); acc1[2 * i + 1] = ffma1(x1[(i + 1) & 3], y1[(i >> 1) & 7], acc1[2 * i + 1]); }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += __uint_as_float((uint32_t)acc2[i]) + __uint_as_float((uint32_t)(acc2[i] >> 32));
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_clock(unsigned long long* out, int iters) {
  unsigned long long t0, c0 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  float a = threadIdx.x * 1e-3f, b = 1.0001f;
  for (int i = 0; i < iters; ++i) a = fmaf(a, b, 1e-6f);
  unsigned long long t1, c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = c1 - c0; out[1] = t1 - t0; out[2] = (unsigned long long)a; }
}
static void measure_clock() {
  unsigned long long* d; CK(cudaMalloc(&d, 32));
  k_clock<<<148 * 4, 256>>>(d, 400000);
  unsigned long long h[3]; CK(cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost));
  g_clk_mhz = (double)h[0] / (double)h[1] * 1e3;
  printf("{\"probe\":\"clock\",\"sm_mhz_effective\":%.1f}\n", g_clk_mhz);
  CK(cudaFree(d));
}

template <bool PACKED, int NALU, int NLDS, int LDK = 0>
static void run_mix(float* d, int sms) {
  const int grid = sms * 8, block = 256, iters = 20000;
  float ms = time_ms([&] { k_mix<PACKED, NALU, NLDS, LDK><<<grid, block>>>(d, iters, 1.0001f, 0.9999f); });
  const double fmas = (double)grid * block * iters * 32;
  printf("{\"probe\":\"mix\",\"packed\":%d,\"nalu_per_32fma\":%d,\"nlds_per_32fma\":%d,\"ldkind\":%d,\"ms\":%.3f,\"fma_lanes_per_clk_per_sm\":%.1f}\n",
         PACKED ? 1 : 0, NALU, NLDS, LDK, ms, fmas / (ms * 1e-3) / (g_clk_mhz * 1e6) / sms);
}


// 32 scalar FFMA + NLDS distinct LDS.32 + NDF independent DFMA per iteration: can the FP64 pipe work in the
// issue slots the FP32 stream leaves idle (including the cycles lost to shared-memory load write-back)?
template <int NLDS, int NDF>
__global__ void __launch_bounds__(256) k_mix64(float* out, int iters, float x0, float y0, double dx, double dy) {
  __shared__ float smf[256 * 4];
  for (int i = threadIdx.x; i < 1024; i += 256) smf[i] = i * 1e-3f;
  __syncthreads();
  float acc1[32];
  double dacc[16];
  unsigned lsum[8];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc1[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < 16; ++i) dacc[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) lsum[i] = 0u;
#pragma unroll 4
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      acc1[2 * i] = ffma1(acc1[2 * i], x0, y0); acc1[2 * i + 1] = ffma1(acc1[2 * i + 1], x0, y0);
      if (NDF > 0 && (i % (16 / NDF)) == 0) {
        asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(dacc[i]) : "d"(dacc[i]), "d"(dx), "d"(dy));
      }
      if (NLDS > 0 && (i % (16 / NLDS)) == 0) {
        unsigned v;
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "l"(__cvta_generic_to_shared(smf + threadIdx.x + ((it + i) & 3) * 256)) : "memory");
        lsum[(i / (16 / NLDS)) & 7] ^= v;
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc1[i];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += (float)dacc[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += (float)lsum[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NLDS, int NDF>
static void run_mix64(float* d, int sms) {
  const int grid = sms * 8, block = 256, iters = 20000;
  float ms = time_ms([&] { k_mix64<NLDS, NDF><<<grid, block>>>(d, iters, 1.0001f, 0.9999f, 1.0000001, 0.9999999); });
  const double fmas = (double)grid * block * iters * 32;
  const double cyc = ms * 1e-3 * g_clk_mhz * 1e6 / iters / (grid * 8.0 / sms / 4.0);   // cycles per warp-iteration per SMSP
  printf("{\"probe\":\"mix64\",\"nlds32_per_32fma\":%d,\"ndfma_per_32fma\":%d,\"ms\":%.3f,\"fma_lanes_per_clk_per_sm\":%.1f,\"cycles_per_iter\":%.1f}\n",
         NLDS, NDF, ms, fmas / (ms * 1e-3) / (g_clk_mhz * 1e6) / sms, cyc);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  float* d; CK(cudaMalloc(&d, sizeof(float) * 256 * sms * 8));
  for (int i = 0; i < 40; ++i) k_mix<false, 0, 0><<<sms * 8, 256>>>(d, 20000, 1.0001f, 0.9999f);
  CK(cudaDeviceSynchronize());
  measure_clock();
  run_mix<false, 0, 0>(d, sms); run_mix<false, 4, 0>(d, sms); run_mix<false, 8, 0>(d, sms); run_mix<false, 16, 0>(d, sms);
  run_mix<true, 0, 0>(d, sms);  run_mix<true, 4, 0>(d, sms);  run_mix<true, 8, 0>(d, sms);  run_mix<true, 16, 0>(d, sms);
  run_mix<true, 32, 0>(d, sms);
  run_mix<false, 0, 2>(d, sms); run_mix<false, 0, 4>(d, sms); run_mix<false, 0, 8>(d, sms);
  run_mix<true, 0, 2>(d, sms);  run_mix<true, 0, 4>(d, sms);  run_mix<true, 0, 8>(d, sms);
  run_mix<true, 8, 4>(d, sms);  run_mix<true, 16, 8>(d, sms);
  printf("{\"note\":\"ldkind 0=LDS.64 distinct 1=LDS.32 distinct 2=LDS.32 broadcast 3=LDS.128 distinct 4=LDS.128 broadcast\"}\n");
  run_mix<false, 0, 2, 1>(d, sms); run_mix<false, 0, 4, 1>(d, sms); run_mix<false, 0, 8, 1>(d, sms);
  run_mix<false, 0, 2, 2>(d, sms); run_mix<false, 0, 4, 2>(d, sms); run_mix<false, 0, 8, 2>(d, sms);
  run_mix<false, 0, 2, 3>(d, sms); run_mix<false, 0, 4, 3>(d, sms);
  run_mix<false, 0, 2, 4>(d, sms); run_mix<false, 0, 4, 4>(d, sms);
  run_mix64<0, 0>(d, sms); run_mix64<0, 2>(d, sms); run_mix64<0, 4>(d, sms); run_mix64<0, 8>(d, sms); run_mix64<0, 16>(d, sms);
  run_mix64<4, 0>(d, sms); run_mix64<4, 2>(d, sms); run_mix64<4, 4>(d, sms); run_mix64<4, 8>(d, sms); run_mix64<4, 16>(d, sms);
  measure_clock();
  {
    const int grid = sms * 8, block = 256, iters = 20000;
    float ms = time_ms([&] { k_distinct<false><<<grid, block>>>(d, iters, 1.0001f, 0.9999f); });
    double fmas = (double)grid * block * iters * 32;
    printf("{\"probe\":\"distinct_src\",\"packed\":0,\"ms\":%.3f,\"fma_lanes_per_clk_per_sm\":%.1f}\n", ms, fmas / (ms * 1e-3) / (g_clk_mhz * 1e6) / sms);
    ms = time_ms([&] { k_distinct<true><<<grid, block>>>(d, iters, 1.0001f, 0.9999f); });
    printf("{\"probe\":\"distinct_src\",\"packed\":1,\"ms\":%.3f,\"fma_lanes_per_clk_per_sm\":%.1f}\n", ms, fmas / (ms * 1e-3) / (g_clk_mhz * 1e6) / sms);
  }
  CK(cudaFree(d));
  return 0;
}
