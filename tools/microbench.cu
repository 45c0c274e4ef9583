// Pipe-rate microbenchmarks for sm_100a (B200).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench tools/microbench.cu
// Prints one JSON line per probe.  These numbers size the K1 (C(t) lag kernel) register tile and
// give the measured FP32/FP64 CUDA-core denominators that MEASURED_PEAKS.json does not carry.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

static int g_sms = 148;

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

// ---- 1. plain FFMA, 3 distinct registers, 16 independent chains --------------------------------
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float x0, float y0) {
  float acc[16], x[4], y[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) { x[i] = x0 + i * 1e-3f; y[i] = y0 + i * 1e-3f + threadIdx.x * 1e-6f; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(x[i & 3], y[(i >> 2) & 3], acc[i]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- 2. the K1 inner pattern: acc[j] += (a . w[j])^2, R lags in registers, no loads ------------
template <int R>
__global__ void __launch_bounds__(256) k_p2pattern(float* out, int iters, float seed) {
  float wx[R], wy[R], wz[R], acc[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    wx[j] = seed + 0.01f * j + threadIdx.x * 1e-4f; wy[j] = seed - 0.02f * j; wz[j] = 0.5f * seed + 0.003f * j;
    acc[j] = 0.f;
  }
  float ax = seed * 0.3f, ay = seed * 0.4f, az = seed * 0.5f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < R; ++k) {
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int s = (k + j) % R;
        float d = ax * wx[s];
        d = fmaf(ay, wy[s], d);
        d = fmaf(az, wz[s], d);
        acc[j] = fmaf(d, d, acc[j]);
      }
      // fake slide so the compiler cannot hoist: perturb slot k and the left vector
      wx[k] += 1e-7f; ax += 1e-8f;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < R; ++j) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- 2b. the K1 pattern WITH its shared-memory traffic: broadcast left vector + lane-strided window load
// MODE 0: LDS.128 for both (the kernel as shipped in round 1)   1: no loads   2: window load only
//      3: left-vector load only   4: SoA planes, 3 x LDS.32 each   5: left vector via SHFL from a per-lane
//      register copy (one coalesced LDS.128 per 32 steps), window LDS.128   6: window via SHFL from the neighbour lane
template <int R, int MODE>
__global__ void __launch_bounds__(256, 2) k_p2pattern_lds(float* out, int iters) {
  extern __shared__ float4 smv[];
  float* smf = reinterpret_cast<float*>(smv);
  const int n = 4096;
  for (int i = threadIdx.x; i < n; i += blockDim.x) smv[i] = make_float4(0.3f + 1e-4f * i, 0.4f - 1e-4f * i, 0.5f, 0.f);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int o = lane * R;
  float wx[R], wy[R], wz[R], acc[R];
#pragma unroll
  for (int j = 0; j < R; ++j) { float4 v = smv[o + j]; wx[j] = v.x; wy[j] = v.y; wz[j] = v.z; acc[j] = 0.f; }
  int s = warp * 64;
  float4 areg = smv[lane];
  float4 a = make_float4(0.3f, 0.4f, 0.5f, 0.f), nx = make_float4(0.31f, 0.41f, 0.51f, 0.f);
  float4 a2 = a, nx2 = nx;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 5 && (it & 1) == 0) areg = smv[(s + lane) & 2047];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (MODE == 0 || MODE == 3) a = smv[(s + k) & 2047];
      if (MODE == 0 || MODE == 2 || MODE == 5) nx = smv[((s + k) & 1023) + o + R];
      if (MODE == 4) {
        const int ia = (s + k) & 2047, iw = ((s + k) & 1023) + o + R;
        a.x = smf[ia]; a.y = smf[4096 + ia]; a.z = smf[8192 + ia];
        nx.x = smf[iw]; nx.y = smf[4096 + iw]; nx.z = smf[8192 + iw];
      }
      if (MODE == 7) {   // left vector SoA scalar broadcast, window LDS.128
        const int ia = (s + k) & 2047;
        a.x = smf[ia]; a.y = smf[4096 + ia]; a.z = smf[8192 + ia];
        nx = smv[((s + k) & 1023) + o + R];
      }
      if (MODE == 8) {   // left vector LDS.128, window SoA scalars
        const int iw = ((s + k) & 1023) + o + R;
        a = smv[(s + k) & 2047];
        nx.x = smf[iw]; nx.y = smf[4096 + iw]; nx.z = smf[8192 + iw];
      }
      if (MODE == 9) {   // AoS float4 in memory, but read as LDS.64 (xy) + LDS.32 (z): no fourth register written
        const int ia = (s + k) & 2047, iw = ((s + k) & 1023) + o + R;
        const float2 axy = *reinterpret_cast<const float2*>(&smv[ia]); a.x = axy.x; a.y = axy.y; a.z = smf[4 * ia + 2];
        const float2 wxy = *reinterpret_cast<const float2*>(&smv[iw]); nx.x = wxy.x; nx.y = wxy.y; nx.z = smf[4 * iw + 2];
      }
      if (MODE == 10) {  // left vector SoA scalar broadcast, window via SHFL from the neighbour lane
        const int ia = (s + k) & 2047;
        a.x = smf[ia]; a.y = smf[4096 + ia]; a.z = smf[8192 + ia];
        nx.x = __shfl_down_sync(0xffffffffu, wx[k], 1); nx.y = __shfl_down_sync(0xffffffffu, wy[k], 1);
        nx.z = __shfl_down_sync(0xffffffffu, wz[k], 1);
        if (lane == 31) { const int iw = ((s + k) & 1023) + o + R; nx.x = smf[iw]; nx.y = smf[4096 + iw]; nx.z = smf[8192 + iw]; }
      }
      if (MODE == 11) {  // SoA, two steps per load: LDS.64 for both (R odd: window pairs alternate alignment, emulate with even k only)
        if ((k & 1) == 0) {
          const int ia = (s + k) & 2046, iw = (((s + k) & 1022) + o + R) & ~1;
          const float2 ax2 = *reinterpret_cast<const float2*>(smf + ia), ay2 = *reinterpret_cast<const float2*>(smf + 4096 + ia),
                       az2 = *reinterpret_cast<const float2*>(smf + 8192 + ia);
          const float2 wx2 = *reinterpret_cast<const float2*>(smf + iw), wy2 = *reinterpret_cast<const float2*>(smf + 4096 + iw),
                       wz2 = *reinterpret_cast<const float2*>(smf + 8192 + iw);
          a.x = ax2.x; a.y = ay2.x; a.z = az2.x; a2.x = ax2.y; a2.y = ay2.y; a2.z = az2.y;
          nx.x = wx2.x; nx.y = wy2.x; nx.z = wz2.x; nx2.x = wx2.y; nx2.y = wy2.y; nx2.z = wz2.y;
        } else { a = a2; nx = nx2; }
      }
      if (MODE == 5) {
        const int src = (k + (it & 1) * R) & 31;
        a.x = __shfl_sync(0xffffffffu, areg.x, src); a.y = __shfl_sync(0xffffffffu, areg.y, src);
        a.z = __shfl_sync(0xffffffffu, areg.z, src);
      }
      if (MODE == 6) {
        a = smv[(s + k) & 2047];
        nx.x = __shfl_down_sync(0xffffffffu, wx[k], 1); nx.y = __shfl_down_sync(0xffffffffu, wy[k], 1);
        nx.z = __shfl_down_sync(0xffffffffu, wz[k], 1);
        if (lane == 31) { const float4 t4 = smv[((s + k) & 1023) + o + R]; nx.x = t4.x; nx.y = t4.y; nx.z = t4.z; }
      }
      if (MODE == 1 || MODE == 3) { nx.x += 1e-7f; nx.y -= 1e-7f; nx.z += 2e-7f; }
      if (MODE == 1 || MODE == 2) { a.x += 1e-8f; a.y -= 1e-8f; a.z += 2e-8f; }
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int sl = (k + j) % R;
        float d = a.x * wx[sl];
        d = fmaf(a.y, wy[sl], d);
        d = fmaf(a.z, wz[sl], d);
        acc[j] = fmaf(d, d, acc[j]);
      }
      wx[k] = nx.x; wy[k] = nx.y; wz[k] = nx.z;
    }
    s += R;
  }
  float t = 0.f;
#pragma unroll
  for (int j = 0; j < R; ++j) t += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

// ---- 3. packed FFMA2 (fma.rn.f32x2), 16 independent 64-bit chains ------------------------------
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  return (uint64_t)__float_as_uint(lo) | ((uint64_t)__float_as_uint(hi) << 32);
}
__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters, float x0, float y0) {
  uint64_t acc[16], x[4], y[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = pack2(threadIdx.x * 1e-3f + i, 1.f + i);
#pragma unroll
  for (int i = 0; i < 4; ++i) { x[i] = pack2(x0 + i * 1e-3f, x0 - i * 1e-3f); y[i] = pack2(y0 + i * 1e-3f, y0 + threadIdx.x * 1e-6f); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = ffma2(x[i & 3], y[(i >> 2) & 3], acc[i]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += __uint_as_float((uint32_t)acc[i]) + __uint_as_float((uint32_t)(acc[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// ---- 3b. the K1 pattern with packed FFMA2: window frames as aligned pairs, two accumulator pairings
// (A: left vector of step t, B: left vector of step t+1), R lags per lane, no loads
template <int R>
__global__ void __launch_bounds__(256) k_p2pattern_ffma2(float* out, int iters, float seed) {
  constexpr int H = R / 2;
  uint64_t X2[H], Y2[H], Z2[H], accA[H], accB[H];
#pragma unroll
  for (int m = 0; m < H; ++m) {
    X2[m] = pack2(seed + 0.01f * m + threadIdx.x * 1e-4f, seed - 0.01f * m);
    Y2[m] = pack2(seed - 0.02f * m, seed + 0.02f * m);
    Z2[m] = pack2(0.5f * seed + 0.003f * m, 0.4f * seed);
    accA[m] = 0ull; accB[m] = 0ull;
  }
  float ax = seed * 0.3f, ay = seed * 0.4f, az = seed * 0.5f, bx = seed * 0.2f, by = seed * 0.1f, bz = seed * 0.6f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < H; ++k) {       // one step pair (t, t+1) per k; window slides by one aligned pair
      const uint64_t AX = pack2(ax, ax), AY = pack2(ay, ay), AZ = pack2(az, az);
      const uint64_t BX = pack2(bx, bx), BY = pack2(by, by), BZ = pack2(bz, bz);
#pragma unroll
      for (int m = 0; m < H; ++m) {
        const int s = (k + m) % H;
        uint64_t d = fmul2(AX, X2[s]);
        d = ffma2(AY, Y2[s], d);
        d = ffma2(AZ, Z2[s], d);
        accA[m] = ffma2(d, d, accA[m]);
        uint64_t e = fmul2(BX, X2[s]);
        e = ffma2(BY, Y2[s], e);
        e = ffma2(BZ, Z2[s], e);
        accB[m] = ffma2(e, e, accB[m]);
      }
      X2[k] += 0x0000000100000001ull;   // fake slide
      ax += 1e-8f; bx += 1e-8f;
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int m = 0; m < H; ++m)
    sum += __uint_as_float((uint32_t)accA[m]) + __uint_as_float((uint32_t)(accA[m] >> 32)) +
           __uint_as_float((uint32_t)accB[m]) + __uint_as_float((uint32_t)(accB[m] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

// ---- 4. DFMA chains ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double x0, double y0) {
  double acc[8], x[4], y[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = x0 + i * 1e-3;
  y[0] = y0; y[1] = y0 + threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fma(x[i & 3], y[i >> 2], acc[i]);
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- 5. F2D + DADD flush pattern --------------------------------------------------------------
__global__ void __launch_bounds__(256) k_flush(double* out, int iters, float seed) {
  float f[8]; double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = seed + i + threadIdx.x * 1e-3f; acc[i] = 0.; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] += (double)f[i]; f[i] += 1.0f; }
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- 6. LDS.128 throughput: lane stride of STRIDE float4 (odd = conflict-free) ----------------
template <int STRIDE>
__global__ void __launch_bounds__(256) k_lds128(float* out, int iters) {
  extern __shared__ float4 sm[];
  const int n = 32 * STRIDE * 8 + 64;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = make_float4(i, 1.f, 2.f, 3.f);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float4 s = make_float4(0, 0, 0, 0);
  int base = w * 32 * STRIDE / 8 + lane * STRIDE;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float4 v = sm[base + k];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    base = (base + 1) & 1023;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s.x + s.y + s.z + s.w;
}

// ---- 7. shared-memory atomics, 2592 bins, pseudo-random vs clustered bins ---------------------
template <int MODE>   // 0 = uniform random bins, 1 = clustered (random walk over ~40 bins), 2 = all lanes same bin
__global__ void __launch_bounds__(256) k_atoms(unsigned* out, int iters) {
  __shared__ unsigned hist[2592];
  for (int i = threadIdx.x; i < 2592; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  for (int it = 0; it < iters; ++it) {
    s = s * 1664525u + 1013904223u;
    unsigned bin;
    if (MODE == 0) bin = (s >> 8) % 2592u;
    else if (MODE == 1) bin = 1000u + ((s >> 10) % 6u) * 36u + ((s >> 16) % 7u);
    else bin = 1234u + (it & 1);
    atomicAdd(&hist[bin], 1u);
  }
  __syncthreads();
  unsigned t = 0;
  for (int i = threadIdx.x; i < 2592; i += blockDim.x) t += hist[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

// effective SM clock: cycles (clock64) per wall ns (globaltimer) over a busy FFMA loop
__global__ void k_clock(unsigned long long* out, int iters) {
  unsigned long long t0, c0 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  float a = threadIdx.x * 1e-3f, b = 1.0001f;
  for (int i = 0; i < iters; ++i) a = fmaf(a, b, 1e-6f);
  unsigned long long t1, c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = c1 - c0; out[1] = t1 - t0; out[2] = (unsigned long long)a; }
}
static void report_clock(const char* tag) {
  unsigned long long* d; CK(cudaMalloc(&d, 32));
  k_clock<<<148 * 4, 256>>>(d, 400000);
  unsigned long long h[3]; CK(cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost));
  printf("{\"probe\":\"clock\",\"after\":\"%s\",\"sm_mhz_effective\":%.1f}\n", tag, (double)h[0] / (double)h[1] * 1e3);
  CK(cudaFree(d));
}

// ---- 1b. FFMA with two shared sources (reuse-cache friendly) and immediate form ----------------
__global__ void __launch_bounds__(256) k_ffma_shared(float* out, int iters, float x0, float y0) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], x0, y0);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_ffma_imm(float* out, int iters) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], 0.999f, 0.25f);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  g_sms = p.multiProcessorCount;
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("{\"probe\":\"device\",\"name\":\"%s\",\"sms\":%d,\"clock_mhz\":%d}\n", p.name, g_sms, clk_khz / 1000);
  float* d; CK(cudaMalloc(&d, sizeof(double) * 256 * 148 * 16));
  const int grid = g_sms * 8, block = 256;
  const double nthr = (double)grid * block;
  // warm the clocks for ~1 s before the first probe
  for (int i = 0; i < 60; ++i) k_ffma<<<grid, block>>>(d, 20000, 1.0001f, 0.9999f);
  CK(cudaDeviceSynchronize());
  report_clock("warmup");
  {
    int iters = 20000;
    float ms = time_ms([&] { k_ffma_shared<<<grid, block>>>(d, iters, 1.0001f, 0.9999f); });
    printf("{\"probe\":\"ffma_shared_src\",\"ms\":%.3f,\"tflops\":%.2f}\n", ms, nthr * iters * 32 / ms * 1e-9);
    ms = time_ms([&] { k_ffma_imm<<<grid, block>>>(d, iters); });
    printf("{\"probe\":\"ffma_imm\",\"ms\":%.3f,\"tflops\":%.2f}\n", ms, nthr * iters * 32 / ms * 1e-9);
    report_clock("ffma_variants");
  }
  {
    int iters = 20000;
    float ms = time_ms([&] { k_ffma<<<grid, block>>>(d, iters, 1.0001f, 0.9999f); });
    double fl = nthr * iters * 16 * 2;
    printf("{\"probe\":\"ffma\",\"ms\":%.3f,\"tflops\":%.2f}\n", ms, fl / ms * 1e-9);
  }
  {
    int iters = 1500;
    float ms = time_ms([&] { k_p2pattern<15><<<grid, block>>>(d, iters, 0.7f); });
    double pairs = nthr * iters * 15 * 15;
    printf("{\"probe\":\"p2pattern_R15\",\"ms\":%.3f,\"pairs_per_s\":%.4g,\"tflops7\":%.2f}\n", ms, pairs / ms * 1e3, pairs * 7 / ms * 1e-9);
    ms = time_ms([&] { k_p2pattern<9><<<grid, block>>>(d, iters, 0.7f); });
    pairs = nthr * iters * 9 * 9;
    printf("{\"probe\":\"p2pattern_R9\",\"ms\":%.3f,\"pairs_per_s\":%.4g,\"tflops7\":%.2f}\n", ms, pairs / ms * 1e3, pairs * 7 / ms * 1e-9);
    ms = time_ms([&] { k_p2pattern<16><<<grid, block>>>(d, iters, 0.7f); });
    pairs = nthr * iters * 16 * 16;
    printf("{\"probe\":\"p2pattern_R16\",\"ms\":%.3f,\"pairs_per_s\":%.4g,\"tflops7\":%.2f}\n", ms, pairs / ms * 1e3, pairs * 7 / ms * 1e-9);
    {
      size_t smem = 4096 * 16 + 1024;
      const int g2 = g_sms * 2;
      const double nthr2 = (double)g2 * block;
      int it2 = 6000;
#define LDS_PROBE(MODE)                                                                                         \
      CK(cudaFuncSetAttribute(k_p2pattern_lds<15, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      ms = time_ms([&] { k_p2pattern_lds<15, MODE><<<g2, block, smem>>>(d, it2); });                              \
      pairs = nthr2 * it2 * 15 * 15;                                                                              \
      printf("{\"probe\":\"p2pattern_lds_R15_mode%d\",\"ms\":%.3f,\"pairs_per_s\":%.4g,\"tflops7\":%.2f}\n", MODE, ms, pairs / ms * 1e3, pairs * 7 / ms * 1e-9);
      LDS_PROBE(0) LDS_PROBE(3) LDS_PROBE(4) LDS_PROBE(6) LDS_PROBE(7) LDS_PROBE(10) LDS_PROBE(11)
#undef LDS_PROBE
    }
  }
  {
    int iters = 3000;
    float ms = time_ms([&] { k_p2pattern_ffma2<12><<<grid, block>>>(d, iters, 0.7f); });
    double pairs = nthr * iters * 6.0 * 6 * 4;      // H step-pairs x H window pairs x (2 lags x 2 left vectors)
    printf("{\"probe\":\"p2pattern_ffma2_R12\",\"ms\":%.3f,\"pairs_per_s\":%.4g,\"tflops7\":%.2f}\n", ms, pairs / ms * 1e3, pairs * 7 / ms * 1e-9);
    ms = time_ms([&] { k_p2pattern_ffma2<16><<<grid, block>>>(d, iters, 0.7f); });
    pairs = nthr * iters * 8.0 * 8 * 4;
    printf("{\"probe\":\"p2pattern_ffma2_R16\",\"ms\":%.3f,\"pairs_per_s\":%.4g,\"tflops7\":%.2f}\n", ms, pairs / ms * 1e3, pairs * 7 / ms * 1e-9);
  }
  {
    int iters = 20000;
    float ms = time_ms([&] { k_ffma2<<<grid, block>>>(d, iters, 1.0001f, 0.9999f); });
    double fl = nthr * iters * 16 * 4;
    printf("{\"probe\":\"ffma2\",\"ms\":%.3f,\"tflops\":%.2f}\n", ms, fl / ms * 1e-9);
  }
  {
    int iters = 20000;
    float ms = time_ms([&] { k_dfma<<<grid, block>>>((double*)d, iters, 1.0000001, 0.9999999); });
    double fl = nthr * iters * 8 * 2;
    printf("{\"probe\":\"dfma\",\"ms\":%.3f,\"tflops\":%.2f}\n", ms, fl / ms * 1e-9);
    report_clock("dfma");
  }
  {
    int iters = 20000;
    float ms = time_ms([&] { k_flush<<<grid, block>>>((double*)d, iters, 1.0f); });
    double ops = nthr * iters * 8;
    printf("{\"probe\":\"f2d_dadd_fadd\",\"ms\":%.3f,\"gflush_per_s\":%.2f}\n", ms, ops / ms * 1e-6);
  }
  {
    int iters = 4000;
    size_t smem = (32 * 15 * 8 + 64 + 1024) * 16;
    CK(cudaFuncSetAttribute(k_lds128<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_lds128<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_lds128<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float ms = time_ms([&] { k_lds128<15><<<g_sms * 2, block, smem>>>(d, iters); });
    double bytes = (double)g_sms * 2 * block * iters * 8 * 16;
    printf("{\"probe\":\"lds128_stride15\",\"ms\":%.3f,\"B_per_clk_per_sm_at_max\":%.1f,\"TBps\":%.2f}\n", ms, bytes / (ms * 1e-3) / g_sms / (clk_khz * 1e3), bytes / ms * 1e-9);
    ms = time_ms([&] { k_lds128<16><<<g_sms * 2, block, smem>>>(d, iters); });
    printf("{\"probe\":\"lds128_stride16\",\"ms\":%.3f,\"TBps\":%.2f}\n", ms, bytes / ms * 1e-9);
    ms = time_ms([&] { k_lds128<1><<<g_sms * 2, block, smem>>>(d, iters); });
    printf("{\"probe\":\"lds128_stride1\",\"ms\":%.3f,\"TBps\":%.2f}\n", ms, bytes / ms * 1e-9);
  }
  {
    int iters = 20000;
    float ms = time_ms([&] { k_atoms<0><<<grid, block>>>((unsigned*)d, iters); });
    printf("{\"probe\":\"atoms_uniform\",\"ms\":%.3f,\"gatom_per_s\":%.2f}\n", ms, nthr * iters / ms * 1e-6);
    ms = time_ms([&] { k_atoms<1><<<grid, block>>>((unsigned*)d, iters); });
    printf("{\"probe\":\"atoms_clustered42\",\"ms\":%.3f,\"gatom_per_s\":%.2f}\n", ms, nthr * iters / ms * 1e-6);
    ms = time_ms([&] { k_atoms<2><<<grid, block>>>((unsigned*)d, iters); });
    printf("{\"probe\":\"atoms_same\",\"ms\":%.3f,\"gatom_per_s\":%.2f}\n", ms, nthr * iters / ms * 1e-6);
  }
  CK(cudaFree(d));
  return 0;
}
