// Register-file operand bandwidth probes for sm_100a: how fast does an FFMA/FMUL stream run when the operands cannot
// come from the operand-reuse cache?  (K1's FMA stream runs at ~71 % of the FFMA peak even with every shared-memory
// load removed, tools/tune_ct.py DIAG=4.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench3 tools/microbench3.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

#define FMA(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define MUL(d, a, b) asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))

constexpr int N = 16;

// MODE 0: acc[i] = a[i]*b[i] + acc[i]      three distinct registers per instruction, nothing reusable
// MODE 1: acc[i] = a[i]*X + acc[i]          two distinct + one reusable
// MODE 2: acc[i] = a[i]*a[i] + acc[i]       same register in two operand slots
// MODE 3: K1 order per lag: d = X*w0[i]; d = Y*w1[i] + d; d = Z*w2[i] + d; acc[i] = d*d + acc[i]   (lag after lag)
// MODE 4: K1 in phases: all muls (X reusable), all Y fmas, all Z fmas, all accumulates
// MODE 5: as 4 with the multiply written as an FMA with a zero addend
// MODE 6: as 3 with the multiply written as an FMA with a zero addend
// body-size probe: the K1-ordered stream (mode 3) with the step loop unrolled UNR times -> 64*UNR FMA-pipe instructions
// (16 bytes each) per loop body.  K1's own body is 1558 instructions = 25 KB.
template <int UNR>
__global__ void __launch_bounds__(384, 1) k_body(float* out, int iters, float x0) {
  float a[N], b[N], c[N], acc[N], d[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { a[i] = x0 + i * 1e-3f + threadIdx.x * 1e-6f; b[i] = 1.f - i * 1e-3f; c[i] = 0.5f + i * 1e-4f; acc[i] = i; d[i] = 0.f; }
  float X = x0, Y = x0 * 0.5f, Z = x0 * 0.25f;
#pragma unroll 1
  for (int it = 0; it < iters; it += UNR) {
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        MUL(d[i], X, a[i]);
        FMA(d[i], Y, b[i], d[i]);
        FMA(d[i], Z, c[i], d[i]);
        FMA(acc[i], d[i], d[i], acc[i]);
      }
      X += 1e-7f; Y -= 1e-7f; Z += 2e-7f;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) s += acc[i] + d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int UNR>
void run_body(float* d, int sms) {
  const int iters = 19200;
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(a));
    k_body<UNR><<<sms, 384>>>(d, iters, 1.0001f);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double n = (double)iters * 67 * 384 * sms;
  printf("{\"probe\":\"body\",\"unroll\":%d,\"body_instr\":%d,\"body_kb\":%.1f,\"ms\":%.3f,\"fma_lanes_per_clk_per_sm_at_1965\":%.1f}\n", UNR,
         67 * UNR, 67 * UNR * 16 / 1024.0, best, n / (best * 1e-3) / sms / 1.965e9);
}

// register-footprint probe: the K1-ordered stream over NL lags with a[], b[], c[], acc[] all live (4 NL + NL temporaries)
template <int NL, int NT>
__global__ void __launch_bounds__(NT, 1) k_regs(float* out, int iters, float x0) {
  float a[NL], b[NL], c[NL], acc[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) { a[i] = x0 + i * 1e-3f + threadIdx.x * 1e-6f; b[i] = 1.f - i * 1e-3f + threadIdx.x * 1e-6f; c[i] = 0.5f + i * 1e-4f + threadIdx.x * 1e-6f; acc[i] = i; }
  float X = x0, Y = x0 * 0.5f, Z = x0 * 0.25f;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      float d;
      MUL(d, X, a[i]);
      FMA(d, Y, b[i], d);
      FMA(d, Z, c[i], d);
      FMA(acc[i], d, d, acc[i]);
    }
    X += 1e-7f; Y -= 1e-7f; Z += 2e-7f;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NL; ++i) s += acc[i] + a[i] + b[i] + c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NL, int NT>
void run_regs(float* d, int sms) {
  const int iters = 20000;
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(a));
    k_regs<NL, NT><<<sms, NT>>>(d, iters, 1.0001f);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k_regs<NL, NT>));
  const double n = (double)iters * (4 * NL + 3) * NT * sms;
  printf("{\"probe\":\"regs\",\"lags\":%d,\"threads\":%d,\"regs\":%d,\"ms\":%.3f,\"fma_lanes_per_clk_per_sm_at_1965\":%.1f}\n", NL, NT,
         fa.numRegs, best, n / (best * 1e-3) / sms / 1.965e9);
}

template <int MODE>
__global__ void __launch_bounds__(384, 1) k_rf(float* out, int iters, float x0) {
  float a[N], b[N], c[N], acc[N], d[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { a[i] = x0 + i * 1e-3f + threadIdx.x * 1e-6f; b[i] = 1.f - i * 1e-3f; c[i] = 0.5f + i * 1e-4f; acc[i] = i; d[i] = 0.f; }
  float X = x0, Y = x0 * 0.5f, Z = x0 * 0.25f, zero = 0.f;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < N; ++i) FMA(acc[i], a[i], b[i], acc[i]);
    } else if (MODE == 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < N; ++i) FMA(acc[i], a[i], X, acc[i]);
    } else if (MODE == 2) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < N; ++i) FMA(acc[i], a[i], a[i], acc[i]);
    } else if (MODE == 3 || MODE == 6) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (MODE == 3) MUL(d[i], X, a[i]); else FMA(d[i], X, a[i], zero);
        FMA(d[i], Y, b[i], d[i]);
        FMA(d[i], Z, c[i], d[i]);
        FMA(acc[i], d[i], d[i], acc[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) { if (MODE == 4) MUL(d[i], X, a[i]); else FMA(d[i], X, a[i], zero); }
#pragma unroll
      for (int i = 0; i < N; ++i) FMA(d[i], Y, b[i], d[i]);
#pragma unroll
      for (int i = 0; i < N; ++i) FMA(d[i], Z, c[i], d[i]);
#pragma unroll
      for (int i = 0; i < N; ++i) FMA(acc[i], d[i], d[i], acc[i]);
    }
    X += 1e-7f; Y -= 1e-7f; Z += 2e-7f;      // new left vector every step, like K1
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) s += acc[i] + d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(float* d, int sms, const char* what) {
  const int iters = 20000;
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(a));
    k_rf<MODE><<<sms, 384>>>(d, iters, 1.0001f);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double fma_instr = (double)iters * 64 * 384 * sms;        // lane-instructions on the FMA pipe (64 per iteration)
  const double adds = (double)iters * 3 * 384 * sms;              // the X, Y, Z updates also use the FMA pipe
  int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  printf("{\"probe\":\"rf\",\"mode\":%d,\"what\":\"%s\",\"ms\":%.3f,\"fma_lanes_per_clk_per_sm_at_1965\":%.1f}\n", MODE, what, best,
         (fma_instr + adds) / (best * 1e-3) / sms / 1.965e9);
}

int main() {
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  float* d; CK(cudaMalloc(&d, sizeof(float) * 384 * sms));
  run<0>(d, sms, "a[i]*b[i]+acc[i]: 3 distinct, no reuse");
  run<1>(d, sms, "a[i]*X+acc[i]: one reusable");
  run<2>(d, sms, "a[i]*a[i]+acc[i]: same register twice");
  run<3>(d, sms, "K1 order, lag after lag (mul, fma, fma, fma)");
  run<6>(d, sms, "K1 order, multiply as fma+0");
  run<4>(d, sms, "K1 in phases (16 mul, 16 fma, 16 fma, 16 fma)");
  run<5>(d, sms, "K1 in phases, multiply as fma+0");
  run_regs<8, 384>(d, sms); run_regs<16, 384>(d, sms); run_regs<24, 384>(d, sms); run_regs<32, 384>(d, sms); run_regs<38, 384>(d, sms);
  run_regs<16, 256>(d, sms); run_regs<32, 256>(d, sms); run_regs<48, 256>(d, sms); run_regs<56, 256>(d, sms);
  run_regs<16, 512>(d, sms); run_regs<24, 512>(d, sms); run_regs<16, 1024>(d, sms);
  run_body<1>(d, sms); run_body<4>(d, sms); run_body<8>(d, sms); run_body<12>(d, sms); run_body<16>(d, sms);
  run_body<24>(d, sms); run_body<32>(d, sms); run_body<48>(d, sms); run_body<64>(d, sms); run_body<96>(d, sms);
  return 0;
}
