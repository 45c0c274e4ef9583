// K4: quaternion-displacement statistics  dq(t) = conj(q(t)) * q(t+delta)  over lag windows, reduced to
// the second moments sum_t v v^T of the vector part, per lag and per sub-chunk.
// Reference arithmetic: obtain_self_dq calculate-dq-distribution.py:102-109 (quat_invert / quat_mult_simd /
// quat_reduce_simd, transforms3d_supplement.py:185,163-183,219-227), average_anisotropic_tensor[_chunk]
// :118-144, average_LegendreP1quat[_chunk] :111-135.  Inputs are float32 (the PLUMED reader rounds every
// field to float32, plumedcolvario.py:68); products and sums are float64 exactly as NumPy promotes them.
//
// The sign image of dq (w >= 0) flips v -> -v, which leaves v v^T and |v|^2 unchanged, so the moments
// never need it; sr_dq_self applies it for callers that want the displacement quaternions themselves.
#include "common.cuh"

namespace {

constexpr int kDqThreads = 256;
constexpr int kDqPerThread = 16;
constexpr int kDqTile = kDqThreads * kDqPerThread;   // frames per CTA per lag

struct Vec3d { double x, y, z; };

__device__ __forceinline__ Vec3d dq_vector(const float4 a, const float4 b) {
  // q stored (w,x,y,z) -> float4 (x=w, y=x, z=y, w=z).  Vector part of conj(a) * b.
  const double w1 = a.x, x1 = a.y, y1 = a.z, z1 = a.w;
  const double w2 = b.x, x2 = b.y, y2 = b.z, z2 = b.w;
  Vec3d v;
  v.x = (w1 * x2 - w2 * x1) + (z1 * y2 - y1 * z2);
  v.y = (w1 * y2 - w2 * y1) + (x1 * z2 - z1 * x2);
  v.z = (w1 * z2 - w2 * z1) + (y1 * x2 - x1 * y2);
  return v;
}

// grid.x = lagIndex * tilesMax + tile
__global__ void __launch_bounds__(kDqThreads)
dq_moments_kernel(const float4* __restrict__ q, long long N, const long long* __restrict__ lags, int nLags, int nCh,
                  int tilesMax, double* __restrict__ M) {
  const int li = blockIdx.x / tilesMax;
  const int tile = blockIdx.x - li * tilesMax;
  const long long delta = lags[li];
  const long long n = N - delta;
  const long long lo = (long long)tile * kDqTile;
  if (lo >= n) return;
  const long long hi = min(n, lo + kDqTile);
  const long long nb = (n + nCh - 1) / nCh;   // ceil(n / nchunk), calculate-dq-distribution.py:129
  __shared__ double red[kDqThreads / 32][6];

  for (long long k = lo / nb; k * nb < hi; ++k) {
    const long long a = max(lo, k * nb), b = min(hi, (k + 1) * nb);
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (long long t = a + threadIdx.x; t < b; t += kDqThreads) {
      const Vec3d v = dq_vector(__ldg(q + t), __ldg(q + t + delta));
      s[0] = fma(v.x, v.x, s[0]); s[1] = fma(v.x, v.y, s[1]); s[2] = fma(v.x, v.z, s[2]);
      s[3] = fma(v.y, v.y, s[3]); s[4] = fma(v.y, v.z, s[4]); s[5] = fma(v.z, v.z, s[5]);
    }
#pragma unroll
    for (int m = 0; m < 6; ++m) s[m] = sr_warp_sum(s[m]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int m = 0; m < 6; ++m) red[warp][m] = s[m];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < kDqThreads / 32; ++w) tot += red[w][threadIdx.x];
      atomicAdd(&M[((long long)li * nCh + k) * 6 + threadIdx.x], tot);
    }
  }
}

__global__ void __launch_bounds__(256)
dq_self_kernel(const float4* __restrict__ q, long long N, long long delta, double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N - delta) return;
  const float4 a = __ldg(q + t), b = __ldg(q + t + delta);
  const double w1 = a.x, x1 = a.y, y1 = a.z, z1 = a.w;
  const double w2 = b.x, x2 = b.y, y2 = b.z, z2 = b.w;
  // quat_mult_simd(conj(a), b): w = w1 w2 - (conj v1).v2
  double w = w1 * w2 - ((-x1) * x2 + (-y1) * y2 + (-z1) * z2);
  Vec3d v = dq_vector(a, b);
  const double sgn = (w < 0.0) ? -1.0 : 1.0;   // quat_reduce_simd: sign(q.qref), 0 counts as +
  double4* o = reinterpret_cast<double4*>(out) + t;
  *o = make_double4(w * sgn, v.x * sgn, v.y * sgn, v.z * sgn);
}

// second moments of an arbitrary (n,3) float64 vector list, split in nCh consecutive blocks of ceil(n/nCh)
__global__ void __launch_bounds__(kDqThreads)
vec_moments_kernel(const double* __restrict__ v, long long n, int nCh, double* __restrict__ M) {
  const long long nb = (n + nCh - 1) / nCh;
  const long long lo = (long long)blockIdx.x * kDqTile;
  if (lo >= n) return;
  const long long hi = min(n, lo + kDqTile);
  __shared__ double red[kDqThreads / 32][6];
  for (long long k = lo / nb; k * nb < hi; ++k) {
    const long long a = max(lo, k * nb), b = min(hi, (k + 1) * nb);
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (long long t = a + threadIdx.x; t < b; t += kDqThreads) {
      const double x = v[3 * t], y = v[3 * t + 1], z = v[3 * t + 2];
      s[0] = fma(x, x, s[0]); s[1] = fma(x, y, s[1]); s[2] = fma(x, z, s[2]);
      s[3] = fma(y, y, s[3]); s[4] = fma(y, z, s[4]); s[5] = fma(z, z, s[5]);
    }
#pragma unroll
    for (int m = 0; m < 6; ++m) s[m] = sr_warp_sum(s[m]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int m = 0; m < 6; ++m) red[warp][m] = s[m];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < kDqThreads / 32; ++w) tot += red[w][threadIdx.x];
      atomicAdd(&M[(long long)k * 6 + threadIdx.x], tot);
    }
  }
}

}  // namespace

extern "C" int sr_dq_moments(const float* d_q, long long N, const long long* d_lags, int nLags, long long min_lag,
                             int nCh, double* d_M, void* stream) {
  SR_REQUIRE(d_q && d_lags && d_M, "sr_dq_moments: null pointer");
  SR_REQUIRE(N >= 2 && nLags > 0 && nCh >= 1, "sr_dq_moments: bad shape (N=%lld nLags=%d nCh=%d)", N, nLags, nCh);
  SR_REQUIRE(min_lag >= 1 && min_lag < N, "sr_dq_moments: min_lag %lld outside [1, N)", min_lag);
  const long long tilesMax = (N - min_lag + kDqTile - 1) / kDqTile;
  const long long blocks = tilesMax * nLags;
  SR_REQUIRE(blocks < (1LL << 31), "sr_dq_moments: %lld blocks exceed the grid limit; split the lag list", blocks);
  SR_CUDA(cudaMemsetAsync(d_M, 0, sizeof(double) * 6 * (size_t)nLags * nCh, (cudaStream_t)stream));
  dq_moments_kernel<<<(unsigned)blocks, kDqThreads, 0, (cudaStream_t)stream>>>((const float4*)d_q, N, d_lags, nLags, nCh,
                                                                              (int)tilesMax, d_M);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_dq_self(const float* d_q, long long N, long long delta, double* d_out, void* stream) {
  SR_REQUIRE(d_q && d_out, "sr_dq_self: null pointer");
  SR_REQUIRE(delta >= 1 && delta < N, "sr_dq_self: delta %lld outside [1, N)", delta);
  const long long n = N - delta;
  dq_self_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)d_q, N, delta, d_out);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_vec_second_moments(const double* d_v, long long n, int nCh, double* d_M, void* stream) {
  SR_REQUIRE(d_v && d_M, "sr_vec_second_moments: null pointer");
  SR_REQUIRE(n >= 1 && nCh >= 1, "sr_vec_second_moments: bad shape");
  SR_CUDA(cudaMemsetAsync(d_M, 0, sizeof(double) * 6 * (size_t)nCh, (cudaStream_t)stream));
  const long long blocks = (n + kDqTile - 1) / kDqTile;
  vec_moments_kernel<<<(unsigned)blocks, kDqThreads, 0, (cudaStream_t)stream>>>(d_v, n, nCh, d_M);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}
