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

#include <algorithm>
#include <mutex>

namespace {

constexpr int kDqThreads = 256;
constexpr int kDqPerThread = 16;
constexpr int kDqTile = kDqThreads * kDqPerThread;   // frames per CTA (vec_moments_kernel)
constexpr int kDqLagTile = 16;                       // lags per CTA (dq_moments_kernel)
constexpr int kDqFrameTile = 1024;                   // frames per CTA (dq_moments_kernel)
constexpr int kDqU = 4;                              // q(t + delta) loads in flight per thread (x2: double buffered)
constexpr int kDqSuper = 8;                          // frame tiles per super tile of the consecutive-lag kernel
constexpr int kDqSuperFrames = kDqSuper * kDqFrameTile;

// A (lag tile, super tile) is "interior" when the 16 lags of the tile are consecutive integers d0 .. d0 + 15, every
// pair (t, t + delta) of the 8192-frame super tile exists for all of them, and for each lag the whole super tile
// falls into one sub-chunk.  dq_moments_consec_kernel reduces interior super tiles with its shared-memory fast path
// and everything else (edges of the trajectory, sub-chunk boundaries) with the generic tile routine.
__device__ __forceinline__ bool dq_lag_interior(long long N, long long delta, long long lo, int nCh, int nRep, int replica) {
  const long long n = N - delta;
  if (lo + kDqSuperFrames > n) return false;
  const long long nb = (n * nRep + nCh - 1) / nCh;
  const long long pos = (long long)replica * n + lo;
  return pos / nb == (pos + kDqSuperFrames - 1) / nb;
}

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

// CTA = (tile of 16 entries of the lag list, tile of 1024 frames).  The left quaternions q(t) of the frame tile
// are converted to double once and staged in shared memory (34 KB), so each of them is reused by the 16 lags;
// thread (lag j = tid >> 4, frame lane = tid & 15) streams q(t + delta_j) as coalesced float4 (consecutive
// lags hit the same L1 lines) and keeps its six second-moment sums in registers.  Per pair: 4 F2F + 18
// FP64 instructions, 16 bytes of L1/L2 traffic.  grid.x = lagTile * tilesMax + frameTile.
__device__ void dq_tile_generic(const float4* __restrict__ q, long long N, const long long* __restrict__ lags, int nLags,
                                int nCh, int lt, int tile, int replica, int nRep, double* __restrict__ M, double* sh_a) {
  // left quaternions of the frame tile as four planes of doubles (w | x | y | z): a half-warp reads 16
  // consecutive doubles per plane (one wavefront, broadcast to the other half-warp); an array of double4
  // would cost 8 wavefronts per LDS.128 (32-byte lane stride) and saturate the shared-memory data pipe
  constexpr int kPlane = kDqFrameTile + 16 * kDqU;
  const long long lo = (long long)tile * kDqFrameTile;
  const int jl = threadIdx.x >> 4, fl = threadIdx.x & 15;
  const int li = lt * kDqLagTile + jl;
  const unsigned halfmask = 0xffffu << (threadIdx.x & 16);
  // the smallest lag of the tile decides whether this frame tile holds any pair at all
  long long dmin = lags[lt * kDqLagTile];
  for (int j = 1; j < kDqLagTile && lt * kDqLagTile + j < nLags; ++j) dmin = min(dmin, lags[lt * kDqLagTile + j]);
  if (lo >= N - dmin) return;
  for (int i = threadIdx.x; i < kDqFrameTile + 16 * kDqU; i += kDqThreads) {   // zero tail: see the pipeline below
    const long long t = lo + i;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < N && i < kDqFrameTile) a = __ldg(q + t);
    sh_a[i] = (double)a.x; sh_a[kPlane + i] = (double)a.y; sh_a[2 * kPlane + i] = (double)a.z; sh_a[3 * kPlane + i] = (double)a.w;
  }
  __syncthreads();
  if (li >= nLags) return;
  const long long delta = lags[li];
  const long long n = N - delta;
  if (lo >= n) return;
  const long long hi = min(n, lo + kDqFrameTile);
  // Sub-chunks are consecutive blocks of ceil(n_pooled / nchunk) samples of the POOLED sample list
  // (calculate-dq-distribution.py:129; with several replica trajectories the samples of replica r occupy
  // [r n, (r + 1) n), calculate-dq-distribution-multi.py:533-539), so a block may straddle replicas.
  const long long nb = (n * nRep + nCh - 1) / nCh;
  const long long off = (long long)replica * n;          // pooled index of this replica's sample t = 0
  const float4* __restrict__ qb = q + delta;

  for (long long k = (off + lo) / nb; k * nb < off + hi; ++k) {
    const long long a0 = max(lo, k * nb - off), b0 = min(hi, (k + 1) * nb - off);
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0;
    // software pipeline: the next four q(t + delta) are in flight while the current four are consumed; a zero
    // quaternion (past the end of the segment) yields v = 0 and needs no predicate
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    long long t = a0 + fl;
    float4 bn[kDqU];
#pragma unroll
    for (int u = 0; u < kDqU; ++u) bn[u] = (t + 16 * u < b0) ? __ldg(qb + t + 16 * u) : zero4;
    while (t < b0) {
      float4 bc[kDqU];
#pragma unroll
      for (int u = 0; u < kDqU; ++u) bc[u] = bn[u];
      const long long tn = t + 16 * kDqU;
#pragma unroll
      for (int u = 0; u < kDqU; ++u) bn[u] = (tn + 16 * u < b0) ? __ldg(qb + tn + 16 * u) : zero4;
      const double* ap = sh_a + (t - lo);
#pragma unroll
      for (int u = 0; u < kDqU; ++u) {
        const double4 a = make_double4(ap[16 * u], ap[kPlane + 16 * u], ap[2 * kPlane + 16 * u], ap[3 * kPlane + 16 * u]);
        const double w2 = bc[u].x, x2 = bc[u].y, y2 = bc[u].z, z2 = bc[u].w;
        // vector part of conj(a) * b (quat_mult_simd(quat_invert(a), b)); the sign image w >= 0 leaves v v^T unchanged
        const double vx = fma(a.x, x2, fma(-w2, a.y, fma(a.w, y2, -(a.z * z2))));
        const double vy = fma(a.x, y2, fma(-w2, a.z, fma(a.y, z2, -(a.w * x2))));
        const double vz = fma(a.x, z2, fma(-w2, a.w, fma(a.z, x2, -(a.y * y2))));
        s0 = fma(vx, vx, s0); s1 = fma(vx, vy, s1); s2 = fma(vx, vz, s2);
        s3 = fma(vy, vy, s3); s4 = fma(vy, vz, s4); s5 = fma(vz, vz, s5);
      }
      t = tn;
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(halfmask, s0, o); s1 += __shfl_xor_sync(halfmask, s1, o);
      s2 += __shfl_xor_sync(halfmask, s2, o); s3 += __shfl_xor_sync(halfmask, s3, o);
      s4 += __shfl_xor_sync(halfmask, s4, o); s5 += __shfl_xor_sync(halfmask, s5, o);
    }
    if (fl == 0) {
      double* m = M + ((long long)li * nCh + k) * 6;
      atomicAdd(m + 0, s0); atomicAdd(m + 1, s1); atomicAdd(m + 2, s2);
      atomicAdd(m + 3, s3); atomicAdd(m + 4, s4); atomicAdd(m + 5, s5);
    }
  }
}

__global__ void __launch_bounds__(kDqThreads)
dq_moments_kernel(const float4* __restrict__ q, long long N, const long long* __restrict__ lags, int nLags, int nCh,
                  int tilesMax, int replica, int nRep, double* __restrict__ M) {
  extern __shared__ __align__(16) double sh_a[];
  const int lt = blockIdx.x / tilesMax;
  dq_tile_generic(q, N, lags, nLags, nCh, lt, blockIdx.x - lt * tilesMax, replica, nRep, M, sh_a);
}

// Consecutive lags (the "all windows" lag list 1, 2, 3, ...): CTA = (tile of 16 consecutive lags d0 .. d0 + 15, super tile
// of 8192 frames).  Interior super tiles take the fast path: both operands come from shared memory as float64 planes
// converted ONCE per CTA -- q(t) for a 1024-frame tile and q(t + d0 ...) for the 1024 + 16 frames the 16 lags reach -- so
// the loop has no global load and no F2F.  Thread = (lag group g of 4 lags, frame lane owning 16 consecutive frames): it
// walks its frames two at a time and keeps a sliding window of six right quaternions in registers, so a step of 8
// pairs (2 frames x 4 lags, 144 FP64 instructions) loads one new pair of left and one new pair of right quaternions:
// 16 shared-memory bytes per pair.  Planes are padded by 2 doubles every 16 (lane stride 18 doubles: an LDS.128 of a
// quarter warp touches all 32 banks once).  The 24 moment sums of a thread live in registers for the whole super
// tile; warp shuffle reduction, then one FP64 atomic per (lag, moment) and warp.
__host__ __device__ constexpr int dq_pad(int i) { return i + 2 * (i >> 4); }
constexpr int kDqcPlaneA = dq_pad(kDqFrameTile);
constexpr int kDqcPlaneB = dq_pad(kDqFrameTile + 2 * kDqLagTile);     // window reaches 16 lags + the look-ahead pair
constexpr int kDqcSmemDoubles = 4 * kDqcPlaneA + 4 * kDqcPlaneB;
static_assert(kDqcSmemDoubles >= 4 * (kDqFrameTile + 16 * kDqU), "the generic tile routine borrows this buffer");
static_assert(kDqcPlaneA % 2 == 0 && kDqcPlaneB % 2 == 0, "planes must keep 16-byte alignment");

struct DqQuatPair { double w[2], x[2], y[2], z[2]; };

__device__ __forceinline__ DqQuatPair dq_load_pair(const double* __restrict__ plane0, int planeStride, int i) {
  const double* p = plane0 + dq_pad(i);
  const double2 w = *reinterpret_cast<const double2*>(p);
  const double2 x = *reinterpret_cast<const double2*>(p + planeStride);
  const double2 y = *reinterpret_cast<const double2*>(p + 2 * planeStride);
  const double2 z = *reinterpret_cast<const double2*>(p + 3 * planeStride);
  DqQuatPair r;
  r.w[0] = w.x; r.w[1] = w.y; r.x[0] = x.x; r.x[1] = x.y; r.y[0] = y.x; r.y[1] = y.y; r.z[0] = z.x; r.z[1] = z.y;
  return r;
}

__global__ void __launch_bounds__(kDqThreads, 2)
dq_moments_consec_kernel(const float4* __restrict__ q, long long N, const long long* __restrict__ lags, long long d_first,
                         int nLags, int nCh, int superMax, int replica, int nRep, double* __restrict__ M) {
  extern __shared__ __align__(16) double sh[];
  double* const A = sh;
  double* const B = sh + 4 * kDqcPlaneA;
  const int lt = blockIdx.x / superMax;
  const int st = blockIdx.x - lt * superMax;
  const long long d0 = d_first + (long long)lt * kDqLagTile;
  const long long lo0 = (long long)st * kDqSuperFrames;
  if (lo0 >= N - d0) return;                                         // no pair at all in this super tile
  // interior test: one lag per thread (two 64-bit divisions each), combined over the CTA
  int ok = 1;
  if (threadIdx.x < kDqLagTile)
    ok = (lt * kDqLagTile + (int)threadIdx.x < nLags) && dq_lag_interior(N, d0 + threadIdx.x, lo0, nCh, nRep, replica);
  if (!__syncthreads_and(ok)) {
    // edge of the trajectory, sub-chunk boundary or ragged last lag tile: the generic routine, tile by tile
    for (int sub = 0; sub < kDqSuper; ++sub) {
      dq_tile_generic(q, N, lags, nLags, nCh, lt, st * kDqSuper + sub, replica, nRep, M, sh);
      __syncthreads();
    }
    return;
  }
  const int g = threadIdx.x >> 6, lane = threadIdx.x & 63;          // lag group (4 lags), frame lane (16 frames)
  double acc[4][6];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int m = 0; m < 6; ++m) acc[j][m] = 0.0;

  for (int sub = 0; sub < kDqSuper; ++sub) {
    const long long lo = lo0 + (long long)sub * kDqFrameTile;
    __syncthreads();                                                // previous tile fully consumed
    {
      // all global loads of the tile first (one exposed L2 latency instead of five), then the conversions
      constexpr int kNB = (kDqFrameTile + 2 * kDqLagTile + kDqThreads - 1) / kDqThreads;     // 5
      constexpr int kNA = kDqFrameTile / kDqThreads;                                         // 4
      float4 rb[kNB], ra[kNA];
#pragma unroll
      for (int u = 0; u < kNB; ++u) {
        const int i = threadIdx.x + u * kDqThreads;
        const long long tb = lo + d0 + i;                           // interior: tb < N for i < 1024 + 16; the look-ahead
        rb[u] = (i < kDqFrameTile + 2 * kDqLagTile && tb < N) ? __ldg(q + tb) : make_float4(0.f, 0.f, 0.f, 0.f);   // pair beyond is unused
      }
#pragma unroll
      for (int u = 0; u < kNA; ++u) ra[u] = __ldg(q + lo + threadIdx.x + u * kDqThreads);
#pragma unroll
      for (int u = 0; u < kNB; ++u) {
        const int i = threadIdx.x + u * kDqThreads;
        if (i < kDqFrameTile + 2 * kDqLagTile) {
          const int ib = dq_pad(i);
          B[ib] = (double)rb[u].x; B[kDqcPlaneB + ib] = (double)rb[u].y; B[2 * kDqcPlaneB + ib] = (double)rb[u].z;
          B[3 * kDqcPlaneB + ib] = (double)rb[u].w;
        }
      }
#pragma unroll
      for (int u = 0; u < kNA; ++u) {
        const int ia = dq_pad(threadIdx.x + u * kDqThreads);
        A[ia] = (double)ra[u].x; A[kDqcPlaneA + ia] = (double)ra[u].y; A[2 * kDqcPlaneA + ia] = (double)ra[u].z;
        A[3 * kDqcPlaneA + ia] = (double)ra[u].w;
      }
    }
    __syncthreads();
    const int f0 = 16 * lane;                                       // this thread's frames f0 .. f0 + 15 of the tile
    const int b0 = f0 + 4 * g;                                      // B[i] = q(lo + d0 + i): lag 4 g + j of frame f is B[f + 4 g + j]
    DqQuatPair w0 = dq_load_pair(B, kDqcPlaneB, b0), w1 = dq_load_pair(B, kDqcPlaneB, b0 + 2), w2;
#pragma unroll
    for (int step = 0; step < 8; ++step) {
      const DqQuatPair a = dq_load_pair(A, kDqcPlaneA, f0 + 2 * step);
      w2 = dq_load_pair(B, kDqcPlaneB, b0 + 2 * step + 4);
      // window: right quaternions e + j for frame e in {0, 1} and lag j in 0..3 -> entries 0..4 of (w0, w1, w2)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = e + j;                                      // 0 .. 4
          const DqQuatPair& wp = k < 2 ? w0 : (k < 4 ? w1 : w2);
          const double w2_ = wp.w[k & 1], x2 = wp.x[k & 1], y2 = wp.y[k & 1], z2 = wp.z[k & 1];
          // vector part of conj(a) * b, same expression tree as the generic routine
          const double vx = fma(a.w[e], x2, fma(-w2_, a.x[e], fma(a.z[e], y2, -(a.y[e] * z2))));
          const double vy = fma(a.w[e], y2, fma(-w2_, a.y[e], fma(a.x[e], z2, -(a.z[e] * x2))));
          const double vz = fma(a.w[e], z2, fma(-w2_, a.z[e], fma(a.y[e], x2, -(a.x[e] * y2))));
          acc[j][0] = fma(vx, vx, acc[j][0]); acc[j][1] = fma(vx, vy, acc[j][1]); acc[j][2] = fma(vx, vz, acc[j][2]);
          acc[j][3] = fma(vy, vy, acc[j][3]); acc[j][4] = fma(vy, vz, acc[j][4]); acc[j][5] = fma(vz, vz, acc[j][5]);
        }
      }
      w0 = w1; w1 = w2;
    }
  }
  // a warp holds one lag group: reduce its 24 sums and add them to the sub-chunk this super tile lies in
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long delta = d0 + 4 * g + j;
    const long long n = N - delta;
    const long long nb = (n * nRep + nCh - 1) / nCh;
    const long long k = ((long long)replica * n + lo0) / nb;
#pragma unroll
    for (int m = 0; m < 6; ++m) {
      const double v = sr_warp_sum(acc[j][m]);
      if ((threadIdx.x & 31) == 0) atomicAdd(M + (((long long)lt * kDqLagTile + 4 * g + j) * nCh + k) * 6 + m, v);
    }
  }
}

// flag = 1 unless the lag list is d, d + 1, d + 2, ...
__global__ void dq_check_consecutive_kernel(const long long* __restrict__ lags, int nLags, long long first, int* flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nLags) return;
  if (lags[i] != first + i) *flag = 1;
}

// one int of device scratch per call in flight: a small ring allocated once per device
int* dq_flag_slot() {
  static std::mutex mu;
  static int* ring[64] = {nullptr};
  static unsigned next[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!ring[dev] && cudaMalloc(&ring[dev], 256 * sizeof(int)) != cudaSuccess) return nullptr;
  return ring[dev] + (next[dev]++ & 255u);
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

// 3-D histogram of the vector part of dq over [-1,1]^3 (the --hist option, calculate-dq-distribution.py:633-647:
// np.histogramdd(v_dq, range=((-1,1),)*3, bins=(nb,nb,nb))).  dq is formed exactly like dq_self_kernel; the bin of
// every component is located against the SAME float64 edge array NumPy builds (np.linspace), so counts are
// identical to NumPy's unless a component lies within 4 ulp of an edge -- those samples (and NaNs) are listed for
// the host instead of being counted.  Counts go straight to global memory: nb^3 bins (4 MB for nb = 101) live in L2.
__device__ __forceinline__ int locate_bin(double x, const double* __restrict__ edges, int nb, bool& ambiguous) {
  // np.histogramdd: searchsorted(edges, x, 'right') - 1, samples equal to the last edge go to the last bin,
  // samples outside [edges[0], edges[nb]] are dropped (-1)
  if (!(x >= edges[0] && x <= edges[nb])) { ambiguous = ambiguous || (x != x); return -1; }
  int c = (int)floor((x - edges[0]) / (edges[nb] - edges[0]) * nb);
  c = max(0, min(nb - 1, c));
  while (c > 0 && x < edges[c]) --c;
  while (c < nb - 1 && x >= edges[c + 1]) ++c;
  const double tol = 4.0 * 2.220446049250313e-16 * fmax(fabs(x), 1e-3);
  if (fabs(x - edges[c]) <= tol || fabs(x - edges[c + 1]) <= tol) ambiguous = true;
  return c;
}

__global__ void __launch_bounds__(256)
dq_hist3d_kernel(const float4* __restrict__ q, long long N, long long delta, const double* __restrict__ edges, int nb,
                 unsigned int* __restrict__ counts, long long* __restrict__ amb_idx, int amb_capacity,
                 int* __restrict__ amb_count) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N - delta) return;
  const float4 a = __ldg(q + t), b = __ldg(q + t + delta);
  const double w1 = a.x, x1 = a.y, y1 = a.z, z1 = a.w;
  const double w2 = b.x, x2 = b.y, y2 = b.z, z2 = b.w;
  const double w = w1 * w2 - ((-x1) * x2 + (-y1) * y2 + (-z1) * z2);
  const Vec3d v = dq_vector(a, b);
  const double sgn = (w < 0.0) ? -1.0 : 1.0;
  bool amb = false;
  const int i = locate_bin(v.x * sgn, edges, nb, amb);
  const int j = locate_bin(v.y * sgn, edges, nb, amb);
  const int k = locate_bin(v.z * sgn, edges, nb, amb);
  if (amb) {
    const int slot = atomicAdd(amb_count, 1);
    if (slot < amb_capacity) amb_idx[slot] = t;
    return;
  }
  if (i < 0 || j < 0 || k < 0) return;
  atomicAdd(&counts[((long long)i * nb + j) * nb + k], 1u);
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

// Objective of the reference's 1-parameter Powell fits (powell_expdecay, calculate-dq-distribution.py:199-203):
// mean_i (C0 exp(-x_i / A) + C1 - y_i)^2 over a decay curve that stays on the device between the ~100 evaluations of
// a fit.  Fixed-shape two-level reduction (every CTA sums a contiguous slice in a fixed order, the last CTA to finish
// adds the partial sums in index order), so the value does not depend on scheduling.
constexpr int kChiThreads = 256;
constexpr int kChiBlocks = 128;

__global__ void __launch_bounds__(kChiThreads)
expdecay_chi2_kernel(const double* __restrict__ x, const double* __restrict__ y, long long n, double C0, double C1, double A,
                     double* __restrict__ work, double* __restrict__ out) {
  __shared__ double red[kChiThreads / 32];
  __shared__ bool last;
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long a = per * blockIdx.x, b = min(n, a + per);
  double s = 0.0;
  for (long long i = a + threadIdx.x; i < b; i += kChiThreads) {
    const double d = C0 * exp(-x[i] / A) + C1 - y[i];
    s = fma(d, d, s);
  }
  s = sr_warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kChiThreads / 32; ++w) t += red[w];
    work[1 + blockIdx.x] = t;
    __threadfence();
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned int*>(work), 1u);
    last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double t = 0.0;
    for (unsigned i = 0; i < gridDim.x; ++i) t += reinterpret_cast<volatile double*>(work)[1 + i];
    *out = t / (double)n;
    *reinterpret_cast<unsigned int*>(work) = 0u;          // ready for the next evaluation
  }
}

}  // namespace

extern "C" int sr_expdecay_chi2(const double* d_x, const double* d_y, long long n, double C0, double C1, double A,
                                double* d_work, int work_doubles, double* d_out, void* stream) {
  SR_REQUIRE(d_x && d_y && d_work && d_out, "sr_expdecay_chi2: null pointer");
  SR_REQUIRE(n >= 1, "sr_expdecay_chi2: empty curve");
  SR_REQUIRE(work_doubles >= kChiBlocks + 1, "sr_expdecay_chi2: workspace of %d doubles, need %d (zero-initialised)", work_doubles,
             kChiBlocks + 1);
  const int blocks = (int)std::min<long long>(kChiBlocks, (n + kChiThreads - 1) / kChiThreads);
  expdecay_chi2_kernel<<<blocks, kChiThreads, 0, (cudaStream_t)stream>>>(d_x, d_y, n, C0, C1, A, d_work, d_out);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_dq_moments_pooled(const float* d_q, long long N, const long long* d_lags, int nLags, long long min_lag,
                                    int nCh, int replica, int nReplicas, int accumulate, double* d_M, void* stream) {
  SR_REQUIRE(d_q && d_lags && d_M, "sr_dq_moments: null pointer");
  SR_REQUIRE(nReplicas >= 1 && replica >= 0 && replica < nReplicas, "sr_dq_moments: replica %d outside [0, %d)", replica,
             nReplicas);
  SR_REQUIRE(N >= 2 && nLags > 0 && nCh >= 1, "sr_dq_moments: bad shape (N=%lld nLags=%d nCh=%d)", N, nLags, nCh);
  SR_REQUIRE(min_lag >= 1 && min_lag < N, "sr_dq_moments: min_lag %lld outside [1, N)", min_lag);
  const long long tilesMax = (N - min_lag + kDqFrameTile - 1) / kDqFrameTile;
  const long long lagTiles = (nLags + kDqLagTile - 1) / kDqLagTile;
  const long long blocks = tilesMax * lagTiles;
  SR_REQUIRE(blocks < (1LL << 31), "sr_dq_moments: %lld blocks exceed the grid limit; split the lag list", blocks);
  if (!accumulate) SR_CUDA(cudaMemsetAsync(d_M, 0, sizeof(double) * 6 * (size_t)nLags * nCh, (cudaStream_t)stream));
  // a lag list of consecutive integers (all windows: 1, 2, 3, ...) takes the shared-memory kernel for the interior
  // of the (lag, frame) plane; the list is tested on the device so that the call stays asynchronous
  // A lag list of consecutive integers (all windows: 1, 2, 3, ...) takes the shared-memory kernel.  The list lives on
  // the device, so it is tested there and the one-word verdict is read back (the only synchronisation of this call:
  // a few microseconds next to a reduction that runs for milliseconds).
  int not_consecutive = 1;
  if (nLags >= 2 * kDqLagTile) {
    int* flag = dq_flag_slot();
    SR_REQUIRE(flag != nullptr, "sr_dq_moments: cannot allocate device scratch");
    SR_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), (cudaStream_t)stream));
    dq_check_consecutive_kernel<<<(nLags + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_lags, nLags, min_lag, flag);
    SR_CUDA(cudaMemcpyAsync(&not_consecutive, flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    SR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  }
  if (not_consecutive) {
    const int smem = (kDqFrameTile + 16 * kDqU) * (int)sizeof(double4);
    SR_CUDA(cudaFuncSetAttribute(dq_moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dq_moments_kernel<<<(unsigned)blocks, kDqThreads, smem, (cudaStream_t)stream>>>((const float4*)d_q, N, d_lags, nLags,
                                                                                   nCh, (int)tilesMax, replica, nReplicas, d_M);
  } else {
    const long long superMax = (tilesMax + kDqSuper - 1) / kDqSuper;
    const long long cblocks = superMax * lagTiles;
    const int csmem = kDqcSmemDoubles * (int)sizeof(double);
    SR_CUDA(cudaFuncSetAttribute(dq_moments_consec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, csmem));
    dq_moments_consec_kernel<<<(unsigned)cblocks, kDqThreads, csmem, (cudaStream_t)stream>>>(
        (const float4*)d_q, N, d_lags, min_lag, nLags, nCh, (int)superMax, replica, nReplicas, d_M);
  }
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_dq_moments(const float* d_q, long long N, const long long* d_lags, int nLags, long long min_lag,
                             int nCh, double* d_M, void* stream) {
  return sr_dq_moments_pooled(d_q, N, d_lags, nLags, min_lag, nCh, 0, 1, 0, d_M, stream);
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

extern "C" int sr_dq_hist3d(const float* d_q, long long N, long long delta, const double* d_edges, int nb,
                            unsigned int* d_counts, long long* d_amb_idx, int amb_capacity, int* d_amb_count, void* stream) {
  SR_REQUIRE(d_q && d_edges && d_counts && d_amb_idx && d_amb_count, "sr_dq_hist3d: null pointer");
  SR_REQUIRE(delta >= 1 && delta < N, "sr_dq_hist3d: delta %lld outside [1, N)", delta);
  SR_REQUIRE(nb >= 1 && nb <= 1024, "sr_dq_hist3d: %d bins per axis outside [1, 1024]", nb);
  const long long n = N - delta;
  dq_hist3d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)d_q, N, delta, d_edges, nb,
                                                                                d_counts, d_amb_idx, amb_capacity, d_amb_count);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}
