// K3: PAF rotation + Lambert-cylindrical (phi, cos theta) histogram of unit bond vectors.
// Reference arithmetic: calculate-Ct-from-traj.py:567 (rotate_vector_simd), :588 (gm.xyz_to_rtp),
// :600-626 (transpose, cos(theta), np.histogramdd with bins (nbx, nby) over ((-pi,pi),(-1,1))).
//
// Bit-exact counts without reproducing NumPy's libm, in two passes:
//   1. sphere_hist_kernel (hot, FP32): rotates, proposes a bin and accepts it only if the sample is farther
//      than a float32 margin from all four edges of that bin -- phi through the sign of the cross products
//      with the tabulated edge directions, cos(theta) by direct comparison.  Misses (~1e-4 of the samples)
//      are appended to a retry list.
//   2. sphere_hist_resolve_kernel (rare, FP64): re-examines the retry list with the same tests in double
//      precision and a ~1e-11 margin (rotated stream), or drops NaN / zero vectors (float32 reference stream).
//      What is still undecided keeps its sample id in the list for the host, which re-bins those few samples
//      with the reference's own NumPy formula; everything else in the list is overwritten with -1.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

namespace {

constexpr int kHistThreads = 384;   // threads per CTA (2 CTAs per SM: 85 registers per thread)
constexpr int kHistGroup = 16;      // vectors per CTA (4 per thread)
constexpr int kHistFP = kHistThreads / 4;   // frames covered by one pass of the CTA

struct HistParams {
  double R[9];        // rotation matrix (row major), identity when no rotation
  double tol_phi;     // angular margin (rad) of the FP64 test
  double tol_cos;     // cos(theta) margin of the FP64 test
  float Rf[9];        // the same matrix in float32 for the fast path
  float fphi_abs, fphi_rel, fcos;   // fast-path margins (see fast_classify)
  int nbx, nby;
  int f32_reference;  // 1: no rotation, the reference itself works in float32 -> fast-path misses go to the host
};

struct SlowParams { double tol_phi, tol_cos; int nbx, nby; };

__device__ __forceinline__ int classify(double x, double y, double z, const SlowParams p,
                                        const double2* __restrict__ edge_dir,   // (cos e_i, sin e_i), i = 0..nbx
                                        const double* __restrict__ edge_cos,    // e_j, j = 0..nby
                                        int& bin_out) {
  // returns 0 = counted in bin_out, 1 = dropped (NaN / zero vector, as np.histogramdd drops NaN), 2 = ambiguous
  const double r2 = x * x + y * y + z * z;
  if (r2 != r2 || r2 == 0.0) return 1;   // NaN component or zero vector: z/r is NaN, np.histogramdd drops it
  if (!(r2 < 1e300)) return 2;           // overflow / inf: leave to the host formula
  // ---- phi ----
  const double rho1 = fabs(x) + fabs(y);
  if (rho1 == 0.0) return 2;   // atan2(+-0, +-0): let the host apply NumPy's signed-zero rules
  const float phif = atan2f((float)y, (float)x);
  int i0 = (int)floorf((phif + 3.14159265f) * (p.nbx * 0.15915494f));
  i0 = max(0, min(p.nbx - 1, i0));
  const double mphi = rho1 * p.tol_phi;
  int bphi = -1;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int i = i0 + (t == 0 ? 0 : (t == 1 ? -1 : 1));
    if (i < 0 || i >= p.nbx) continue;
    const double2 lo = edge_dir[i], hi = edge_dir[i + 1];
    const double clo = lo.x * y - lo.y * x;   // rho sin(phi - e_i)
    const double chi = hi.x * y - hi.y * x;   // rho sin(phi - e_{i+1})
    if (clo >= mphi && chi <= -mphi) { bphi = i; break; }
  }
  if (bphi < 0) return 2;
  // ---- cos(theta) = z / r : compare s(z/r) = z|z|/r2 with s(e) = e|e| ----
  const float cf = (float)z * rsqrtf((float)r2);
  int j0 = (int)floorf((cf + 1.0f) * (0.5f * p.nby));
  j0 = max(0, min(p.nby - 1, j0));
  const double zs = z * fabs(z);
  int bcos = -1;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int j = j0 + (t == 0 ? 0 : (t == 1 ? -1 : 1));
    if (j < 0 || j >= p.nby) continue;
    const double elo = edge_cos[j], ehi = edge_cos[j + 1];
    const double mlo = (2.0 * fabs(elo) * p.tol_cos + p.tol_cos * p.tol_cos) * r2;
    const double mhi = (2.0 * fabs(ehi) * p.tol_cos + p.tol_cos * p.tol_cos) * r2;
    if (zs >= elo * fabs(elo) * r2 + mlo && zs <= ehi * fabs(ehi) * r2 - mhi) { bcos = j; break; }
  }
  if (bcos < 0) return 2;
  bin_out = bphi * p.nby + bcos;
  return 0;
}

// FP32 fast path margins.  fphi_abs covers the float32 rounding of the rotated coordinates (absolute, rotated
// path); fphi_rel / fcos the reference's own float32 arctan2/arccos/cos error (unrotated path).
//
// Hot kernel.  grid.x = vector groups (fastest, so CTAs sharing a frame range run together and the 64-byte
// DRAM blocks that straddle two groups are fetched once), grid.y = frame blocks.  A CTA owns kHistGroup
// vectors; thread t works on the four vectors 4*(t&3) .. +3 of frame t>>2 (+128 per pass), i.e. on 48
// contiguous bytes of the reference's AoS layout -- three LDG.128 when a frame row is a multiple of 16 bytes
// (nR % 4 == 0), twelve scalar loads otherwise -- and always has the next pass's 48 bytes in flight while it
// classifies the current ones.  Bins are privatised in shared memory as packed 16-bit counters (an even
// number of bins per vector; a CTA sees < 65536 frames) and flushed with global atomics at the end.
template <bool VEC4>
__global__ void __launch_bounds__(kHistThreads, 2)
sphere_hist_kernel(const float* __restrict__ vecs, long long nFrames, int nR, int framesPerBlock,
                   HistParams p, const double2* __restrict__ edge_dir,
                   unsigned int* __restrict__ counts, long long* __restrict__ amb_idx, int amb_capacity,
                   int* __restrict__ amb_count) {
  extern __shared__ __align__(16) unsigned int sh_raw[];
  float4* const sh_edge = reinterpret_cast<float4*>(sh_raw);
  unsigned int* const sh_hist = sh_raw + 4 * p.nbx;
  const int nbx = p.nbx, nby = p.nby;
  const int nbins = nbx * nby;
  const int wordsPerVec = (nbins + 1) >> 1;
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * kHistGroup;
  const int nv = min(kHistGroup, nR - r0);
  for (int i = tid; i < nv * wordsPerVec; i += kHistThreads) sh_hist[i] = 0u;
  for (int i = tid; i < nbx; i += kHistThreads) {
    const double2 lo = edge_dir[i], hi = edge_dir[i + 1];
    sh_edge[i] = make_float4((float)lo.x, (float)lo.y, (float)hi.x, (float)hi.y);
  }
  __syncthreads();

  const long long f0 = (long long)blockIdx.y * framesPerBlock;
  const int nfl = (int)min((long long)framesPerBlock, nFrames - f0);
  const int quad = tid & 3;
  const int nvq = max(0, min(4, nv - 4 * quad));            // valid vectors of this thread
  int fl = tid >> 2;
  const uint32_t edge_addr = sr_smem_u32(sh_edge);
  const uint32_t hist_addr = sr_smem_u32(sh_hist + 4 * quad * wordsPerVec);
  const uint32_t hist_step = (uint32_t)wordsPerVec * 4u;
  const float q0 = p.Rf[0], q1 = p.Rf[1], q2 = p.Rf[2], q3 = p.Rf[3], q4 = p.Rf[4], q5 = p.Rf[5], q6 = p.Rf[6],
              q7 = p.Rf[7], q8 = p.Rf[8];
  const float fphi_abs = p.fphi_abs, fphi_rel = p.fphi_rel, fcos = p.fcos;
  const float phi_scale = nbx * 0.159154943f, cos_scale = 0.5f * nby, wbin = 2.0f / nby;
  const size_t passStride = (size_t)kHistFP * nR * 3;
  const float* src = vecs + ((size_t)(f0 + fl) * nR + r0 + 4 * quad) * 3;
  long long sidx = (f0 + fl) * nR + r0 + 4 * quad;          // sample id of this thread's first vector
  const long long sidxStep = (long long)kHistFP * nR;

  auto load = [&](float (&d)[12], const float* s) {
    if (VEC4) {
      const float4* s4 = reinterpret_cast<const float4*>(s);
      const float4 a = __ldg(s4), b = __ldg(s4 + 1), c = __ldg(s4 + 2);
      d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
      d[8] = c.x; d[9] = c.y; d[10] = c.z; d[11] = c.w;
    } else {
#pragma unroll
      for (int i = 0; i < 12; ++i) d[i] = (i < 3 * nvq) ? __ldg(s + i) : 0.f;
    }
  };
  auto process = [&](const float (&d)[12], long long sid) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float vx = d[3 * u], vy = d[3 * u + 1], vz = d[3 * u + 2];
      const float x = fmaf(q0, vx, fmaf(q1, vy, q2 * vz));
      const float y = fmaf(q3, vx, fmaf(q4, vy, q5 * vz));
      const float z = fmaf(q6, vx, fmaf(q7, vy, q8 * vz));
      // NaN, zero, huge or polar vectors need no explicit test: they fail the margin comparisons below
      // (rsqrtf(0) = inf -> cf = NaN; rho1 = 0 -> |cross| = 0 < mphi) and go to the retry list.
      const float r2 = fmaf(x, x, fmaf(y, y, z * z));
      const float ax = fabsf(x), ay = fabsf(y);
      const float rho1 = ax + ay;
      // phi candidate from a degree-11 odd minimax polynomial of atan on [0,1] (max error 1.8e-6 rad)
      const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
      const float t = __fdividef(mn, mx), t2 = t * t;
      float a = fmaf(t2, -0.01171912346035242f, 0.052647337317466736f);
      a = fmaf(t2, a, -0.1164264902472496f);
      a = fmaf(t2, a, 0.19354039430618286f);
      a = fmaf(t2, a, -0.33262282609939575f);
      a = fmaf(t2, a, 0.9999772310256958f) * t;
      if (ay > ax) a = 1.57079632679f - a;
      if (x < 0.f) a = 3.14159265359f - a;
      if (y < 0.f) a = -a;
      int i = (int)floorf(fmaf(a, phi_scale, 3.14159265359f * phi_scale));
      i = max(0, min(nbx - 1, i));
      float4 e;   // (cos e_i, sin e_i, cos e_{i+1}, sin e_{i+1})
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(e.x), "=f"(e.y), "=f"(e.z), "=f"(e.w) : "r"(edge_addr + i * 16));
      const float mphi = fmaf(fphi_rel, rho1, fphi_abs * (rho1 + fabsf(z))) + 1e-30f;
      const float clo = fmaf(e.x, y, -e.y * x);   // rho sin(phi - e_i)
      const float chi = fmaf(e.z, y, -e.w * x);   // rho sin(phi - e_{i+1})
      const float cf = z * rsqrtf(r2);
      int j = (int)floorf(fmaf(cf, cos_scale, cos_scale));
      j = max(0, min(nby - 1, j));
      const float elo = fmaf((float)j, wbin, -1.0f), ehi = elo + wbin;
      const bool ok = (clo >= mphi) && (chi <= -mphi) && (cf - elo >= fcos) && (ehi - cf >= fcos);   // NaN compares false
      const int bin = i * nby + j;
      if (u < nvq) {
        if (ok) {
          asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hist_addr + u * hist_step + ((bin >> 1) << 2)),
                       "r"(1u << ((bin & 1) << 4)) : "memory");
        } else {   // rare: hand the sample to sphere_hist_resolve_kernel
          const int slot = atomicAdd(amb_count, 1);
          if (slot < amb_capacity) amb_idx[slot] = sid + u;
        }
      }
    }
  };

  if (nvq > 0) {
    float A[12], B[12];
    if (fl < nfl) load(A, src);
    while (fl < nfl) {
      const bool moreB = fl + kHistFP < nfl;
      if (moreB) load(B, src + passStride);
      process(A, sidx);
      if (!moreB) break;
      const bool moreA = fl + 2 * kHistFP < nfl;
      if (moreA) load(A, src + 2 * passStride);
      process(B, sidx + sidxStep);
      fl += 2 * kHistFP; src += 2 * passStride; sidx += 2 * sidxStep;
    }
  }
  __syncthreads();
  for (int i = tid; i < nv * wordsPerVec; i += kHistThreads) {
    const unsigned int w = sh_hist[i];
    if (w == 0u) continue;
    const int v = i / wordsPerVec, k = i - v * wordsPerVec;
    const long long g = (long long)(r0 + v) * nbins + 2 * k;
    if (w & 0xffffu) atomicAdd(&counts[g], w & 0xffffu);
    if (w >> 16) atomicAdd(&counts[g + 1], w >> 16);
  }
}

// remembers where this call's retry entries start (the list may already hold entries of earlier calls)
__global__ void sphere_hist_mark_kernel(const int* __restrict__ amb_count, int amb_capacity, int* __restrict__ amb_start) {
  *amb_start = min(*amb_count, amb_capacity);
}

// Second pass over the fast path's misses (typically ~1e-4 of the samples): FP64 re-examination for the rotated
// stream, NaN / zero-vector drop for the float32 reference stream.  Resolved entries are overwritten with -1,
// entries that stay ambiguous keep their sample id for the host tie-break (hist.py::_reference_bins).
__global__ void __launch_bounds__(256)
sphere_hist_resolve_kernel(const float* __restrict__ vecs, int nR, HistParams p, const double2* __restrict__ edge_dir,
                           const double* __restrict__ edge_cos, unsigned int* __restrict__ counts,
                           long long* __restrict__ amb_idx, int amb_capacity, const int* __restrict__ amb_count,
                           const int* __restrict__ amb_start) {
  // only the entries appended by THIS call's hot pass: [*amb_start, *amb_count)
  const int n = min(*amb_count, amb_capacity);
  const int nbins = p.nbx * p.nby;
  for (int e = *amb_start + blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const long long sidx = amb_idx[e];
    if (sidx < 0) continue;
    const float* v = vecs + sidx * 3;
    const float vx = v[0], vy = v[1], vz = v[2];
    int bin = 0, cls;
    if (p.f32_reference) {
      const float n2 = vx * vx + vy * vy + vz * vz;
      cls = (n2 != n2 || n2 == 0.f) ? 1 : 2;
    } else {
      const double dx = vx, dy = vy, dz = vz;
      SlowParams sp; sp.tol_phi = p.tol_phi; sp.tol_cos = p.tol_cos; sp.nbx = p.nbx; sp.nby = p.nby;
      cls = classify(p.R[0] * dx + p.R[1] * dy + p.R[2] * dz, p.R[3] * dx + p.R[4] * dy + p.R[5] * dz,
                     p.R[6] * dx + p.R[7] * dy + p.R[8] * dz, sp, edge_dir, edge_cos, bin);
    }
    if (cls == 0) atomicAdd(&counts[(sidx % nR) * nbins + bin], 1u);
    if (cls != 2) amb_idx[e] = -1;
  }
}

}  // namespace

extern "C" int sr_sphere_hist_table_doubles(int nbx, int nby) { return 2 * (nbx + 1) + (nby + 1); }

extern "C" int sr_sphere_hist(const float* d_vecs, long long nFrames, int nR, const double* h_q_rot, int nbx, int nby,
                              const double* d_edge_table, double tol_phi, double tol_cos, unsigned int* d_counts,
                              long long* d_amb_idx, int amb_capacity, int* d_amb_count, void* stream) {
  SR_REQUIRE(d_vecs && d_edge_table && d_counts && d_amb_idx && d_amb_count, "sr_sphere_hist: null pointer");
  SR_REQUIRE(nFrames > 0 && nR > 0 && nbx > 0 && nby > 0, "sr_sphere_hist: empty shape");
  SR_REQUIRE(tol_phi > 0 && tol_cos > 0, "sr_sphere_hist: tolerances must be positive");
  SR_REQUIRE(amb_capacity > 0, "sr_sphere_hist: the retry / ambiguous list needs a positive capacity");
  HistParams p;
  p.nbx = nbx; p.nby = nby; p.tol_phi = tol_phi; p.tol_cos = tol_cos;
  double q[4] = {1, 0, 0, 0};
  if (h_q_rot) {
    const double n = sqrt(h_q_rot[0] * h_q_rot[0] + h_q_rot[1] * h_q_rot[1] + h_q_rot[2] * h_q_rot[2] +
                          h_q_rot[3] * h_q_rot[3]);
    SR_REQUIRE(n > 0, "sr_sphere_hist: zero rotation quaternion");
    for (int i = 0; i < 4; ++i) q[i] = h_q_rot[i] / n;
  }
  {  // rotation matrix of v -> v + 2 q_v x (q_v x v + q_w v)
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    p.R[0] = 1 - 2 * (y * y + z * z); p.R[1] = 2 * (x * y - w * z);     p.R[2] = 2 * (x * z + w * y);
    p.R[3] = 2 * (x * y + w * z);     p.R[4] = 1 - 2 * (x * x + z * z); p.R[5] = 2 * (y * z - w * x);
    p.R[6] = 2 * (x * z - w * y);     p.R[7] = 2 * (y * z + w * x);     p.R[8] = 1 - 2 * (x * x + y * y);
  }
  for (int i = 0; i < 9; ++i) p.Rf[i] = (float)p.R[i];
  p.f32_reference = h_q_rot ? 0 : 1;
  if (h_q_rot) {   // float32 rounding of the rotated coordinates: ~3e-7 |v| absolute
    p.fphi_abs = 1.5e-6f; p.fphi_rel = 0.f; p.fcos = 2.0e-6f;
  } else {         // the caller's tolerances describe the reference's own float32 error
    p.fphi_abs = 1e-12f; p.fphi_rel = (float)tol_phi; p.fcos = (float)tol_cos;
  }
  const int nbins = nbx * nby;
  int dev = 0, max_smem = 0, sms = 0;
  SR_CUDA(cudaGetDevice(&dev));
  SR_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  SR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t wordsPerVec = ((size_t)nbins + 1) / 2;
  const size_t smem = (size_t)kHistGroup * wordsPerVec * 4 + (size_t)nbx * 16;
  SR_REQUIRE(smem <= (size_t)max_smem, "sr_sphere_hist: %d bins do not fit in shared memory", nbins);
  const int nGroups = (nR + kHistGroup - 1) / kHistGroup;
  // one wave of 2 CTAs per SM; a CTA must see fewer than 65536 frames (16-bit privatised counters)
  long long nFB = std::max(1LL, (2LL * sms) / nGroups);
  long long fpb = (nFrames + nFB - 1) / nFB;
  fpb = std::min(std::max(fpb, (long long)kHistFP), 32768LL);
  nFB = (nFrames + fpb - 1) / fpb;
  SR_REQUIRE(nFB <= 65535, "sr_sphere_hist: %lld frame blocks exceed the grid limit", nFB);
  dim3 grid((unsigned)nGroups, (unsigned)nFB);
  const double2* edge_dir = (const double2*)d_edge_table;
  const double* edge_cos = d_edge_table + 2 * (nbx + 1);
  cudaStream_t st = (cudaStream_t)stream;
  // one int of device scratch per call in flight: a small ring allocated once per device.  (cudaMallocAsync is not
  // an option here: releasing the pool at the caller's next synchronisation costs ~0.4 s next to torch's allocator.)
  static std::mutex ring_mu;
  static int* ring[64] = {nullptr};
  static unsigned ring_next[64] = {0};
  int* d_amb_start = nullptr;
  {
    std::lock_guard<std::mutex> lock(ring_mu);
    SR_REQUIRE(dev >= 0 && dev < 64, "sr_sphere_hist: device ordinal %d not supported", dev);
    if (!ring[dev]) SR_CUDA(cudaMalloc(&ring[dev], 256 * sizeof(int)));
    d_amb_start = ring[dev] + (ring_next[dev]++ & 255u);
  }
  sphere_hist_mark_kernel<<<1, 1, 0, st>>>(d_amb_count, amb_capacity, d_amb_start);
  if (nR % 4 == 0 && ((uintptr_t)d_vecs & 15) == 0) {
    SR_CUDA(cudaFuncSetAttribute(sphere_hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sphere_hist_kernel<true><<<grid, kHistThreads, smem, st>>>(d_vecs, nFrames, nR, (int)fpb, p, edge_dir, d_counts,
                                                              d_amb_idx, amb_capacity, d_amb_count);
  } else {
    SR_CUDA(cudaFuncSetAttribute(sphere_hist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sphere_hist_kernel<false><<<grid, kHistThreads, smem, st>>>(d_vecs, nFrames, nR, (int)fpb, p, edge_dir, d_counts,
                                                               d_amb_idx, amb_capacity, d_amb_count);
  }
  SR_CUDA(cudaGetLastError());
  sphere_hist_resolve_kernel<<<sms, 256, 0, st>>>(d_vecs, nR, p, edge_dir, edge_cos, d_counts, d_amb_idx, amb_capacity,
                                                  d_amb_count, d_amb_start);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

// Host-buffer entry point: builds NumPy's edges (np.linspace(-pi, pi, nbx + 1), np.linspace(-1, 1, nby + 1)) and the
// edge-direction table itself, uploads the vectors, runs both passes and returns the counts plus the sample ids
// the device left undecided (what a caller without NumPy cannot reproduce bit for bit is exactly that short list).
extern "C" int sr_sphere_hist_host(const float* h_vecs, long long nFrames, int nR, const double* h_q_rot, int nbx, int nby,
                                   unsigned int* h_counts, long long* h_amb_idx, int amb_capacity, int* h_n_amb) {
  SR_REQUIRE(h_vecs && h_counts && h_amb_idx && h_n_amb, "sr_sphere_hist_host: null pointer");
  SR_REQUIRE(nFrames > 0 && nR > 0 && nbx > 0 && nby > 0 && amb_capacity > 0, "sr_sphere_hist_host: empty shape");
  const double kPi = 3.141592653589793;
  std::vector<double> table((size_t)2 * (nbx + 1) + (nby + 1));
  for (int i = 0; i <= nbx; ++i) {           // np.linspace: start + i * step, last point = stop
    const double e = (i == nbx) ? kPi : -kPi + i * ((kPi - (-kPi)) / nbx);
    table[2 * i] = cos(e); table[2 * i + 1] = sin(e);
  }
  for (int j = 0; j <= nby; ++j) table[2 * (nbx + 1) + j] = (j == nby) ? 1.0 : -1.0 + j * (2.0 / nby);
  const size_t in_bytes = (size_t)nFrames * nR * 3 * sizeof(float);
  const size_t cnt_bytes = (size_t)nR * nbx * nby * sizeof(unsigned int);
  float* d_in = nullptr; double* d_table = nullptr; unsigned int* d_counts = nullptr; long long* d_amb = nullptr; int* d_namb = nullptr;
  int rc = SR_OK;
  cudaError_t e = cudaSuccess;
  if ((e = cudaMalloc(&d_in, in_bytes)) != cudaSuccess || (e = cudaMalloc(&d_table, table.size() * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(&d_counts, cnt_bytes)) != cudaSuccess || (e = cudaMalloc(&d_amb, (size_t)amb_capacity * sizeof(long long))) != cudaSuccess ||
      (e = cudaMalloc(&d_namb, sizeof(int))) != cudaSuccess ||
      (e = cudaMemcpy(d_in, h_vecs, in_bytes, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(d_table, table.data(), table.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemset(d_counts, 0, cnt_bytes)) != cudaSuccess || (e = cudaMemset(d_namb, 0, sizeof(int))) != cudaSuccess) {
    sr_set_error("sr_sphere_hist_host: allocation / upload failed: %s", cudaGetErrorString(e));
    rc = SR_ERR_CUDA;
  }
  // margins: FP64 resolve pass for the rotated stream; the reference's own float32 error for the unrotated one
  const double tol_phi = h_q_rot ? 1e-11 : 4e-6, tol_cos = h_q_rot ? 1e-11 : 2e-6;
  if (!rc) rc = sr_sphere_hist(d_in, nFrames, nR, h_q_rot, nbx, nby, d_table, tol_phi, tol_cos, d_counts, d_amb, amb_capacity,
                               d_namb, nullptr);
  int n = 0;
  std::vector<long long> list;
  if (!rc && ((e = cudaMemcpy(h_counts, d_counts, cnt_bytes, cudaMemcpyDeviceToHost)) != cudaSuccess ||
              (e = cudaMemcpy(&n, d_namb, sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess)) {
    sr_set_error("sr_sphere_hist_host: download failed: %s", cudaGetErrorString(e));
    rc = SR_ERR_CUDA;
  }
  if (!rc && n > amb_capacity) {
    sr_set_error("sr_sphere_hist_host: %d retry samples exceed the list capacity %d", n, amb_capacity);
    rc = SR_ERR_OVERFLOW;
  }
  if (!rc && n > 0) {
    list.resize(n);
    if ((e = cudaMemcpy(list.data(), d_amb, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost)) != cudaSuccess) {
      sr_set_error("sr_sphere_hist_host: download failed: %s", cudaGetErrorString(e));
      rc = SR_ERR_CUDA;
    }
  }
  if (!rc) {
    int k = 0;
    for (int i = 0; i < n; ++i)
      if (list[i] >= 0) h_amb_idx[k++] = list[i];      // -1 = resolved on the device
    *h_n_amb = k;
  }
  cudaFree(d_in); cudaFree(d_table); cudaFree(d_counts); cudaFree(d_amb); cudaFree(d_namb);
  return rc;
}
