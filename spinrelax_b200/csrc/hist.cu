// K3: PAF rotation + Lambert-cylindrical (phi, cos theta) histogram of unit bond vectors.
// Reference arithmetic: calculate-Ct-from-traj.py:567 (rotate_vector_simd), :588 (gm.xyz_to_rtp),
// :600-626 (transpose, cos(theta), np.histogramdd with bins (nbx, nby) over ((-pi,pi),(-1,1))).
//
// Bit-exact counts without reproducing NumPy's libm: the kernel decides a bin only when the sample is
// provably farther than `tol` from every bin edge -- phi through the sign of the cross product with the
// tabulated edge directions, cos(theta) by comparing z|z| with e|e| r^2 -- and otherwise appends the
// sample index to an "ambiguous" list that the host re-bins with the reference's own NumPy formula.
// tol is ~1e-11 for the float64 (rotated) path and a few float32 ulps for the unrotated float32 path.
#include "common.cuh"

namespace {

constexpr int kHistThreads = 512;

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

// FP32 fast path.  Returns the flat bin, or -1 when the sample is within the fast-path margin of a bin edge
// (or degenerate) and has to be re-examined.  The margin covers the float32 rounding of the rotated
// coordinates (fphi_abs, absolute, rotated path) or the reference's own float32 arctan2/arccos/cos error
// (fphi_rel, relative to the xy-projection, unrotated path).
struct FastParams { float fphi_abs, fphi_rel, fcos; int nbx, nby; };

__device__ __forceinline__ int fast_classify(float x, float y, float z, const FastParams p,
                                             const float4* __restrict__ sh_edge /* (cos_i, sin_i, cos_i+1, sin_i+1) */) {
  // NaN, zero, huge or polar vectors need no explicit test: they fail the margin comparisons below
  // (rsqrtf(0) = inf -> cf = NaN; rho1 = 0 -> |cross| = 0 < mphi) and drop to the slow path.
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
  int i = (int)floorf((a + 3.14159265359f) * (p.nbx * 0.159154943f));
  i = max(0, min(p.nbx - 1, i));
  const float4 e = sh_edge[i];
  const float mphi = fmaf(p.fphi_rel, rho1, p.fphi_abs * (rho1 + fabsf(z))) + 1e-30f;
  const float clo = fmaf(e.x, y, -e.y * x);   // rho sin(phi - e_i)
  const float chi = fmaf(e.z, y, -e.w * x);   // rho sin(phi - e_{i+1})
  const bool ok_phi = (clo >= mphi) && (chi <= -mphi);
  // cos(theta)
  const float cf = z * rsqrtf(r2);
  const float wbin = 2.0f / p.nby;
  int j = (int)floorf((cf + 1.0f) * (0.5f * p.nby));
  j = max(0, min(p.nby - 1, j));
  const float elo = fmaf((float)j, wbin, -1.0f), ehi = elo + wbin;
  const bool ok_cos = (cf - elo >= p.fcos) && (ehi - cf >= p.fcos);   // NaN compares false
  return (ok_phi && ok_cos) ? i * p.nby + j : -1;
}

// Rare path, kept out of line so the hot loop stays small: FP64 re-examination (rotated stream) or hand-over
// to the host (float32 reference stream), NaN / zero vectors dropped as np.histogramdd drops them.
__device__ __noinline__ void slow_sample(float vx, float vy, float vz, int f32_reference, double r0, double r1, double r2,
                                         double r3, double r4, double r5, double r6, double r7, double r8,
                                         double tol_phi, double tol_cos, int nbx, int nby,
                                         const double2* __restrict__ edge_dir, const double* __restrict__ edge_cos,
                                         unsigned int* sh_vec_hist, long long sample_idx, long long* __restrict__ amb_idx,
                                         int amb_capacity, int* __restrict__ amb_count) {
  int bin = 0, cls;
  if (f32_reference) {
    const float n2 = vx * vx + vy * vy + vz * vz;
    cls = (n2 != n2 || n2 == 0.f) ? 1 : 2;
  } else {
    const double dx = vx, dy = vy, dz = vz;
    SlowParams sp; sp.tol_phi = tol_phi; sp.tol_cos = tol_cos; sp.nbx = nbx; sp.nby = nby;
    cls = classify(r0 * dx + r1 * dy + r2 * dz, r3 * dx + r4 * dy + r5 * dz, r6 * dx + r7 * dy + r8 * dz, sp, edge_dir,
                   edge_cos, bin);
  }
  if (cls == 0) {
    atomicAdd(&sh_vec_hist[bin >> 1], 1u << ((bin & 1) << 4));
  } else if (cls == 2) {
    const int slot = atomicAdd(amb_count, 1);
    if (slot < amb_capacity) amb_idx[slot] = sample_idx;
  }
}

constexpr int kHistU = 4;

// grid.x = vector groups (fastest, so CTAs sharing a frame range run together and the 64-byte DRAM blocks
// that straddle two groups are fetched once), grid.y = frame blocks.  A CTA owns `group` = 1 << gshift
// vectors; thread t always works on vector t & (group-1), so its shared-memory histogram base and its global
// pointer stride are loop invariants.  Bins are privatised in shared memory as packed 16-bit counters
// (an even number of bins per vector; a CTA sees < 65536 frames).
__global__ void __launch_bounds__(kHistThreads, 2)
sphere_hist_kernel(const float* __restrict__ vecs, long long nFrames, int nR, int gshift, int framesPerBlock,
                   HistParams p, const double2* __restrict__ edge_dir, const double* __restrict__ edge_cos,
                   unsigned int* __restrict__ counts, long long* __restrict__ amb_idx, int amb_capacity,
                   int* __restrict__ amb_count) {
  extern __shared__ unsigned int sh_raw[];
  float4* sh_edge = reinterpret_cast<float4*>(sh_raw);
  unsigned int* sh_hist = sh_raw + 4 * p.nbx;
  const int nbins = p.nbx * p.nby;
  const int wordsPerVec = (nbins + 1) >> 1;
  const int group = 1 << gshift;
  const int r0 = blockIdx.x * group;
  const int nv = min(group, nR - r0);
  for (int i = threadIdx.x; i < nv * wordsPerVec; i += kHistThreads) sh_hist[i] = 0u;
  for (int i = threadIdx.x; i < p.nbx; i += kHistThreads) {
    const double2 lo = edge_dir[i], hi = edge_dir[i + 1];
    sh_edge[i] = make_float4((float)lo.x, (float)lo.y, (float)hi.x, (float)hi.y);
  }
  __syncthreads();

  const long long f0 = (long long)blockIdx.y * framesPerBlock;
  const int nfl = (int)min((long long)framesPerBlock, nFrames - f0);
  const int vl = threadIdx.x & (group - 1);
  const int FP = kHistThreads >> gshift;          // frames covered by one pass of the CTA
  const size_t rowStride = (size_t)nR * 3;
  if (vl < nv) {
    unsigned int* const myhist = sh_hist + vl * wordsPerVec;
    const long long sidx0 = f0 * nR + r0 + vl;
    int fl = threadIdx.x >> gshift;
    const float* src = vecs + ((size_t)(f0 + fl) * nR + r0 + vl) * 3;
    FastParams fp;
    fp.fphi_abs = p.fphi_abs; fp.fphi_rel = p.fphi_rel; fp.fcos = p.fcos; fp.nbx = p.nbx; fp.nby = p.nby;
    const float q0 = p.Rf[0], q1 = p.Rf[1], q2 = p.Rf[2], q3 = p.Rf[3], q4 = p.Rf[4], q5 = p.Rf[5], q6 = p.Rf[6],
                q7 = p.Rf[7], q8 = p.Rf[8];
    auto fast = [&](float vx, float vy, float vz) {
      const float x = fmaf(q0, vx, fmaf(q1, vy, q2 * vz));
      const float y = fmaf(q3, vx, fmaf(q4, vy, q5 * vz));
      const float z = fmaf(q6, vx, fmaf(q7, vy, q8 * vz));
      return fast_classify(x, y, z, fp, sh_edge);
    };
    const size_t stepU = (size_t)FP * rowStride;
    for (; fl + (kHistU - 1) * FP < nfl; fl += kHistU * FP) {
      float vx[kHistU], vy[kHistU], vz[kHistU];
      int bin[kHistU];
#pragma unroll
      for (int u = 0; u < kHistU; ++u) {
        vx[u] = __ldg(src); vy[u] = __ldg(src + 1); vz[u] = __ldg(src + 2);
        src += stepU;
      }
      int worst = 0;
#pragma unroll
      for (int u = 0; u < kHistU; ++u) { bin[u] = fast(vx[u], vy[u], vz[u]); worst = min(worst, bin[u]); }
#pragma unroll
      for (int u = 0; u < kHistU; ++u)
        if (bin[u] >= 0) atomicAdd(&myhist[bin[u] >> 1], 1u << ((bin[u] & 1) << 4));
      if (worst < 0) {
#pragma unroll      // static indices keep vx/vy/vz/bin in registers
        for (int u = 0; u < kHistU; ++u)
          if (bin[u] < 0)
            slow_sample(vx[u], vy[u], vz[u], p.f32_reference, p.R[0], p.R[1], p.R[2], p.R[3], p.R[4], p.R[5], p.R[6], p.R[7],
                        p.R[8], p.tol_phi, p.tol_cos, p.nbx, p.nby, edge_dir, edge_cos, myhist,
                        sidx0 + (long long)(fl + u * FP) * nR, amb_idx, amb_capacity, amb_count);
      }
    }
    for (; fl < nfl; fl += FP, src += stepU) {
      const float vx = __ldg(src), vy = __ldg(src + 1), vz = __ldg(src + 2);
      const int bin = fast(vx, vy, vz);
      if (bin >= 0) atomicAdd(&myhist[bin >> 1], 1u << ((bin & 1) << 4));
      else slow_sample(vx, vy, vz, p.f32_reference, p.R[0], p.R[1], p.R[2], p.R[3], p.R[4], p.R[5], p.R[6], p.R[7], p.R[8],
                       p.tol_phi, p.tol_cos, p.nbx, p.nby, edge_dir, edge_cos, myhist, sidx0 + (long long)fl * nR, amb_idx,
                       amb_capacity, amb_count);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nv * wordsPerVec; i += kHistThreads) {
    const unsigned int w = sh_hist[i];
    const int v = i / wordsPerVec, k = i - v * wordsPerVec;
    const long long g = (long long)(r0 + v) * nbins + 2 * k;
    if (w & 0xffffu) atomicAdd(&counts[g], w & 0xffffu);
    if (w >> 16) atomicAdd(&counts[g + 1], w >> 16);
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
  // two CTAs per SM: at most ~100 KB of privatised 16-bit bins each; group is a power of two <= 16
  int gshift = 4;
  const size_t wordsPerVec = ((size_t)nbins + 1) / 2;
  while (gshift > 0 && ((size_t)(1 << gshift) * wordsPerVec) * 4 > (size_t)100 * 1024) --gshift;
  while (gshift > 0 && (1 << (gshift - 1)) >= nR) --gshift;
  const int group = 1 << gshift;
  const size_t smem = (size_t)group * wordsPerVec * 4 + (size_t)nbx * 16;
  SR_REQUIRE(smem <= (size_t)max_smem, "sr_sphere_hist: %d bins do not fit in shared memory", nbins);
  SR_CUDA(cudaFuncSetAttribute(sphere_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nGroups = (nR + group - 1) / group;
  long long nFB = (6LL * 2 * sms + nGroups - 1) / nGroups;     // ~6 waves of 2 CTAs/SM
  long long fpb = (nFrames + nFB - 1) / nFB;
  if (fpb < 64) fpb = 64;
  if (fpb > 32768) fpb = 32768;                                 // 16-bit counters: fewer than 65536 frames per CTA
  nFB = (nFrames + fpb - 1) / fpb;
  SR_REQUIRE(nFB <= 65535, "sr_sphere_hist: %lld frame blocks exceed the grid limit", nFB);
  dim3 grid((unsigned)nGroups, (unsigned)nFB);
  const double2* edge_dir = (const double2*)d_edge_table;
  const double* edge_cos = d_edge_table + 2 * (nbx + 1);
  sphere_hist_kernel<<<grid, kHistThreads, smem, (cudaStream_t)stream>>>(d_vecs, nFrames, nR, gshift, (int)fpb, p, edge_dir,
                                                                         edge_cos, d_counts, d_amb_idx, amb_capacity,
                                                                         d_amb_count);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}
