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

constexpr int kHistThreads = 256;

struct HistParams {
  double R[9];        // rotation matrix (row major), identity when no rotation
  double tol_phi;     // angular margin (rad)
  double tol_cos;     // cos(theta) margin
  int nbx, nby;
};

__device__ __forceinline__ int classify(double x, double y, double z, const HistParams& p,
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

// grid.x = vector groups (fastest, so CTAs sharing a frame range run together and the 64-byte DRAM
// blocks that straddle two groups are fetched once), grid.y = frame blocks.  `group` = 1 << gshift vectors
// per CTA; bins are privatised in shared memory as packed 16-bit counters (a CTA sees < 65536 frames).
__global__ void __launch_bounds__(kHistThreads)
sphere_hist_kernel(const float* __restrict__ vecs, long long nFrames, int nR, int gshift, int framesPerBlock,
                   HistParams p, const double2* __restrict__ edge_dir, const double* __restrict__ edge_cos,
                   unsigned int* __restrict__ counts, long long* __restrict__ amb_idx, int amb_capacity,
                   int* __restrict__ amb_count) {
  extern __shared__ unsigned int sh_hist[];
  const int nbins = p.nbx * p.nby;
  const int group = 1 << gshift;
  const int r0 = blockIdx.x * group;
  const int nv = min(group, nR - r0);
  const int nwords = (nv * nbins + 1) >> 1;
  for (int i = threadIdx.x; i < nwords; i += kHistThreads) sh_hist[i] = 0u;
  __syncthreads();

  const long long f0 = (long long)blockIdx.y * framesPerBlock;
  const int nfl = (int)min((long long)framesPerBlock, nFrames - f0);
  const int nSamp = nfl << gshift;
  const float* base = vecs + (f0 * nR + r0) * 3;
  const int rowStride = nR * 3;
  constexpr int U = 4;
  for (int s0 = threadIdx.x; s0 < nSamp; s0 += kHistThreads * U) {
    float vx[U], vy[U], vz[U];
    bool on[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int s = s0 + u * kHistThreads;
      const int vl = s & (group - 1), fl = s >> gshift;
      on[u] = (s < nSamp) && (vl < nv);
      vx[u] = vy[u] = vz[u] = 0.f;
      if (on[u]) {
        const float* src = base + (long long)fl * rowStride + vl * 3;
        vx[u] = __ldg(src); vy[u] = __ldg(src + 1); vz[u] = __ldg(src + 2);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!on[u]) continue;
      const int s = s0 + u * kHistThreads;
      const int vl = s & (group - 1), fl = s >> gshift;
      const double dx = vx[u], dy = vy[u], dz = vz[u];
      const double x = p.R[0] * dx + p.R[1] * dy + p.R[2] * dz;
      const double y = p.R[3] * dx + p.R[4] * dy + p.R[5] * dz;
      const double z = p.R[6] * dx + p.R[7] * dy + p.R[8] * dz;
      int bin = 0;
      const int cls = classify(x, y, z, p, edge_dir, edge_cos, bin);
      if (cls == 0) {
        const int k = vl * nbins + bin;
        atomicAdd(&sh_hist[k >> 1], 1u << ((k & 1) << 4));
      } else if (cls == 2) {
        const int slot = atomicAdd(amb_count, 1);
        if (slot < amb_capacity) amb_idx[slot] = (f0 + fl) * nR + r0 + vl;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nwords; i += kHistThreads) {
    const unsigned int w = sh_hist[i];
    if (w & 0xffffu) atomicAdd(&counts[(long long)r0 * nbins + 2 * i], w & 0xffffu);
    if (w >> 16) atomicAdd(&counts[(long long)r0 * nbins + 2 * i + 1], w >> 16);
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
  const int nbins = nbx * nby;
  int dev = 0, max_smem = 0, sms = 0;
  SR_CUDA(cudaGetDevice(&dev));
  SR_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  SR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // two CTAs per SM: at most ~100 KB of privatised 16-bit bins each; group is a power of two <= 16
  int gshift = 4;
  while (gshift > 0 && (((size_t)(1 << gshift) * nbins + 1) / 2) * 4 > (size_t)100 * 1024) --gshift;
  while (gshift > 0 && (1 << (gshift - 1)) >= nR) --gshift;
  const int group = 1 << gshift;
  const size_t smem = (((size_t)group * nbins + 1) / 2) * 4;
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

// ------------------------------------------------------------------------------------------------
// rotate_vector_simd (transforms3d_supplement.py:270-296) for float32 vectors and one float64 quaternion:
// float64 result, every product and sum rounded separately in NumPy's order so the output is bit-identical.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
rotate_f32_f64_kernel(const float* __restrict__ v, long long n, double qw, double qx, double qy, double qz,
                      double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double vx = v[3 * i], vy = v[3 * i + 1], vz = v[3 * i + 2];
  // a = cross(q_v, v) + q_w * v
  const double ax = __dadd_rn(__dsub_rn(__dmul_rn(qy, vz), __dmul_rn(qz, vy)), __dmul_rn(qw, vx));
  const double ay = __dadd_rn(__dsub_rn(__dmul_rn(qz, vx), __dmul_rn(qx, vz)), __dmul_rn(qw, vy));
  const double az = __dadd_rn(__dsub_rn(__dmul_rn(qx, vy), __dmul_rn(qy, vx)), __dmul_rn(qw, vz));
  // b = cross(q_v, a) ; out = b + b + v
  const double bx = __dsub_rn(__dmul_rn(qy, az), __dmul_rn(qz, ay));
  const double by = __dsub_rn(__dmul_rn(qz, ax), __dmul_rn(qx, az));
  const double bz = __dsub_rn(__dmul_rn(qx, ay), __dmul_rn(qy, ax));
  out[3 * i] = __dadd_rn(__dadd_rn(bx, bx), vx);
  out[3 * i + 1] = __dadd_rn(__dadd_rn(by, by), vy);
  out[3 * i + 2] = __dadd_rn(__dadd_rn(bz, bz), vz);
}
}  // namespace

extern "C" int sr_rotate_vectors_f32_f64(const float* d_v, long long n, const double* h_q, double* d_out, void* stream) {
  SR_REQUIRE(d_v && h_q && d_out, "sr_rotate_vectors_f32_f64: null pointer");
  SR_REQUIRE(n >= 1, "sr_rotate_vectors_f32_f64: empty input");
  rotate_f32_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_v, n, h_q[0], h_q[1], h_q[2],
                                                                                     h_q[3], d_out);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}
