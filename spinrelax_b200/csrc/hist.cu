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
constexpr int kMaxGroup = 8;   // vectors per CTA (shared-memory privatised histograms)

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

// grid.x = frame blocks, grid.y = vector groups of `group` vectors
__global__ void __launch_bounds__(kHistThreads)
sphere_hist_kernel(const float* __restrict__ vecs, long long nFrames, int nR, int group, long long framesPerBlock,
                   HistParams p, const double2* __restrict__ edge_dir, const double* __restrict__ edge_cos,
                   unsigned int* __restrict__ counts, long long* __restrict__ amb_idx, int amb_capacity,
                   int* __restrict__ amb_count) {
  extern __shared__ unsigned int sh_hist[];
  const int nbins = p.nbx * p.nby;
  const int r0 = blockIdx.y * group;
  const int nv = min(group, nR - r0);
  for (int i = threadIdx.x; i < nv * nbins; i += kHistThreads) sh_hist[i] = 0u;
  __syncthreads();

  const long long f0 = (long long)blockIdx.x * framesPerBlock;
  const long long f1 = min(nFrames, f0 + framesPerBlock);
  const long long nSamp = (f1 - f0) * nv;
  for (long long s = threadIdx.x; s < nSamp; s += kHistThreads) {
    const int vl = (int)(s % nv);
    const long long f = f0 + s / nv;
    const float* src = vecs + (f * nR + r0 + vl) * 3;
    const double vx = (double)__ldg(src), vy = (double)__ldg(src + 1), vz = (double)__ldg(src + 2);
    const double x = p.R[0] * vx + p.R[1] * vy + p.R[2] * vz;
    const double y = p.R[3] * vx + p.R[4] * vy + p.R[5] * vz;
    const double z = p.R[6] * vx + p.R[7] * vy + p.R[8] * vz;
    int bin = 0;
    const int cls = classify(x, y, z, p, edge_dir, edge_cos, bin);
    if (cls == 0) {
      atomicAdd(&sh_hist[vl * nbins + bin], 1u);
    } else if (cls == 2) {
      const int slot = atomicAdd(amb_count, 1);
      if (slot < amb_capacity) amb_idx[slot] = f * nR + r0 + vl;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nv * nbins; i += kHistThreads) {
    const unsigned int c = sh_hist[i];
    if (c) atomicAdd(&counts[(long long)r0 * nbins + i], c);
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
  // two CTAs per SM: at most ~100 KB of privatised bins each
  int group = (int)((size_t)100 * 1024 / ((size_t)nbins * 4));
  if (group > kMaxGroup) group = kMaxGroup;
  if (group > nR) group = nR;
  if (group < 1) {
    group = 1;
    SR_REQUIRE((size_t)nbins * 4 <= (size_t)max_smem, "sr_sphere_hist: %d bins do not fit in shared memory", nbins);
  }
  const size_t smem = (size_t)group * nbins * 4;
  SR_CUDA(cudaFuncSetAttribute(sphere_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nGroups = (nR + group - 1) / group;
  long long nFB = (4LL * 2 * sms + nGroups - 1) / nGroups;     // ~4 waves of 2 CTAs/SM
  long long fpb = (nFrames + nFB - 1) / nFB;
  if (fpb < 256) fpb = 256;
  nFB = (nFrames + fpb - 1) / fpb;
  SR_REQUIRE(nGroups <= 65535, "sr_sphere_hist: too many vector groups");
  dim3 grid((unsigned)nFB, (unsigned)nGroups);
  const double2* edge_dir = (const double2*)d_edge_table;
  const double* edge_cos = d_edge_table + 2 * (nbx + 1);
  sphere_hist_kernel<<<grid, kHistThreads, smem, (cudaStream_t)stream>>>(d_vecs, nFrames, nR, group, fpb, p, edge_dir,
                                                                         edge_cos, d_counts, d_amb_idx, amb_capacity,
                                                                         d_amb_count);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}
