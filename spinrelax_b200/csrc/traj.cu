// X-H bond-vector extraction from Cartesian trajectories, with and without the per-frame least-squares
// superposition onto a reference structure: the step immediately before the C(t) hot path.
// Reference: obtain_XHvecs, calculate-Ct-from-traj.py:64-86 (np.take differences + qs.vecnorm_NDarray,
// transforms3d_supplement.py:40-52) and the mdtraj calls trj.center_coordinates(); trj.superpose(ref, frame=0,
// atom_indices=fit_indices) at :466-467 (split read :437-438).  Only the rotation of the superposition matters
// for bond vectors, so the coordinates themselves are never rewritten.
//
// Compiled without -ftz / with IEEE sqrt and division: the normalisation is bit-identical to NumPy's float32
// v / sqrt((x*x + y*y) + z*z) followed by nan_to_num.
#include "common.cuh"

#include <cfloat>

namespace {

__device__ __forceinline__ float nan_to_num_f32(float v) {
  if (v != v) return 0.f;
  if (isinf(v)) return v > 0.f ? FLT_MAX : -FLT_MAX;
  return v;
}

// vecnorm_NDarray for one float32 vector, NumPy's operation order, no FMA contraction
__device__ __forceinline__ void unit_f32(float x, float y, float z, float* out) {
  const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  const float n = __fsqrt_rn(n2);
  out[0] = nan_to_num_f32(__fdiv_rn(x, n));
  out[1] = nan_to_num_f32(__fdiv_rn(y, n));
  out[2] = nan_to_num_f32(__fdiv_rn(z, n));
}

// ---- plain extraction: one thread per (frame, bond) -------------------------------------------------------
__global__ void __launch_bounds__(256)
xh_vectors_kernel(const float* __restrict__ xyz, long long nFrames, int nAtoms, const int* __restrict__ idxH,
                  const int* __restrict__ idxX, int nR, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nFrames * nR) return;
  const long long f = i / nR;
  const int r = (int)(i - f * nR);
  const float* fr = xyz + f * (long long)nAtoms * 3;
  const float* h = fr + 3LL * __ldg(idxH + r);
  const float* x = fr + 3LL * __ldg(idxX + r);
  unit_f32(__fsub_rn(__ldg(h), __ldg(x)), __fsub_rn(__ldg(h + 1), __ldg(x + 1)), __fsub_rn(__ldg(h + 2), __ldg(x + 2)),
           out + 3 * i);
}

// ---- cyclic Jacobi for a symmetric 4x4 matrix, everything in registers (static indices) ------------------
struct Sym4 { double a[4][4]; double v[4][4]; };

template <int P, int Q>
__device__ __forceinline__ void jacobi_rotate(Sym4& m) {
  const double apq = m.a[P][Q];
  if (fabs(apq) < 1e-300) return;
  const double theta = (m.a[Q][Q] - m.a[P][P]) / (2.0 * apq);
  const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
  const double c = rsqrt(t * t + 1.0), s = t * c;
#pragma unroll
  for (int k = 0; k < 4; ++k) {          // A <- A J (columns p, q)
    const double akp = m.a[k][P], akq = m.a[k][Q];
    m.a[k][P] = c * akp - s * akq;
    m.a[k][Q] = s * akp + c * akq;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {          // A <- J^T A (rows p, q)
    const double apk = m.a[P][k], aqk = m.a[Q][k];
    m.a[P][k] = c * apk - s * aqk;
    m.a[Q][k] = s * apk + c * aqk;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {          // V <- V J
    const double vkp = m.v[k][P], vkq = m.v[k][Q];
    m.v[k][P] = c * vkp - s * vkq;
    m.v[k][Q] = s * vkp + c * vkq;
  }
}

// Horn (1987): the rotation R minimising sum |R x_i - y_i|^2 is the rotation of the unit quaternion that is the
// dominant eigenvector of the symmetric 4x4 matrix N built from S_ab = sum_i x_ia y_ib.  Always a proper
// rotation (det = +1), like the quaternion-based QCP solver behind mdtraj's superpose.
__device__ __forceinline__ void horn_rotation(const double S[9], double R[9]) {
  const double Sxx = S[0], Sxy = S[1], Sxz = S[2], Syx = S[3], Syy = S[4], Syz = S[5], Szx = S[6], Szy = S[7], Szz = S[8];
  Sym4 m;
  m.a[0][0] = Sxx + Syy + Szz; m.a[0][1] = Syz - Szy;       m.a[0][2] = Szx - Sxz;        m.a[0][3] = Sxy - Syx;
  m.a[1][1] = Sxx - Syy - Szz; m.a[1][2] = Sxy + Syx;       m.a[1][3] = Szx + Sxz;
  m.a[2][2] = -Sxx + Syy - Szz; m.a[2][3] = Syz + Szy;
  m.a[3][3] = -Sxx - Syy + Szz;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < i) m.a[i][j] = m.a[j][i];
      m.v[i][j] = (i == j) ? 1.0 : 0.0;
    }
#pragma unroll 1
  for (int sweep = 0; sweep < 10; ++sweep) {
    jacobi_rotate<0, 1>(m); jacobi_rotate<0, 2>(m); jacobi_rotate<0, 3>(m);
    jacobi_rotate<1, 2>(m); jacobi_rotate<1, 3>(m); jacobi_rotate<2, 3>(m);
  }
  // dominant eigenvector (static selects keep m in registers)
  double lam = m.a[0][0], qw = m.v[0][0], qx = m.v[1][0], qy = m.v[2][0], qz = m.v[3][0];
#pragma unroll
  for (int j = 1; j < 4; ++j)
    if (m.a[j][j] > lam) { lam = m.a[j][j]; qw = m.v[0][j]; qx = m.v[1][j]; qy = m.v[2][j]; qz = m.v[3][j]; }
  const double n = rsqrt(qw * qw + qx * qx + qy * qy + qz * qz);
  qw *= n; qx *= n; qy *= n; qz *= n;
  R[0] = 1 - 2 * (qy * qy + qz * qz); R[1] = 2 * (qx * qy - qw * qz);     R[2] = 2 * (qx * qz + qw * qy);
  R[3] = 2 * (qx * qy + qw * qz);     R[4] = 1 - 2 * (qx * qx + qz * qz); R[5] = 2 * (qy * qz - qw * qx);
  R[6] = 2 * (qx * qz - qw * qy);     R[7] = 2 * (qy * qz + qw * qx);     R[8] = 1 - 2 * (qx * qx + qy * qy);
}

// ---- superposed extraction: one warp per frame -------------------------------------------------------------
// ref_fit holds the reference coordinates of the fit atoms with their centroid removed, so
// S = sum_i x_i y_i^T needs no centring of the frame (sum_i y_i = 0 removes the centroid term).
__global__ void __launch_bounds__(256)
xh_superposed_kernel(const float* __restrict__ xyz, long long nFrames, int nAtoms, const int* __restrict__ fitIdx,
                     const double* __restrict__ refFit, int nFit, const int* __restrict__ idxH,
                     const int* __restrict__ idxX, int nR, float* __restrict__ out, double* __restrict__ rot) {
  const int lane = threadIdx.x & 31;
  const long long f = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= nFrames) return;
  const float* fr = xyz + f * (long long)nAtoms * 3;
  double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = lane; i < nFit; i += 32) {
    const float* a = fr + 3LL * __ldg(fitIdx + i);
    const double x0 = __ldg(a), x1 = __ldg(a + 1), x2 = __ldg(a + 2);
    const double y0 = refFit[3 * i], y1 = refFit[3 * i + 1], y2 = refFit[3 * i + 2];
    S[0] = fma(x0, y0, S[0]); S[1] = fma(x0, y1, S[1]); S[2] = fma(x0, y2, S[2]);
    S[3] = fma(x1, y0, S[3]); S[4] = fma(x1, y1, S[4]); S[5] = fma(x1, y2, S[5]);
    S[6] = fma(x2, y0, S[6]); S[7] = fma(x2, y1, S[7]); S[8] = fma(x2, y2, S[8]);
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) S[k] = sr_warp_sum(S[k]);      // xor butterfly: every lane ends with the same bits
  double R[9];
  horn_rotation(S, R);                                        // redundantly on all lanes (no divergence, no broadcast)
  if (rot && lane == 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k) rot[f * 9 + k] = R[k];
  }
  for (int r = lane; r < nR; r += 32) {
    const float* h = fr + 3LL * __ldg(idxH + r);
    const float* x = fr + 3LL * __ldg(idxX + r);
    const double d0 = __fsub_rn(__ldg(h), __ldg(x)), d1 = __fsub_rn(__ldg(h + 1), __ldg(x + 1)),
                 d2 = __fsub_rn(__ldg(h + 2), __ldg(x + 2));
    const float w0 = (float)(R[0] * d0 + R[1] * d1 + R[2] * d2);
    const float w1 = (float)(R[3] * d0 + R[4] * d1 + R[5] * d2);
    const float w2 = (float)(R[6] * d0 + R[7] * d1 + R[8] * d2);
    unit_f32(w0, w1, w2, out + (f * nR + r) * 3);
  }
}

int check_indices(const char* what, int n) {
  if (n <= 0) { sr_set_error("%s: empty index list", what); return SR_ERR_ARG; }
  return SR_OK;
}

}  // namespace

extern "C" int sr_xh_vectors(const float* d_xyz, long long nFrames, int nAtoms, const int* d_idxH, const int* d_idxX,
                             int nR, float* d_out, void* stream) {
  SR_REQUIRE(d_xyz && d_idxH && d_idxX && d_out, "sr_xh_vectors: null pointer");
  SR_REQUIRE(nFrames > 0 && nAtoms > 0, "sr_xh_vectors: empty trajectory (nFrames=%lld nAtoms=%d)", nFrames, nAtoms);
  if (int rc = check_indices("sr_xh_vectors", nR)) return rc;
  const long long n = nFrames * nR;
  SR_REQUIRE((n + 255) / 256 < (1LL << 31), "sr_xh_vectors: too many (frame, bond) pairs for one launch");
  xh_vectors_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_xyz, nFrames, nAtoms, d_idxH, d_idxX,
                                                                                 nR, d_out);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_xh_vectors_superposed(const float* d_xyz, long long nFrames, int nAtoms, const int* d_fitIdx,
                                        const double* d_refFitCentred, int nFit, const int* d_idxH, const int* d_idxX,
                                        int nR, float* d_out, double* d_rot, void* stream) {
  SR_REQUIRE(d_xyz && d_fitIdx && d_refFitCentred && d_idxH && d_idxX && d_out, "sr_xh_vectors_superposed: null pointer");
  SR_REQUIRE(nFrames > 0 && nAtoms > 0, "sr_xh_vectors_superposed: empty trajectory (nFrames=%lld nAtoms=%d)", nFrames,
             nAtoms);
  SR_REQUIRE(nFit >= 3, "sr_xh_vectors_superposed: need at least 3 fit atoms (got %d)", nFit);
  if (int rc = check_indices("sr_xh_vectors_superposed", nR)) return rc;
  const long long blocks = (nFrames + 7) / 8;
  SR_REQUIRE(blocks < (1LL << 31), "sr_xh_vectors_superposed: too many frames for one launch");
  xh_superposed_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_xyz, nFrames, nAtoms, d_fitIdx,
                                                                          d_refFitCentred, nFit, d_idxH, d_idxX, nR, d_out,
                                                                          d_rot);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}
