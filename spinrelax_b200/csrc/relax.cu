// K6/K7: spectral densities and relaxation rates.
//   K7 jomega_kernel      npufunc.Jomega(x,y) = x/(x*x+y*y), Jomega/Jomega.c:30-104 (float, double loops)
//   K6a a_moments_kernel  per-residue weighted mean (3) and covariance (6) of the A_J coefficients over the
//                         histogram bins (update_A_coefficients spectral_densities.py:503-523, weights = counts)
//   K6b relax_eval_kernel J(omega) (calc_Jomega_one :552-557 / _do_Jsum :1961-1972; isotropic :430-443),
//                         R1/R2/NOE (:824-829, :859-864, :888-892) and their weighted mean / sigma over the
//                         bin vectors (check_and_calculate_average :751-763) for every
//                         (residue, field, CSA grid point).
// R1, R2 and NOE-1 are linear in J and J is linear in A_J(bin), so mean_b w R = Abar . c and
// sigma^2 = c^T Cov_w(A) c with c a 3-vector of Lorentzians: the (bins x 3)(3 x 5) contraction of the
// reference collapses to one pass over the weights plus O(1) work per evaluation.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) jomega_f64_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                             double* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double a = x[i], b = y[i];
    out[i] = __ddiv_rn(a, __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));   // no FMA contraction: matches the C loop
  }
}
__global__ void __launch_bounds__(256) jomega_f32_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float a = x[i], b = y[i];
    out[i] = __fdiv_rn(a, __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)));
  }
}

// one block per residue: weighted moments of A over the bins
__global__ void __launch_bounds__(256)
a_moments_kernel(const double* __restrict__ A /*[B][3] or [nR][B][3]*/, int per_residue_A,
                 const double* __restrict__ W /*[nR][B]*/, int B, double* __restrict__ out /*[nR][10]*/) {
  const int r = blockIdx.x;
  const double* w = W + (long long)r * B;
  const double* a = A + (per_residue_A ? (long long)r * B * 3 : 0);
  __shared__ double red[8][9];
  __shared__ double mean[4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // pass 1: sum w, sum w A
  double s[4] = {0, 0, 0, 0};
  for (int b = threadIdx.x; b < B; b += 256) {
    const double wb = w[b];
    s[0] += wb; s[1] = fma(wb, a[3 * b], s[1]); s[2] = fma(wb, a[3 * b + 1], s[2]); s[3] = fma(wb, a[3 * b + 2], s[3]);
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) s[m] = sr_warp_sum(s[m]);
  if (lane == 0) { red[warp][0] = s[0]; red[warp][1] = s[1]; red[warp][2] = s[2]; red[warp][3] = s[3]; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    mean[threadIdx.x] = t;
  }
  __syncthreads();
  const double sw = mean[0];
  const double m0 = mean[1] / sw, m1 = mean[2] / sw, m2 = mean[3] / sw;
  __syncthreads();
  // pass 2: centred second moments
  double c[6] = {0, 0, 0, 0, 0, 0};
  for (int b = threadIdx.x; b < B; b += 256) {
    const double wb = w[b];
    const double d0 = a[3 * b] - m0, d1 = a[3 * b + 1] - m1, d2 = a[3 * b + 2] - m2;
    c[0] = fma(wb * d0, d0, c[0]); c[1] = fma(wb * d0, d1, c[1]); c[2] = fma(wb * d0, d2, c[2]);
    c[3] = fma(wb * d1, d1, c[3]); c[4] = fma(wb * d1, d2, c[4]); c[5] = fma(wb * d2, d2, c[5]);
  }
#pragma unroll
  for (int m = 0; m < 6; ++m) c[m] = sr_warp_sum(c[m]);
  if (lane == 0) {
#pragma unroll
    for (int m = 0; m < 6; ++m) red[warp][m] = c[m];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double t = 0;
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    out[(long long)r * 10 + 4 + threadIdx.x] = t / sw;
  }
  if (threadIdx.x == 0) { out[(long long)r * 10] = sw; out[(long long)r * 10 + 1] = m0; out[(long long)r * 10 + 2] = m1; out[(long long)r * 10 + 3] = m2; }
}

struct RelaxConsts {
  double zeta, time_fact, gammaA, gammaB, f_dd;
  double D_J[3];
  int iso;          // 1: isotropic tumbling, D_J[0] = D_iso, no vector averaging
  int max_comp;
};

__device__ __forceinline__ double lorentz(double d, double om) { return d / (d * d + om * om); }

// one thread per (residue, field, csa)
__global__ void __launch_bounds__(128)
relax_eval_kernel(RelaxConsts k, int nR, int nField, int nCSA, int csa_per_residue,
                  const double* __restrict__ amom /*[nR][10]*/, const double* __restrict__ S2,
                  const double* __restrict__ C /*[nR][max_comp]*/, const double* __restrict__ tau,
                  const int* __restrict__ nComp, const double* __restrict__ omega /*[nField][5]*/,
                  const double* __restrict__ f_csa /*[nField][nCSA] or [nField][nR]*/,
                  double* __restrict__ out /*[nR][nField][nCSA][6] = R1,R2,NOE,sR1,sR2,sNOE*/) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)nR * nField * nCSA;
  if (idx >= total) return;
  const int ic = (int)(idx % nCSA);
  const int f = (int)((idx / nCSA) % nField);
  const int r = (int)(idx / ((long long)nCSA * nField));
  const double* om = omega + f * 5;
  const double fc = csa_per_residue ? f_csa[(long long)f * nR + r] : f_csa[(long long)f * nCSA + ic];
  const int nc = nComp[r];
  const double s2 = S2[r];
  double* o = out + idx * 6;
  const double tf = k.time_fact, fdd = k.f_dd;

  if (k.iso) {
    // J_k = zeta [ S2 tg/(1+(om tg)^2) + sum_c C_c kc/(kc^2+om^2) ], tg = 1/(6D), kc = 1/tg + 1/tau_c  (:430-443)
    const double tg = 1.0 / (6.0 * k.D_J[0]);
    double J[5];
#pragma unroll
    for (int w = 0; w < 5; ++w) {
      double j = k.zeta * s2 * tg / (1.0 + (om[w] * tg) * (om[w] * tg));
      for (int c = 0; c < nc; ++c) {
        const double kc = 1.0 / tg + 1.0 / tau[(long long)r * k.max_comp + c];
        j += k.zeta * C[(long long)r * k.max_comp + c] * kc / (kc * kc + om[w] * om[w]);
      }
      J[w] = j;
    }
    const double R1 = tf * (fdd * (J[2] + 3 * J[1] + 6 * J[4]) + fc * J[1]);
    const double R2 = tf * (0.5 * fdd * (4 * J[0] + J[2] + 3 * J[1] + 6 * J[4] + 6 * J[3]) + 1.0 / 6.0 * fc * (4 * J[0] + 3 * J[1]));
    const double NOE = 1.0 + tf * k.gammaB / (k.gammaA * R1) * fdd * (6 * J[4] - J[2]);
    o[0] = R1; o[1] = R2; o[2] = NOE; o[3] = 0; o[4] = 0; o[5] = 0;
    return;
  }
  // T[j][w] = zeta [ S2 L(D_j, om_w) + sum_c C_c L(D_j + 1/tau_c, om_w) ]
  double c1[3], c2[3], cn[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    double T[5];
#pragma unroll
    for (int w = 0; w < 5; ++w) {
      double t = s2 * lorentz(k.D_J[j], om[w]);
      for (int c = 0; c < nc; ++c)
        t += C[(long long)r * k.max_comp + c] * lorentz(k.D_J[j] + 1.0 / tau[(long long)r * k.max_comp + c], om[w]);
      T[w] = k.zeta * t;
    }
    c1[j] = tf * (fdd * (T[2] + 3 * T[1] + 6 * T[4]) + fc * T[1]);
    c2[j] = tf * (0.5 * fdd * (4 * T[0] + T[2] + 3 * T[1] + 6 * T[4] + 6 * T[3]) + 1.0 / 6.0 * fc * (4 * T[0] + 3 * T[1]));
    cn[j] = fdd * (6 * T[4] - T[2]);
  }
  const double* m = amom + (long long)r * 10;
  const double a0 = m[1], a1 = m[2], a2 = m[3];
  auto quad = [&](const double* c) {
    const double q = c[0] * c[0] * m[4] + c[1] * c[1] * m[7] + c[2] * c[2] * m[9] +
                     2.0 * (c[0] * c[1] * m[5] + c[0] * c[2] * m[6] + c[1] * c[2] * m[8]);
    return q > 0.0 ? sqrt(q) : 0.0;
  };
  const double R1 = a0 * c1[0] + a1 * c1[1] + a2 * c1[2];
  const double R2 = a0 * c2[0] + a1 * c2[1] + a2 * c2[2];
  const double pre = tf * k.gammaB / (k.gammaA * R1);     // NOE uses the bin-averaged R1 (quirk G8, :885-886,:901)
  const double NOE = 1.0 + pre * (a0 * cn[0] + a1 * cn[1] + a2 * cn[2]);
  o[0] = R1; o[1] = R2; o[2] = NOE;
  o[3] = quad(c1); o[4] = quad(c2); o[5] = fabs(pre) * quad(cn);
}

}  // namespace

extern "C" int sr_jomega_f64(const double* d_x, const double* d_y, double* d_out, long long n, void* stream) {
  SR_REQUIRE(d_x && d_y && d_out, "sr_jomega_f64: null pointer");
  if (n <= 0) return SR_OK;
  jomega_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_x, d_y, d_out, n);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_jomega_f32(const float* d_x, const float* d_y, float* d_out, long long n, void* stream) {
  SR_REQUIRE(d_x && d_y && d_out, "sr_jomega_f32: null pointer");
  if (n <= 0) return SR_OK;
  jomega_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_x, d_y, d_out, n);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

// Host-strided form for a NumPy ufunc inner loop (Jomega/Jomega.c:49-66 has the signature
// loop(char** args, npy_intp* dimensions, npy_intp* steps, void*)): gathers the three strided streams, runs the kernel
// and scatters the result.  Zero strides (broadcast scalars) are honoured.
namespace {
// grow-only staging buffers per host thread (NumPy may call a ufunc loop many times with a short inner dimension)
struct JomegaScratch {
  void* h = nullptr; void* d = nullptr; size_t cap = 0; int dev = -1;
  ~JomegaScratch() { if (h) cudaFreeHost(h); if (d) cudaFree(d); }
};
thread_local JomegaScratch g_jomega_scratch;

template <typename T, typename K>
int jomega_host_strided(const char* x, long long sx, const char* y, long long sy, char* out, long long so, long long n,
                        K launch) {
  if (n <= 0) return SR_OK;
  JomegaScratch& sc = g_jomega_scratch;
  int dev = 0;
  SR_CUDA(cudaGetDevice(&dev));
  const size_t need = sizeof(T) * 3 * (size_t)n;
  cudaError_t e = cudaSuccess;
  if (sc.cap < need || sc.dev != dev) {
    if (sc.h) cudaFreeHost(sc.h);
    if (sc.d) cudaFree(sc.d);
    sc.h = sc.d = nullptr; sc.cap = 0; sc.dev = dev;
    const size_t cap = need < 4096 ? 4096 : need;
    if ((e = cudaMallocHost(&sc.h, cap)) != cudaSuccess || (e = cudaMalloc(&sc.d, cap)) != cudaSuccess) {
      if (sc.h) cudaFreeHost(sc.h);
      sc.h = nullptr;
      sr_set_error("sr_jomega_host: allocation of %lld elements failed: %s", n, cudaGetErrorString(e));
      return SR_ERR_CUDA;
    }
    sc.cap = cap;
  }
  T* h = static_cast<T*>(sc.h);
  T* d = static_cast<T*>(sc.d);
  for (long long i = 0; i < n; ++i) {
    h[i] = *reinterpret_cast<const T*>(x + i * sx);
    h[n + i] = *reinterpret_cast<const T*>(y + i * sy);
  }
  int rc = SR_OK;
  if ((e = cudaMemcpy(d, h, sizeof(T) * 2 * (size_t)n, cudaMemcpyHostToDevice)) == cudaSuccess) {
    rc = launch(d, d + n, d + 2 * n, n);
    if (rc == SR_OK) e = cudaMemcpy(h + 2 * n, d + 2 * n, sizeof(T) * (size_t)n, cudaMemcpyDeviceToHost);
  }
  if (e != cudaSuccess) { sr_set_error("sr_jomega_host: copy failed: %s", cudaGetErrorString(e)); rc = SR_ERR_CUDA; }
  if (rc == SR_OK)
    for (long long i = 0; i < n; ++i) *reinterpret_cast<T*>(out + i * so) = h[2 * n + i];
  return rc;
}
}  // namespace

extern "C" int sr_jomega_host_f64(const char* x, long long sx, const char* y, long long sy, char* out, long long so, long long n) {
  SR_REQUIRE(x && y && out, "sr_jomega_host_f64: null pointer");
  return jomega_host_strided<double>(x, sx, y, sy, out, so, n, [](const double* a, const double* b, double* c, long long m) {
    return sr_jomega_f64(a, b, c, m, nullptr); });
}

extern "C" int sr_jomega_host_f32(const char* x, long long sx, const char* y, long long sy, char* out, long long so, long long n) {
  SR_REQUIRE(x && y && out, "sr_jomega_host_f32: null pointer");
  return jomega_host_strided<float>(x, sx, y, sy, out, so, n, [](const float* a, const float* b, float* c, long long m) {
    return sr_jomega_f32(a, b, c, m, nullptr); });
}

extern "C" int sr_relax_a_moments(const double* d_A, int per_residue_A, const double* d_W, int nR, int B,
                                  double* d_amom, void* stream) {
  SR_REQUIRE(d_A && d_W && d_amom, "sr_relax_a_moments: null pointer");
  SR_REQUIRE(nR > 0 && B > 0, "sr_relax_a_moments: empty shape");
  a_moments_kernel<<<nR, 256, 0, (cudaStream_t)stream>>>(d_A, per_residue_A, d_W, B, d_amom);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_relax_eval(int iso, const double* h_D_J, double zeta, double time_fact, double gammaA, double gammaB,
                             double f_dd, int nR, int nField, int nCSA, int csa_per_residue, int max_comp,
                             const double* d_amom, const double* d_S2, const double* d_C, const double* d_tau,
                             const int* d_nComp, const double* d_omega, const double* d_f_csa, double* d_out,
                             void* stream) {
  SR_REQUIRE(h_D_J && d_S2 && d_C && d_tau && d_nComp && d_omega && d_f_csa && d_out, "sr_relax_eval: null pointer");
  SR_REQUIRE(iso || d_amom, "sr_relax_eval: A-moments required for axisymmetric tumbling");
  SR_REQUIRE(nR > 0 && nField > 0 && nCSA > 0 && max_comp >= 0, "sr_relax_eval: empty shape");
  RelaxConsts k;
  k.zeta = zeta; k.time_fact = time_fact; k.gammaA = gammaA; k.gammaB = gammaB; k.f_dd = f_dd;
  k.D_J[0] = h_D_J[0]; k.D_J[1] = iso ? 0.0 : h_D_J[1]; k.D_J[2] = iso ? 0.0 : h_D_J[2];
  k.iso = iso; k.max_comp = max_comp;
  const long long total = (long long)nR * nField * nCSA;
  relax_eval_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      k, nR, nField, nCSA, csa_per_residue, d_amom, d_S2, d_C, d_tau, d_nComp, d_omega, d_f_csa, d_out);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}
