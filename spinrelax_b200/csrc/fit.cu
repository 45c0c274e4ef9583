// K5: batched bounded Levenberg-Marquardt fit of C(t) = S2 + sum_i C_i exp(-t/tau_i), one CTA per residue.
// Replaces the scipy.optimize.curve_fit call of autoCorrelationModel.conduct_curve_fitting
// (fitting_Ct_functions.py:306-345; model curvefit_exponential :419-427; bounds :412-416).
// Parameter vector as in the reference: p = (C_1..C_nc, tau_1..tau_nc [, S2]); with an even count
// S2 = 1 - sum C.  Residuals are (model - y)/sigma (curve_fit with sigma, absolute_sigma=False).
// The kernel returns the optimum, the Gauss-Newton matrix J^T J at the optimum (the host forms pcov from
// it the way SciPy does from the SVD of J) and the cost 0.5 sum r^2.  Model selection (the 2,3,5,7,9
// ladder, chi^2 ratio, over-fitting flags) stays in Python, verbatim (fitct.py).
#include "common.cuh"

namespace {

constexpr int kMaxP = 9;
constexpr int kNRed = kMaxP * (kMaxP + 1) / 2 + kMaxP + 1;   // packed JtJ + Jtr + cost
constexpr int kFitThreads = 128;

struct FitShared {
  double p[kMaxP], ptry[kMaxP], lo[kMaxP], hi[kMaxP];
  double JtJ[kMaxP * kMaxP], Jtr[kMaxP], cost;
  double tJtJ[kMaxP * kMaxP], tJtr[kMaxP], tcost;
  double red[kFitThreads / 32][kNRed];
  int flag, trunc;
};

// accumulate cost, J^T r and J^T J of the model at parameters q over this thread's points, then block-reduce
template <int nP>
__device__ void evaluate(const double* __restrict__ t, const double* __restrict__ y, const double* __restrict__ sig,
                         int L, const double* q, FitShared& sh, double* outJtJ, double* outJtr, double* outCost) {
  constexpr int nc = nP / 2;
  constexpr bool free_s2 = (nP & 1);
  double C[4], itau[4];
  double sumC = 0.0;
#pragma unroll
  for (int i = 0; i < nc; ++i) { C[i] = q[i]; itau[i] = 1.0 / q[nc + i]; sumC += C[i]; }   // one division per tau, not per point
  const double S2 = free_s2 ? q[nP - 1] : 1.0 - sumC;
  constexpr int nUsed = nP * (nP + 1) / 2 + nP + 1;
  double acc[nUsed];
#pragma unroll
  for (int i = 0; i < nUsed; ++i) acc[i] = 0.0;
  for (int k = threadIdx.x; k < L; k += kFitThreads) {
    const double tk = t[k];
    const double w = sig ? 1.0 / sig[k] : 1.0;
    double g[nP];
    double f = S2;
#pragma unroll
    for (int i = 0; i < nc; ++i) {
      const double e = exp(-tk * itau[i]);
      f += C[i] * e;
      g[i] = (e - (free_s2 ? 0.0 : 1.0)) * w;
      g[nc + i] = C[i] * e * tk * (itau[i] * itau[i]) * w;
    }
    if (free_s2) g[nP - 1] = w;
    const double r = (f - y[k]) * w;
    int m = 0;
#pragma unroll
    for (int a = 0; a < nP; ++a)
#pragma unroll
      for (int b = a; b < nP; ++b) acc[m++] += g[a] * g[b];
#pragma unroll
    for (int a = 0; a < nP; ++a) acc[m++] += g[a] * r;
    acc[m] += 0.5 * r * r;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < nUsed; ++i) {
    const double v = sr_warp_sum(acc[i]);
    if (lane == 0) sh.red[warp][i] = v;
  }
  __syncthreads();
  // one thread per reduced quantity combines the warps' partial sums (the solver iterates one residue per CTA, so
  // the time of a fit is the latency of this chain, not its throughput)
  if (threadIdx.x < nUsed) {
    const int m = threadIdx.x;
    double s2 = 0.0;
#pragma unroll
    for (int w2 = 0; w2 < kFitThreads / 32; ++w2) s2 += sh.red[w2][m];
    constexpr int nTri = nP * (nP + 1) / 2;
    if (m < nTri) {
      int a = 0, rem = m;                      // m -> (a, b >= a) of the packed upper triangle
      while (rem >= nP - a) { rem -= nP - a; ++a; }
      const int b = a + rem;
      outJtJ[a * nP + b] = s2; outJtJ[b * nP + a] = s2;
    } else if (m < nTri + nP) {
      outJtr[m - nTri] = s2;
    } else {
      *outCost = s2;
    }
  }
  __syncthreads();
}

// solve (A + lam*diag(A)) d = -g on the free set by Cholesky; returns false if not positive definite
__device__ bool lm_step(const double* A, const double* g, const bool* fixed, int n, double lam, double* d) {
  double M[kMaxP * kMaxP], b[kMaxP];
  int idx[kMaxP], m = 0;
  for (int i = 0; i < n; ++i) { d[i] = 0.0; if (!fixed[i]) idx[m++] = i; }
  if (m == 0) return true;
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j < m; ++j) M[i * m + j] = A[idx[i] * n + idx[j]];
    const double dii = A[idx[i] * n + idx[i]];
    M[i * m + i] += lam * (dii > 0.0 ? dii : 1.0);
    b[i] = -g[idx[i]];
  }
  double inv[kMaxP];                           // 1 / L_jj: one rsqrt per column instead of a division per element
  for (int j = 0; j < m; ++j) {
    double s = M[j * m + j];
    for (int k = 0; k < j; ++k) s -= M[j * m + k] * M[j * m + k];
    if (!(s > 0.0) || !(s < 1e300)) return false;
    const double rj = rsqrt(s);
    inv[j] = rj;
    M[j * m + j] = s * rj;
    for (int i = j + 1; i < m; ++i) {
      double v = M[i * m + j];
      for (int k = 0; k < j; ++k) v -= M[i * m + k] * M[j * m + k];
      M[i * m + j] = v * rj;
    }
  }
  for (int i = 0; i < m; ++i) {
    double v = b[i];
    for (int k = 0; k < i; ++k) v -= M[i * m + k] * b[k];
    b[i] = v * inv[i];
  }
  for (int i = m - 1; i >= 0; --i) {
    double v = b[i];
    for (int k = i + 1; k < m; ++k) v -= M[k * m + i] * b[k];
    b[i] = v * inv[i];
  }
  for (int i = 0; i < m; ++i) d[idx[i]] = b[i];
  return true;
}

template <int nP>
__global__ void __launch_bounds__(kFitThreads)
ct_fit_lm_kernel(const double* __restrict__ T, const double* __restrict__ Y, const double* __restrict__ SIG, int L,
                 const double* __restrict__ P0, const double* __restrict__ LO, const double* __restrict__ HI, int max_iter,
                 double ftol, double* __restrict__ POPT, double* __restrict__ JTJ, double* __restrict__ COST,
                 int* __restrict__ STATUS) {
  __shared__ FitShared sh;
  const int r = blockIdx.x;
  const double* t = T + (long long)r * L;
  const double* y = Y + (long long)r * L;
  const double* sig = SIG ? SIG + (long long)r * L : nullptr;
  constexpr int nc = nP / 2;
  if (threadIdx.x < nP) {
    const int i = threadIdx.x;
    double lo = LO[(long long)r * nP + i], hi = HI[(long long)r * nP + i];
    if (i >= nc && i < 2 * nc) lo = fmax(lo, 1e-12 * hi);     // tau strictly positive (SciPy TRF iterates are interior)
    sh.lo[i] = lo; sh.hi[i] = hi;
    const double eps = 1e-10 * (hi - lo);                     // strictly feasible start (scipy make_strictly_feasible)
    sh.p[i] = fmin(fmax(P0[(long long)r * nP + i], lo + eps), hi - eps);
  }
  __syncthreads();
  evaluate<nP>(t, y, sig, L, sh.p, sh, sh.JtJ, sh.Jtr, &sh.cost);
  double lam = 1e-3;
  int status = 0, it = 0, small = 0;
  for (; it < max_iter; ++it) {
    if (threadIdx.x == 0) {
      // active set: a parameter sitting on a bound with the gradient pushing outwards is frozen
      bool fixed[kMaxP];
      for (int i = 0; i < nP; ++i) {
        const double span = sh.hi[i] - sh.lo[i];
        const bool at_lo = sh.p[i] - sh.lo[i] <= 1e-12 * span, at_hi = sh.hi[i] - sh.p[i] <= 1e-12 * span;
        fixed[i] = (at_lo && sh.Jtr[i] > 0.0) || (at_hi && sh.Jtr[i] < 0.0);
      }
      double d[kMaxP];
      int ok = lm_step(sh.JtJ, sh.Jtr, fixed, nP, lam, d) ? 1 : 0;
      // per-coordinate fraction-to-boundary rule: a coordinate whose step would leave the box moves 99.5% of the
      // way to that bound instead (the trial point stays strictly inside, as SciPy's TRF iterates do, so a wild
      // step can never park tau on 0 where the model has no gradient); the other coordinates keep their step.
      int tiny = 1, truncated = 0;          // tiny: every |step_i| < 1e-15 |p_i|
      for (int i = 0; i < nP; ++i) {
        double q = sh.p[i] + d[i];
        if (q < sh.lo[i]) { q = sh.p[i] - 0.995 * (sh.p[i] - sh.lo[i]); truncated = 1; }
        else if (q > sh.hi[i]) { q = sh.p[i] + 0.995 * (sh.hi[i] - sh.p[i]); truncated = 1; }
        if (!(fabs(q - sh.p[i]) < 1e-15 * (fabs(sh.p[i]) + 1e-300))) tiny = 0;
        sh.ptry[i] = q;
      }
      sh.trunc = truncated;
      sh.flag = ok ? (tiny ? 2 : 1) : 0;
    }
    __syncthreads();
    const int flag = sh.flag, trunc = sh.trunc;
    __syncthreads();
    if (flag == 0) { lam *= 10.0; if (lam > 1e20) { status = 3; break; } continue; }
    if (flag == 2 && !trunc) { status = 2; break; }      // step below machine precision
    evaluate<nP>(t, y, sig, L, sh.ptry, sh, sh.tJtJ, sh.tJtr, &sh.tcost);
    const double c0 = sh.cost, c1 = sh.tcost;
    const bool accept = (c1 <= c0) && (c1 == c1);
    __syncthreads();
    if (accept) {
      if (threadIdx.x == 0) {
        for (int i = 0; i < nP; ++i) { sh.p[i] = sh.ptry[i]; sh.Jtr[i] = sh.tJtr[i]; }
        for (int i = 0; i < nP * nP; ++i) sh.JtJ[i] = sh.tJtJ[i];
        sh.cost = c1;
      }
      lam = fmax(lam * 0.3, 1e-12);
      __syncthreads();
      // converged: three consecutive negligible decreases taken with (almost) undamped Gauss-Newton steps.
      // A tiny decrease under heavy damping only means the step was short (flat multi-exponential valleys).
      small = (c0 - c1 <= ftol * c0 && !trunc) ? small + 1 : 0;    // a truncated step says nothing about convergence
      if (small >= 3 && lam <= 1e-7) { status = 1; ++it; break; }
    } else {
      lam *= 4.0;
      if (lam > 1e20) { status = 3; break; }
    }
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < nP; ++i) POPT[(long long)r * nP + i] = sh.p[i];
    for (int i = 0; i < nP * nP; ++i) JTJ[(long long)r * nP * nP + i] = sh.JtJ[i];
    COST[r] = sh.cost;
    STATUS[2 * r] = status; STATUS[2 * r + 1] = it;
  }
}

}  // namespace

extern "C" int sr_ct_fit_lm(const double* d_t, const double* d_y, const double* d_sigma, int nR, long long L, int nParams,
                            const double* d_p0, const double* d_lo, const double* d_hi, int max_iter, double ftol,
                            double* d_popt, double* d_JtJ, double* d_cost, int* d_status, void* stream) {
  SR_REQUIRE(d_t && d_y && d_p0 && d_lo && d_hi && d_popt && d_JtJ && d_cost && d_status, "sr_ct_fit_lm: null pointer");
  SR_REQUIRE(nR > 0 && L > 0 && L < (1LL << 31), "sr_ct_fit_lm: bad shape (nR=%d L=%lld)", nR, L);
  SR_REQUIRE(nParams >= 2 && nParams <= kMaxP, "sr_ct_fit_lm: nParams %d outside [2, %d]", nParams, kMaxP);
  SR_REQUIRE(max_iter > 0 && ftol >= 0, "sr_ct_fit_lm: bad solver settings");
#define SR_FIT_CASE(NP)                                                                                     \
  case NP:                                                                                                  \
    ct_fit_lm_kernel<NP><<<nR, kFitThreads, 0, (cudaStream_t)stream>>>(d_t, d_y, d_sigma, (int)L, d_p0, d_lo, d_hi, \
                                                                       max_iter, ftol, d_popt, d_JtJ, d_cost, d_status); \
    break;
  switch (nParams) {
    SR_FIT_CASE(2) SR_FIT_CASE(3) SR_FIT_CASE(4) SR_FIT_CASE(5) SR_FIT_CASE(6) SR_FIT_CASE(7) SR_FIT_CASE(8) SR_FIT_CASE(9)
  }
#undef SR_FIT_CASE
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}
