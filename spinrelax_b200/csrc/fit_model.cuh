// Per-point model arithmetic of K5 and the per-column Householder algebra of its QR; shared by the CUDA kernel
// (csrc/fit.cu) and the CPU test harness (tests/cpu_harness), so no CUDA-only construct here.
//
// Model: curvefit_exponential, fitting_Ct_functions.py:419-427 -- p = (C_1..C_nc, tau_1..tau_nc [, S2]),
// f(t) = S2 + sum_i C_i exp(-t / tau_i), S2 = 1 - sum C when the parameter count is even.  Residual as
// scipy.optimize.curve_fit forms it with sigma: (1/sigma) * (f - y).
#pragma once
#include "trf_core.cuh"

namespace srfit {

template <int N>
struct ModelPars {
  double C[N / 2 > 0 ? N / 2 : 1], tau[N / 2 > 0 ? N / 2 : 1], S2;
};

template <int N>
SR_HD void prepare(const double* x, ModelPars<N>& mp) {
  constexpr int nc = N / 2;
  double sumC = 0.0;
  for (int i = 0; i < nc; ++i) { mp.C[i] = x[i]; mp.tau[i] = x[nc + i]; sumC += x[i]; }
  mp.S2 = (N & 1) ? x[N - 1] : 1.0 - sumC;
}

// residual (1/sigma)(f - y) at one point and, if row != nullptr, the row of the analytic Jacobian of the residual
template <int N>
SR_HD double residual_and_row(const ModelPars<N>& mp, double t, double y, double w, double* row) {
  constexpr int nc = N / 2;
  constexpr bool free_s2 = (N & 1);
  double sum = 0.0;
  for (int i = 0; i < nc; ++i) {
    const double e = exp(-t / mp.tau[i]);
    const double ce = mp.C[i] * e;
    sum = (i == 0) ? ce : sum + ce;
    if (row) {
      row[i] = (e - (free_s2 ? 0.0 : 1.0)) * w;
      row[nc + i] = (e > 0.0) ? ce * t / (mp.tau[i] * mp.tau[i]) * w : 0.0;
    }
  }
  if (row && free_s2) row[N - 1] = w;
  return w * ((mp.S2 + sum) - y);
}

// One Householder step k of the QR of an M x (N+1) matrix whose rows are spread over the threads of a CTA.
// Inputs: row[j] = A[k][j] (the pivot row) and sums[j] = sum_{i>k} A[i][k] A[i][j] for j = k..N.
// Outputs: rrow[j] = R[k][j] (j = k..N; entry N is (Q^T rhs)[k]); every row i > k is then updated as
// A[i][j] -= tw[j] * (A[i][k] * vscale) for j = k+1..N.
template <int N>
struct HouseholderCol {
  double rrow[N + 1], tw[N + 1], vscale;
};

template <int N>
SR_HD void householder_column(int k, const double* row, const double* sums, HouseholderCol<N>& h) {
  const double alpha = row[k], xnorm2 = sums[k];
  if (xnorm2 == 0.0) {                       // nothing below the diagonal: H = I
    for (int j = k; j <= N; ++j) { h.rrow[j] = row[j]; h.tw[j] = 0.0; }
    h.vscale = 0.0;
    return;
  }
  const double nrm = sqrt(alpha * alpha + xnorm2);
  const double beta = (alpha >= 0.0) ? -nrm : nrm;
  const double tau = (beta - alpha) / beta;
  h.vscale = 1.0 / (alpha - beta);           // v = (1, A[i][k] * vscale)
  h.rrow[k] = beta; h.tw[k] = 0.0;
  for (int j = k + 1; j <= N; ++j) {
    const double wj = row[j] + sums[j] * h.vscale;     // v^T A[:, j]
    h.tw[j] = tau * wj;
    h.rrow[j] = row[j] - h.tw[j];
  }
}

}  // namespace srfit
