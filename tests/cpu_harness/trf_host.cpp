// TEST INFRASTRUCTURE ONLY.  Compiles the scalar solver core of K5 (spinrelax_b200/csrc/trf_core.cuh, the code the
// CUDA kernel runs on its leader thread) with g++ and drives it with serial stand-ins for the CTA-parallel parts
// (residuals, Jacobian, Householder QR), so the solver logic can be compared with SciPy's TRF on a machine without a
// GPU.  Never loaded by the spinrelax_b200 package.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../spinrelax_b200/csrc/trf_core.cuh"
#include "../../spinrelax_b200/csrc/fit_model.cuh"

namespace {

// Householder QR of the M x (N+1) column-major matrix A (last column = right-hand side); returns R (row major, upper)
// and the first N entries of Q^T rhs
template <int N>
void qr_host(std::vector<double>& A, int M, double* R, double* qtf) {
  for (int k = 0; k < N; ++k) {
    double sums[N + 1];
    for (int j = k; j <= N; ++j) {
      double s = 0.0;
      for (int i = k + 1; i < M; ++i) s += A[(size_t)k * M + i] * A[(size_t)j * M + i];
      sums[j] = s;
    }
    srfit::HouseholderCol<N> h;
    double row[N + 1];
    for (int j = k; j <= N; ++j) row[j] = A[(size_t)j * M + k];
    srfit::householder_column<N>(k, row, sums, h);
    for (int j = k; j < N; ++j) R[k * N + j] = h.rrow[j];
    for (int j = 0; j < k; ++j) R[k * N + j] = 0.0;
    qtf[k] = h.rrow[N];
    for (int i = k + 1; i < M; ++i) {
      const double vi = A[(size_t)k * M + i] * h.vscale;
      for (int j = k + 1; j <= N; ++j) A[(size_t)j * M + i] -= h.tw[j] * vi;
    }
  }
}

// the warp-cooperative SVD of trf_core.cuh, its 32 lanes run one after the other (phases separated as on the device)
template <int N>
void svd_host(srtrf::Core<N>& c, const double* qtf) {
  double W[N * N];
  const bool warm = (c.svd_calls % srtrf::kSvdRestart) != 0;
  c.svd_calls += 1;
  for (int lane = 0; lane < 32; ++lane) srtrf::svd_init<N>(c, W, lane, warm);
  for (int sweep = 0; sweep < srtrf::kSvdMaxSweeps; ++sweep) {
    bool rotated = false;
    for (int round = 0; round < srtrf::svd_rounds<N>(); ++round) {
      srtrf::SvdRot rot[32];
      for (int lane = 0; lane < 32; ++lane) rot[lane] = srtrf::svd_pair<N>(W, round, lane);
      for (int lane = 0; lane < 32; ++lane) { srtrf::svd_apply<N>(rot[lane], W, c.V, lane); rotated = rotated || rot[lane].p >= 0; }
    }
    if (!rotated) break;
  }
  for (int lane = 0; lane < 32; ++lane) srtrf::svd_values<N>(c, W, qtf, lane);
  srtrf::svd_finish<N>(c);
}

template <int N>
void fit_one(const double* t, const double* y, const double* sig, int L, const double* p0, const double* lo,
             const double* hi, int max_nfev, double* popt, double* Rout, double* cost, int* status) {
  srtrf::Core<N> c;
  std::vector<double> w(L), f(L), fn(L), J((size_t)L * N);
  for (int k = 0; k < L; ++k) w[k] = sig ? 1.0 / sig[k] : 1.0;
  bool feasible = true;
  for (int i = 0; i < N; ++i) {
    c.x[i] = p0[i]; c.lb[i] = lo[i]; c.ub[i] = hi[i];
    feasible = feasible && p0[i] >= lo[i] && p0[i] <= hi[i];
  }
  status[0] = -3; status[1] = 0; *cost = INFINITY;
  for (int i = 0; i < N; ++i) popt[i] = p0[i];
  for (int i = 0; i < N * N; ++i) Rout[i] = 0.0;
  if (!feasible) return;                                    // least_squares: "`x0` is infeasible" (ValueError)
  srtrf::make_strictly_feasible<N>(c.x, c.lb, c.ub, 1e-10);
  auto eval = [&](const double* x, std::vector<double>& r, double* Jm, double* g) {
    double cst = 0.0;
    srfit::ModelPars<N> mp; srfit::prepare<N>(x, mp);
    if (g) for (int i = 0; i < N; ++i) g[i] = 0.0;
    for (int k = 0; k < L; ++k) {
      double row[N];
      r[k] = srfit::residual_and_row<N>(mp, t[k], y[k], w[k], Jm ? row : nullptr);
      cst += r[k] * r[k];
      if (Jm) for (int i = 0; i < N; ++i) { Jm[(size_t)i * L + k] = row[i]; g[i] += row[i] * r[k]; }
    }
    return 0.5 * cst;
  };
  c.cost = eval(c.x, f, J.data(), c.g);
  if (!std::isfinite(c.cost)) { status[0] = -4; return; }   // "Residuals are not finite in the initial point"
  srtrf::begin<N>(c, L, max_nfev, 1e-8, 1e-8, 1e-8);
  const int M = L + N;
  std::vector<double> A((size_t)M * (N + 1));
  double cost_new = 0.0;
  while (srtrf::outer_begin<N>(c)) {
    for (int j = 0; j < N; ++j) {
      for (int k = 0; k < L; ++k) A[(size_t)j * M + k] = J[(size_t)j * L + k] * c.d[j];
      for (int i = 0; i < N; ++i) A[(size_t)j * M + L + i] = (i == j) ? std::sqrt(c.diag_h[j]) : 0.0;
    }
    for (int k = 0; k < L; ++k) A[(size_t)N * M + k] = f[k];
    for (int i = 0; i < N; ++i) A[(size_t)N * M + L + i] = 0.0;
    double qtf[N];
    qr_host<N>(A, M, c.R, qtf);
    svd_host<N>(c, qtf);
    while (srtrf::inner_propose<N>(c)) {
      cost_new = eval(c.x_new, fn, nullptr, nullptr);
      if (srtrf::inner_judge<N>(c, cost_new, std::isfinite(cost_new))) break;
    }
    if (srtrf::outer_end<N>(c, cost_new)) c.cost = eval(c.x, f, J.data(), c.g);
  }
  // R factor of the unscaled Jacobian at the solution (what curve_fit takes the SVD of for pcov)
  std::vector<double> B((size_t)L * (N + 1));
  std::memcpy(B.data(), J.data(), sizeof(double) * (size_t)L * N);
  for (int k = 0; k < L; ++k) B[(size_t)N * L + k] = f[k];
  double qtf[N];
  qr_host<N>(B, L, Rout, qtf);
  for (int i = 0; i < N; ++i) popt[i] = c.x[i];
  *cost = c.cost;
  status[0] = c.status; status[1] = c.nfev;
}

}  // namespace

extern "C" int trf_host_fit(const double* t, const double* y, const double* sig, int nR, int L, int nP, const double* p0,
                            const double* lo, const double* hi, int max_nfev, double* popt, double* R, double* cost,
                            int* status) {
  for (int r = 0; r < nR; ++r) {
    const double* tr = t + (size_t)r * L; const double* yr = y + (size_t)r * L;
    const double* sr = sig ? sig + (size_t)r * L : nullptr;
    const size_t o = (size_t)r * nP;
#define CASE(NP) case NP: fit_one<NP>(tr, yr, sr, L, p0 + o, lo + o, hi + o, max_nfev > 0 ? max_nfev : 100 * NP, popt + o, \
                                      R + o * nP, cost + r, status + 2 * r); break;
    switch (nP) { CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) default: return -1; }
#undef CASE
  }
  return 0;
}
