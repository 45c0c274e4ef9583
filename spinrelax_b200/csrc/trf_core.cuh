// Scalar core of the bounded trust-region-reflective least-squares solver used by K5 (csrc/fit.cu).
//
// What it restates.  The reference fits every C(t) curve with scipy.optimize.curve_fit(..., bounds=...)
// (fitting_Ct_functions.py:322-324), i.e. SciPy's least_squares(method='trf', tr_solver='exact', x_scale=1,
// ftol=xtol=gtol=1e-8, max_nfev=100*n) -- SURVEY Appendix B; SciPy 1.18.1 in this image, unpinned upstream
// (requirements.txt:2).  The *result* of that solver (where it stops, whether it reports success) decides the
// reference's model-selection ladder, so the kernel follows the same published algorithm step for step
// (Branch, Coleman & Li 1999; More' 1978 for the exact subproblem; as laid out in scipy/optimize/_lsq/trf.py
// `trf_bounds` and _lsq/common.py): Coleman-Li scaling vector, augmented Jacobian, trust-region subproblem from the
// singular values of the augmented Jacobian, reflected / constrained-Cauchy step selection, radius update and the
// ftol/xtol/gtol tests.  The O(L) work (residuals, Jacobian, Householder QR of the augmented Jacobian) is done by the
// whole CTA in fit.cu; everything here is O(n^3) with n <= 9 and runs on one thread.  The file has no CUDA-only
// construct so that tests/cpu_harness can compile the very same code with g++ and compare it with SciPy on a CPU.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SR_HD __host__ __device__
#else
#define SR_HD
#endif

namespace srtrf {

constexpr double kEps = 2.220446049250313e-16;

template <int N>
struct Core {
  // current point
  double x[N], lb[N], ub[N], g[N];
  double cost;
  // Coleman-Li scaling at x
  double v[N], dv[N], d[N], diag_h[N], g_h[N];
  double g_norm, theta;
  // subproblem in the scaled ("hat") variables: R from the QR of [J d; diag(sqrt(diag_h))], then its SVD
  double R[N * N];      // upper triangle, row major
  double s[N], V[N * N], suf[N];   // singular values, right vectors (columns), s_i * (U^T f)_i
  int full_rank;
  double s_max, s_min;
  double Delta, alpha;
  // trial step
  double step[N], step_h[N], x_new[N];
  double predicted, step_h_norm, step_norm, actual;
  int nfev, max_nfev, status, iteration, m;
  int svd_calls;        // decompositions since the start of the solve (warm-start bookkeeping)
  double ftol, xtol, gtol;
};

template <int N>
SR_HD double norm2(const double* a) {
  double s = 0.0;
  for (int i = 0; i < N; ++i) s += a[i] * a[i];
  return sqrt(s);
}

template <int N>
SR_HD double dot(const double* a, const double* b) {
  double s = 0.0;
  for (int i = 0; i < N; ++i) s += a[i] * b[i];
  return s;
}

// y = R s for the upper-triangular R (row major); returns ||R s||^2 = ||J_h s||^2 + s . diag_h . s
template <int N>
SR_HD void rmul(const double* R, const double* s, double* y) {
  for (int i = 0; i < N; ++i) {
    double a = 0.0;
    for (int j = i; j < N; ++j) a += R[i * N + j] * s[j];
    y[i] = a;
  }
}

// scipy _lsq/common.py make_strictly_feasible
template <int N>
SR_HD void make_strictly_feasible(double* x, const double* lb, const double* ub, double rstep) {
  for (int i = 0; i < N; ++i) {
    int active = 0;
    if (rstep == 0.0) {
      if (x[i] <= lb[i]) active = -1;
      if (x[i] >= ub[i]) active = 1;
    } else {
      const double lower_dist = x[i] - lb[i], upper_dist = ub[i] - x[i];
      const double lt = rstep * fmax(1.0, fabs(lb[i])), ut = rstep * fmax(1.0, fabs(ub[i]));
      if (isfinite(lb[i]) && lower_dist <= fmin(upper_dist, lt)) active = -1;
      if (isfinite(ub[i]) && upper_dist <= fmin(lower_dist, ut)) active = 1;
    }
    if (active == -1) x[i] = (rstep == 0.0) ? nextafter(lb[i], ub[i]) : lb[i] + rstep * fmax(1.0, fabs(lb[i]));
    if (active == 1) x[i] = (rstep == 0.0) ? nextafter(ub[i], lb[i]) : ub[i] - rstep * fmax(1.0, fabs(ub[i]));
    if (x[i] < lb[i] || x[i] > ub[i]) x[i] = 0.5 * (lb[i] + ub[i]);
  }
}

// CL_scaling_vector, and the derived hat-space quantities of one outer iteration (x_scale = 1)
template <int N>
SR_HD void scaling(Core<N>& c) {
  double gn = 0.0;
  for (int i = 0; i < N; ++i) {
    double v = 1.0, dv = 0.0;
    if (c.g[i] < 0.0 && isfinite(c.ub[i])) { v = c.ub[i] - c.x[i]; dv = -1.0; }
    if (c.g[i] > 0.0 && isfinite(c.lb[i])) { v = c.x[i] - c.lb[i]; dv = 1.0; }
    c.v[i] = v; c.dv[i] = dv;
    gn = fmax(gn, fabs(c.g[i] * v));
    c.d[i] = sqrt(v);
    c.diag_h[i] = c.g[i] * dv;
    c.g_h[i] = c.d[i] * c.g[i];
  }
  c.g_norm = gn;
  c.theta = fmax(0.995, 1.0 - gn);
}

// One-sided (Hestenes) Jacobi SVD of the n x n upper-triangular R:  R V = W with orthogonal columns,
// s_i = ||W_i||, and s_i (U^T qtf)_i = W_i . qtf.  Accurate to eps * s_max like LAPACK's driver; the order of the
// singular values is irrelevant to every use below except their extremes.
//
// The sweep is organised for one warp: a round-robin tournament pairs the columns so that the (up to four) column
// pairs of a round are disjoint and rotate at the same time, one group of eight lanes per pair, each lane updating
// one row of W and V (lane 0 of the group also row 8).  Every round has a read phase (svd_pair: the lanes of a group
// form the same three column dot products and the same rotation) and a write phase (svd_apply); the caller puts a
// warp barrier after each.  tests/cpu_harness runs the same two functions lane by lane.
//
// Warm start: consecutive outer iterations decompose nearly the same matrix, so the sweep starts from W = R V_prev
// (V_prev = the right vectors of the previous iteration, orthogonal by construction) and needs 2-3 sweeps instead of
// 6-8.  Every kSvdRestart-th call starts again from V = I so that rounding in V cannot accumulate over a long solve.
constexpr int kSvdRestart = 24;

template <int N>
SR_HD void svd_init(Core<N>& c, double* W, int lane, bool warm) {
  if (!warm) {
    for (int j = lane; j < N; j += 32)
      for (int i = 0; i < N; ++i) {
        W[j * N + i] = (i <= j) ? c.R[i * N + j] : 0.0;     // column major: W[j*N + i]
        c.V[j * N + i] = (i == j) ? 1.0 : 0.0;              // V column j
      }
    return;
  }
  for (int e = lane; e < N * N; e += 32) {                  // W[:, j] = R V_prev[:, j], R upper triangular
    const int j = e / N, i = e - j * N;
    double a = 0.0;
    for (int k = i; k < N; ++k) a += c.R[i * N + k] * c.V[j * N + k];
    W[j * N + i] = a;
  }
}

constexpr int kSvdMaxSweeps = 60;

template <int N>
SR_HD int svd_rounds() { return (N % 2 == 0) ? N - 1 : N; }

struct SvdRot { int p, q; double cs, sn; };    // p < 0: nothing to do for this lane's group

// read phase of round `round`: the rotation for the column pair of this lane's group
template <int N>
SR_HD SvdRot svd_pair(const double* W, int round, int lane) {
  constexpr int ne = (N % 2 == 0) ? N : N + 1;            // players of the tournament (one dummy when N is odd)
  constexpr int m = ne - 1;
  SvdRot r; r.p = -1; r.q = -1; r.cs = 1.0; r.sn = 0.0;
  const int g = lane >> 3;
  // pair k of the round: k = 0 -> (ne - 1, round), k >= 1 -> ((round + k) % m, (round - k + m) % m); the pair holding
  // the dummy column is skipped, the others are numbered 0, 1, 2, 3 for the four lane groups
  int k = (N % 2 == 0) ? g : g + 1;
  if (k >= ne / 2) return r;
  int a = (k == 0) ? ne - 1 : (round + k) % m;
  int b = (k == 0) ? round : (round - k + m) % m;
  const int p = a < b ? a : b, q = a < b ? b : a;
  double aa = 0.0, bb = 0.0, gg = 0.0;
  for (int i = 0; i < N; ++i) {
    const double wp = W[p * N + i], wq = W[q * N + i];
    aa += wp * wp; bb += wq * wq; gg += wp * wq;
  }
  if (gg == 0.0 || fabs(gg) <= kEps * sqrt(aa * bb)) return r;
  const double zeta = (bb - aa) / (2.0 * gg);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  r.cs = 1.0 / sqrt(1.0 + t * t);
  r.sn = r.cs * t;
  r.p = p; r.q = q;
  return r;
}

// write phase: this lane's rows of the two columns, in W and in V
template <int N>
SR_HD void svd_apply(const SvdRot& r, double* W, double* V, int lane) {
  if (r.p < 0) return;
  const int row0 = lane & 7;
  for (int i = row0; i < N; i += 8) {
    const double wp = W[r.p * N + i], wq = W[r.q * N + i];
    W[r.p * N + i] = r.cs * wp - r.sn * wq;
    W[r.q * N + i] = r.sn * wp + r.cs * wq;
    const double vp = V[r.p * N + i], vq = V[r.q * N + i];
    V[r.p * N + i] = r.cs * vp - r.sn * vq;
    V[r.q * N + i] = r.sn * vp + r.cs * vq;
  }
}

// after the sweeps: singular values and s * (U^T qtf), one column per lane; then the extremes (lane 0)
template <int N>
SR_HD void svd_values(Core<N>& c, const double* W, const double* qtf, int lane) {
  for (int j = lane; j < N; j += 32) {
    double a = 0.0, u = 0.0;
    for (int i = 0; i < N; ++i) { a += W[j * N + i] * W[j * N + i]; u += W[j * N + i] * qtf[i]; }
    c.s[j] = sqrt(a);
    c.suf[j] = u;
  }
}

template <int N>
SR_HD void svd_finish(Core<N>& c) {
  double smax = 0.0, smin = INFINITY;
  for (int j = 0; j < N; ++j) { smax = fmax(smax, c.s[j]); smin = fmin(smin, c.s[j]); }
  c.s_max = smax; c.s_min = smin;
  c.full_rank = (c.m >= N) && (smin > kEps * c.m * smax);
}

// solve_lsq_trust_region (More' iteration on alpha with the SVD); p_h returned, c.alpha updated
template <int N>
SR_HD void solve_tr(Core<N>& c, double* p_h) {
  const double Delta = c.Delta;
  double tmp[N];
  if (c.full_rank) {
    for (int j = 0; j < N; ++j) tmp[j] = c.suf[j] / (c.s[j] * c.s[j]);       // uf / s
    for (int i = 0; i < N; ++i) {
      double a = 0.0;
      for (int j = 0; j < N; ++j) a += c.V[j * N + i] * tmp[j];
      p_h[i] = -a;
    }
    if (norm2<N>(p_h) <= Delta) { c.alpha = 0.0; return; }
  }
  double alpha_upper = norm2<N>(c.suf) / Delta;
  double alpha_lower = 0.0;
  auto phi_and_derivative = [&](double al, double& phi, double& phi_prime) {
    double pn2 = 0.0, dsum = 0.0;
    for (int j = 0; j < N; ++j) {
      const double denom = c.s[j] * c.s[j] + al;
      const double q = c.suf[j] / denom;
      pn2 += q * q;
      dsum += c.suf[j] * c.suf[j] / (denom * denom * denom);
    }
    const double p_norm = sqrt(pn2);
    phi = p_norm - Delta;
    phi_prime = -dsum / p_norm;
  };
  if (c.full_rank) {
    double phi, phi_prime;
    phi_and_derivative(0.0, phi, phi_prime);
    alpha_lower = -phi / phi_prime;
  }
  double alpha = c.alpha;
  if (!c.full_rank && alpha == 0.0) alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
  for (int it = 0; it < 10; ++it) {
    if (alpha < alpha_lower || alpha > alpha_upper) alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
    double phi, phi_prime;
    phi_and_derivative(alpha, phi, phi_prime);
    if (phi < 0.0) alpha_upper = alpha;
    const double ratio = phi / phi_prime;
    alpha_lower = fmax(alpha_lower, alpha - ratio);
    alpha -= (phi + Delta) * ratio / Delta;
    if (fabs(phi) < 0.01 * Delta) break;
  }
  for (int j = 0; j < N; ++j) tmp[j] = c.suf[j] / (c.s[j] * c.s[j] + alpha);
  for (int i = 0; i < N; ++i) {
    double a = 0.0;
    for (int j = 0; j < N; ++j) a += c.V[j * N + i] * tmp[j];
    p_h[i] = -a;
  }
  const double f = Delta / norm2<N>(p_h);
  for (int i = 0; i < N; ++i) p_h[i] *= f;
  c.alpha = alpha;
}

// step_size_to_bound: smallest positive multiple of s that reaches a bound, and which bounds are hit (signed)
template <int N>
SR_HD double step_size_to_bound(const double* x, const double* s, const double* lb, const double* ub, int* hits) {
  double steps[N], mn = INFINITY;
  for (int i = 0; i < N; ++i) {
    steps[i] = INFINITY;
    if (s[i] != 0.0) steps[i] = fmax((lb[i] - x[i]) / s[i], (ub[i] - x[i]) / s[i]);
    mn = fmin(mn, steps[i]);
  }
  if (hits)
    for (int i = 0; i < N; ++i) hits[i] = (steps[i] == mn) ? (s[i] > 0.0 ? 1 : (s[i] < 0.0 ? -1 : 0)) : 0;
  return mn;
}

SR_HD inline void minimize_quadratic_1d(double a, double b, double lo, double hi, double c0, double& t_best, double& y_best) {
  double t[3] = {lo, hi, 0.0};
  int nt = 2;
  if (a != 0.0) {
    const double ext = -0.5 * b / a;
    if (lo < ext && ext < hi) t[nt++] = ext;
  }
  t_best = t[0]; y_best = t[0] * (a * t[0] + b) + c0;
  for (int i = 1; i < nt; ++i) {
    const double y = t[i] * (a * t[i] + b) + c0;
    if (y < y_best) { y_best = y; t_best = t[i]; }       // argmin keeps the first minimum
  }
}

// select_step of trf.py: the trust-region step if it stays inside the box, otherwise the best of the truncated
// step, its reflection at the first bound hit and the constrained Cauchy step.  Fills c.step, c.step_h, c.predicted.
template <int N>
SR_HD void select_step(Core<N>& c, double* p_h) {
  double p[N], y[N];
  bool inside = true;
  for (int i = 0; i < N; ++i) {
    p[i] = c.d[i] * p_h[i];
    const double xn = c.x[i] + p[i];
    inside = inside && (xn >= c.lb[i]) && (xn <= c.ub[i]);
  }
  if (inside) {
    rmul<N>(c.R, p_h, y);
    const double val = 0.5 * dot<N>(y, y) + dot<N>(p_h, c.g_h);
    for (int i = 0; i < N; ++i) { c.step[i] = p[i]; c.step_h[i] = p_h[i]; }
    c.predicted = -val;
    return;
  }
  int hits[N];
  const double p_stride = step_size_to_bound<N>(c.x, p, c.lb, c.ub, hits);
  double r_h[N], r[N], x_on_bound[N];
  for (int i = 0; i < N; ++i) {
    r_h[i] = hits[i] ? -p_h[i] : p_h[i];
    r[i] = c.d[i] * r_h[i];
    p[i] *= p_stride; p_h[i] *= p_stride;
    x_on_bound[i] = c.x[i] + p[i];
  }
  // intersect_trust_region(p_h, r_h, Delta): larger root t2 of ||p_h + t r_h|| = Delta
  double to_tr;
  {
    const double a = dot<N>(r_h, r_h), b = dot<N>(p_h, r_h), cc = dot<N>(p_h, p_h) - c.Delta * c.Delta;
    const double disc = sqrt(b * b - a * cc);
    const double q = -(b + copysign(disc, b));
    const double t1 = q / a, t2 = cc / q;
    to_tr = (t1 < t2) ? t2 : t1;
  }
  const double to_bound = step_size_to_bound<N>(x_on_bound, r, c.lb, c.ub, nullptr);
  const double theta = c.theta;
  double r_stride = fmin(to_bound, to_tr), r_stride_l, r_stride_u;
  if (r_stride > 0.0) {
    r_stride_l = (1.0 - theta) * p_stride / r_stride;
    r_stride_u = (r_stride == to_bound) ? theta * to_bound : to_tr;
  } else {
    r_stride_l = 0.0; r_stride_u = -1.0;
  }
  double r_value = INFINITY;
  double Rp[N], Rr[N];
  if (r_stride_l <= r_stride_u) {
    // build_quadratic_1d(J_h, g_h, r_h, s0 = p_h, diag = diag_h) through R
    rmul<N>(c.R, r_h, Rr);
    rmul<N>(c.R, p_h, Rp);
    const double a = 0.5 * dot<N>(Rr, Rr);
    const double b = dot<N>(c.g_h, r_h) + dot<N>(Rp, Rr);
    const double c0 = 0.5 * dot<N>(Rp, Rp) + dot<N>(c.g_h, p_h);
    minimize_quadratic_1d(a, b, r_stride_l, r_stride_u, c0, r_stride, r_value);
    for (int i = 0; i < N; ++i) { r_h[i] = r_h[i] * r_stride + p_h[i]; r[i] = r_h[i] * c.d[i]; }
  }
  // truncated step made strictly interior
  for (int i = 0; i < N; ++i) { p[i] *= theta; p_h[i] *= theta; }
  rmul<N>(c.R, p_h, Rp);
  const double p_value = 0.5 * dot<N>(Rp, Rp) + dot<N>(p_h, c.g_h);
  // constrained Cauchy step along the anti-gradient
  double ag_h[N], ag[N];
  for (int i = 0; i < N; ++i) { ag_h[i] = -c.g_h[i]; ag[i] = c.d[i] * ag_h[i]; }
  const double ag_to_tr = c.Delta / norm2<N>(ag_h);
  const double ag_to_bound = step_size_to_bound<N>(c.x, ag, c.lb, c.ub, nullptr);
  double ag_stride = (ag_to_bound < ag_to_tr) ? theta * ag_to_bound : ag_to_tr;
  double ag_value;
  {
    rmul<N>(c.R, ag_h, Rr);
    const double a = 0.5 * dot<N>(Rr, Rr), b = dot<N>(c.g_h, ag_h);
    minimize_quadratic_1d(a, b, 0.0, ag_stride, 0.0, ag_stride, ag_value);
  }
  for (int i = 0; i < N; ++i) { ag_h[i] *= ag_stride; ag[i] *= ag_stride; }
  const double* bs; const double* bsh; double val;
  if (p_value < r_value && p_value < ag_value) { bs = p; bsh = p_h; val = p_value; }
  else if (r_value < p_value && r_value < ag_value) { bs = r; bsh = r_h; val = r_value; }
  else { bs = ag; bsh = ag_h; val = ag_value; }
  for (int i = 0; i < N; ++i) { c.step[i] = bs[i]; c.step_h[i] = bsh[i]; }
  c.predicted = -val;
}

// ---- the driver, split at the points where the CTA has to do O(L) work ------------------------------------------

// after the first residual/Jacobian evaluation at the strictly feasible x0 (c.x, c.g, c.cost set by the caller)
template <int N>
SR_HD void begin(Core<N>& c, int m, int max_nfev, double ftol, double xtol, double gtol) {
  c.m = m; c.max_nfev = max_nfev; c.ftol = ftol; c.xtol = xtol; c.gtol = gtol;
  c.nfev = 1; c.status = -1; c.iteration = 0; c.alpha = 0.0; c.svd_calls = 0;
  scaling<N>(c);
  double a = 0.0;
  for (int i = 0; i < N; ++i) { const double q = c.x[i] / sqrt(c.v[i]); a += q * q; }
  c.Delta = sqrt(a);
  if (c.Delta == 0.0) c.Delta = 1.0;
}

// top of the outer loop: returns false when the solve is over (c.status final), true when the caller must now build
// R and qtf from [J d; diag(sqrt(diag_h))] (c.d, c.diag_h valid) and call svd_setup
template <int N>
SR_HD bool outer_begin(Core<N>& c) {
  scaling<N>(c);
  if (c.g_norm < c.gtol) c.status = 1;
  if (c.status != -1 || c.nfev == c.max_nfev) {
    if (c.status == -1) c.status = 0;
    return false;
  }
  c.actual = -1.0;
  return true;
}

// inner loop head: false = inner loop finished without a trial (nfev exhausted or reduction achieved)
template <int N>
SR_HD bool inner_propose(Core<N>& c) {
  if (!(c.actual <= 0.0 && c.nfev < c.max_nfev)) return false;
  double p_h[N];
  solve_tr<N>(c, p_h);
  select_step<N>(c, p_h);
  for (int i = 0; i < N; ++i) c.x_new[i] = c.x[i] + c.step[i];
  make_strictly_feasible<N>(c.x_new, c.lb, c.ub, 0.0);
  return true;
}

// after the residuals at x_new: returns true if the inner loop must stop because a termination test fired
template <int N>
SR_HD bool inner_judge(Core<N>& c, double cost_new, bool finite_new) {
  c.nfev += 1;
  c.step_h_norm = norm2<N>(c.step_h);
  if (!finite_new) { c.Delta = 0.25 * c.step_h_norm; return false; }
  c.actual = c.cost - cost_new;
  // update_tr_radius
  double ratio;
  if (c.predicted > 0.0) ratio = c.actual / c.predicted;
  else if (c.predicted == 0.0 && c.actual == 0.0) ratio = 1.0;
  else ratio = 0.0;
  double Delta_new = c.Delta;
  if (ratio < 0.25) Delta_new = 0.25 * c.step_h_norm;
  else if (ratio > 0.75 && c.step_h_norm > 0.95 * c.Delta) Delta_new = c.Delta * 2.0;
  c.step_norm = norm2<N>(c.step);
  // check_termination
  const bool f_ok = (c.actual < c.ftol * c.cost) && (ratio > 0.25);
  const bool x_ok = c.step_norm < c.xtol * (c.xtol + norm2<N>(c.x));
  if (f_ok && x_ok) c.status = 4;
  else if (f_ok) c.status = 2;
  else if (x_ok) c.status = 3;
  if (c.status != -1) return true;
  c.alpha *= c.Delta / Delta_new;
  c.Delta = Delta_new;
  return false;
}

// end of an outer iteration: true = step accepted (caller re-evaluates J and g at c.x, cost already moved)
template <int N>
SR_HD bool outer_end(Core<N>& c, double cost_new) {
  c.iteration += 1;
  if (c.actual > 0.0) {
    for (int i = 0; i < N; ++i) c.x[i] = c.x_new[i];
    c.cost = cost_new;
    return true;
  }
  c.step_norm = 0.0; c.actual = 0.0;
  return false;
}

}  // namespace srtrf
