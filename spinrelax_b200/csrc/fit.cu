// K5: batched bounded least-squares fit of C(t) = S2 + sum_i C_i exp(-t/tau_i), one CTA per residue.
// Replaces the scipy.optimize.curve_fit call of autoCorrelationModel.conduct_curve_fitting
// (fitting_Ct_functions.py:306-345; model curvefit_exponential :419-427; bounds :412-416).
//
// The reference's model-selection ladder (:278-304) depends on *where SciPy's solver stops* and on whether it
// reports success (a RuntimeError from curve_fit is swallowed at :325-328 and ends the ladder), so the kernel runs
// the same algorithm as SciPy's least_squares(method='trf', tr_solver='exact', x_scale=1, ftol=xtol=gtol=1e-8,
// max_nfev=100 n): see trf_core.cuh for the solver logic (one thread) and fit_model.cuh for the model arithmetic.
// This file holds the O(L) parts, done by the whole CTA:
//   * residuals (1/sigma)(f - y) and the analytic Jacobian, written column-major into the work matrix A (shared
//     memory when the curve fits, a global workspace otherwise), with the gradient J^T f and the cost block-reduced
//     in the same pass;
//   * Householder QR of the augmented Jacobian [J d; diag(sqrt(diag_h))] with the residuals as an extra column:
//     one block reduction per column (the sums of A[:,k] . A[:,j] for all j >= k give the column norm and every
//     v^T A[:,j] at once), each thread updating the rows it owns -- 9 barriers for a 9-parameter fit;
//   * at the solution, the QR of the unscaled Jacobian: its R factor has the singular values and right vectors of
//     J, from which the host forms pcov exactly as curve_fit does from svd(J).
#include "common.cuh"
#include "fit_model.cuh"

namespace {

constexpr int kMaxP = 9;
constexpr int kFitThreads = 128;
constexpr int kFitWarps = kFitThreads / 32;
constexpr size_t kFitSmemLimit = 200 * 1024;

enum : int { kFlagStop = 0, kFlagQR = 1, kFlagTrial = 2, kFlagAccept = 3, kFlagReject = 4 };

template <int N>
struct FitShared {
  srtrf::Core<N> core;
  double red[2][kFitWarps][N + 2];
  double qtf[N];
  double W[N * N];          // working copy of R for the Jacobi SVD (column major)
  int flag;
};

// SVD of core.R by warp 0 (all 32 lanes call this); results in core.s, core.V, core.suf, core.full_rank
template <int N>
__device__ void warp_svd(FitShared<N>& sh) {
  const int lane = threadIdx.x & 31;
  const bool warm = (sh.core.svd_calls % srtrf::kSvdRestart) != 0;
  __syncwarp();
  if (lane == 0) sh.core.svd_calls += 1;
  srtrf::svd_init<N>(sh.core, sh.W, lane, warm);
  __syncwarp();
  for (int sweep = 0; sweep < srtrf::kSvdMaxSweeps; ++sweep) {
    bool rotated = false;
    for (int round = 0; round < srtrf::svd_rounds<N>(); ++round) {
      const srtrf::SvdRot rot = srtrf::svd_pair<N>(sh.W, round, lane);
      __syncwarp();
      srtrf::svd_apply<N>(rot, sh.W, sh.core.V, lane);
      rotated = rotated || rot.p >= 0;
      __syncwarp();
    }
    if (!__any_sync(0xffffffffu, rotated)) break;
  }
  srtrf::svd_values<N>(sh.core, sh.W, sh.qtf, lane);
  __syncwarp();
  if (lane == 0) srtrf::svd_finish<N>(sh.core);
  __syncwarp();
}

// block-wide sum of NV per-thread values; every thread gets all NV totals.  `parity` alternates between calls so
// that one barrier per reduction is enough.
template <int N, int NV>
__device__ __forceinline__ void block_sum(double* acc, FitShared<N>& sh, int& parity) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double v = sr_warp_sum(acc[i]);
    if (lane == 0) sh.red[parity][warp][i] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kFitWarps; ++w) s += sh.red[parity][w][i];
    acc[i] = s;
  }
  parity ^= 1;
}

template <int N>
struct Problem {
  const double* t; const double* y; const double* w;     // w = 1/sigma (shared-memory copy) or nullptr
  const double* sig;                                       // global sigma, used when w == nullptr
  int L;
  __device__ __forceinline__ double weight(int k) const { return w ? w[k] : (sig ? 1.0 / sig[k] : 1.0); }
};

// residuals and Jacobian at x into A (column-major, pitch Mp; column N = residuals); returns cost, fills g
// CHI: also sum (f - y)^2 / sigma = r^2 / w, the quantity the reference averages into its chi^2 (calc_chiSq,
// fitting_Ct_functions.py:272-276, quirk G6: divided by sigma, not sigma^2)
template <int N, bool CHI = false>
__device__ void eval_jac(const Problem<N>& pb, const double* x, double* A, int Mp, FitShared<N>& sh, int& parity,
                         double* g, double& cost, double* chi_sum = nullptr) {
  srfit::ModelPars<N> mp;
  srfit::prepare<N>(x, mp);
  double acc[N + 2];
#pragma unroll
  for (int i = 0; i < N + 2; ++i) acc[i] = 0.0;
  for (int k = threadIdx.x; k < pb.L; k += kFitThreads) {
    double row[N];
    const double w = pb.weight(k);
    const double r = srfit::residual_and_row<N>(mp, pb.t[k], pb.y[k], w, row);
#pragma unroll
    for (int i = 0; i < N; ++i) { A[(size_t)i * Mp + k] = row[i]; acc[i] += row[i] * r; }
    A[(size_t)N * Mp + k] = r;
    acc[N] += r * r;
    if (CHI) acc[N + 1] += r * r / w;
  }
  block_sum<N, CHI ? N + 2 : N + 1>(acc, sh, parity);
#pragma unroll
  for (int i = 0; i < N; ++i) g[i] = acc[i];
  cost = 0.5 * acc[N];
  if (CHI) *chi_sum = acc[N + 1];
}

template <int N>
__device__ double eval_cost(const Problem<N>& pb, const double* x, FitShared<N>& sh, int& parity) {
  srfit::ModelPars<N> mp;
  srfit::prepare<N>(x, mp);
  double acc[1] = {0.0};
  for (int k = threadIdx.x; k < pb.L; k += kFitThreads) {
    const double r = srfit::residual_and_row<N>(mp, pb.t[k], pb.y[k], pb.weight(k), nullptr);
    acc[0] += r * r;
  }
  block_sum<N, 1>(acc, sh, parity);
  return 0.5 * acc[0];
}

// Householder QR of the M x (N+1) matrix A; thread 0 stores R (row major) and the first N entries of Q^T rhs
template <int N>
__device__ void block_qr(double* A, int M, int Mp, FitShared<N>& sh, int& parity, double* R, double* qtf) {
  for (int k = 0; k < N; ++k) {
    double sums[N + 1];
#pragma unroll
    for (int j = 0; j <= N; ++j) sums[j] = 0.0;
    int i0 = threadIdx.x;
    while (i0 <= k) i0 += kFitThreads;
    for (int i = i0; i < M; i += kFitThreads) {
      const double aik = A[(size_t)k * Mp + i];
#pragma unroll
      for (int j = 0; j <= N; ++j)
        if (j >= k) sums[j] += aik * A[(size_t)j * Mp + i];
    }
    block_sum<N, N + 1>(sums, sh, parity);
    double row[N + 1];
#pragma unroll
    for (int j = 0; j <= N; ++j) row[j] = (j >= k) ? A[(size_t)j * Mp + k] : 0.0;
    srfit::HouseholderCol<N> h;
    srfit::householder_column<N>(k, row, sums, h);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int j = 0; j < N; ++j) R[k * N + j] = (j >= k) ? h.rrow[j] : 0.0;
      qtf[k] = h.rrow[N];
    }
    if (h.vscale != 0.0) {
      for (int i = i0; i < M; i += kFitThreads) {
        const double vi = A[(size_t)k * Mp + i] * h.vscale;
#pragma unroll
        for (int j = 0; j <= N; ++j)
          if (j > k) A[(size_t)j * Mp + i] -= h.tw[j] * vi;
      }
    }
  }
}

template <int N>
__global__ void __launch_bounds__(kFitThreads)
ct_fit_trf_kernel(const double* __restrict__ T, const double* __restrict__ Y, const double* __restrict__ SIG, int L,
                  const double* __restrict__ P0, const double* __restrict__ LO, const double* __restrict__ HI,
                  int max_nfev, double ftol, double xtol, double gtol, double* __restrict__ POPT,
                  double* __restrict__ ROUT, double* __restrict__ COST, int* __restrict__ STATUS,
                  double* __restrict__ CHI, double* work, int in_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FitShared<N>& sh = *reinterpret_cast<FitShared<N>*>(smem_raw);
  srtrf::Core<N>& c = sh.core;
  const int r = blockIdx.x;
  const int M = L + N, Mp = M;
  Problem<N> pb;
  pb.L = L;
  pb.sig = SIG ? SIG + (size_t)r * L : nullptr;
  double* A;
  if (in_smem) {
    double* base = reinterpret_cast<double*>(smem_raw + ((sizeof(FitShared<N>) + 15) / 16) * 16);
    double* ts = base; double* ys = ts + L; double* ws = ys + L;
    A = ws + L;
    for (int k = threadIdx.x; k < L; k += kFitThreads) {
      ts[k] = T[(size_t)r * L + k];
      ys[k] = Y[(size_t)r * L + k];
      ws[k] = pb.sig ? 1.0 / pb.sig[k] : 1.0;
    }
    pb.t = ts; pb.y = ys; pb.w = ws;
  } else {
    A = work + (size_t)r * Mp * (N + 1);
    pb.t = T + (size_t)r * L; pb.y = Y + (size_t)r * L; pb.w = nullptr;
  }
  int parity = 0;
  if (threadIdx.x == 0) {
    bool feasible = true;
    for (int i = 0; i < N; ++i) {
      c.x[i] = P0[(size_t)r * N + i]; c.lb[i] = LO[(size_t)r * N + i]; c.ub[i] = HI[(size_t)r * N + i];
      feasible = feasible && (c.x[i] >= c.lb[i]) && (c.x[i] <= c.ub[i]);
    }
    if (feasible) srtrf::make_strictly_feasible<N>(c.x, c.lb, c.ub, 1e-10);
    sh.flag = feasible ? kFlagQR : kFlagStop;
  }
  __syncthreads();
  if (sh.flag == kFlagStop) {                     // least_squares raises "`x0` is infeasible": a failed fit upstream
    if (threadIdx.x == 0) {
      for (int i = 0; i < N; ++i) POPT[(size_t)r * N + i] = c.x[i];
      for (int i = 0; i < N * N; ++i) ROUT[(size_t)r * N * N + i] = 0.0;
      COST[r] = INFINITY; STATUS[2 * r] = -3; STATUS[2 * r + 1] = 0;
      if (CHI) CHI[r] = INFINITY;
    }
    return;
  }
  double g[N], cost;
  eval_jac<N>(pb, c.x, A, Mp, sh, parity, g, cost);
  if (!isfinite(cost)) {                          // "Residuals are not finite in the initial point"
    if (threadIdx.x == 0) {
      for (int i = 0; i < N; ++i) POPT[(size_t)r * N + i] = c.x[i];
      for (int i = 0; i < N * N; ++i) ROUT[(size_t)r * N * N + i] = 0.0;
      COST[r] = INFINITY; STATUS[2 * r] = -4; STATUS[2 * r + 1] = 1;
      if (CHI) CHI[r] = INFINITY;
    }
    return;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < N; ++i) c.g[i] = g[i];
    c.cost = cost;
    srtrf::begin<N>(c, L, max_nfev, ftol, xtol, gtol);
    sh.flag = srtrf::outer_begin<N>(c) ? kFlagQR : kFlagStop;
  }
  __syncthreads();
  while (sh.flag == kFlagQR) {
    // augmented, scaled Jacobian: data rows times d, then the n rows diag(sqrt(diag_h)); rhs rows below L are zero
    for (int i = threadIdx.x; i < M; i += kFitThreads) {
      if (i < L) {
#pragma unroll
        for (int j = 0; j < N; ++j) A[(size_t)j * Mp + i] *= c.d[j];
      } else {
#pragma unroll
        for (int j = 0; j < N; ++j) A[(size_t)j * Mp + i] = (j == i - L) ? sqrt(c.diag_h[j]) : 0.0;
        A[(size_t)N * Mp + i] = 0.0;
      }
    }
    block_qr<N>(A, M, Mp, sh, parity, c.R, sh.qtf);
    double cost_new = 0.0;
    __syncthreads();                                           // R and qtf (written by thread 0) visible to warp 0
    if (threadIdx.x < 32) warp_svd<N>(sh);
    if (threadIdx.x == 0)
      sh.flag = srtrf::inner_propose<N>(c) ? kFlagTrial : (srtrf::outer_end<N>(c, 0.0) ? kFlagAccept : kFlagReject);
    __syncthreads();
    while (sh.flag == kFlagTrial) {
      cost_new = eval_cost<N>(pb, c.x_new, sh, parity);       // barrier inside: every thread has read the flag
      if (threadIdx.x == 0) {
        const bool stop = srtrf::inner_judge<N>(c, cost_new, isfinite(cost_new));
        if (!stop && srtrf::inner_propose<N>(c)) sh.flag = kFlagTrial;
        else sh.flag = srtrf::outer_end<N>(c, cost_new) ? kFlagAccept : kFlagReject;
      }
      __syncthreads();
    }
    const bool accepted = (sh.flag == kFlagAccept);
    __syncthreads();                                           // flag read by all before thread 0 rewrites it
    if (accepted) {
      eval_jac<N>(pb, c.x, A, Mp, sh, parity, g, cost);
      if (threadIdx.x == 0)
        for (int i = 0; i < N; ++i) c.g[i] = g[i];
    }
    if (threadIdx.x == 0) sh.flag = srtrf::outer_begin<N>(c) ? kFlagQR : kFlagStop;
    __syncthreads();
    if (sh.flag == kFlagQR && !accepted) {                     // cannot happen (a rejected outer step ends the solve)
      if (threadIdx.x == 0) c.status = 0;
      break;
    }
  }
  // R factor of the unscaled Jacobian at the solution
  double chi_sum = 0.0;
  eval_jac<N, true>(pb, c.x, A, Mp, sh, parity, g, cost, &chi_sum);
  block_qr<N>(A, L, Mp, sh, parity, c.R, sh.qtf);
  if (threadIdx.x == 0) {
    for (int i = 0; i < N; ++i) POPT[(size_t)r * N + i] = c.x[i];
    for (int i = 0; i < N * N; ++i) ROUT[(size_t)r * N * N + i] = c.R[i];
    COST[r] = c.cost;
    STATUS[2 * r] = c.status; STATUS[2 * r + 1] = c.nfev;
    if (CHI) CHI[r] = chi_sum / (double)L;
  }
}

template <int N>
size_t fit_smem_bytes(long long L) {
  return ((sizeof(FitShared<N>) + 15) / 16) * 16 + sizeof(double) * (size_t)(3 * L + (L + N) * (N + 1));
}

template <int N>
int launch_fit(const double* d_t, const double* d_y, const double* d_sigma, int nR, long long L, const double* d_p0,
               const double* d_lo, const double* d_hi, int max_nfev, double ftol, double xtol, double gtol,
               double* d_popt, double* d_R, double* d_cost, int* d_status, double* d_chi, void* d_work, size_t work_bytes,
               cudaStream_t stream) {
  size_t smem = fit_smem_bytes<N>(L);
  int in_smem = smem <= kFitSmemLimit;
  if (!in_smem) {
    const size_t need = sizeof(double) * (size_t)nR * (L + N) * (N + 1);
    if (!d_work || work_bytes < need) {
      sr_set_error("sr_ct_fit_trf: curves of %lld points need a workspace of %zu bytes (got %zu)", L, need, work_bytes);
      return SR_ERR_WORKSPACE;
    }
    smem = ((sizeof(FitShared<N>) + 15) / 16) * 16;
  }
  SR_CUDA(cudaFuncSetAttribute(ct_fit_trf_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFitSmemLimit));
  ct_fit_trf_kernel<N><<<nR, kFitThreads, smem, stream>>>(d_t, d_y, d_sigma, (int)L, d_p0, d_lo, d_hi, max_nfev, ftol,
                                                           xtol, gtol, d_popt, d_R, d_cost, d_status, d_chi,
                                                           (double*)d_work, in_smem);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

}  // namespace

extern "C" size_t sr_ct_fit_workspace_bytes(int nR, long long L, int nParams) {
  if (nR <= 0 || L <= 0 || nParams < 2 || nParams > kMaxP) return 0;
  const size_t smem = ((sizeof(FitShared<kMaxP>) + 15) / 16) * 16 + sizeof(double) * (size_t)(3 * L + (L + nParams) * (nParams + 1));
  if (smem <= kFitSmemLimit - 1024) return 0;      // conservative: never reports 0 for a shape the launcher sends to the workspace
  return sizeof(double) * (size_t)nR * (L + nParams) * (nParams + 1);
}

extern "C" int sr_ct_fit_trf(const double* d_t, const double* d_y, const double* d_sigma, int nR, long long L, int nParams,
                             const double* d_p0, const double* d_lo, const double* d_hi, int max_nfev, double ftol,
                             double xtol, double gtol, double* d_popt, double* d_R, double* d_cost, int* d_status,
                             double* d_chi, void* d_work, size_t work_bytes, void* stream) {
  SR_REQUIRE(d_t && d_y && d_p0 && d_lo && d_hi && d_popt && d_R && d_cost && d_status, "sr_ct_fit_trf: null pointer");
  SR_REQUIRE(nR > 0 && L > 0 && L < (1LL << 30), "sr_ct_fit_trf: bad shape (nR=%d L=%lld)", nR, L);
  SR_REQUIRE(nParams >= 2 && nParams <= kMaxP, "sr_ct_fit_trf: nParams %d outside [2, %d]", nParams, kMaxP);
  SR_REQUIRE(ftol >= 0 && xtol >= 0 && gtol >= 0, "sr_ct_fit_trf: negative tolerance");
  if (max_nfev <= 0) max_nfev = 100 * nParams;     // SciPy's default for method='trf'
#define SR_FIT_CASE(NP)                                                                                            \
  case NP:                                                                                                         \
    return launch_fit<NP>(d_t, d_y, d_sigma, nR, L, d_p0, d_lo, d_hi, max_nfev, ftol, xtol, gtol, d_popt, d_R, d_cost, \
                          d_status, d_chi, d_work, work_bytes, (cudaStream_t)stream);
  switch (nParams) {
    SR_FIT_CASE(2) SR_FIT_CASE(3) SR_FIT_CASE(4) SR_FIT_CASE(5) SR_FIT_CASE(6) SR_FIT_CASE(7) SR_FIT_CASE(8) SR_FIT_CASE(9)
  }
#undef SR_FIT_CASE
  return SR_ERR_ARG;
}
