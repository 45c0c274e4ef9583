/*
 * spinrelax_b200.h -- C ABI of libspinrelax_b200.so (CUDA, sm_100a only).
 *
 * Drop-in boundary for the trajectory-analysis hot path of zharmad/SpinRelax.  The reference has no
 * FFI for this path except the `npufunc.Jomega` ufunc (Jomega/Jomega.c:49-66); everything else is
 * NumPy inside the stage scripts.  Each entry point below names the reference function (file:line in
 * the SpinRelax tree) whose arithmetic it replaces.  The Python host mirror (the .py files of spinrelax_b200/)
 * binds these with ctypes; INTEGRATION.md shows the stub a SpinRelax maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void* (NULL = default stream)
 *   - pointers named d_* are device pointers, h_* are host pointers
 *   - every function returns SR_OK (0) or a negative SR_ERR_* code; sr_last_error() gives the text
 *   - no CPU fallback: without a CUDA device every compute entry point returns SR_ERR_CUDA
 */
#ifndef SPINRELAX_B200_H
#define SPINRELAX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SR_OK 0
#define SR_ERR_ARG (-1)
#define SR_ERR_CUDA (-2)
#define SR_ERR_WORKSPACE (-3)
#define SR_ERR_OVERFLOW (-4)

#define SR_ABI_VERSION 2

/* library / device introspection */
int sr_abi_version(void);
const char* sr_last_error(void);
int sr_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* global_mem_bytes);

/* ------------------------------------------------------------------------------------------------
 * C(t) = < P2( u(t) . u(t+delta) ) >, Palmer block averaging.
 * Replaces calculate_Ct_Palmer(), calculate-Ct-from-traj.py:200-238 (hot loop :222-228).
 *
 * Input layout is the reference's: vecs (nC, nF, nR, 3) float32, C-contiguous (chunks, frames per
 * chunk, bond vectors, xyz).  Output Ct, dCt are (L, nR) float32 with L = nF/2, first row is
 * delta = 1 (lag 0 is never computed, :222).
 * ---------------------------------------------------------------------------------------------- */

/* row pitch (frames, multiple of 4) of the packed stream: nF plus the zero padding K1's tiles read past the end */
long long sr_ct_row_pitch(long long nF);

/* bytes of device scratch needed by sr_ct_palmer_device: packed SoA stream (12 B per frame) + FP64 lag sums */
size_t sr_ct_workspace_bytes(int nC, long long nF, int nR);

/* K2: (nC, nF, nR, 3) float32 AoS -> vector-major structure-of-arrays stream U[nR][nC][3][pitch] float32
 * (one X, Y and Z plane per (vector, chunk) row; 16-byte aligned, nR*nC*3*pitch*4 bytes), planes
 * zero-padded to `pitch` frames.  If q_rot != NULL (host pointer to 4 doubles, w x y z) the vectors
 * are rotated by the normalised quaternion first (rotate_vector_simd, transforms3d_supplement.py:270-296). */
int sr_pack_vectors_f32(const float* d_vecs, int nC, long long nF, int nR, const double* h_q_rot,
                        void* d_packed, long long pitch, void* stream);

/* The same for the chunk range [c0, c0 + nCsub) only: d_vecs_c0 points at chunk c0 of the input, the rows written are
 * the ones sr_pack_vectors_f32 would write for those chunks.  Lets a host pipeline overlap the H2D copy of chunk
 * c + 1 with the kernels of chunk c. */
int sr_pack_vectors_f32_chunks(const float* d_vecs_c0, int nC, int c0, int nCsub, long long nF, int nR,
                               const double* h_q_rot, void* d_packed, long long pitch, void* stream);

/* K1: lag sums S[(r*nC + c)*L + (delta-1)] = sum_t ( u(t) . u(t+delta) )^2, t in [0, nF-delta),
 * FP32 products with FP64 accumulation.  d_packed is the output of sr_pack_vectors_f32. */
int sr_ct_lag_sums(const void* d_packed, long long pitch, int nC, long long nF, int nR, long long L,
                   double* d_S, void* stream);

/* K1 restricted to the chunk range [c0, c0 + nCsub); d_packed and d_S are the full buffers. */
int sr_ct_lag_sums_chunks(const void* d_packed, long long pitch, int nC, int c0, int nCsub, long long nF, int nR,
                          long long L, double* d_S, void* stream);

/* per-chunk mean -> mean and std/(sqrt(nC)-1) over chunks (:226-228). */
int sr_ct_palmer_finalize(const double* d_S, int nC, long long nF, int nR, long long L, float* d_Ct,
                          float* d_dCt, void* stream);

/* K2 + K1 + finalize on device buffers. d_workspace >= sr_ct_workspace_bytes(). */
int sr_ct_palmer_device(const float* d_vecs, int nC, long long nF, int nR, float* d_Ct, float* d_dCt,
                        void* d_workspace, size_t workspace_bytes, void* stream);

/* Same through host (pageable) buffers: pinned staging, H2D of chunk c + 1 overlapped with the kernels of chunk c,
 * D2H of Ct/dCt.  Scratch (device buffers, pinned staging, streams) is cached between calls; sr_release_host_cache()
 * frees it.  Calls are serialised by an internal mutex. */
int sr_ct_palmer_host(const float* h_vecs, int nC, long long nF, int nR, float* h_Ct, float* h_dCt);
void sr_release_host_cache(void);

/* ------------------------------------------------------------------------------------------------
 * K3: PAF rotation + Lambert-cylindrical histogram, replaces calculate-Ct-from-traj.py:567 (rotation,
 * transforms3d_supplement.py:270-296), :588 (gm.xyz_to_rtp, general_maths.py:143-158), :600-626
 * (np.histogramdd per vector, bins (nbx, nby), range ((-pi,pi),(-1,1))).
 *
 * d_vecs: (nFrames, nR, 3) float32 AoS (the reference layout after the reshape at :535-536).
 * h_q_rot: host pointer to (w,x,y,z) or NULL for no rotation (float32 path of the reference).
 * d_edge_table: 2*(nbx+1) doubles (cos e_i, sin e_i of the phi edges) followed by the nby+1 cos(theta)
 * edges, as np.histogramdd builds them (np.linspace over (-pi,pi) and (-1,1): the hot pass assumes uniform bins;
 * the table serves the FP64 pass).  Counts are ADDED into d_counts [nR][nbx][nby] (uint32).
 * Two passes (three launches with the marker of the retry list): the FP32 hot pass counts every sample it can
 * place with a float32 error margin and appends the flat sample index (frame*nR + r) of the others (~3e-4 of the
 * stream) to d_amb_idx (up to amb_capacity;
 * *d_amb_count keeps counting beyond it, so amb_capacity bounds the retry list, not only the final list); the
 * FP64 resolve pass then re-examines that list, counts what is farther than tol_phi (rad) / tol_cos from every
 * bin edge and overwrites those entries with -1.  Entries that are still >= 0 afterwards are samples the caller
 * must bin with the reference's exact NumPy formula -- this is what makes the counts bit-identical to the
 * reference without sharing its libm.  Only the entries appended by this call are re-examined, so counts and the
 * list may be accumulated over several calls (sample indices are relative to each call's d_vecs).
 * ---------------------------------------------------------------------------------------------- */
int sr_sphere_hist_table_doubles(int nbx, int nby);
int sr_sphere_hist(const float* d_vecs, long long nFrames, int nR, const double* h_q_rot, int nbx, int nby,
                   const double* d_edge_table, double tol_phi, double tol_cos, unsigned int* d_counts,
                   long long* d_amb_idx, int amb_capacity, int* d_amb_count, void* stream);

/* The same through host (pageable) buffers.  Builds np.histogramdd's edges itself; h_counts [nR][nbx][nby] receives
 * the counts of every sample the two device passes could place, h_amb_idx[0 .. *h_n_amb) the flat ids (frame*nR + r)
 * of the few that lie within the tie-break tolerance of a bin edge (to be binned by the caller with the reference's
 * formula; spinrelax_b200/hist.py::_reference_bins is that formula). */
int sr_sphere_hist_host(const float* h_vecs, long long nFrames, int nR, const double* h_q_rot, int nbx, int nby,
                        unsigned int* h_counts, long long* h_amb_idx, int amb_capacity, int* h_n_amb);

/* ------------------------------------------------------------------------------------------------
 * X-H bond vectors from Cartesian coordinates: the step in front of the hot path.
 * Replaces obtain_XHvecs(), calculate-Ct-from-traj.py:64-86: np.take(xyz, indexH, 1) - np.take(xyz, indexX, 1)
 * followed by qs.vecnorm_NDarray (transforms3d_supplement.py:40-52; float32, bit-identical to NumPy including
 * nan_to_num of a zero vector).
 * d_xyz: (nFrames, nAtoms, 3) float32 (mdtraj's traj.xyz); d_idxH / d_idxX: nR atom indices each (int32);
 * d_out: (nFrames, nR, 3) float32 unit vectors, the layout calculate_Ct_Palmer's input is built from.
 * ---------------------------------------------------------------------------------------------- */
int sr_xh_vectors(const float* d_xyz, long long nFrames, int nAtoms, const int* d_idxH, const int* d_idxX, int nR,
                  float* d_out, void* stream);

/* The same after trj.center_coordinates(); trj.superpose(ref, frame=0, atom_indices=fit) (:466-467, mdtraj): every
 * frame is rotated by the proper rotation that minimises the RMSD of its fit atoms to the reference (Horn's
 * quaternion solution, float64; translations do not matter for bond vectors, so the coordinates are not rewritten).
 * d_fitIdx: nFit atom indices; d_refFitCentred: (nFit, 3) float64 reference coordinates of those atoms minus their
 * centroid; d_rot: optional (nFrames, 9) float64 row-major rotation matrices (NULL to skip). */
int sr_xh_vectors_superposed(const float* d_xyz, long long nFrames, int nAtoms, const int* d_fitIdx,
                             const double* d_refFitCentred, int nFit, const int* d_idxH, const int* d_idxX, int nR,
                             float* d_out, double* d_rot, void* stream);

/* Per (block of framesPerBlock frames, vector) sums of x,y,z,xx,xy,xz,yy,yz,zz (FP64) of (nFrames, nR, 3) float32
 * vectors: the reductions behind --vecAvg (calculate-Ct-from-traj.py:579-583) and --S2
 * (calculate_S2_by_outerProduct :96-145).  d_out [nBlocks][nR][9], nBlocks = ceil(nFrames/framesPerBlock). */
int sr_vec_block_moments(const float* d_vecs, long long nFrames, int nR, long long framesPerBlock, double* d_out,
                         void* stream);

/* qs.rotate_vector_simd(v, q) (transforms3d_supplement.py:270-296) for n float32 vectors and one float64
 * quaternion h_q (w,x,y,z; the caller normalises it as vecnorm_NDarray does): float64 output, bit-identical
 * to NumPy (separately rounded products/sums in NumPy's order). */
int sr_rotate_vectors_f32_f64(const float* d_v, long long n, const double* h_q, double* d_out, void* stream);

/* gm.xyz_to_rtp(uv, vaxis=-1, bUnit) (general_maths.py:118-158) for n vectors stored (n,3), computed in the
 * precision of the input as NumPy does.  unit_form == 0: d_out (n,3) = (|v|, atan2(y,x), acos(z/|v|)), |v|
 * bit-identical to np.linalg.norm(uv, axis=-1).  unit_form != 0: d_out (n,2) = (phi, acos(z/phi)), the shipped
 * bUnit branch (:131-133 divides by phi).  Called on (frames, nR, 3) by calculate-Ct-from-traj.py:588 for the
 * _vecPhiTheta.{npz,dat} outputs; phi/theta agree with libm to <= 2 ulp. */
int sr_xyz_to_rtp_f32(const float* d_v, long long n, float* d_out, int unit_form, void* stream);
int sr_xyz_to_rtp_f64(const double* d_v, long long n, double* d_out, int unit_form, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K4: quaternion-displacement statistics over lag windows.
 * Replaces the per-lag body of calculate-dq-distribution.py:554-625: obtain_self_dq (:102-109),
 * average_anisotropic_tensor[_chunk] (:118-144) and average_LegendreP1quat[_chunk] (:111-135).
 *
 * d_q: (N, 4) float32 (w,x,y,z) orientation quaternions (float32 as read by plumedcolvario.py:68).
 * d_lags: nLags frame lags (device, int64), each in [min_lag, N).  nCh >= 1 consecutive sub-chunks of
 * ceil((N-lag)/nCh) samples (:129).  Output d_M [nLags][nCh][6] = sum_t of (xx, xy, xz, yy, yz, zz) of the
 * vector part of conj(q_t) q_{t+lag}, float64.  Full-trajectory sums are the sum over chunks; means, the
 * shipped iso moment 1-(2/3)sum|v|^2 (quirk G1), the intended 1-2<|v|^2> and R M R^T all derive on the host.
 * ---------------------------------------------------------------------------------------------- */
int sr_dq_moments(const float* d_q, long long N, const long long* d_lags, int nLags, long long min_lag, int nCh,
                  double* d_M, void* stream);

/* The same for replica `replica` of `nReplicas` equally long trajectories whose displacement samples are pooled
 * (calculate-dq-distribution-multi.py:529-540): sub-chunk boundaries are taken on the pooled sample index, and
 * with accumulate != 0 the sums are added to d_M so that one call per replica (or one all-reduce over ranks that
 * each hold one replica) yields the pooled moments. */
int sr_dq_moments_pooled(const float* d_q, long long N, const long long* d_lags, int nLags, long long min_lag, int nCh,
                         int replica, int nReplicas, int accumulate, double* d_M, void* stream);

/* Objective of the 1-parameter Powell fits of the decay curves (powell_expdecay, calculate-dq-distribution.py:199-203):
 * *d_out = mean_i (C0 exp(-x_i / A) + C1 - y_i)^2 for a curve (d_x, d_y, n doubles) resident on the device.  d_work:
 * at least 129 doubles, zero-initialised once by the caller and reusable for every further evaluation on that stream.
 * SciPy's fmin_powell stays on the host (:205) and calls this once per trial tau. */
int sr_expdecay_chi2(const double* d_x, const double* d_y, long long n, double C0, double C1, double A, double* d_work,
                     int work_doubles, double* d_out, void* stream);

/* obtain_self_dq(q, delta): d_out (N-delta, 4) float64, imaged so that w >= 0 (quat_reduce_simd). */
int sr_dq_self(const float* d_q, long long N, long long delta, double* d_out, void* stream);

/* 3-D histogram of the vector part of dq(t) for one lag over the cube spanned by d_edges (nb + 1 float64 edges,
 * the same on all three axes: np.linspace(-1, 1, nb + 1)), the --hist option of calculate-dq-distribution.py:633-647
 * (np.histogramdd semantics).  Counts are ADDED into d_counts [nb][nb][nb]; samples within 4 ulp of an edge are
 * not counted but listed in d_amb_idx (frame index t) for the caller to bin with NumPy. */
int sr_dq_hist3d(const float* d_q, long long N, long long delta, const double* d_edges, int nb, unsigned int* d_counts,
                 long long* d_amb_idx, int amb_capacity, int* d_amb_count, void* stream);

/* second moments of a float64 (n,3) vector list in nCh consecutive blocks: d_M [nCh][6]. */
int sr_vec_second_moments(const double* d_v, long long n, int nCh, double* d_M, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K7: the npufunc.Jomega ufunc, Jomega/Jomega.c:49-66 (double loop) and :68-85 (float loop):
 * out[i] = x[i] / (x[i]*x[i] + y[i]*y[i]) on n contiguous elements (the Python binding does the NumPy
 * broadcasting / .outer); products, sum and quotient are rounded separately like the C loop.
 * The reference's native loop signature is void(char **args, npy_intp *dimensions, npy_intp *steps, void*).
 * ---------------------------------------------------------------------------------------------- */
int sr_jomega_f64(const double* d_x, const double* d_y, double* d_out, long long n, void* stream);
int sr_jomega_f32(const float* d_x, const float* d_y, float* d_out, long long n, void* stream);

/* The same for the inner loop of a NumPy ufunc (the reference registers double_Jomega / float_Jomega with
 * PyUFunc_FromFuncAndData, Jomega/Jomega.c:49-104, 135-156): host pointers with byte strides exactly as NumPy hands them
 * to a loop function (stride 0 = broadcast scalar).  Gather, H2D, kernel, D2H, scatter.  Bound by
 * spinrelax_b200/csrc/npufunc_module.c, the extension module that makes `npufunc.Jomega` a real numpy.ufunc. */
int sr_jomega_host_f64(const char* x, long long sx, const char* y, long long sy, char* out, long long so, long long n);
int sr_jomega_host_f32(const char* x, long long sx, const char* y, long long sy, char* out, long long so, long long n);

/* ------------------------------------------------------------------------------------------------
 * K6: J(omega) -> R1 / R2 / NOE with weighted averaging over the bond-vector distribution.
 * Replaces spectral_densities.py: update_A_coefficients :503-523 (caller builds A), calc_Jomega_one :552-557,
 * _do_Jsum :1961-1972, isotropic calc_Jomega_one :430-443, spinRelaxationR1/R2/NOE.func :824-829, :859-864,
 * :888-892 and check_and_calculate_average :751-763.
 *
 * sr_relax_a_moments: d_A (B,3) [or (nR,B,3) if per_residue_A] A_J coefficients of the bin vectors,
 *   d_W (nR,B) weights (histogram counts).  d_amom (nR,10) = sum w, mean A (3), weighted covariance of A
 *   (xx,xy,xz,yy,yz,zz).
 * sr_relax_eval: one evaluation per (residue, field, CSA point).  h_D_J: host (3) D_J coefficients
 *   (axisymmetric) or (1) D_iso when iso != 0.  Models: d_S2 (nR), d_C/d_tau (nR,max_comp), d_nComp (nR).
 *   d_omega (nField,5) rad per time unit; d_f_csa (nField,nCSA), or (nField,nR) when csa_per_residue != 0
 *   (then nCSA must be 1).  d_out (nR,nField,nCSA,6) = R1, R2, NOE, sigma_R1, sigma_R2, sigma_NOE; NOE uses
 *   the bin-averaged R1 (:885-886); sigmas are 0 for isotropic tumbling.
 * ---------------------------------------------------------------------------------------------- */
int sr_relax_a_moments(const double* d_A, int per_residue_A, const double* d_W, int nR, int B, double* d_amom,
                       void* stream);
int sr_relax_eval(int iso, const double* h_D_J, double zeta, double time_fact, double gammaA, double gammaB,
                  double f_dd, int nR, int nField, int nCSA, int csa_per_residue, int max_comp,
                  const double* d_amom, const double* d_S2, const double* d_C, const double* d_tau,
                  const int* d_nComp, const double* d_omega, const double* d_f_csa, double* d_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K5: batched bounded least-squares fit of C(t) = S2 + sum_i C_i exp(-t/tau_i), one CTA per residue.
 * Replaces the scipy.optimize.curve_fit call in autoCorrelationModel.conduct_curve_fitting,
 * fitting_Ct_functions.py:322-324 (model :419-427, bounds :412-416, initial guess :359-374 built by caller).
 * The solver is the algorithm curve_fit runs when bounds are given -- least_squares(method='trf', tr_solver='exact',
 * x_scale=1) -- restated step for step (csrc/trf_core.cuh), because the reference's model selection depends on where
 * that solver stops and on whether it reports success (:325-328).
 * d_t, d_y, d_sigma: (nR, L) float64 (d_sigma may be NULL = unweighted).  Parameters ordered as the
 * reference: C_1..C_nc, tau_1..tau_nc [, S2]; d_p0, d_lo, d_hi, d_popt are (nR, nParams).
 * max_nfev <= 0 selects SciPy's default 100*nParams; ftol = xtol = gtol = 1e-8 are curve_fit's values.
 * d_R (nR, nParams, nParams): upper-triangular R factor of the sigma-weighted Jacobian at the solution (same singular
 * values and right singular vectors as the Jacobian curve_fit decomposes for pcov); d_cost (nR) = 0.5 sum r^2;
 * d_status (nR, 2) = {SciPy termination status, nfev}: 1 gtol, 2 ftol, 3 xtol, 4 ftol and xtol, 0 max_nfev reached
 * (curve_fit raises RuntimeError), -3 p0 outside the bounds, -4 residuals not finite at p0 (both ValueError upstream).
 * d_chi (nR, may be NULL): mean_k (f(t_k) - y_k)^2 / sigma_k at the solution -- the reference's chi^2 (calc_chiSq,
 * fitting_Ct_functions.py:272-276; divided by sigma, not sigma^2) for a model whose zeta is 1; inf for failed solves.
 * Curves too long for shared memory need d_work of sr_ct_fit_workspace_bytes() bytes (0 = not needed).
 * ---------------------------------------------------------------------------------------------- */
size_t sr_ct_fit_workspace_bytes(int nR, long long L, int nParams);
int sr_ct_fit_trf(const double* d_t, const double* d_y, const double* d_sigma, int nR, long long L, int nParams,
                  const double* d_p0, const double* d_lo, const double* d_hi, int max_nfev, double ftol, double xtol,
                  double gtol, double* d_popt, double* d_R, double* d_cost, int* d_status, double* d_chi, void* d_work,
                  size_t work_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPINRELAX_B200_H */
