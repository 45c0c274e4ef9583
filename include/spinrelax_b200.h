/*
 * spinrelax_b200.h -- C ABI of libspinrelax_b200.so (CUDA, sm_100a only).
 *
 * Drop-in boundary for the trajectory-analysis hot path of zharmad/SpinRelax.  The reference has no
 * FFI for this path except the `npufunc.Jomega` ufunc (Jomega/Jomega.c:49-66); everything else is
 * NumPy inside the stage scripts.  Each entry point below names the reference function (file:line in
 * the SpinRelax tree) whose arithmetic it replaces.  The Python host mirror (spinrelax_b200/*.py)
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

#define SR_ABI_VERSION 1

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

/* frames of zero padding appended to every (vector, chunk) row of the packed stream */
long long sr_ct_row_pitch(long long nF);

/* bytes of device scratch needed by sr_ct_palmer_device: packed float4 stream + FP64 lag sums */
size_t sr_ct_workspace_bytes(int nC, long long nF, int nR);

/* K2: (nC, nF, nR, 3) float32 AoS -> vector-major float4 stream U[nR][nC][pitch] (x,y,z,0), rows
 * zero-padded to `pitch` frames.  If q_rot != NULL (host pointer to 4 doubles, w x y z) the vectors
 * are rotated by the normalised quaternion first (rotate_vector_simd, transforms3d_supplement.py:270-296). */
int sr_pack_vectors_f32(const float* d_vecs, int nC, long long nF, int nR, const double* h_q_rot,
                        void* d_packed, long long pitch, void* stream);

/* K1: lag sums S[(r*nC + c)*L + (delta-1)] = sum_t ( u(t) . u(t+delta) )^2, t in [0, nF-delta),
 * FP32 products with FP64 accumulation.  d_packed is the output of sr_pack_vectors_f32. */
int sr_ct_lag_sums(const void* d_packed, long long pitch, int nC, long long nF, int nR, long long L,
                   double* d_S, void* stream);

/* per-chunk mean -> mean and std/(sqrt(nC)-1) over chunks (:226-228). */
int sr_ct_palmer_finalize(const double* d_S, int nC, long long nF, int nR, long long L, float* d_Ct,
                          float* d_dCt, void* stream);

/* K2 + K1 + finalize on device buffers. d_workspace >= sr_ct_workspace_bytes(). */
int sr_ct_palmer_device(const float* d_vecs, int nC, long long nF, int nR, float* d_Ct, float* d_dCt,
                        void* d_workspace, size_t workspace_bytes, void* stream);

/* Same through host buffers: H2D of vecs, compute, D2H of Ct/dCt (allocates its own scratch). */
int sr_ct_palmer_host(const float* h_vecs, int nC, long long nF, int nR, float* h_Ct, float* h_dCt);

#ifdef __cplusplus
}
#endif
#endif /* SPINRELAX_B200_H */
