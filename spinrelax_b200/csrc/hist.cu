// K3: PAF rotation + Lambert-cylindrical (phi, cos theta) histogram of unit bond vectors.
// Reference arithmetic: calculate-Ct-from-traj.py:567 (rotate_vector_simd), :588 (gm.xyz_to_rtp),
// :600-626 (transpose, cos(theta), np.histogramdd with bins (nbx, nby) over ((-pi,pi),(-1,1))).
//
// Bit-exact counts without reproducing NumPy's libm, in two passes:
//   1. sphere_hist_kernel (hot, FP32): rotates, locates the sample in bin units along both axes and accepts the
//      bin only if the sample is farther than a float32 error bound from all four edges of that bin.  Misses
//      (~2e-4 of the samples) are appended to a retry list.
//   2. sphere_hist_resolve_kernel (rare, FP64): re-examines the retry list with the same tests in double
//      precision and a ~1e-11 margin (rotated stream), or drops NaN / zero vectors (float32 reference stream).
//      What is still undecided keeps its sample id in the list for the host, which re-bins those few samples
//      with the reference's own NumPy formula; everything else in the list is overwritten with -1.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <utility>
#include <vector>

namespace {

constexpr int kHistThreads = 768;   // default threads per CTA, one CTA per SM (80 registers per thread)

struct HistParams {
  double R[9];        // rotation matrix (row major), identity when no rotation
  double tol_phi;     // angular margin (rad) of the FP64 test
  double tol_cos;     // cos(theta) margin of the FP64 test
  float Rf[9];        // the same matrix in float32 for the fast path
  float pk[6];        // atan polynomial scaled to phi-bin units: atan(t) nbx/(2 pi) ~ t (pk0 + pk1 t^2 + ... + pk5 t^10)
  float quarter, half;              // nbx/4, nbx/2: a quarter and a half turn in phi-bin units
  float phi_c0, phi_m1;             // fast phi test: |frac - 1/2| rho1 <= phi_c0 rho1 - phi_m1 (rho1 + |z|)
  float cos_scale, cos_c0;          // nby/2 ; fast cos test: |frac - 1/2| <= cos_c0
  int nbx, nby;
  int f32_reference;  // 1: no rotation, the reference itself works in float32 -> fast-path misses go to the host
};

struct SlowParams { double tol_phi, tol_cos; int nbx, nby; };

__device__ __forceinline__ int classify(double x, double y, double z, const SlowParams p,
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

// Hot kernel (FP32).  For every sample it computes the position of the rotated vector in *bin units* along both
// histogram axes, u_phi = (atan2(y, x) + pi) nbx / (2 pi) from a degree-11 minimax polynomial of atan on [0,1]
// (|error| < 2.1e-5 bins at nbx = 72, float32 evaluation included) and u_cos = (z / r + 1) nby / 2, and accepts the
// bin (floor u_phi, floor u_cos) only if both fractional parts are farther from 0 and 1 than a margin that
// bounds every float32 effect on them:
//   phi:  |frac - 1/2| rho1 <= (1/2 - m0) rho1 - m1 (rho1 + |z|),  rho1 = |x| + |y|
//         (frac - 1/2 is taken as (u - 1/2) - rint(u - 1/2): a tie of the rounding is a sample on an edge, rejected)
//         m0 = polynomial + evaluation error, m1 (rho1 + |z|) / rho1 >= the angle error caused by the float32
//         rounding of the rotated x, y (absolute ~3e-7 |v|, so it grows towards the poles);
//   cos:  |frac - 1/2| <= 1/2 - mc.
// The unrotated stream has no coordinate rounding (m1 = 0); there m0 / mc hold the reference's own float32
// arctan2 / arccos / cos error (the caller's tolerances).  NaN, zero, huge and polar vectors need no explicit test:
// they make one of the comparisons false.  Misses (~2e-4 of the samples) go to the retry list, once per thread and
// pass rather than per sample, so the common path is branch-free.
//
// grid.x = vector groups (fastest, so CTAs sharing a frame range run together and the 64-byte DRAM blocks that
// straddle two groups are fetched once), grid.y = frame blocks.  A CTA owns `4 << qshift` vectors; thread t works
// on four consecutive vectors of one frame, i.e. on 48 contiguous bytes of the reference's AoS layout -- three
// LDG.128 when a frame row is a multiple of 16 bytes (nR % 4 == 0), twelve scalar loads otherwise -- and always has
// the next pass's 48 bytes in flight while it classifies the current ones.  Bins are privatised in shared memory
// as 32-bit counters (one CTA per SM owns 16 x 2592 of them at the default 72 x 36) and flushed with global
// atomics at the end.
//
// NBX/NBY > 0 compiles the bin counts (and with them the polynomial, the turn constants and the per-vector
// histogram offsets) into immediates -- the reference's default 72 x 36; 0 keeps them as kernel parameters.
constexpr double kAtanK[6] = {0.9999772310256958, -0.33262282609939575, 0.19354039430618286,
                              -0.1164264902472496, 0.052647337317466736, -0.01171912346035242};
constexpr double kPiD = 3.141592653589793;
constexpr int kHistPadLo = 4;   // words before the first histogram: a rejected sample may carry bin -1 (added value 0)

template <int OFF>
__device__ __forceinline__ void red_shared_add(uint32_t addr, unsigned v) {
  asm volatile("red.shared.add.u32 [%0+%2], %1;" ::"r"(addr), "r"(v), "n"(OFF) : "memory");
}

template <bool VEC4, int NBX, int NBY, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
sphere_hist_kernel(const float* __restrict__ vecs, long long nFrames, int nR, int framesPerBlock, int qshift,
                   const __grid_constant__ HistParams p, unsigned int* __restrict__ counts,
                   long long* __restrict__ amb_idx, int amb_capacity, int* __restrict__ amb_count) {
  extern __shared__ __align__(16) unsigned int sh_raw[];
  unsigned int* const sh_hist = sh_raw + kHistPadLo;
  const int nbx = NBX ? NBX : p.nbx, nby = NBY ? NBY : p.nby;
  const int nbins = nbx * nby;
  const int tid = threadIdx.x;
  const int group = 4 << qshift;                            // vectors per CTA
  const int framesPerPass = THREADS >> qshift;
  const int r0 = blockIdx.x * group;
  const int nv = min(group, nR - r0);
  for (int i = tid; i < nv * nbins; i += THREADS) sh_hist[i] = 0u;
  __syncthreads();

  const long long f0 = (long long)blockIdx.y * framesPerBlock;
  const int nfl = (int)min((long long)framesPerBlock, nFrames - f0);
  const int quad = tid & ((1 << qshift) - 1);
  const int nvq = max(0, min(4, nv - 4 * quad));            // valid vectors of this thread
  int fl = tid >> qshift;
  const uint32_t hist_addr = sr_smem_u32(sh_hist + 4 * quad * nbins);
  const uint32_t hist_step = (uint32_t)nbins * 4u;
  const float q0 = p.Rf[0], q1 = p.Rf[1], q2 = p.Rf[2], q3 = p.Rf[3], q4 = p.Rf[4], q5 = p.Rf[5], q6 = p.Rf[6],
              q7 = p.Rf[7], q8 = p.Rf[8];
  constexpr double kScale = NBX / (2.0 * kPiD);
  const float k0 = NBX ? (float)(kAtanK[0] * kScale) : p.pk[0], k1 = NBX ? (float)(kAtanK[1] * kScale) : p.pk[1],
              k2 = NBX ? (float)(kAtanK[2] * kScale) : p.pk[2], k3 = NBX ? (float)(kAtanK[3] * kScale) : p.pk[3],
              k4 = NBX ? (float)(kAtanK[4] * kScale) : p.pk[4], k5 = NBX ? (float)(kAtanK[5] * kScale) : p.pk[5];
  const float quarter = NBX ? 0.25f * NBX : p.quarter, half = NBX ? 0.5f * NBX : p.half;
  const float cos_scale = NBY ? 0.5f * NBY : p.cos_scale, fnby = NBY ? (float)NBY : (float)p.nby;
  const float halfm = half - 0.5f, cos_scalem = cos_scale - 0.5f;      // exact
  const float phi_c1 = p.phi_c0 - p.phi_m1, phi_m1 = p.phi_m1, cos_c0 = p.cos_c0;
  const size_t passStride = (size_t)framesPerPass * nR * 3;
  const float* src = vecs + ((size_t)(f0 + fl) * nR + r0 + 4 * quad) * 3;
  long long sidx = (f0 + fl) * nR + r0 + 4 * quad;          // sample id of this thread's first vector
  const long long sidxStep = (long long)framesPerPass * nR;
  const unsigned validMask = (1u << nvq) - 1u;

  auto load = [&](float (&d)[12], const float* s) {
    if (VEC4) {
      const float4* s4 = reinterpret_cast<const float4*>(s);
      const float4 a = __ldg(s4), b = __ldg(s4 + 1), c = __ldg(s4 + 2);
      d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
      d[8] = c.x; d[9] = c.y; d[10] = c.z; d[11] = c.w;
    } else {
#pragma unroll
      for (int i = 0; i < 12; ++i) d[i] = (i < 3 * nvq) ? __ldg(s + i) : 0.f;
    }
  };
  auto process = [&](const float (&d)[12], long long sid) {
    unsigned miss = 0u;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float vx = d[3 * u], vy = d[3 * u + 1], vz = d[3 * u + 2];
      const float x = fmaf(q0, vx, fmaf(q1, vy, q2 * vz));
      const float y = fmaf(q3, vx, fmaf(q4, vy, q5 * vz));
      const float z = fmaf(q6, vx, fmaf(q7, vy, q8 * vz));
      const float r2 = fmaf(x, x, fmaf(y, y, z * z));
      const float ax = fabsf(x), ay = fabsf(y);
      const float rho1 = ax + ay;
      // ---- phi, in bin units ----
      const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
      const float t = __fdividef(mn, mx), t2 = t * t;
      float a = fmaf(t2, k5, k4);
      a = fmaf(t2, a, k3);
      a = fmaf(t2, a, k2);
      a = fmaf(t2, a, k1);
      a = fmaf(t2, a, k0) * t;                              // [0, nbx/8]
      if (ay > ax) a = quarter - a;                         // [0, nbx/4]
      if (x < 0.f) a = half - a;                            // [0, nbx/2]
      // u - 1/2 along each axis: its nearest integer is the bin, its distance from that integer is |frac - 1/2|
      const float uphi = copysignf(a, y) + halfm;           // [-1/2, nbx - 1/2]
      const float fphi = rintf(uphi);
      const float dphi = fabsf(uphi - fphi);
      const bool okphi = dphi * rho1 <= fmaf(-phi_m1, fabsf(z), phi_c1 * rho1);
      // ---- cos(theta), in bin units ----
      const float ucos = fmaf(z * rsqrtf(r2), cos_scale, cos_scalem);  // [-1/2 - eps, nby - 1/2 + eps]
      const float fcos = rintf(ucos);
      const bool okcos = fabsf(ucos - fcos) <= cos_c0;
      const bool ok = okphi && okcos;                        // NaN compares false
      // ok implies 0 <= bin < nbins (the fractional parts at and beyond the ends of both ranges fail the margins);
      // a rejected sample adds 0 to a word inside [-1, nbins + nby] of its vector's histogram (padding both ends).
      const int bin = (int)fmaf(fphi, fnby, fcos);
      const uint32_t addr = hist_addr + ((uint32_t)bin << 2);
      if (NBX && NBY) {
        const unsigned one = ok ? 1u : 0u;                   // the offset of vector u is an immediate of the instruction
        if (u == 0) red_shared_add<0>(addr, one);
        if (u == 1) red_shared_add<4 * NBX * NBY>(addr, one);
        if (u == 2) red_shared_add<8 * NBX * NBY>(addr, one);
        if (u == 3) red_shared_add<12 * NBX * NBY>(addr, one);
      } else {
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr + u * hist_step), "r"(ok ? 1u : 0u) : "memory");
      }
      if (!ok) miss |= 1u << u;
    }
    miss &= validMask;
    if (miss) {     // rare: hand the samples to sphere_hist_resolve_kernel
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (miss & (1u << u)) {
          const int slot = atomicAdd(amb_count, 1);
          if (slot < amb_capacity) amb_idx[slot] = sid + u;
        }
    }
  };

  if (nvq > 0) {
    float A[12], B[12];
    if (fl < nfl) load(A, src);
    while (fl < nfl) {
      const bool moreB = fl + framesPerPass < nfl;
      if (moreB) load(B, src + passStride);
      process(A, sidx);
      if (!moreB) break;
      const bool moreA = fl + 2 * framesPerPass < nfl;
      if (moreA) load(A, src + 2 * passStride);
      process(B, sidx + sidxStep);
      fl += 2 * framesPerPass; src += 2 * passStride; sidx += 2 * sidxStep;
    }
  }
  __syncthreads();
  for (int i = tid; i < nv * nbins; i += THREADS) {
    const unsigned int w = sh_hist[i];
    if (w) atomicAdd(&counts[(long long)r0 * nbins + i], w);
  }
}

// remembers where this call's retry entries start (the list may already hold entries of earlier calls)
__global__ void sphere_hist_mark_kernel(const int* __restrict__ amb_count, int amb_capacity, int* __restrict__ amb_start) {
  *amb_start = min(*amb_count, amb_capacity);
}

// Second pass over the fast path's misses (typically ~1e-4 of the samples): FP64 re-examination for the rotated
// stream, NaN / zero-vector drop for the float32 reference stream.  Resolved entries are overwritten with -1,
// entries that stay ambiguous keep their sample id for the host tie-break (hist.py::_reference_bins).
__global__ void __launch_bounds__(256)
sphere_hist_resolve_kernel(const float* __restrict__ vecs, int nR, HistParams p, const double2* __restrict__ edge_dir,
                           const double* __restrict__ edge_cos, unsigned int* __restrict__ counts,
                           long long* __restrict__ amb_idx, int amb_capacity, const int* __restrict__ amb_count,
                           const int* __restrict__ amb_start) {
  // only the entries appended by THIS call's hot pass: [*amb_start, *amb_count)
  const int n = min(*amb_count, amb_capacity);
  const int nbins = p.nbx * p.nby;
  for (int e = *amb_start + blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const long long sidx = amb_idx[e];
    if (sidx < 0) continue;
    const float* v = vecs + sidx * 3;
    const float vx = v[0], vy = v[1], vz = v[2];
    int bin = 0, cls;
    if (p.f32_reference) {
      const float n2 = vx * vx + vy * vy + vz * vz;
      cls = (n2 != n2 || n2 == 0.f) ? 1 : 2;
    } else {
      const double dx = vx, dy = vy, dz = vz;
      SlowParams sp; sp.tol_phi = p.tol_phi; sp.tol_cos = p.tol_cos; sp.nbx = p.nbx; sp.nby = p.nby;
      cls = classify(p.R[0] * dx + p.R[1] * dy + p.R[2] * dz, p.R[3] * dx + p.R[4] * dy + p.R[5] * dz,
                     p.R[6] * dx + p.R[7] * dy + p.R[8] * dz, sp, edge_dir, edge_cos, bin);
    }
    if (cls == 0) atomicAdd(&counts[(sidx % nR) * nbins + bin], 1u);
    if (cls != 2) amb_idx[e] = -1;
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
  SR_REQUIRE(amb_capacity > 0, "sr_sphere_hist: the retry / ambiguous list needs a positive capacity");
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
  for (int i = 0; i < 9; ++i) p.Rf[i] = (float)p.R[i];
  p.f32_reference = h_q_rot ? 0 : 1;
  {  // fast-path constants (derivation above sphere_hist_kernel)
    static const double atan_k[6] = {0.9999772310256958, -0.33262282609939575, 0.19354039430618286,
                                     -0.1164264902472496, 0.052647337317466736, -0.01171912346035242};
    const double kPi = 3.141592653589793, scale = nbx / (2.0 * kPi);
    for (int i = 0; i < 6; ++i) p.pk[i] = (float)(atan_k[i] * scale);
    p.quarter = 0.25f * nbx; p.half = 0.5f * nbx; p.cos_scale = 0.5f * nby;
    const double eval_phi = 3.0e-7 * nbx, eval_cos = 1.5e-7 * nby;     // float32 evaluation of u_phi / u_cos, 2x
    double m0, m1, mc;
    if (h_q_rot) {   // polynomial 1.7e-6 rad + evaluation 0.7e-6 rad, doubled; coordinate rounding 6e-7 |v| / rho1 rad, doubled
      m0 = 4.8e-6 * scale + eval_phi; m1 = 1.2e-6 * scale; mc = 2.0e-6 * p.cos_scale + eval_cos;
    } else {         // the caller's tolerances describe the reference's own float32 error; ours (2.4e-6 rad) on top
      m0 = (tol_phi + 2.4e-6) * scale + eval_phi; m1 = 0.0; mc = tol_cos * p.cos_scale + eval_cos;
    }
    SR_REQUIRE(m0 + m1 < 0.25 && mc < 0.25, "sr_sphere_hist: bins too narrow for the float32 pass (%d x %d)", nbx, nby);
    p.phi_c0 = (float)(0.5 - m0); p.phi_m1 = (float)m1; p.cos_c0 = (float)(0.5 - mc);
  }
  const int nbins = nbx * nby;
  int dev = 0, max_smem = 0, sms = 0;
  SR_CUDA(cudaGetDevice(&dev));
  SR_REQUIRE(dev >= 0 && dev < 64, "sr_sphere_hist: device ordinal %d not supported", dev);
  {  // device limits are looked up once per device: a short histogram call is launch-latency bound
    static std::mutex attr_mu;
    static int attr_smem[64] = {0}, attr_sms[64] = {0};
    std::lock_guard<std::mutex> lock(attr_mu);
    if (!attr_sms[dev]) {
      SR_CUDA(cudaDeviceGetAttribute(&attr_smem[dev], cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
      SR_CUDA(cudaDeviceGetAttribute(&attr_sms[dev], cudaDevAttrMultiProcessorCount, dev));
    }
    max_smem = attr_smem[dev]; sms = attr_sms[dev];
  }
  // vectors per CTA: 16 when their 32-bit bins fit in shared memory (72 x 36: 162 KB), else 8 or 4
  int qshift = 2;
  while (qshift > 0 && ((size_t)(4 << qshift) * nbins + kHistPadLo + nby + 4) * 4 > (size_t)max_smem) --qshift;
  const int group = 4 << qshift;
  const size_t smem = ((size_t)group * nbins + kHistPadLo + nby + 4) * 4;      // padding: see the kernel's bin comment
  SR_REQUIRE(smem <= (size_t)max_smem, "sr_sphere_hist: %d bins do not fit in shared memory", nbins);
  static const int threads_env = [] { const char* e = getenv("SR_HIST_THREADS"); return e ? atoi(e) : 0; }();
  const int threads = (threads_env == 640 || threads_env == 512 || threads_env == 896 || threads_env == 1024) ? threads_env : kHistThreads;   // tuning hook
  const int framesPerPass = threads >> qshift;
  const int nGroups = (nR + group - 1) / group;
  // One CTA per SM at a time.  The number of frame blocks minimises waves x (samples per CTA + fixed cost per CTA):
  // the grid comes out close to a whole number of waves, and unequal CTAs (a last group with fewer vectors)
  // backfill when there are several waves.  The fixed cost -- zeroing and flushing the CTA's counters, ~3 us --
  // is worth ~10^4 samples; a CTA always sees at least 16 passes.
  long long nFB = 1;
  {
    const long long maxFB = std::max(1LL, std::min(65535LL, nFrames / (16LL * framesPerPass)));
    const long long hi = std::min(maxFB, (32LL * sms) / nGroups + 1);
    double best = 1e300;
    for (long long cand = 1; cand <= hi; ++cand) {
      const long long total = cand * nGroups, waves = (total + sms - 1) / sms;
      const double cost = (double)waves * ((double)group * (double)((nFrames + cand - 1) / cand) + 1.0e4);
      if (cost < best * (1.0 - 1e-9)) { best = cost; nFB = cand; }
    }
  }
  long long fpb = (nFrames + nFB - 1) / nFB;
  fpb = std::max(fpb, (long long)framesPerPass);
  nFB = (nFrames + fpb - 1) / fpb;
  SR_REQUIRE(nFB <= 65535 && fpb <= 2147483647LL, "sr_sphere_hist: %lld frame blocks exceed the grid limit", nFB);
  dim3 grid((unsigned)nGroups, (unsigned)nFB);
  const double2* edge_dir = (const double2*)d_edge_table;
  const double* edge_cos = d_edge_table + 2 * (nbx + 1);
  cudaStream_t st = (cudaStream_t)stream;
  // one int of device scratch per call in flight: a small ring allocated once per device.  (cudaMallocAsync is not
  // an option here: releasing the pool at the caller's next synchronisation costs ~0.4 s next to torch's allocator.)
  static std::mutex ring_mu;
  static int* ring[64] = {nullptr};
  static unsigned ring_next[64] = {0};
  int* d_amb_start = nullptr;
  {
    std::lock_guard<std::mutex> lock(ring_mu);
    if (!ring[dev]) SR_CUDA(cudaMalloc(&ring[dev], 256 * sizeof(int)));
    d_amb_start = ring[dev] + (ring_next[dev]++ & 255u);
  }
  sphere_hist_mark_kernel<<<1, 1, 0, st>>>(d_amb_count, amb_capacity, d_amb_start);
  const bool vec4 = nR % 4 == 0 && ((uintptr_t)d_vecs & 15) == 0;
  const bool dflt = nbx == 72 && nby == 36;                // the reference's default --histBin
  void (*kern)(const float*, long long, int, int, int, const HistParams, unsigned int*, long long*, int, int*);
  if (threads == 1024) kern = vec4 ? (dflt ? sphere_hist_kernel<true, 72, 36, 1024> : sphere_hist_kernel<true, 0, 0, 1024>)
                                   : (dflt ? sphere_hist_kernel<false, 72, 36, 1024> : sphere_hist_kernel<false, 0, 0, 1024>);
  else if (threads == 896) kern = vec4 ? (dflt ? sphere_hist_kernel<true, 72, 36, 896> : sphere_hist_kernel<true, 0, 0, 896>)
                                       : (dflt ? sphere_hist_kernel<false, 72, 36, 896> : sphere_hist_kernel<false, 0, 0, 896>);
  else if (threads == 640) kern = vec4 ? (dflt ? sphere_hist_kernel<true, 72, 36, 640> : sphere_hist_kernel<true, 0, 0, 640>)
                                  : (dflt ? sphere_hist_kernel<false, 72, 36, 640> : sphere_hist_kernel<false, 0, 0, 640>);
  else if (threads == 512) kern = vec4 ? (dflt ? sphere_hist_kernel<true, 72, 36, 512> : sphere_hist_kernel<true, 0, 0, 512>)
                                       : (dflt ? sphere_hist_kernel<false, 72, 36, 512> : sphere_hist_kernel<false, 0, 0, 512>);
  else kern = vec4 ? (dflt ? sphere_hist_kernel<true, 72, 36, 768> : sphere_hist_kernel<true, 0, 0, 768>)
                   : (dflt ? sphere_hist_kernel<false, 72, 36, 768> : sphere_hist_kernel<false, 0, 0, 768>);
  {  // opt in to the large dynamic shared memory; the limit of a (kernel, device) is only ever raised
    static std::mutex fa_mu;
    static std::vector<std::pair<std::pair<const void*, int>, size_t>> limit;
    std::lock_guard<std::mutex> lock(fa_mu);
    const std::pair<const void*, int> key((const void*)kern, dev);
    auto it = std::find_if(limit.begin(), limit.end(), [&](const auto& e) { return e.first == key; });
    if (it == limit.end() || it->second < smem) {
      SR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      if (it == limit.end()) limit.push_back({key, smem}); else it->second = smem;
    }
  }
  kern<<<grid, threads, smem, st>>>(d_vecs, nFrames, nR, (int)fpb, qshift, p, d_counts, d_amb_idx, amb_capacity,
                                    d_amb_count);
  SR_CUDA(cudaGetLastError());
  sphere_hist_resolve_kernel<<<sms, 256, 0, st>>>(d_vecs, nR, p, edge_dir, edge_cos, d_counts, d_amb_idx, amb_capacity,
                                                  d_amb_count, d_amb_start);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

// Host-buffer entry point: builds NumPy's edges (np.linspace(-pi, pi, nbx + 1), np.linspace(-1, 1, nby + 1)) and the
// edge-direction table itself, uploads the vectors, runs both passes and returns the counts plus the sample ids
// the device left undecided (what a caller without NumPy cannot reproduce bit for bit is exactly that short list).
extern "C" int sr_sphere_hist_host(const float* h_vecs, long long nFrames, int nR, const double* h_q_rot, int nbx, int nby,
                                   unsigned int* h_counts, long long* h_amb_idx, int amb_capacity, int* h_n_amb) {
  SR_REQUIRE(h_vecs && h_counts && h_amb_idx && h_n_amb, "sr_sphere_hist_host: null pointer");
  SR_REQUIRE(nFrames > 0 && nR > 0 && nbx > 0 && nby > 0 && amb_capacity > 0, "sr_sphere_hist_host: empty shape");
  const double kPi = 3.141592653589793;
  std::vector<double> table((size_t)2 * (nbx + 1) + (nby + 1));
  for (int i = 0; i <= nbx; ++i) {           // np.linspace: start + i * step, last point = stop
    const double e = (i == nbx) ? kPi : -kPi + i * ((kPi - (-kPi)) / nbx);
    table[2 * i] = cos(e); table[2 * i + 1] = sin(e);
  }
  for (int j = 0; j <= nby; ++j) table[2 * (nbx + 1) + j] = (j == nby) ? 1.0 : -1.0 + j * (2.0 / nby);
  const size_t in_bytes = (size_t)nFrames * nR * 3 * sizeof(float);
  const size_t cnt_bytes = (size_t)nR * nbx * nby * sizeof(unsigned int);
  float* d_in = nullptr; double* d_table = nullptr; unsigned int* d_counts = nullptr; long long* d_amb = nullptr; int* d_namb = nullptr;
  int rc = SR_OK;
  cudaError_t e = cudaSuccess;
  if ((e = cudaMalloc(&d_in, in_bytes)) != cudaSuccess || (e = cudaMalloc(&d_table, table.size() * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(&d_counts, cnt_bytes)) != cudaSuccess || (e = cudaMalloc(&d_amb, (size_t)amb_capacity * sizeof(long long))) != cudaSuccess ||
      (e = cudaMalloc(&d_namb, sizeof(int))) != cudaSuccess ||
      (e = cudaMemcpy(d_in, h_vecs, in_bytes, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(d_table, table.data(), table.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemset(d_counts, 0, cnt_bytes)) != cudaSuccess || (e = cudaMemset(d_namb, 0, sizeof(int))) != cudaSuccess) {
    sr_set_error("sr_sphere_hist_host: allocation / upload failed: %s", cudaGetErrorString(e));
    rc = SR_ERR_CUDA;
  }
  // margins: FP64 resolve pass for the rotated stream; the reference's own float32 error for the unrotated one
  const double tol_phi = h_q_rot ? 1e-11 : 4e-6, tol_cos = h_q_rot ? 1e-11 : 2e-6;
  if (!rc) rc = sr_sphere_hist(d_in, nFrames, nR, h_q_rot, nbx, nby, d_table, tol_phi, tol_cos, d_counts, d_amb, amb_capacity,
                               d_namb, nullptr);
  int n = 0;
  std::vector<long long> list;
  if (!rc && ((e = cudaMemcpy(h_counts, d_counts, cnt_bytes, cudaMemcpyDeviceToHost)) != cudaSuccess ||
              (e = cudaMemcpy(&n, d_namb, sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess)) {
    sr_set_error("sr_sphere_hist_host: download failed: %s", cudaGetErrorString(e));
    rc = SR_ERR_CUDA;
  }
  if (!rc && n > amb_capacity) {
    sr_set_error("sr_sphere_hist_host: %d retry samples exceed the list capacity %d", n, amb_capacity);
    rc = SR_ERR_OVERFLOW;
  }
  if (!rc && n > 0) {
    list.resize(n);
    if ((e = cudaMemcpy(list.data(), d_amb, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost)) != cudaSuccess) {
      sr_set_error("sr_sphere_hist_host: download failed: %s", cudaGetErrorString(e));
      rc = SR_ERR_CUDA;
    }
  }
  if (!rc) {
    int k = 0;
    for (int i = 0; i < n; ++i)
      if (list[i] >= 0) h_amb_idx[k++] = list[i];      // -1 = resolved on the device
    *h_n_amb = k;
  }
  cudaFree(d_in); cudaFree(d_table); cudaFree(d_counts); cudaFree(d_amb); cudaFree(d_namb);
  return rc;
}
