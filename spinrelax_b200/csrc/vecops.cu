// Element-wise / reduction helpers on bond-vector streams that must follow IEEE semantics exactly (compiled
// without -ftz, unlike hist.cu): the bit-identical PAF rotation and the block moments behind --vecAvg / --S2.
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// rotate_vector_simd (transforms3d_supplement.py:270-296) for float32 vectors and one float64 quaternion:
// float64 result, every product and sum rounded separately in NumPy's order so the output is bit-identical.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
rotate_f32_f64_kernel(const float* __restrict__ v, long long n, double qw, double qx, double qy, double qz,
                      double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double vx = v[3 * i], vy = v[3 * i + 1], vz = v[3 * i + 2];
  // a = cross(q_v, v) + q_w * v
  const double ax = __dadd_rn(__dsub_rn(__dmul_rn(qy, vz), __dmul_rn(qz, vy)), __dmul_rn(qw, vx));
  const double ay = __dadd_rn(__dsub_rn(__dmul_rn(qz, vx), __dmul_rn(qx, vz)), __dmul_rn(qw, vy));
  const double az = __dadd_rn(__dsub_rn(__dmul_rn(qx, vy), __dmul_rn(qy, vx)), __dmul_rn(qw, vz));
  // b = cross(q_v, a) ; out = b + b + v
  const double bx = __dsub_rn(__dmul_rn(qy, az), __dmul_rn(qz, ay));
  const double by = __dsub_rn(__dmul_rn(qz, ax), __dmul_rn(qx, az));
  const double bz = __dsub_rn(__dmul_rn(qx, ay), __dmul_rn(qy, ax));
  out[3 * i] = __dadd_rn(__dadd_rn(bx, bx), vx);
  out[3 * i + 1] = __dadd_rn(__dadd_rn(by, by), vy);
  out[3 * i + 2] = __dadd_rn(__dadd_rn(bz, bz), vz);
}
}  // namespace

extern "C" int sr_rotate_vectors_f32_f64(const float* d_v, long long n, const double* h_q, double* d_out, void* stream) {
  SR_REQUIRE(d_v && h_q && d_out, "sr_rotate_vectors_f32_f64: null pointer");
  SR_REQUIRE(n >= 1, "sr_rotate_vectors_f32_f64: empty input");
  rotate_f32_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_v, n, h_q[0], h_q[1], h_q[2],
                                                                                     h_q[3], d_out);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

// ------------------------------------------------------------------------------------------------
// First and second moments of the bond vectors per (block of frames, vector): the reductions behind
// --vecAvg (calculate-Ct-from-traj.py:579-583) and --S2 (calculate_S2_by_outerProduct :96-145), which
// run-all.bash:481 always requests together with --Ct.  out[block][r][9] = sum x,y,z,xx,xy,xz,yy,yz,zz (FP64).
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
vec_block_moments_kernel(const float* __restrict__ vecs, long long nFrames, int nR, long long framesPerBlock,
                         int tilesPerBlock, double* __restrict__ out) {
  // grid.x = vector groups of 16, grid.y = block * tilesPerBlock + tile
  __shared__ double red[9][256];
  const int r0 = blockIdx.x * 16;
  const int vl = threadIdx.x & 15, fsub = threadIdx.x >> 4;       // 16 frames per pass
  const long long blk = blockIdx.y / tilesPerBlock;
  const int tile = blockIdx.y - (int)blk * tilesPerBlock;
  const long long tileLen = (framesPerBlock + tilesPerBlock - 1) / tilesPerBlock;
  const long long fa = blk * framesPerBlock + (long long)tile * tileLen;
  const long long fb = min(min(nFrames, (blk + 1) * framesPerBlock), fa + tileLen);
  double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (r0 + vl < nR) {
    for (long long f = fa + fsub; f < fb; f += 16) {
      const float* src = vecs + (f * nR + r0 + vl) * 3;
      const double x = __ldg(src), y = __ldg(src + 1), z = __ldg(src + 2);
      s[0] += x; s[1] += y; s[2] += z;
      s[3] = fma(x, x, s[3]); s[4] = fma(x, y, s[4]); s[5] = fma(x, z, s[5]);
      s[6] = fma(y, y, s[6]); s[7] = fma(y, z, s[7]); s[8] = fma(z, z, s[8]);
    }
  }
#pragma unroll
  for (int m = 0; m < 9; ++m) red[m][threadIdx.x] = s[m];
  __syncthreads();
  if (threadIdx.x < 16 * 9) {
    const int v = threadIdx.x & 15, m = threadIdx.x >> 4;
    double t = 0.0;
    for (int k = 0; k < 16; ++k) t += red[m][k * 16 + v];
    if (r0 + v < nR) atomicAdd(&out[(blk * nR + r0 + v) * 9 + m], t);
  }
}
}  // namespace

extern "C" int sr_vec_block_moments(const float* d_vecs, long long nFrames, int nR, long long framesPerBlock,
                                    double* d_out, void* stream) {
  SR_REQUIRE(d_vecs && d_out, "sr_vec_block_moments: null pointer");
  SR_REQUIRE(nFrames > 0 && nR > 0 && framesPerBlock > 0, "sr_vec_block_moments: empty shape");
  const long long nBlocks = (nFrames + framesPerBlock - 1) / framesPerBlock;
  const int groups = (nR + 15) / 16;
  long long tiles = (2048 + groups * nBlocks - 1) / (groups * nBlocks);       // ~2048 CTAs in flight
  const long long maxTiles = (framesPerBlock + 255) / 256;
  if (tiles > maxTiles) tiles = maxTiles;
  if (tiles < 1) tiles = 1;
  SR_REQUIRE(nBlocks * tiles <= 65535, "sr_vec_block_moments: too many frame blocks");
  SR_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * 9 * (size_t)nBlocks * nR, (cudaStream_t)stream));
  dim3 grid((unsigned)groups, (unsigned)(nBlocks * tiles));
  vec_block_moments_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_vecs, nFrames, nR, framesPerBlock, (int)tiles, d_out);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

// ------------------------------------------------------------------------------------------------
// gm.xyz_to_rtp (general_maths.py:118-158), last-axis layout, in the precision of the input like NumPy:
//   full form  (r, phi, theta) = (|v|, atan2(y, x), acos(z / r)); |v| = sqrt((x*x + y*y) + z*z) with every
//              operation rounded separately (np.linalg.norm over an axis of length 3) -> r is bit-identical;
//   bUnit form (phi, acos(z / phi)) -- the reference divides by phi there (:131-133) and so does this.
// phi / theta differ from glibc / SVML by the last ulp or two (CUDA atan2/acos are <= 2 ulp), never more.
// ------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ float sr_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double sr_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float sr_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double sr_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sr_sqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sr_sqrt(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float sr_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double sr_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float sr_atan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double sr_atan2(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float sr_acos(float a) { return acosf(a); }
__device__ __forceinline__ double sr_acos(double a) { return acos(a); }

template <typename T>
__device__ __forceinline__ void rtp_one(T x, T y, T z, int unitForm, T& r, T& phi, T& theta) {
  phi = sr_atan2(y, x);
  if (unitForm) {
    r = phi;
    theta = sr_acos(sr_div(z, phi));
  } else {
    r = sr_sqrt(sr_add(sr_add(sr_mul(x, x), sr_mul(y, y)), sr_mul(z, z)));
    theta = sr_acos(sr_div(z, r));
  }
}

// Main kernel: a thread owns 4 consecutive vectors = 12 values, moved as 16-byte loads / stores straight from and
// to registers (3 x LDG.128 for float, 6 for double); 48-96 B in flight per thread keeps HBM busy without staging.
template <typename T>
__global__ void __launch_bounds__(256)
xyz_to_rtp_vec4_kernel(const T* __restrict__ v, long long nGroups, T* __restrict__ out, int unitForm) {
  constexpr int kPer16 = 16 / (int)sizeof(T);            // values per 16-byte chunk
  constexpr int kIn = 12 / kPer16;                       // chunks in, 3 or 6
  const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
  if (g >= nGroups) return;
  union { int4 q[kIn]; T s[12]; } in, res;
  const int4* src = reinterpret_cast<const int4*>(v + g * 12);
#pragma unroll
  for (int k = 0; k < kIn; ++k) in.q[k] = __ldg(src + k);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    T r, phi, theta;
    rtp_one<T>(in.s[3 * k], in.s[3 * k + 1], in.s[3 * k + 2], unitForm, r, phi, theta);
    if (unitForm) {
      res.s[2 * k] = phi; res.s[2 * k + 1] = theta;
    } else {
      res.s[3 * k] = r; res.s[3 * k + 1] = phi; res.s[3 * k + 2] = theta;
    }
  }
  if (unitForm) {
    int4* dst = reinterpret_cast<int4*>(out + g * 8);
#pragma unroll
    for (int k = 0; k < 8 / kPer16; ++k) dst[k] = res.q[k];
  } else {
    int4* dst = reinterpret_cast<int4*>(out + g * 12);
#pragma unroll
    for (int k = 0; k < kIn; ++k) dst[k] = res.q[k];
  }
}

// Tail (< 4 vectors) and unaligned buffers: one vector per thread.
template <typename T>
__global__ void __launch_bounds__(256)
xyz_to_rtp_scalar_kernel(const T* __restrict__ v, long long first, long long n, T* __restrict__ out, int unitForm) {
  const long long i = first + (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  T r, phi, theta;
  rtp_one<T>(v[3 * i], v[3 * i + 1], v[3 * i + 2], unitForm, r, phi, theta);
  if (unitForm) {
    out[2 * i] = phi; out[2 * i + 1] = theta;
  } else {
    out[3 * i] = r; out[3 * i + 1] = phi; out[3 * i + 2] = theta;
  }
}

template <typename T>
int xyz_to_rtp_launch(const T* d_v, long long n, T* d_out, int unitForm, void* stream) {
  SR_REQUIRE(d_v && d_out, "sr_xyz_to_rtp: null pointer");
  SR_REQUIRE(n >= 1, "sr_xyz_to_rtp: empty input");
  SR_REQUIRE(n / 1024 + 1 <= 2147483647LL, "sr_xyz_to_rtp: too many vectors");
  const bool aligned = ((uintptr_t)d_v % 16 == 0) && ((uintptr_t)d_out % 16 == 0);
  const long long nGroups = aligned ? n / 4 : 0;
  if (nGroups > 0) {
    xyz_to_rtp_vec4_kernel<T><<<(unsigned)((nGroups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_v, nGroups, d_out,
                                                                                                  unitForm);
    SR_CUDA(cudaGetLastError());
  }
  const long long rest = n - nGroups * 4;
  if (rest > 0) {
    SR_REQUIRE((rest + 255) / 256 <= 2147483647LL, "sr_xyz_to_rtp: too many vectors");
    xyz_to_rtp_scalar_kernel<T><<<(unsigned)((rest + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_v, nGroups * 4, n,
                                                                                                 d_out, unitForm);
    SR_CUDA(cudaGetLastError());
  }
  return SR_OK;
}
}  // namespace

extern "C" int sr_xyz_to_rtp_f32(const float* d_v, long long n, float* d_out, int unit_form, void* stream) {
  return xyz_to_rtp_launch<float>(d_v, n, d_out, unit_form, stream);
}
extern "C" int sr_xyz_to_rtp_f64(const double* d_v, long long n, double* d_out, int unit_form, void* stream) {
  return xyz_to_rtp_launch<double>(d_v, n, d_out, unit_form, stream);
}
