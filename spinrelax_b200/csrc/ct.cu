// C(t) = <P2(u(t).u(t+delta))> hot path: K2 pack, K1 lag-tiled autocorrelation, Palmer finalize.
// Reference arithmetic: calculate_Ct_Palmer, calculate-Ct-from-traj.py:200-238.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// K1 geometry.  One CTA = one (vector, chunk, lag tile).  Inside the CTA every lane owns R consecutive
// lags (lane l: d0 + l*R + j), so a warp covers TL = 32*R lags and the left vector u(t) is a
// shared-memory broadcast.  R is odd so that the strided LDS.128 of the sliding window
// (lane stride = R float4) is bank-conflict free.  The 8 warps split each staged frame tile of
// TF = 8 * R * MB frames; each lane keeps its R window vectors in registers and slides them by one
// frame per step (static circular indexing, fully unrolled over R), so one step costs
// 2 LDS.128 + 4R FMA-pipe instructions.
// ------------------------------------------------------------------------------------------------
// Tile geometry is a compile-time configuration; kDefaultVariant is what the product launches, the other
// instantiations exist so that tools/tune_ct.py can time them on the GPU.
template <int R_, int MB_, int FB_, int NW_, int MINB_>
struct CtCfg {
  static constexpr int R = R_;            // lags per lane (odd)
  static constexpr int MB = MB_;          // R-step blocks per warp per frame tile
  static constexpr int FB = FB_;          // blocks between FP32 -> FP64 flushes (R*FB terms per FP32 partial sum)
  static constexpr int NW = NW_;          // warps per CTA
  static constexpr int MINB = MINB_;      // CTAs per SM promised to the compiler
  static constexpr int TFW = R * MB;      // frames per warp per tile
  static constexpr int TF = NW * TFW;     // frames per tile
  static constexpr int TL = 32 * R;       // lags per tile
  static constexpr int StageVecs = TF + (TF + TL);   // left range + window range, float4 each
  static constexpr int StageBytes = StageVecs * 16;
  static constexpr int SmemBytes = 2 * StageBytes;
  static_assert(R % 2 == 1, "R must be odd: lane stride of the window LDS.128 has to be conflict free");
  static_assert(MB % FB == 0, "flush interval must divide the warp tile");
  static_assert(NW * TL * 8 <= SmemBytes, "epilogue reduction buffer must fit in the stage buffers");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NW * 32, Cfg::MINB)
ct_lag_kernel(const float4* __restrict__ U, long long pitch, int nF, int L, int nLT, double* __restrict__ S) {
  constexpr int kR = Cfg::R, kMB = Cfg::MB, kFB = Cfg::FB, kNW = Cfg::NW, kTFW = Cfg::TFW, kTF = Cfg::TF,
                kTL = Cfg::TL, kStageVecs = Cfg::StageVecs, kStageBytes = Cfg::StageBytes;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[2];
  float4* const stage_base = reinterpret_cast<float4*>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long item = blockIdx.x;
  const long long rc = item / nLT;
  const int lt = (int)(item % nLT);
  const int d0 = 1 + lt * kTL;      // smallest lag of this tile
  const int nSteps = nF - d0;       // t in [0, nSteps) has at least one valid pair in this tile
  const int nTiles = (nSteps + kTF - 1) / kTF;
  const float4* const row = U + rc * pitch;

  if (tid == 0) {
    sr_mbar_init(&full_bar[0], 1);
    sr_mbar_init(&full_bar[1], 1);
    sr_fence_barrier_init();
  }
  __syncthreads();

  auto issue = [&](int k) {
    const int st = k & 1;
    float4* Ls = stage_base + st * kStageVecs;
    float4* Ws = Ls + kTF;
    sr_mbar_expect_tx(&full_bar[st], kStageBytes);
    sr_tma_load_1d(Ls, row + (long long)k * kTF, kTF * 16, &full_bar[st]);
    sr_tma_load_1d(Ws, row + (long long)k * kTF + d0, (kTF + kTL) * 16, &full_bar[st]);
  };

  double acc64[kR];
#pragma unroll
  for (int j = 0; j < kR; ++j) acc64[j] = 0.0;

  if (tid == 0 && nTiles > 0) issue(0);
  const int o = lane * kR;   // lag offset of this lane inside the tile
  const int s0 = warp * kTFW;

  for (int k = 0; k < nTiles; ++k) {
    if (tid == 0 && k + 1 < nTiles) issue(k + 1);   // stage (k+1)&1 was released by the barrier ending tile k-1
    sr_mbar_wait(&full_bar[k & 1], (k >> 1) & 1);
    const float4* __restrict__ Ls = stage_base + (k & 1) * kStageVecs;
    const float4* __restrict__ Ws = Ls + kTF;

    if (k * kTF + s0 < nSteps) {   // warp-uniform: whole warp range is past the last valid pair otherwise
      float wx[kR], wy[kR], wz[kR];
#pragma unroll
      for (int j = 0; j < kR; ++j) {
        const float4 v = Ws[s0 + o + j];
        wx[j] = v.x; wy[j] = v.y; wz[j] = v.z;
      }
#pragma unroll 1
      for (int b = 0; b < kMB; b += kFB) {
        float acc[kR];
#pragma unroll
        for (int j = 0; j < kR; ++j) acc[j] = 0.f;
#pragma unroll 1
        for (int bb = 0; bb < kFB; ++bb) {
          const int s = s0 + (b + bb) * kR;
#pragma unroll
          for (int kk = 0; kk < kR; ++kk) {
            const float4 a = Ls[s + kk];                 // u(t), broadcast
            const float4 nx = Ws[s + kk + o + kR];       // frame entering the window
#pragma unroll
            for (int j = 0; j < kR; ++j) {
              const int sl = (kk + j) % kR;              // slot holding frame t + d0 + o + j
              float d = a.x * wx[sl];
              d = fmaf(a.y, wy[sl], d);
              d = fmaf(a.z, wz[sl], d);
              acc[j] = fmaf(d, d, acc[j]);
            }
            wx[kk] = nx.x; wy[kk] = nx.y; wz[kk] = nx.z;
          }
        }
#pragma unroll
        for (int j = 0; j < kR; ++j) acc64[j] += (double)acc[j];
      }
    }
    sr_fence_proxy_async();
    __syncthreads();
  }

  // cross-warp reduction through the (now idle) stage buffers, one store per lag
  double* red = reinterpret_cast<double*>(smem_raw);
#pragma unroll
  for (int j = 0; j < kR; ++j) red[warp * kTL + o + j] = acc64[j];
  __syncthreads();
  for (int i = tid; i < kTL; i += kNW * 32) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kNW; ++w) s += red[w * kTL + i];
    const int lag = d0 + i;
    if (lag <= L) S[rc * (long long)L + (lag - 1)] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// K2: AoS (chunk, frame, vector, xyz) float32 -> vector-major float4 rows with zero padding.
// Tile = 64 frames x 32 vectors through shared memory so that both sides are coalesced.
// ------------------------------------------------------------------------------------------------
constexpr int kPF = 64, kPV = 32;

__global__ void __launch_bounds__(256)
pack_kernel(const float* __restrict__ vecs, int nC, long long nF, int nR, float4* __restrict__ U, long long pitch,
            int do_rot, double qw, double qx, double qy, double qz) {
  __shared__ float tile[kPF][kPV * 3 + 1];
  const int c = blockIdx.z;
  const long long f0 = (long long)blockIdx.x * kPF;
  const int r0 = blockIdx.y * kPV;
  const int nv = min(kPV, nR - r0);
  const int tid = threadIdx.x;

  if (f0 < nF) {
    const int rowlen = nv * 3;
    for (int e = tid; e < kPF * rowlen; e += 256) {
      const int fl = e / rowlen, col = e - fl * rowlen;
      const long long f = f0 + fl;
      float v = 0.f;
      if (f < nF) v = vecs[(((long long)c * nF + f) * nR + r0) * 3 + col];
      tile[fl][col] = v;
    }
  }
  __syncthreads();
  for (int e = tid; e < kPF * nv; e += 256) {
    const int rl = e / kPF, fl = e - rl * kPF;
    const long long f = f0 + fl;
    if (f >= pitch) continue;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f < nF) {
      float x = tile[fl][rl * 3 + 0], y = tile[fl][rl * 3 + 1], z = tile[fl][rl * 3 + 2];
      if (do_rot) {
        // a = q_v x v + q_w v ; b = q_v x a ; out = b + b + v   (transforms3d_supplement.py:283-296)
        const double vx = x, vy = y, vz = z;
        const double ax = qy * vz - qz * vy + qw * vx;
        const double ay = qz * vx - qx * vz + qw * vy;
        const double az = qx * vy - qy * vx + qw * vz;
        const double bx = qy * az - qz * ay, by = qz * ax - qx * az, bz = qx * ay - qy * ax;
        x = (float)(bx + bx + vx); y = (float)(by + by + vy); z = (float)(bz + bz + vz);
      }
      o = make_float4(x, y, z, 0.f);
    }
    U[((long long)(r0 + rl) * nC + c) * pitch + f] = o;
  }
}

// ------------------------------------------------------------------------------------------------
// Palmer finalize: chunk means of P2, then mean and std/(sqrt(nC)-1) over chunks (:226-228).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ct_finalize_kernel(const double* __restrict__ S, int nC, int nF, int nR, int L, float* __restrict__ Ct,
                   float* __restrict__ dCt) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)L * nR) return;
  const int r = (int)(idx / L);
  const int di = (int)(idx - (long long)r * L);   // delta - 1
  const double nVals = (double)(nF - (di + 1));
  double mean = 0.0;
  for (int c = 0; c < nC; ++c) {
    const double m = -0.5 + 1.5 * (S[((long long)r * nC + c) * L + di] / nVals);
    mean += m;
  }
  mean /= nC;
  double var = 0.0;
  for (int c = 0; c < nC; ++c) {
    const double m = -0.5 + 1.5 * (S[((long long)r * nC + c) * L + di] / nVals);
    var += (m - mean) * (m - mean);
  }
  var /= nC;
  Ct[(long long)di * nR + r] = (float)mean;
  dCt[(long long)di * nR + r] = (float)(sqrt(var) / (sqrt((double)nC) - 1.0));
}

// ================================================================================================
// C ABI
// ================================================================================================
// variants (R, MB, FB, NW, CTAs/SM); tools/tune_ct.py times them, kDefaultVariant is the product path
using CtV0 = CtCfg<15, 8, 1, 8, 2>;
using CtV1 = CtCfg<15, 8, 2, 8, 2>;
using CtV2 = CtCfg<15, 8, 4, 8, 2>;
using CtV3 = CtCfg<15, 8, 2, 8, 1>;
using CtV4 = CtCfg<15, 8, 2, 12, 1>;
using CtV5 = CtCfg<13, 8, 2, 8, 2>;
using CtV6 = CtCfg<17, 8, 2, 8, 2>;
using CtV7 = CtCfg<15, 12, 2, 8, 2>;
constexpr int kNumVariants = 8;
constexpr int kDefaultVariant = 0;   // flush every 15 terms: dCt stays within 1e-5 of the float64 reference
constexpr int kMaxTF = 1440, kMaxTL = 32 * 17;   // padding must cover the largest tile of any variant

template <class Cfg>
int launch_ct_lag(const float4* U, long long pitch, long long nF, int nRC, long long L, double* S, cudaStream_t st) {
  const long long nLT = (L + Cfg::TL - 1) / Cfg::TL;
  const long long items = (long long)nRC * nLT;
  SR_REQUIRE(items < (1LL << 31), "sr_ct_lag_sums: %lld work items exceed the grid limit", items);
  SR_CUDA(cudaFuncSetAttribute(ct_lag_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SmemBytes));
  ct_lag_kernel<Cfg><<<(unsigned)items, Cfg::NW * 32, Cfg::SmemBytes, st>>>(U, pitch, (int)nF, (int)L, (int)nLT, S);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

}  // namespace

extern "C" long long sr_ct_row_pitch(long long nF) { return sr_round_up(nF + kMaxTF + kMaxTL + 64, 8); }

extern "C" size_t sr_ct_workspace_bytes(int nC, long long nF, int nR) {
  const long long pitch = sr_ct_row_pitch(nF);
  const long long L = nF / 2;
  size_t packed = (size_t)nR * nC * pitch * 16;
  size_t sums = (size_t)nR * nC * L * 8;
  return sr_round_up((long long)packed, 256) + sr_round_up((long long)sums, 256);
}

extern "C" int sr_pack_vectors_f32(const float* d_vecs, int nC, long long nF, int nR, const double* h_q_rot,
                                   void* d_packed, long long pitch, void* stream) {
  SR_REQUIRE(d_vecs && d_packed, "sr_pack_vectors_f32: null pointer");
  SR_REQUIRE(nC > 0 && nF > 0 && nR > 0, "sr_pack_vectors_f32: empty shape (nC=%d nF=%lld nR=%d)", nC, nF, nR);
  SR_REQUIRE(pitch >= nF, "sr_pack_vectors_f32: pitch %lld < nF %lld", pitch, nF);
  SR_REQUIRE(nC <= 65535, "sr_pack_vectors_f32: nC %d exceeds grid limit", nC);
  double q[4] = {1, 0, 0, 0};
  int do_rot = 0;
  if (h_q_rot) {
    const double n = sqrt(h_q_rot[0] * h_q_rot[0] + h_q_rot[1] * h_q_rot[1] + h_q_rot[2] * h_q_rot[2] +
                          h_q_rot[3] * h_q_rot[3]);
    SR_REQUIRE(n > 0, "sr_pack_vectors_f32: zero rotation quaternion");
    for (int i = 0; i < 4; ++i) q[i] = h_q_rot[i] / n;   // vecnorm_NDarray(q), transforms3d_supplement.py:280
    do_rot = 1;
  }
  dim3 grid((unsigned)((pitch + kPF - 1) / kPF), (unsigned)((nR + kPV - 1) / kPV), (unsigned)nC);
  pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_vecs, nC, nF, nR, (float4*)d_packed, pitch, do_rot, q[0], q[1],
                                                      q[2], q[3]);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_ct_lag_sums_variant(const void* d_packed, long long pitch, int nC, long long nF, int nR, long long L,
                                      double* d_S, int variant, void* stream) {
  SR_REQUIRE(d_packed && d_S, "sr_ct_lag_sums: null pointer");
  SR_REQUIRE(nC > 0 && nR > 0 && nF >= 2, "sr_ct_lag_sums: bad shape (nC=%d nF=%lld nR=%d)", nC, nF, nR);
  SR_REQUIRE(L >= 1 && L <= nF - 1, "sr_ct_lag_sums: L=%lld outside [1, nF-1]", L);
  SR_REQUIRE(nF < (1LL << 30), "sr_ct_lag_sums: nF too large for 32-bit tile indices");
  SR_REQUIRE(pitch >= nF + kMaxTF + kMaxTL, "sr_ct_lag_sums: pitch %lld lacks %d frames of zero padding", pitch,
             kMaxTF + kMaxTL);
  SR_REQUIRE((long long)nR * nC < (1LL << 31), "sr_ct_lag_sums: too many (vector, chunk) rows");
  const float4* U = (const float4*)d_packed;
  cudaStream_t st = (cudaStream_t)stream;
  const int nRC = nR * nC;
  switch (variant) {
    case 0: return launch_ct_lag<CtV0>(U, pitch, nF, nRC, L, d_S, st);
    case 1: return launch_ct_lag<CtV1>(U, pitch, nF, nRC, L, d_S, st);
    case 2: return launch_ct_lag<CtV2>(U, pitch, nF, nRC, L, d_S, st);
    case 3: return launch_ct_lag<CtV3>(U, pitch, nF, nRC, L, d_S, st);
    case 4: return launch_ct_lag<CtV4>(U, pitch, nF, nRC, L, d_S, st);
    case 5: return launch_ct_lag<CtV5>(U, pitch, nF, nRC, L, d_S, st);
    case 6: return launch_ct_lag<CtV6>(U, pitch, nF, nRC, L, d_S, st);
    case 7: return launch_ct_lag<CtV7>(U, pitch, nF, nRC, L, d_S, st);
    default: break;
  }
  sr_set_error("sr_ct_lag_sums_variant: unknown variant %d (have %d)", variant, kNumVariants);
  return SR_ERR_ARG;
}

extern "C" int sr_ct_lag_sums(const void* d_packed, long long pitch, int nC, long long nF, int nR, long long L,
                              double* d_S, void* stream) {
  return sr_ct_lag_sums_variant(d_packed, pitch, nC, nF, nR, L, d_S, kDefaultVariant, stream);
}

extern "C" int sr_ct_palmer_finalize(const double* d_S, int nC, long long nF, int nR, long long L, float* d_Ct,
                                     float* d_dCt, void* stream) {
  SR_REQUIRE(d_S && d_Ct && d_dCt, "sr_ct_palmer_finalize: null pointer");
  SR_REQUIRE(nC > 0 && nR > 0 && L >= 1 && L < nF, "sr_ct_palmer_finalize: bad shape");
  const long long n = L * nR;
  ct_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_S, nC, (int)nF, nR, (int)L, d_Ct,
                                                                                    d_dCt);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_ct_palmer_device(const float* d_vecs, int nC, long long nF, int nR, float* d_Ct, float* d_dCt,
                                   void* d_workspace, size_t workspace_bytes, void* stream) {
  SR_REQUIRE(d_vecs && d_Ct && d_dCt && d_workspace, "sr_ct_palmer_device: null pointer");
  SR_REQUIRE(nF >= 2, "sr_ct_palmer_device: need at least 2 frames per chunk (nF=%lld)", nF);
  const size_t need = sr_ct_workspace_bytes(nC, nF, nR);
  if (workspace_bytes < need) {
    sr_set_error("sr_ct_palmer_device: workspace %zu < required %zu bytes", workspace_bytes, need);
    return SR_ERR_WORKSPACE;
  }
  const long long pitch = sr_ct_row_pitch(nF);
  const long long L = nF / 2;
  char* ws = (char*)d_workspace;
  void* packed = ws;
  double* S = (double*)(ws + sr_round_up((long long)((size_t)nR * nC * pitch * 16), 256));
  int rc = sr_pack_vectors_f32(d_vecs, nC, nF, nR, nullptr, packed, pitch, stream);
  if (rc) return rc;
  rc = sr_ct_lag_sums(packed, pitch, nC, nF, nR, L, S, stream);
  if (rc) return rc;
  return sr_ct_palmer_finalize(S, nC, nF, nR, L, d_Ct, d_dCt, stream);
}

extern "C" int sr_ct_palmer_host(const float* h_vecs, int nC, long long nF, int nR, float* h_Ct, float* h_dCt) {
  SR_REQUIRE(h_vecs && h_Ct && h_dCt, "sr_ct_palmer_host: null pointer");
  SR_REQUIRE(nC > 0 && nR > 0 && nF >= 2, "sr_ct_palmer_host: bad shape");
  const long long L = nF / 2;
  const size_t in_bytes = (size_t)nC * nF * nR * 3 * sizeof(float);
  const size_t out_bytes = (size_t)L * nR * sizeof(float);
  const size_t ws_bytes = sr_ct_workspace_bytes(nC, nF, nR);
  float *d_in = nullptr, *d_Ct = nullptr, *d_dCt = nullptr;
  void* d_ws = nullptr;
  int rc = SR_OK;
  cudaError_t e;
  if ((e = cudaMalloc(&d_in, in_bytes)) != cudaSuccess || (e = cudaMalloc(&d_Ct, out_bytes)) != cudaSuccess ||
      (e = cudaMalloc(&d_dCt, out_bytes)) != cudaSuccess || (e = cudaMalloc(&d_ws, ws_bytes)) != cudaSuccess) {
    sr_set_error("sr_ct_palmer_host: cudaMalloc failed: %s", cudaGetErrorString(e));
    rc = SR_ERR_CUDA;
  }
  if (!rc && (e = cudaMemcpy(d_in, h_vecs, in_bytes, cudaMemcpyHostToDevice)) != cudaSuccess) {
    sr_set_error("sr_ct_palmer_host: H2D failed: %s", cudaGetErrorString(e));
    rc = SR_ERR_CUDA;
  }
  if (!rc) rc = sr_ct_palmer_device(d_in, nC, nF, nR, d_Ct, d_dCt, d_ws, ws_bytes, nullptr);
  if (!rc && ((e = cudaMemcpy(h_Ct, d_Ct, out_bytes, cudaMemcpyDeviceToHost)) != cudaSuccess ||
              (e = cudaMemcpy(h_dCt, d_dCt, out_bytes, cudaMemcpyDeviceToHost)) != cudaSuccess)) {
    sr_set_error("sr_ct_palmer_host: D2H failed: %s", cudaGetErrorString(e));
    rc = SR_ERR_CUDA;
  }
  cudaFree(d_in); cudaFree(d_Ct); cudaFree(d_dCt); cudaFree(d_ws);
  return rc;
}
