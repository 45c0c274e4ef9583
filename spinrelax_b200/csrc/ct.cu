// C(t) = <P2(u(t).u(t+delta))> hot path: K2 pack, K1 lag-tiled autocorrelation, Palmer finalize.
// Reference arithmetic: calculate_Ct_Palmer, calculate-Ct-from-traj.py:200-238.
#include "common.cuh"

#include <algorithm>
#include <mutex>

namespace {

// ------------------------------------------------------------------------------------------------
// Packed stream layout (K2 -> K1): one row per (vector, chunk) made of three float32 planes
//   row = [ X[pitch] | Y[pitch] | Z[pitch] ],  frames >= nF are zero,  pitch % 4 == 0  (16-byte planes for TMA)
// i.e. 12 algorithmic bytes per frame, structure-of-arrays.  SoA matters for K1: the left vector u(t) is read
// with three broadcast LDS.32 and the window frame with three lane-strided LDS.32, which costs measurably
// fewer issue cycles than two LDS.128 of (x,y,z,0) records (tools/microbench.cu, modes 0 vs 4: +8 %).
//
// K1 geometry.  One CTA = one (vector, chunk, lag tile).  Tile lt covers the lags lt*TL .. lt*TL + TL - 1
// (lag 0 of tile 0 is computed and dropped: starting tiles at multiples of TL keeps every TMA source
// 16-byte aligned).  Inside the CTA every lane owns R consecutive lags (lane l: D0 + l*R + j), so a warp
// covers TL = 32*R lags and the left vector u(t) is a shared-memory broadcast.  R is odd so that the
// lane-strided window loads (stride R floats) are bank-conflict free.  The NW warps split each staged frame
// tile of TF = NW * R * MB frames; each lane keeps its R window vectors in registers and slides them by one
// frame per step (static circular indexing, fully unrolled over R), so one step costs
// 6 LDS.32 + 4R FMA-pipe instructions.
// ------------------------------------------------------------------------------------------------
// Tile geometry is a compile-time configuration; CtLong / CtShort are what the product launches.  Building with
// -DSR_TUNING (tools/build_tune.py -> tools/libct_tune.so, never the product library) adds the table of experimental
// configurations and the entry point sr_ct_lag_sums_variant that tools/tune_ct.py times on the GPU.
//
// FLUSH selects how the FP32 partial sums reach the FP64 lag sums:
//   0  block flush: every FB blocks all R partial sums are converted and added to FP64 accumulators held in
//      registers, then zeroed (R x (F2F + DADD + MOV) in one burst);
//   1  staggered flush, FP64 accumulators in registers: in the last block of every FB, iteration kk of the unrolled
//      window loop flushes partial sum kk and restarts it with a plain multiply (the FFMA of that term becomes an
//      FMUL, so the reset costs nothing) -- one F2F + DADD per step instead of a burst, no MOVs;
//   2  staggered flush, FP64 accumulators in shared memory (one array of TL doubles per warp): frees 2R registers
//      per thread for a wider window (R up to 33 at 12 warps), at one LDS.64 + STS.64 per flushed sum.
// DIAG (tuning builds only) removes pieces of the loop to attribute their cost: 1 no flush, 2 no window loads,
// 3 no left-vector loads, 4 no shared-memory loads at all.  Results are wrong by construction.
template <int R_, int MB_, int FB_, int NW_, int MINB_, int NS_ = 2, int FLUSH_ = 0, int DIAG_ = 0, int ORDER_ = 0,
          int SYNC_ = 0>
struct CtCfg {
  // SYNC = 1: the warps that share an SM sub-partition (warp % 4) meet at a named barrier before every flush interval, so
  // that they walk the 25 KB loop body together and share the instruction lines fetched for it
  static constexpr int SYNC = SYNC_;
  static constexpr int ORDER = ORDER_;    // 0: lag after lag; 1: in phases (R multiplies, R + R dot FMAs, R accumulates); 2: phases, order pinned
  static constexpr int NS = NS_;          // TMA stages in the ring
  static constexpr int R = R_;            // lags per lane (odd)
  static constexpr int MB = MB_;          // R-step blocks per warp per frame tile
  static constexpr int FB = FB_;          // blocks between FP32 -> FP64 flushes (R*FB terms per FP32 partial sum)
  static constexpr int NW = NW_;          // warps per CTA
  static constexpr int MINB = MINB_;      // CTAs per SM promised to the compiler
  static constexpr int FLUSH = FLUSH_;
  static constexpr int DIAG = DIAG_;
  static constexpr int TFW = R * MB;      // frames per warp per tile
  static constexpr int TF = NW * TFW;     // frames per tile
  static constexpr int TL = 32 * R;       // lags per tile
  static constexpr int TW = TF + TL;      // window frames per tile
  static constexpr int StageFloats = 3 * TF + 3 * TW;   // LX LY LZ WX WY WZ
  static constexpr int StageBytes = StageFloats * 4;
  static constexpr int AccBytes = FLUSH == 2 ? NW * TL * 8 : 0;   // per-warp FP64 lag sums in shared memory
  static constexpr int SmemBytes = NS * StageBytes + AccBytes;
  static_assert(R % 2 == 1, "R must be odd: lane stride of the window loads has to be conflict free");
  static_assert(MB % FB == 0, "flush interval must divide the warp tile");
  static_assert(TF % 4 == 0 && TL % 4 == 0, "TMA sources and destinations must stay 16-byte aligned");
  static_assert(NW * TL * 8 <= NS * StageBytes, "epilogue reduction buffer must fit in the stage buffers");
  static_assert(StageBytes % 16 == 0, "the FP64 accumulator arrays follow the stages and must stay aligned");
};

// One block of R steps for one lane: R x R (left frame, lag) pairs.  DOFLUSH: this is the last block of a flush
// interval (modes 1 and 2).
template <class Cfg, bool DOFLUSH>
__device__ __forceinline__ void ct_block(const float* __restrict__ LX, const float* __restrict__ LY,
                                         const float* __restrict__ LZ, const float* __restrict__ WX,
                                         const float* __restrict__ WY, const float* __restrict__ WZ, int s, int o,
                                         float (&wx)[Cfg::R], float (&wy)[Cfg::R], float (&wz)[Cfg::R],
                                         float (&acc)[Cfg::R], double (&acc64)[Cfg::FLUSH == 2 ? 1 : Cfg::R],
                                         double* __restrict__ sacc) {
  constexpr int kR = Cfg::R;
#pragma unroll
  for (int kk = 0; kk < kR; ++kk) {
    // SYNC = G + 1 >= 2: the warps of a sub-partition re-align every G steps (they execute identical instruction streams)
    // (no memory clobber: the barrier aligns the warps in time only, the loads of the tile may move across it)
    if (Cfg::SYNC >= 2 && Cfg::SYNC < 10 && kk % (Cfg::SYNC - 1) == 0)
      asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)((threadIdx.x >> 5) & 3)), "r"(32 * (Cfg::NW / 4)));
    if (Cfg::SYNC == 10)   // every step, memory clobber kept (reference point)
      asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)((threadIdx.x >> 5) & 3)), "r"(32 * (Cfg::NW / 4)) : "memory");
    float ax, ay, az, nx, ny, nz;
    if (Cfg::DIAG == 3 || Cfg::DIAG == 4) { ax = wx[(kk + 1) % kR]; ay = wy[(kk + 2) % kR]; az = wz[(kk + 3) % kR]; }
    else { ax = LX[s + kk]; ay = LY[s + kk]; az = LZ[s + kk]; }                  // u(t), broadcast
    const int iw = s + kk + o + kR;                                              // frame entering the window
    if (Cfg::DIAG == 2 || Cfg::DIAG == 4) { nx = wx[kk] + 1e-7f; ny = wy[kk]; nz = wz[kk]; }
    else { nx = WX[iw]; ny = WY[iw]; nz = WZ[iw]; }
    if (Cfg::SYNC == 12)   // re-align after the step's loads have been issued
      asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)((threadIdx.x >> 5) & 3)), "r"(32 * (Cfg::NW / 4)));
    if (DOFLUSH && Cfg::DIAG != 1) {
      if (Cfg::FLUSH == 1) acc64[kk] += (double)acc[kk];
      if (Cfg::FLUSH == 2) sacc[o + kk] += (double)acc[kk];
    }
    if (Cfg::ORDER == 0) {
#pragma unroll
      for (int j = 0; j < kR; ++j) {
        if (Cfg::SYNC == 11 && (j == 0 || j == kR / 2))
          asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)((threadIdx.x >> 5) & 3)), "r"(32 * (Cfg::NW / 4)));
        const int sl = (kk + j) % kR;              // slot holding frame t + d0 + o + j
        float d = ax * wx[sl];
        d = fmaf(ay, wy[sl], d);
        d = fmaf(az, wz[sl], d);
        if (DOFLUSH && j == kk) acc[j] = d * d;    // partial sum kk was just flushed: restart it
        else acc[j] = fmaf(d, d, acc[j]);
      }
    } else {
      // the same arithmetic in four phases: within a phase the left component sits in the operand-reuse cache
      float d[kR];
      if (Cfg::ORDER == 1) {
#pragma unroll
        for (int j = 0; j < kR; ++j) d[j] = ax * wx[(kk + j) % kR];
#pragma unroll
        for (int j = 0; j < kR; ++j) d[j] = fmaf(ay, wy[(kk + j) % kR], d[j]);
#pragma unroll
        for (int j = 0; j < kR; ++j) d[j] = fmaf(az, wz[(kk + j) % kR], d[j]);
#pragma unroll
        for (int j = 0; j < kR; ++j) acc[j] = (DOFLUSH && j == kk) ? d[j] * d[j] : fmaf(d[j], d[j], acc[j]);
      } else {
#pragma unroll
        for (int j = 0; j < kR; ++j) asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d[j]) : "f"(ax), "f"(wx[(kk + j) % kR]));
#pragma unroll
        for (int j = 0; j < kR; ++j) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(d[j]) : "f"(ay), "f"(wy[(kk + j) % kR]));
#pragma unroll
        for (int j = 0; j < kR; ++j) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(d[j]) : "f"(az), "f"(wz[(kk + j) % kR]));
#pragma unroll
        for (int j = 0; j < kR; ++j) {
          if (DOFLUSH && j == kk) acc[j] = d[j] * d[j];
          else asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(acc[j]) : "f"(d[j]));
        }
      }
    }
    wx[kk] = nx; wy[kk] = ny; wz[kk] = nz;
  }
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NW * 32, Cfg::MINB)
ct_lag_kernel(const float* __restrict__ U, long long pitch, int nF, int L, int nLT, int nC, int c0, int nCsub,
              double* __restrict__ S) {
  constexpr int kR = Cfg::R, kMB = Cfg::MB, kFB = Cfg::FB, kNW = Cfg::NW, kTFW = Cfg::TFW, kTF = Cfg::TF,
                kTL = Cfg::TL, kTW = Cfg::TW, kStageFloats = Cfg::StageFloats, kStageBytes = Cfg::StageBytes;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[Cfg::NS], empty_bar[Cfg::NS];
  constexpr int kNS = Cfg::NS;
  float* const stage_base = reinterpret_cast<float*>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long item = blockIdx.x;
  const long long rcs = item / nLT;                    // row inside the launched chunk range [c0, c0 + nCsub)
  const int lt = (int)(item % nLT);
  const long long rc = (rcs / nCsub) * nC + c0 + rcs % nCsub;   // row r * nC + c of the full packed stream / S
  const int d0 = lt * kTL;          // smallest lag of this tile
  const int nSteps = nF - d0;       // t in [0, nSteps) has at least one valid pair in this tile
  const int nTiles = (nSteps + kTF - 1) / kTF;
  const float* const rowX = U + rc * 3 * pitch;
  const float* const rowY = rowX + pitch;
  const float* const rowZ = rowY + pitch;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kNS; ++i) { sr_mbar_init(&full_bar[i], 1); sr_mbar_init(&empty_bar[i], kNW); }
    sr_fence_barrier_init();
  }
  double* const sacc_all = reinterpret_cast<double*>(smem_raw + (size_t)kNS * kStageBytes);
  if (Cfg::FLUSH == 2)
    for (int i = tid; i < kNW * kTL; i += kNW * 32) sacc_all[i] = 0.0;
  __syncthreads();

  auto issue = [&](int k, int st) {
    float* sb = stage_base + st * kStageFloats;
    const long long f = (long long)k * kTF;
    sr_mbar_expect_tx(&full_bar[st], kStageBytes);
    sr_tma_load_1d(sb, rowX + f, kTF * 4, &full_bar[st]);
    sr_tma_load_1d(sb + kTF, rowY + f, kTF * 4, &full_bar[st]);
    sr_tma_load_1d(sb + 2 * kTF, rowZ + f, kTF * 4, &full_bar[st]);
    sr_tma_load_1d(sb + 3 * kTF, rowX + f + d0, kTW * 4, &full_bar[st]);
    sr_tma_load_1d(sb + 3 * kTF + kTW, rowY + f + d0, kTW * 4, &full_bar[st]);
    sr_tma_load_1d(sb + 3 * kTF + 2 * kTW, rowZ + f + d0, kTW * 4, &full_bar[st]);
  };

  double acc64[Cfg::FLUSH == 2 ? 1 : kR];
#pragma unroll
  for (int j = 0; j < (Cfg::FLUSH == 2 ? 1 : kR); ++j) acc64[j] = 0.0;
  float acc[kR];                      // FP32 partial sums; with a staggered flush they live across blocks and tiles
#pragma unroll
  for (int j = 0; j < kR; ++j) acc[j] = 0.f;
  double* const sacc = sacc_all + warp * kTL;

  // Ring of kNS stages.  Thread 0 is the producer: before tile k it refills the stage that tile k-1 used
  // (tile k + kNS - 1), waiting on that stage's "empty" barrier (one arrival per warp).  Consumer warps
  // never meet at a CTA-wide barrier inside the loop, so a warp that finishes its share of a tile early
  // starts the next one at once.
  if (tid == 0)
    for (int i = 0; i < kNS - 1 && i < nTiles; ++i) issue(i, i);
  const int o = lane * kR;   // lag offset of this lane inside the tile
  const int s0 = warp * kTFW;

  int st = 0, ph = 0;                 // stage and full-barrier parity of tile k
  int pst = kNS - 1, pph = 1;         // producer: stage of tile k + kNS - 1; parity of the empty barrier ((kn / kNS) - 1) & 1
  for (int k = 0; k < nTiles; ++k) {
    if (tid == 0) {
      const int kn = k + kNS - 1;
      if (kn < nTiles) {
        if (kn >= kNS) { sr_mbar_wait(&empty_bar[pst], pph); sr_fence_proxy_async(); }
        issue(kn, pst);
      }
    }
    if (++pst == kNS) { pst = 0; pph ^= 1; }
    sr_mbar_wait(&full_bar[st], ph);
    const float* __restrict__ LX = stage_base + st * kStageFloats;
    const float* __restrict__ LY = LX + kTF;
    const float* __restrict__ LZ = LY + kTF;
    const float* __restrict__ WX = LZ + kTF;
    const float* __restrict__ WY = WX + kTW;
    const float* __restrict__ WZ = WY + kTW;

    // warp-uniform: the whole warp range is past the last valid pair otherwise.  With SYNC the test is CTA-uniform (the
    // warps of a sub-partition must reach the same barriers; the extra frames are zero padding and add nothing)
    if (k * kTF + (Cfg::SYNC ? 0 : s0) < nSteps) {
      float wx[kR], wy[kR], wz[kR];
#pragma unroll
      for (int j = 0; j < kR; ++j) {
        wx[j] = WX[s0 + o + j]; wy[j] = WY[s0 + o + j]; wz[j] = WZ[s0 + o + j];
      }
#pragma unroll 1
      for (int b = 0; b < kMB; b += kFB) {
        if (Cfg::SYNC == 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + (warp & 3)), "r"(32 * (kNW / 4)) : "memory");
        if (Cfg::FLUSH == 0 && Cfg::DIAG != 1) {
          float accb[kR];                    // partial sums of this flush interval only
#pragma unroll
          for (int j = 0; j < kR; ++j) accb[j] = 0.f;
#pragma unroll 1
          for (int bb = 0; bb < kFB; ++bb)
            ct_block<Cfg, false>(LX, LY, LZ, WX, WY, WZ, s0 + (b + bb) * kR, o, wx, wy, wz, accb, acc64, sacc);
#pragma unroll
          for (int j = 0; j < kR; ++j) acc64[j] += (double)accb[j];
        } else if (Cfg::FLUSH == 0) {
#pragma unroll 1
          for (int bb = 0; bb < kFB; ++bb)
            ct_block<Cfg, false>(LX, LY, LZ, WX, WY, WZ, s0 + (b + bb) * kR, o, wx, wy, wz, acc, acc64, sacc);
        } else {
#pragma unroll 1
          for (int bb = 0; bb < kFB - 1; ++bb)
            ct_block<Cfg, false>(LX, LY, LZ, WX, WY, WZ, s0 + (b + bb) * kR, o, wx, wy, wz, acc, acc64, sacc);
          ct_block<Cfg, true>(LX, LY, LZ, WX, WY, WZ, s0 + (b + kFB - 1) * kR, o, wx, wy, wz, acc, acc64, sacc);
        }
      }
    }
    __syncwarp();
    if (lane == 0) sr_mbar_arrive(&empty_bar[st]);
    if (++st == kNS) { st = 0; ph ^= 1; }
  }
  if (Cfg::FLUSH != 0 || Cfg::DIAG == 1) {   // what the partial sums still hold after the last flush
#pragma unroll
    for (int j = 0; j < kR; ++j) {
      if (Cfg::FLUSH == 2) sacc[o + j] += (double)acc[j];
      else acc64[j] += (double)acc[j];
    }
  }
  __syncthreads();

  // cross-warp reduction, one store per lag: through the (now idle) stage buffers, or straight from the per-warp
  // shared-memory accumulators
  double* red = Cfg::FLUSH == 2 ? sacc_all : reinterpret_cast<double*>(smem_raw);
  if (Cfg::FLUSH != 2) {
#pragma unroll
    for (int j = 0; j < kR; ++j) red[warp * kTL + o + j] = acc64[j];
    __syncthreads();
  }
  for (int i = tid; i < kTL; i += kNW * 32) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kNW; ++w) s += red[w * kTL + i];
    const int lag = d0 + i;
    if (lag >= 1 && lag <= L) S[rc * (long long)L + (lag - 1)] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// K2: AoS (chunk, frame, vector, xyz) float32 -> vector-major SoA rows (three planes) with zero padding.
// CTA = (chunk, tile of 64 frames, group of 16 vectors).  The 64 x 48-float tile is read with LDG.128 (a
// frame row of the reference layout is a multiple of 16 bytes when nR % 4 == 0; scalar loads otherwise),
// staged in shared memory with a 50-float row pitch, and written out as 48 plane rows of 64 floats with
// STG.128.  A warp stores 8 plane rows x 64 contiguous bytes, which makes the transposed shared-memory reads
// bank-conflict free ((200 q + p) mod 32 is a bijection for q < 4, p < 8).  HBM-bound: 12 B in + 12 B out.
// ------------------------------------------------------------------------------------------------
constexpr int kPF = 64, kPV = 16, kPP = 50, kPT = 256;

template <bool VEC4>
__global__ void __launch_bounds__(kPT)
pack_kernel(const float* __restrict__ vecs, int nC, int c0, long long nF, int nR, float* __restrict__ U, long long pitch,
            int do_rot, double qw, double qx, double qy, double qz) {
  // vecs points at chunk c0 of the input; blockIdx.z counts chunks from there; rows of U are indexed with the
  // absolute chunk number
  __shared__ float tile[kPF * kPP];
  const int cl = blockIdx.z, c = c0 + blockIdx.z;
  // blockIdx.x = vector group (fastest): CTAs that share frame rows run together, so the 32-byte sectors that
  // straddle two groups are fetched from DRAM once
  const long long f0 = (long long)blockIdx.y * kPF;
  const int r0 = blockIdx.x * kPV;
  const int nv = min(kPV, nR - r0);
  const int tid = threadIdx.x;

  // ---- load: 64 frames x (3 nv) floats, zero beyond nF (the padding region of the packed rows) ----
  if (VEC4) {
    const int n4 = (3 * nv) >> 2;                       // float4 per frame row of this group (nv % 4 == 0)
    for (int idx = tid; idx < kPF * 12; idx += kPT) {
      const int fl = idx / 12, k4 = idx - fl * 12;
      const long long f = f0 + fl;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < nF && k4 < n4)
        v = __ldg(reinterpret_cast<const float4*>(vecs + (((long long)cl * nF + f) * nR + r0) * 3) + k4);
      float* t = tile + fl * kPP + 4 * k4;
      t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    }
  } else {
    for (int idx = tid; idx < kPF * 48; idx += kPT) {
      const int fl = idx / 48, k = idx - fl * 48;
      const long long f = f0 + fl;
      float v = 0.f;
      if (f < nF && k < 3 * nv) v = __ldg(vecs + (((long long)cl * nF + f) * nR + r0) * 3 + k);
      tile[fl * kPP + k] = v;
    }
  }
  __syncthreads();
  if (do_rot) {
    // a = q_v x v + q_w v ; b = q_v x a ; out = b + b + v   (transforms3d_supplement.py:283-296), float64 like the reference
    for (int idx = tid; idx < kPF * kPV; idx += kPT) {
      const int fl = idx >> 4, v = idx & 15;
      float* t = tile + fl * kPP + 3 * v;
      const double vx = t[0], vy = t[1], vz = t[2];
      const double ax = qy * vz - qz * vy + qw * vx;
      const double ay = qz * vx - qx * vz + qw * vy;
      const double az = qx * vy - qy * vx + qw * vz;
      const double bx = qy * az - qz * ay, by = qz * ax - qx * az, bz = qx * ay - qy * ax;
      t[0] = (float)(bx + bx + vx); t[1] = (float)(by + by + vy); t[2] = (float)(bz + bz + vz);
    }
    __syncthreads();
  }
  // ---- store: plane row k = 3 v + comp, 16 float4 per row ----
  const int lane = tid & 31, warp = tid >> 5;
  const int prl = lane >> 2, q = lane & 3;
  for (int w = warp; w < 24; w += kPT / 32) {
    const int k = (w % 6) * 8 + prl;                    // tile column
    const int fq = (w / 6) * 4 + q;                     // float4 index inside the 64-frame row
    const int v = k / 3, comp = k - 3 * v;
    const long long f = f0 + 4 * fq;
    if (v < nv && f < pitch) {
      const float* t = tile + (4 * fq) * kPP + k;
      const float4 o = make_float4(t[0], t[kPP], t[2 * kPP], t[3 * kPP]);
      *reinterpret_cast<float4*>(U + (((long long)(r0 + v) * nC + c) * 3 + comp) * pitch + f) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Palmer finalize: chunk means of P2, then mean and std/(sqrt(nC)-1) over chunks (:226-228).
// ------------------------------------------------------------------------------------------------
// CTA = tile of 32 lags x 32 vectors: S is read along the lag axis (256-byte runs of doubles), the results
// are transposed through shared memory so that Ct / dCt rows (lag-major, nR contiguous) are written coalesced.
__global__ void __launch_bounds__(256)
ct_finalize_kernel(const double* __restrict__ S, int nC, int nF, int nR, int L, float* __restrict__ Ct,
                   float* __restrict__ dCt) {
  __shared__ float tm[32][33], ts[32][33];
  const int d0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int dl = threadIdx.x & 31, rl0 = threadIdx.x >> 5;
  const int di = d0 + dl;                                   // delta - 1
  const double inv = di < L ? 1.5 / (double)(nF - (di + 1)) : 0.0;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int rl = rl0 + 8 * rr, r = r0 + rl;
    float m32 = 0.f, s32 = 0.f;
    if (r < nR && di < L) {
      const double* row = S + (long long)r * nC * L + di;
      // one pass over S: shifted sums (shift = first chunk's value) keep sum((m - mean)^2) free of cancellation
      const double m0 = -0.5 + row[0] * inv;
      double s1 = 0.0, s2 = 0.0;
      for (int c = 1; c < nC; ++c) {
        const double d = (-0.5 + row[(long long)c * L] * inv) - m0;
        s1 += d; s2 = fma(d, d, s2);
      }
      const double dm = s1 / nC;
      const double mean = m0 + dm;
      const double var = fmax(0.0, s2 / nC - dm * dm);
      m32 = (float)mean;
      s32 = (float)(sqrt(var) / (sqrt((double)nC) - 1.0));
    }
    tm[dl][rl] = m32; ts[dl][rl] = s32;
  }
  __syncthreads();
  const int rl = threadIdx.x & 31;
#pragma unroll
  for (int dd = 0; dd < 4; ++dd) {
    const int dl2 = (threadIdx.x >> 5) + 8 * dd;
    const int d = d0 + dl2, r = r0 + rl;
    if (d < L && r < nR) {
      Ct[(long long)d * nR + r] = tm[dl2][rl];
      dCt[(long long)d * nR + r] = ts[dl2][rl];
    }
  }
}

// ================================================================================================
// C ABI
// ================================================================================================
// The product launches CtLong for long chunks and CtShort (smaller lag / frame tiles) for short ones.
// R = 23 lags per lane, 12 warps, 1 CTA/SM, 2 stages of 2484 frames; the three warps of every SM sub-partition
// re-aligned at a named barrier once per step (SYNC = 2); the four FMA-pipe instructions of a step issued in phases
// (ORDER = 1); FP64 lag sums in shared memory, flushed once per warp tile (FLUSH = 2, FB = MB: 207 terms per FP32 partial
// sum, S within 4e-8 of the float64 oracle).  Round 1 shipped R = 19 without the barrier, sums in registers flushed every
// 57 terms: 0.613 of peak on the tuning slice.  Barrier alone 0.658; R = 23 0.686-0.688; phased order 0.693; sums in
// shared memory (no spills) 0.697; flushing them once per tile instead of three times 0.7256.  Wider windows lose again
// (R = 25: 0.713, R = 27: 0.668: the loop body outgrows 32 KB of instruction cache).
// tools/tune_ct.py, profiles/r02l..r02y_tune_ct.json.
using CtLong = CtCfg<23, 9, 9, 12, 1, 2, 2, 0, 1, 2>;
using CtShort = CtCfg<15, 8, 1, 8, 2, 3>;      // R = 15, 8 warps, 2 CTAs/SM, 3 stages: 15 terms per FP32 partial sum (few chunk
                                               // means enter dCt when chunks are short, so keep its rounding at the 1e-7 level)
constexpr long long kShortFrames = 8192;
#ifdef SR_TUNING
constexpr int kMaxTF = 2736, kMaxTL = 32 * 45;
#else
constexpr int kMaxTF = 2736, kMaxTL = 32 * 23;   // padding must cover the largest tile of any configuration
#endif

template <class Cfg>
int launch_ct_lag(const float* U, long long pitch, long long nF, int nR, int nC, int c0, int nCsub, long long L, double* S,
                  cudaStream_t st) {
  static_assert(Cfg::TF <= kMaxTF && Cfg::TL <= kMaxTL, "zero padding of the packed rows must cover this tile");
  const long long nLT = (L + Cfg::TL) / Cfg::TL;   // tiles start at lag 0: ceil((L + 1) / TL)
  const long long items = (long long)nR * nCsub * nLT;
  SR_REQUIRE(items < (1LL << 31), "sr_ct_lag_sums: %lld work items exceed the grid limit", items);
  SR_CUDA(cudaFuncSetAttribute(ct_lag_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SmemBytes));
  ct_lag_kernel<Cfg><<<(unsigned)items, Cfg::NW * 32, Cfg::SmemBytes, st>>>(U, pitch, (int)nF, (int)L, (int)nLT, nC, c0,
                                                                            nCsub, S);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

}  // namespace

extern "C" long long sr_ct_row_pitch(long long nF) { return sr_round_up(nF + kMaxTF + kMaxTL + 64, 8); }

extern "C" size_t sr_ct_workspace_bytes(int nC, long long nF, int nR) {
  const long long pitch = sr_ct_row_pitch(nF);
  const long long L = nF / 2;
  size_t packed = (size_t)nR * nC * pitch * 12;
  size_t sums = (size_t)nR * nC * L * 8;
  return sr_round_up((long long)packed, 256) + sr_round_up((long long)sums, 256);
}

extern "C" int sr_pack_vectors_f32_chunks(const float* d_vecs_c0, int nC, int c0, int nCsub, long long nF, int nR,
                                          const double* h_q_rot, void* d_packed, long long pitch, void* stream) {
  SR_REQUIRE(d_vecs_c0 && d_packed, "sr_pack_vectors_f32: null pointer");
  SR_REQUIRE(nC > 0 && nF > 0 && nR > 0, "sr_pack_vectors_f32: empty shape (nC=%d nF=%lld nR=%d)", nC, nF, nR);
  SR_REQUIRE(c0 >= 0 && nCsub > 0 && c0 + nCsub <= nC, "sr_pack_vectors_f32: chunk range [%d, %d) outside [0, %d)", c0,
             c0 + nCsub, nC);
  SR_REQUIRE(pitch >= nF && pitch % 4 == 0, "sr_pack_vectors_f32: pitch %lld must be >= nF %lld and a multiple of 4", pitch, nF);
  SR_REQUIRE(((uintptr_t)d_packed & 15) == 0, "sr_pack_vectors_f32: packed stream must be 16-byte aligned");
  SR_REQUIRE(nCsub <= 65535, "sr_pack_vectors_f32: %d chunks exceed the grid limit", nCsub);
  double q[4] = {1, 0, 0, 0};
  int do_rot = 0;
  if (h_q_rot) {
    const double n = sqrt(h_q_rot[0] * h_q_rot[0] + h_q_rot[1] * h_q_rot[1] + h_q_rot[2] * h_q_rot[2] +
                          h_q_rot[3] * h_q_rot[3]);
    SR_REQUIRE(n > 0, "sr_pack_vectors_f32: zero rotation quaternion");
    for (int i = 0; i < 4; ++i) q[i] = h_q_rot[i] / n;   // vecnorm_NDarray(q), transforms3d_supplement.py:280
    do_rot = 1;
  }
  SR_REQUIRE((pitch + kPF - 1) / kPF <= 65535, "sr_pack_vectors_f32: %lld frames per chunk exceed the grid limit", nF);
  dim3 grid((unsigned)((nR + kPV - 1) / kPV), (unsigned)((pitch + kPF - 1) / kPF), (unsigned)nCsub);
  if (nR % 4 == 0 && ((uintptr_t)d_vecs_c0 & 15) == 0)
    pack_kernel<true><<<grid, kPT, 0, (cudaStream_t)stream>>>(d_vecs_c0, nC, c0, nF, nR, (float*)d_packed, pitch, do_rot,
                                                              q[0], q[1], q[2], q[3]);
  else
    pack_kernel<false><<<grid, kPT, 0, (cudaStream_t)stream>>>(d_vecs_c0, nC, c0, nF, nR, (float*)d_packed, pitch, do_rot,
                                                               q[0], q[1], q[2], q[3]);
  SR_CUDA(cudaGetLastError());
  return SR_OK;
}

extern "C" int sr_pack_vectors_f32(const float* d_vecs, int nC, long long nF, int nR, const double* h_q_rot,
                                   void* d_packed, long long pitch, void* stream) {
  return sr_pack_vectors_f32_chunks(d_vecs, nC, 0, nC, nF, nR, h_q_rot, d_packed, pitch, stream);
}

static int ct_lag_sums_check(const void* d_packed, long long pitch, int nC, int c0, int nCsub, long long nF, int nR,
                             long long L, double* d_S) {
  SR_REQUIRE(d_packed && d_S, "sr_ct_lag_sums: null pointer");
  SR_REQUIRE(nC > 0 && nR > 0 && nF >= 2, "sr_ct_lag_sums: bad shape (nC=%d nF=%lld nR=%d)", nC, nF, nR);
  SR_REQUIRE(L >= 1 && L <= nF - 1, "sr_ct_lag_sums: L=%lld outside [1, nF-1]", L);
  SR_REQUIRE(nF < (1LL << 30), "sr_ct_lag_sums: nF too large for 32-bit tile indices");
  SR_REQUIRE(pitch >= nF + kMaxTF + kMaxTL, "sr_ct_lag_sums: pitch %lld lacks %d frames of zero padding", pitch,
             kMaxTF + kMaxTL);
  SR_REQUIRE((long long)nR * nC < (1LL << 31), "sr_ct_lag_sums: too many (vector, chunk) rows");
  SR_REQUIRE(pitch % 4 == 0 && ((uintptr_t)d_packed & 15) == 0, "sr_ct_lag_sums: packed stream must be 16-byte aligned with pitch %% 4 == 0");
  SR_REQUIRE(c0 >= 0 && nCsub > 0 && c0 + nCsub <= nC, "sr_ct_lag_sums: chunk range [%d, %d) outside [0, %d)", c0, c0 + nCsub, nC);
  return SR_OK;
}

static int ct_lag_sums_impl(const void* d_packed, long long pitch, int nC, int c0, int nCsub, long long nF, int nR,
                            long long L, double* d_S, void* stream) {
  const int rc = ct_lag_sums_check(d_packed, pitch, nC, c0, nCsub, nF, nR, L, d_S);
  if (rc) return rc;
  const float* U = (const float*)d_packed;
  cudaStream_t st = (cudaStream_t)stream;
  if (nF < kShortFrames) return launch_ct_lag<CtShort>(U, pitch, nF, nR, nC, c0, nCsub, L, d_S, st);
  return launch_ct_lag<CtLong>(U, pitch, nF, nR, nC, c0, nCsub, L, d_S, st);
}

#ifdef SR_TUNING
// Experimental configurations (R, MB, FB, NW, CTAs/SM, stages, FLUSH, DIAG), timed by tools/tune_ct.py.
#include "ct_tuning_variants.inc"
#endif

extern "C" int sr_ct_lag_sums(const void* d_packed, long long pitch, int nC, long long nF, int nR, long long L,
                              double* d_S, void* stream) {
  return ct_lag_sums_impl(d_packed, pitch, nC, 0, nC, nF, nR, L, d_S, stream);
}

extern "C" int sr_ct_lag_sums_chunks(const void* d_packed, long long pitch, int nC, int c0, int nCsub, long long nF, int nR,
                                     long long L, double* d_S, void* stream) {
  return ct_lag_sums_impl(d_packed, pitch, nC, c0, nCsub, nF, nR, L, d_S, stream);
}

extern "C" int sr_ct_palmer_finalize(const double* d_S, int nC, long long nF, int nR, long long L, float* d_Ct,
                                     float* d_dCt, void* stream) {
  SR_REQUIRE(d_S && d_Ct && d_dCt, "sr_ct_palmer_finalize: null pointer");
  SR_REQUIRE(nC > 0 && nR > 0 && L >= 1 && L < nF, "sr_ct_palmer_finalize: bad shape");
  dim3 grid((unsigned)((L + 31) / 32), (unsigned)((nR + 31) / 32));
  ct_finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_S, nC, (int)nF, nR, (int)L, d_Ct, d_dCt);
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
  double* S = (double*)(ws + sr_round_up((long long)((size_t)nR * nC * pitch * 12), 256));
  int rc = sr_pack_vectors_f32(d_vecs, nC, nF, nR, nullptr, packed, pitch, stream);
  if (rc) return rc;
  rc = sr_ct_lag_sums(packed, pitch, nC, nF, nR, L, S, stream);
  if (rc) return rc;
  return sr_ct_palmer_finalize(S, nC, nF, nR, L, d_Ct, d_dCt, stream);
}

// Host-buffer entry point.  The caller's array is ordinary pageable memory, so it is staged through two pinned
// buffers on a copy stream; chunk 0 is uploaded first and its K2 / K1 run while the remaining chunks are still on
// their way (the same two-stage schedule as the Python pipeline).  Device buffers, pinned staging buffers, streams
// and events are kept in a grow-only cache between calls (cudaMalloc / cudaFree of gigabytes cost tens of
// milliseconds each); sr_release_host_cache() returns them.
namespace {
struct HostCallCache {
  std::mutex mu;
  int device = -1;
  char* d_buf = nullptr; size_t d_cap = 0;
  char* stage[2] = {nullptr, nullptr}; size_t stage_cap = 0;
  cudaStream_t copy = nullptr, comp = nullptr;
  cudaEvent_t freed[2] = {nullptr, nullptr}, ready[2] = {nullptr, nullptr};
  void release() {
    if (d_buf) cudaFree(d_buf);
    for (int i = 0; i < 2; ++i) {
      if (stage[i]) cudaFreeHost(stage[i]);
      if (freed[i]) cudaEventDestroy(freed[i]);
      if (ready[i]) cudaEventDestroy(ready[i]);
      stage[i] = nullptr; freed[i] = ready[i] = nullptr;
    }
    if (copy) cudaStreamDestroy(copy);
    if (comp) cudaStreamDestroy(comp);
    d_buf = nullptr; d_cap = stage_cap = 0; copy = comp = nullptr; device = -1;
  }
};
// one cache per device: host threads that each drive their own GPU (spinrelax_b200/multigpu.py) do not serialise on
// a shared lock, and switching devices does not throw the buffers away
constexpr int kMaxDevices = 16;
HostCallCache g_host_cache[kMaxDevices];
}  // namespace

extern "C" void sr_release_host_cache(void) {
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < kMaxDevices; ++d) {
    std::lock_guard<std::mutex> lock(g_host_cache[d].mu);
    if (g_host_cache[d].device >= 0) {
      cudaSetDevice(g_host_cache[d].device);
      g_host_cache[d].release();
    }
  }
  cudaSetDevice(cur);
}

extern "C" int sr_ct_palmer_host(const float* h_vecs, int nC, long long nF, int nR, float* h_Ct, float* h_dCt) {
  SR_REQUIRE(h_vecs && h_Ct && h_dCt, "sr_ct_palmer_host: null pointer");
  SR_REQUIRE(nC > 0 && nR > 0 && nF >= 2, "sr_ct_palmer_host: bad shape");
  const long long L = nF / 2;
  const size_t chunk_bytes = (size_t)nF * nR * 3 * sizeof(float);
  const size_t in_bytes = (size_t)sr_round_up((long long)((size_t)nC * chunk_bytes), 256);
  const size_t out_bytes = (size_t)sr_round_up((long long)((size_t)L * nR * sizeof(float)), 256);
  const size_t ws_bytes = sr_ct_workspace_bytes(nC, nF, nR);
  const size_t need = in_bytes + 2 * out_bytes + ws_bytes;
  const size_t stage_bytes = (size_t)32 << 20;
  const long long pitch = sr_ct_row_pitch(nF);
  int rc = SR_OK, dev = 0;
  cudaError_t e = cudaSuccess;
  SR_CUDA(cudaGetDevice(&dev));
  SR_REQUIRE(dev >= 0 && dev < kMaxDevices, "sr_ct_palmer_host: device index %d outside [0, %d)", dev, kMaxDevices);
  HostCallCache& hc = g_host_cache[dev];
  std::lock_guard<std::mutex> lock(hc.mu);
  auto fail = [&](const char* what) {
    sr_set_error("sr_ct_palmer_host: %s failed: %s", what, cudaGetErrorString(e));
    rc = SR_ERR_CUDA;
  };
  hc.device = dev;
  if (hc.d_cap < need) {
    if (hc.d_buf) cudaFree(hc.d_buf);
    hc.d_buf = nullptr; hc.d_cap = 0;
    if ((e = cudaMalloc(&hc.d_buf, need)) != cudaSuccess) { fail("cudaMalloc"); return rc; }
    hc.d_cap = need;
  }
  if (!hc.stage[0]) {
    if ((e = cudaMallocHost(&hc.stage[0], stage_bytes)) != cudaSuccess || (e = cudaMallocHost(&hc.stage[1], stage_bytes)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&hc.copy, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&hc.comp, cudaStreamNonBlocking)) != cudaSuccess) { fail("allocation"); hc.release(); return rc; }
    hc.stage_cap = stage_bytes;
    for (int i = 0; i < 2; ++i)
      if ((e = cudaEventCreateWithFlags(&hc.freed[i], cudaEventDisableTiming)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&hc.ready[i], cudaEventDisableTiming)) != cudaSuccess) { fail("cudaEventCreate"); hc.release(); return rc; }
  }
  float* d_in = (float*)hc.d_buf;
  float* d_Ct = (float*)(hc.d_buf + in_bytes);
  float* d_dCt = (float*)(hc.d_buf + in_bytes + out_bytes);
  char* d_ws = hc.d_buf + in_bytes + 2 * out_bytes;
  void* packed = d_ws;
  double* S = (double*)(d_ws + sr_round_up((long long)((size_t)nR * nC * pitch * 12), 256));
  // upload groups: chunk 0, then chunks 1 .. nC-1
  const int nGroups = nC > 1 ? 2 : 1;
  size_t done = 0;
  int piece = 0;
  for (int g = 0; g < nGroups && !rc; ++g) {
    const int c0 = g == 0 ? 0 : 1, n = g == 0 ? 1 : nC - 1;
    const size_t end = (size_t)(c0 + n) * chunk_bytes;
    while (done < end && !rc) {
      const size_t len = std::min(hc.stage_cap, end - done);
      const int b = piece & 1;
      if (piece >= 2 && (e = cudaEventSynchronize(hc.freed[b])) != cudaSuccess) { fail("cudaEventSynchronize"); break; }
      memcpy(hc.stage[b], (const char*)h_vecs + done, len);
      if ((e = cudaMemcpyAsync((char*)d_in + done, hc.stage[b], len, cudaMemcpyHostToDevice, hc.copy)) != cudaSuccess ||
          (e = cudaEventRecord(hc.freed[b], hc.copy)) != cudaSuccess) { fail("H2D"); break; }
      done += len; ++piece;
    }
    if (rc) break;
    if ((e = cudaEventRecord(hc.ready[g], hc.copy)) != cudaSuccess ||
        (e = cudaStreamWaitEvent(hc.comp, hc.ready[g], 0)) != cudaSuccess) { fail("cudaStreamWaitEvent"); break; }
    rc = sr_pack_vectors_f32_chunks(d_in + (size_t)c0 * (chunk_bytes / sizeof(float)), nC, c0, n, nF, nR, nullptr, packed,
                                    pitch, hc.comp);
    if (!rc) rc = sr_ct_lag_sums_chunks(packed, pitch, nC, c0, n, nF, nR, L, S, hc.comp);
  }
  if (!rc) rc = sr_ct_palmer_finalize(S, nC, nF, nR, L, d_Ct, d_dCt, hc.comp);
  const size_t out_exact = (size_t)L * nR * sizeof(float);
  if (!rc && ((e = cudaMemcpyAsync(h_Ct, d_Ct, out_exact, cudaMemcpyDeviceToHost, hc.comp)) != cudaSuccess ||
              (e = cudaMemcpyAsync(h_dCt, d_dCt, out_exact, cudaMemcpyDeviceToHost, hc.comp)) != cudaSuccess))
    fail("D2H");
  cudaStreamSynchronize(hc.copy);
  if ((e = cudaStreamSynchronize(hc.comp)) != cudaSuccess && !rc) fail("cudaStreamSynchronize");
  return rc;
}
