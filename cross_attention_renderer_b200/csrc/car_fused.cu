// Fused epipolar gather + per-sample encoder for one ray per CTA pair (P == 64).
//
// Replaces, for the tensor-core precisions, the chain
//   k_gather -> GEMM query_encode_latent (+ReLU) -> GEMM query_encode_latent_2
//            -> GEMM latent_value / GEMM key_map (+ReLU)
// (reference models.py:278,317,333-344,487-491) by ONE kernel in which the 576/579-wide
// per-sample activations never leave the SM:
//
//   CTA pair (cluster of 2, tcgen05 cta_group::2, UMMA M = 128): CTA `rank` owns the 64 samples of
//   the ray's epipolar line in context view `rank`.
//   for view v in {0,1}:                         (features of view v at every sample, A.7)
//     GEMM1   acc1[128 x 576] = X_v[128 x 592] · W1^T        X_v produced IN the kernel:
//             gather warps bilinearly sample the NHWC maps (own line: border taps, other view:
//             re-projected zero-padded taps), convert to bf16 hi(+lo) and write 32-channel
//             K-slices straight into the swizzled smem ring that feeds the MMA
//     epilogue acc1 -> +bias, ReLU -> bf16 hi(+lo) -> smem ring (A operand of the next GEMM)
//     GEMM3   acc3[128 x 416] += H1_v[128 x 576] · F_v^T     F = [latent_value; key_map]∘enc2 folded
//   epilogue acc3 -> V (fp32, 288) and relu(key_map) (bf16 hi/lo, 128) per sample -> HBM
//
// TMEM (per CTA, 64 rows in the "2x2" layout => N/2 columns): acc1 288 + acc3 208 = 496 <= 512.
// SMEM: X ring + H ring (A operands), B ring (W1 / F slices via TMA, each CTA loads its half of
// every N chunk), tap table.  All operand tiles are K-major and swizzled; KS K-columns per stage:
//   KS = 32 (64-byte swizzle rows)  hi+lo mode (CAR_PREC_FP32_3XBF16): 8 KB A stages, 36 KB B stages
//   KS = 64 (128-byte swizzle rows) bf16 mode: half as many stages per ray.  The MMA-issuing thread pays
//           ~400 cycles per stage whatever the stage holds (barrier waits, fence, election, descriptor
//           set-up, commits: scripts/fused_stalls.py; a second issuing warp did not help, profiles/README.md),
//           against 200-290 cycles of tensor time per 32-wide stage in bf16 - so the stages are made wider.
//           The H ring then pairs the two lane halves of one accumulator block in one stage, and F is read
//           through a column-permuted copy (car_weights::kv_fold64) whose K order is the epilogue's.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "car_common.cuh"
#include "car_umma.cuh"

namespace car {
extern unsigned long long *g_fused_stats;
int make_tmap_bf16(CUtensorMap *tm, const uint16_t *base, int rows, int K, int ld, int box_rows, int box_k);

namespace {
using namespace ptx;

constexpr int ROWS = 64;               // sample rows per CTA
constexpr int NXMAX = 8;               // X ring depth is Cfg::NX
constexpr int NHMAX = 6;               // H ring depth is Cfg::NH (the ring doubles as the 24 KB / 32 KB acc3 staging area)
constexpr int N1 = 576, N1CH = 192, N1C = 3;      // GEMM1: 3 MMA chunks of 192
constexpr int N3 = 416, N3CH = 208, N3C = 2;      // GEMM3: 2 MMA chunks of 208
constexpr int ACC1_COL = 0, ACC3_COL = 288;
constexpr int PRODUCER_WARPS = 8;
constexpr int BASE_WARPS = 2 + 4 + PRODUCER_WARPS;       // TMA, MMA, 4 epilogue, 8 gather = 448 threads
constexpr int A3_STAGE = 2048;                           // bytes of private staging per dedicated acc3-drain warp
constexpr int MAXB = 6;

struct TapEntry { int off[4]; float w[4]; };              // 32 bytes

// Optional stall accounting (bench/diagnostics): cycles spent in each barrier wait, summed over
// pair 0's leader roles.  Index map: 0 mma:a1_empty 1 mma:x_full 2 mma:b_full(gemm1) 3 mma:a3_empty
// 4 mma:h_full 5 mma:b_full(gemm3) 6 mma:total | 8 epi:a1_full 9 epi:h_empty 10 epi:a3_full 11 epi:total
// | 12 prod:x_empty 13 prod:total 14 prod:barrier | 16 tma:b_empty 17 tma:total
__device__ __forceinline__ void timed_wait(uint64_t *bar, uint32_t parity, unsigned long long *st, int idx) {
  if (st) {
    long long t0 = clock64();
    mbar_wait(bar, parity);
    st[idx] += (unsigned long long)(clock64() - t0);
  } else {
    mbar_wait(bar, parity);
  }
}

struct FusedParams {
  const void *feat[3];
  int H, W, R, g0, g1;
  int P, hpr;                          // samples per line (multiple of 64) and 64-row groups per line (P / 64)
  const float *geom;                   // (rows,32) for rays [g0,g1)
  const float *bias1;                  // [576]
  const float *biasf;                  // [416]
  float *value;                        // (rows,288) fp32
  uint16_t *kh_hi, *kh_lo;             // (rows,128) bf16: relu(key_map)
  int nb;                              // B ring depth
  int cl;                              // cluster size: 2 (one pair) or 4 (two pairs sharing the weight loads)
  unsigned long long *stats;           // optional [32] stall counters (see timed_wait), else null
};

template <int SPLIT, int KS_> struct Cfg {
  static constexpr int KS = KS_;                                   // K per stage (bf16 elements)
  static constexpr int SW = KS * 2;                                // swizzle row bytes: 64 or 128
  static constexpr int KSTEPS = KS / 16;                           // UMMA K = 16 per instruction
  static constexpr int K1_STAGES = (CAR_K_ENC + KS - 1) / KS;      // 19 / 10; the last stage holds 16 valid columns
  static constexpr int K3_STAGES = N1 / KS;                        // 18 / 9
  static constexpr int OPS = SPLIT == 3 ? 2 : 1;                  // hi (+lo) copies
  static constexpr int A_HALF = ROWS * SW;                         // 4096 / 8192 B (hi part of an A stage)
  static constexpr int A_STAGE = A_HALF * OPS;
  static constexpr int W1_CHUNK = (N1CH / 2) * SW;                 // 6144 / 12288 B
  static constexpr int F_CHUNK = (N3CH / 2) * SW;                  // 6656 / 13312 B
  static constexpr int B_HALF = N1C * W1_CHUNK;                    // 18432 / 36864 B (>= N3C*F_CHUNK)
  static constexpr int B_STAGE = B_HALF * OPS;
  static constexpr int NH = KS == 64 ? 3 : (SPLIT == 3 ? 4 : 6);   // NH * A_STAGE >= 24 KB (32 KB with lo copies): acc3 staging
  static constexpr int NX = KS == 64 ? 6 : (SPLIT == 3 ? 6 : 8);
  static constexpr int H_ARRIVALS = KS == 64 ? 8 : 4;              // epilogue warps per H stage (x 2 CTAs)
  // Experiment kept as a switch: four more warps (14..17) drain acc3, so that the acc1 drains of the next ray - which
  // gate GEMM3 - never queue behind the 14 k-cycle acc3 drain.  Measured in bf16 mode: 359 vs 305 ms per step.
  // Registers are allocated in units of 32 per thread, so 576 threads get 96 instead of 128 and the gather / acc1
  // loops spill (1 - 1.4 KB per thread); a build capped at 112 registers does not launch with 576 threads.
  static constexpr bool A3W = false;
  static constexpr int THREADS = (BASE_WARPS + (A3W ? 4 : 0)) * 32;
  static_assert(SW == 64 || SW == 128, "KS must be 32 or 64");
  static_assert(NH * A_STAGE >= (SPLIT == 3 ? 32768 : 24576), "H ring too small to stage acc3");
  static_assert(W1_CHUNK % 1024 == 0 || SW == 64, "128-byte-swizzled chunks must start on 1 KB boundaries");
};

__device__ __forceinline__ float downgrade(float x) {
  if (x > 2147483646.0f || x < -2147483648.0f || !isfinite(x)) return -100.0f;
  return x;
}
// Same arithmetic as car_gather.cu::make_taps (PyTorch CUDA grid_sample formulas).
__device__ __forceinline__ TapEntry make_taps(float gx, float gy, int w, int h, bool border) {
  float ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)w), 1.f), 2.f);
  float iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)h), 1.f), 2.f);
  if (border) {
    ix = fminf((float)(w - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(h - 1), fmaxf(iy, 0.f));
  }
  ix = downgrade(ix);
  iy = downgrade(iy);
  float fx0 = floorf(ix), fy0 = floorf(iy);
  int x0 = (int)fx0, y0 = (int)fy0;
  float wx1 = __fsub_rn(ix, fx0), wx0 = __fsub_rn(__fadd_rn(fx0, 1.f), ix);
  float wy1 = __fsub_rn(iy, fy0), wy0 = __fsub_rn(__fadd_rn(fy0, 1.f), iy);
  TapEntry t;
  t.w[0] = wx0 * wy0; t.w[1] = wx1 * wy0; t.w[2] = wx0 * wy1; t.w[3] = wx1 * wy1;
  bool xin0 = x0 >= 0 && x0 < w, xin1 = x0 + 1 >= 0 && x0 + 1 < w;
  bool yin0 = y0 >= 0 && y0 < h, yin1 = y0 + 1 >= 0 && y0 + 1 < h;
  t.off[0] = (xin0 && yin0) ? y0 * w + x0 : -1;
  t.off[1] = (xin1 && yin0) ? y0 * w + x0 + 1 : -1;
  t.off[2] = (xin0 && yin1) ? (y0 + 1) * w + x0 : -1;
  t.off[3] = (xin1 && yin1) ? (y0 + 1) * w + x0 + 1 : -1;
  return t;
}

// (a, b) -> packed bf16x2 hi = rn(a, b) and, when WITH_LO, lo = rn(a - hi.a, b - hi.b)
template <bool WITH_LO = true>
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<uint32_t *>(&h);
  if (WITH_LO) {
    const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
    __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
    lo = *reinterpret_cast<uint32_t *>(&l);
  } else {
    lo = 0;
  }
}

// registers are allocated per warp in units of 1024 (32 per thread): 128 per thread at 448 threads, 96 at 576
// (a kernel compiled with __maxnreg__(112) does not launch with 576 threads)
template <int SPLIT, typename FT, int KS>
__global__ void __launch_bounds__(Cfg<SPLIT, KS>::THREADS, 1)
k_fused_encode(const __grid_constant__ CUtensorMap tm_w1_hi, const __grid_constant__ CUtensorMap tm_w1_lo,
               const __grid_constant__ CUtensorMap tm_f_hi, const __grid_constant__ CUtensorMap tm_f_lo,
               FusedParams p) {
  using C = Cfg<SPLIT, KS>;
  constexpr int SW = C::SW, K1_STAGES = C::K1_STAGES, K3_STAGES = C::K3_STAGES, THREADS = C::THREADS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *xs = smem;                                       // NX x A_STAGE
  constexpr int NH = C::NH, NX = C::NX;
  uint8_t *hs = xs + NX * C::A_STAGE;                       // NH x A_STAGE
  uint8_t *bs = hs + NH * C::A_STAGE;                       // nb x B_STAGE
  TapEntry *taps = reinterpret_cast<TapEntry *>(bs + (size_t)p.nb * C::B_STAGE);   // [2 buffers][64][3][2]
  float *tanhs = reinterpret_cast<float *>(taps + 2 * ROWS * 3 * 2);               // [2][64][8]
  float *sbias1 = tanhs + 2 * ROWS * 8;                                                // [576]
  float *sbiasf = sbias1 + N1;                                                     // [416]
  uint64_t *bars = reinterpret_cast<uint64_t *>(sbiasf + N3);
  uint64_t *x_full = bars, *x_empty = x_full + NXMAX;
  uint64_t *h_full = x_empty + NXMAX, *h_empty = h_full + NHMAX;
  uint64_t *b_full = h_empty + NHMAX, *b_empty = b_full + MAXB;
  uint64_t *a1_full = b_empty + MAXB, *a1_empty = a1_full + 1;
  uint64_t *a3_full = a1_empty + 1, *a3_empty = a3_full + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a3_empty + 1);
  uint8_t *a3stage = reinterpret_cast<uint8_t *>(tmem_slot + 4);            // [4 warps][A3_STAGE] (A3W only), 16-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // cluster = p.cl CTAs (2 or 4) = p.cl/2 CTA pairs.  A pair (consecutive cluster ranks) shares one
  // ray and the cta_group::2 MMAs; the pairs of a cluster run in lock-step on DIFFERENT rays and share
  // every weight K-slice: pair 0's CTAs load their halves once and TMA-multicast them to the same
  // rank-in-pair of the other pair.
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1;                         // role inside the pair (= context of the rows)
  const uint32_t leader_crank = crank & ~1u;               // cluster rank of this pair's MMA-issuing CTA
  const bool leader = rank == 0;
  const bool loader = (crank >> 1) == 0;                   // this CTA issues the weight TMA loads
  const uint16_t pair_mask = (uint16_t)(0x3u << leader_crank);
  const uint16_t all_mask = (uint16_t)((1u << p.cl) - 1u);
  const uint16_t mc_mask = p.cl == 4 ? (uint16_t)((1u << crank) | (1u << (crank + 2))) : (uint16_t)(1u << crank);
  const int pair = (int)(blockIdx.x / p.cl) * (p.cl / 2) + (int)(crank >> 1), npairs = gridDim.x >> 1;
  // work item = one 64-sample group of one ray: item -> (ray = item / hpr, group = item % hpr); CTA `rank`
  // owns that group on context `rank`.  (Names below still say `ray` for the item index.)
  const int nrays = (p.g1 - p.g0) * p.hpr;
  // every pair runs the same number of iterations (the pairs of a cluster share the weight ring);
  // a pair that runs out of items repeats the last one with its stores suppressed
  const int niter = (nrays + npairs - 1) / npairs;
  auto item_of = [&](int itn) { const int i = pair + itn * npairs; return i < nrays ? i : nrays - 1; };
  auto valid_of = [&](int itn) { return pair + itn * npairs < nrays; };
  auto row_base = [&](int item) -> size_t {
    const int rl = item / p.hpr, gq = item - rl * p.hpr;
    return ((size_t)rl * 2 + rank) * p.P + (size_t)gq * ROWS;
  };
  unsigned long long lst[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long *st = (p.stats && pair == 0 && leader) ? lst : nullptr;
  const long long t_begin = clock64();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_w1_hi);
    prefetch_tmap(&tm_f_hi);
    for (int s = 0; s < NX; ++s) { mbar_init(&x_full[s], 4);   /* 2 warps of the owning group x 2 CTAs */ mbar_init(&x_empty[s], 1); }
    for (int s = 0; s < NH; ++s) { mbar_init(&h_full[s], C::H_ARRIVALS); mbar_init(&h_empty[s], 1); }
    for (int s = 0; s < p.nb; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], p.cl / 2); }
    mbar_init(a1_full, 1); mbar_init(a1_empty, 8);
    mbar_init(a3_full, 1); mbar_init(a3_empty, 8);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < N1; i += THREADS) sbias1[i] = p.bias1[i];
  for (int i = threadIdx.x; i < N3; i += THREADS) sbiasf[i] = p.biasf[i];
  if (warp == 1) tmem_alloc<2>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- acc3 -> V (fp32) and relu(key pre-activation) (bf16 hi/lo), one warp's 32 rows ----
  // A direct store of the TMEM image (lane = row, 1152-byte row pitch) touches 32 different lines per instruction -
  // 6.6 k L1 wavefronts per ray, more than the whole bilinear gather.  The warp therefore transposes its 32 rows in
  // blocks of SRB / 4 columns through a private staging area of 32 x SRB bytes (swizzled, conflict-free both ways)
  // and writes them out row-contiguously: with SRB = 128, 8 lanes per 128-byte row piece = 4 full lines per store
  // instruction.  (A first version staged 96-column rounds with two 128-thread barriers each; a TMA-store version
  // queued behind the weight loads in the SM's TMA unit: profiles/README.md.)
  // Accumulator column n of the pair tile lives at chunk e = n / 208, lane half (n % 208) / 104; warp `sub` of a
  // TMEM-lane quarter set holds rows [32 (sub & 1), +32) of the CTA and the columns of lane half sub >> 1.
  auto drain_acc3 = [&](auto SRB_, int ray, bool valid, uint32_t rq, uint8_t *wbuf, unsigned long long *st_) {
    constexpr int SRB = decltype(SRB_)::value;             // staging row bytes: 128 or 64
    constexpr int MAXC = SRB / 4;                          // fp32 columns per block
    const int sub = warp & 3, half = sub >> 1;
    const uint32_t tlane = tmem_base + ((uint32_t)(sub * 32) << 16);
    timed_wait(a3_full, rq & 1, st_, 2);
    tc_fence_after();
    const size_t grow0 = row_base(ray) + (size_t)((sub & 1) * 32);       // first global row of this warp
    const uint32_t tc0_ = tlane + ACC3_COL, tc1_ = tlane + ACC3_COL + (uint32_t)(N3CH / 2);
    auto wsw = [&](int r_, int c_) {
      return (uint32_t)(r_ * SRB + ((c_ ^ (SRB == 128 ? (r_ & 7) : ((r_ >> 1) & 3))) * 16));
    };
    // CNT (32 / 16 / 8) V columns starting at accumulator column n0 (this lane's TMEM columns [tcol, tcol + CNT))
    auto drain_v = [&](uint32_t tcol, int n0, auto CNT_) {
      constexpr int CNT = decltype(CNT_)::value;
      uint32_t r[CNT];
      if constexpr (CNT == 32) tmem_ld32(tcol, r);
      else if constexpr (CNT == 16) tmem_ld16(tcol, r);
      else tmem_ld8(tcol, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < CNT; i += 4) {
        const float4 bb = *reinterpret_cast<const float4 *>(sbiasf + n0 + i);
        *reinterpret_cast<float4 *>(wbuf + wsw(lane, i >> 2)) =
            make_float4(__uint_as_float(r[i]) + bb.x, __uint_as_float(r[i + 1]) + bb.y,
                        __uint_as_float(r[i + 2]) + bb.z, __uint_as_float(r[i + 3]) + bb.w);
      }
      __syncwarp();
      constexpr int CPR = CNT / 4, RPI = 32 / CPR;         // 16-byte chunks per row, rows per store instruction
      if (valid) {
        const int cj = lane % CPR, r0 = lane / CPR;
#pragma unroll
        for (int i = 0; i < 32 / RPI; ++i) {
          const int rr = r0 + RPI * i;
          *reinterpret_cast<float4 *>(p.value + (grow0 + rr) * CAR_C_LAT + n0 + cj * 4) =
              *reinterpret_cast<const float4 *>(wbuf + wsw(rr, cj));
        }
      }
      __syncwarp();
    };
    // CNT relu(key) columns starting at key column k0 (accumulator column 288 + k0): bf16 hi in the staging row's
    // chunks [0, CNT / 8), lo in chunks [LO, LO + CNT / 8)
    auto drain_k = [&](uint32_t tcol, int k0, auto CNT_) {
      constexpr int CNT = decltype(CNT_)::value;
      constexpr int SPR = SRB / 16, LO = SPR / 2;          // 16-byte slots per staging row; first lo slot
      uint32_t r[CNT];
      if constexpr (CNT == 32) tmem_ld32(tcol, r);
      else if constexpr (CNT == 16) tmem_ld16(tcol, r);
      else tmem_ld8(tcol, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < CNT; i += 8) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          split2<SPLIT == 3>(fmaxf(__uint_as_float(r[i + 2 * j]) + sbiasf[CAR_C_LAT + k0 + i + 2 * j], 0.f),
                             fmaxf(__uint_as_float(r[i + 2 * j + 1]) + sbiasf[CAR_C_LAT + k0 + i + 2 * j + 1], 0.f), hi[j], lo[j]);
        *reinterpret_cast<uint4 *>(wbuf + wsw(lane, i >> 3)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (SPLIT == 3) *reinterpret_cast<uint4 *>(wbuf + wsw(lane, LO + (i >> 3))) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      __syncwarp();
      constexpr int CPR = CNT / 8;                         // 16-byte chunks (8 bf16) per row and copy
      if (valid) {
        // lanes: SPR slots per row (hi then lo), 32 / SPR rows per instruction; slots beyond CPR idle
        const int slot = lane % SPR, r0 = lane / SPR, cj = slot % LO;
        const bool is_lo = slot >= LO;
        if (cj < CPR && (SPLIT == 3 || !is_lo)) {
          uint16_t *dstb = is_lo ? p.kh_lo : p.kh_hi;
#pragma unroll
          for (int i = 0; i < SPR; ++i) {
            const int rr = r0 + (32 / SPR) * i;
            *reinterpret_cast<uint4 *>(dstb + (grow0 + rr) * 128 + k0 + cj * 8) =
                *reinterpret_cast<const uint4 *>(wbuf + wsw(rr, slot));
          }
        }
      }
      __syncwarp();
    };
    // LEN columns as blocks of at most MAXC
    auto span = [&](auto &&fn, uint32_t tcol, int n0, auto LEN_) {
      constexpr int LEN = decltype(LEN_)::value, NFULL = LEN / MAXC, REM = LEN % MAXC;
#pragma unroll
      for (int i = 0; i < NFULL; ++i) fn(tcol + (uint32_t)(i * MAXC), n0 + i * MAXC, std::integral_constant<int, MAXC>{});
      if constexpr (REM >= 16) fn(tcol + (uint32_t)(NFULL * MAXC), n0 + NFULL * MAXC, std::integral_constant<int, 16>{});
      if constexpr (REM % 16 == 8) fn(tcol + (uint32_t)(NFULL * MAXC + (REM >= 16 ? 16 : 0)), n0 + NFULL * MAXC + (REM >= 16 ? 16 : 0),
                                      std::integral_constant<int, 8>{});
    };
    // chunk 0: all V, n = half * 104 + col
    span(drain_v, tc0_, half * (N3CH / 2), std::integral_constant<int, 104>{});
    // chunk 1: lower lane half = V 208..287 (columns 0..79) then key 0..23; upper half = key 24..127
    if (half == 0) {
      span(drain_v, tc1_, 208, std::integral_constant<int, 80>{});
      span(drain_k, tc1_ + 80, 0, std::integral_constant<int, 24>{});
    } else {
      span(drain_k, tc1_, 24, std::integral_constant<int, 104>{});
    }
    // acc3 has been read completely: hand it back
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(a3_empty, leader_crank);
  };

  if (warp == 0) {
    // =========================== B-operand TMA producer ===========================
    if (lane == 0) {
      uint32_t bq = 0;
      for (int itn = 0; itn < niter; ++itn) {
        for (int v = 0; v < 2; ++v) {
          for (int kb = 0; kb < K1_STAGES; ++kb, ++bq) {                    // W1 K-slices
            const int s = bq % p.nb;
            timed_wait(&b_empty[s], ((bq / p.nb) & 1) ^ 1, st, 0);
            uint8_t *st = bs + (size_t)s * C::B_STAGE;
            if (leader) mbar_expect_tx(&b_full[s], (uint32_t)(N1C * C::W1_CHUNK * C::OPS * 2));
            for (int c = 0; c < N1C; ++c) {
              const int n0 = c * N1CH + (int)rank * (N1CH / 2);
              if (!loader) continue;
              tma_load_2d_pair_mc(st + c * C::W1_CHUNK, &tm_w1_hi, &b_full[s], kb * KS, n0, mc_mask);
              if (SPLIT == 3) tma_load_2d_pair_mc(st + C::B_HALF + c * C::W1_CHUNK, &tm_w1_lo, &b_full[s], kb * KS, n0, mc_mask);
            }
          }
          for (int q = 0; q < K3_STAGES; ++q, ++bq) {                       // F_v K-slices, epilogue order
            const int s = bq % p.nb;
            timed_wait(&b_empty[s], ((bq / p.nb) & 1) ^ 1, st, 0);
            uint8_t *st = bs + (size_t)s * C::B_STAGE;
            int k0;
            if (KS == 32) {                                                     // the epilogue's production order
              const int half = q & 1, cj = q >> 1, c = cj / 3, jj = cj - c * 3;
              k0 = v * N1 + c * N1CH + half * (N1CH / 2) + jj * KS;
            } else {
              k0 = v * N1 + q * KS;                                             // kv_fold64 is stored in that order
            }
            if (leader) mbar_expect_tx(&b_full[s], (uint32_t)(N3C * C::F_CHUNK * C::OPS * 2));
            for (int e = 0; e < N3C; ++e) {
              const int n0 = e * N3CH + (int)rank * (N3CH / 2);
              if (!loader) continue;
              tma_load_2d_pair_mc(st + e * C::F_CHUNK, &tm_f_hi, &b_full[s], k0, n0, mc_mask);
              if (SPLIT == 3) tma_load_2d_pair_mc(st + C::B_HALF + e * C::F_CHUNK, &tm_f_lo, &b_full[s], k0, n0, mc_mask);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA) ===========================
    // The whole warp stays converged; one elected lane issues.  Per stage the two operand
    // descriptors are built once and advanced by compile-time constants inside fully unrolled
    // loops, so the issue path is a handful of instructions per tcgen05.mma.
    // A-operand collector reuse: MMAs of one k-step that read the same A tile are issued back to back with
    // collector::a::fill / use / lastuse (measured: fp32 560 -> 549 ms, bf16 305.5 -> 301 ms per step)
    constexpr bool COLL = true;
    if (leader) {
      const uint32_t idesc1 = make_idesc_bf16(128, N1CH), idesc3 = make_idesc_bf16(128, N3CH);
      uint32_t av = 0, rq = 0;
      // ring positions as (stage, phase) counters: the ring depths are run-time values and an integer
      // division per stage is a visible part of the single issuing thread's fixed cost
      int sb = 0, sx = 0, sh = 0;
      uint32_t pb = 0, px = 0, ph = 0;
      const int nbr = p.nb;
      for (int itn = 0; itn < niter; ++itn, ++rq) {
        for (int v = 0; v < 2; ++v, ++av) {
          timed_wait(a1_empty, (av & 1) ^ 1, st, 0);                         // acc1 drained
          tc_fence_after();
          for (int kb = 0; kb < K1_STAGES; ++kb) {
            timed_wait(&x_full[sx], px, st, 1);
            timed_wait(&b_full[sb], pb, st, 2);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t da = make_desc<SW>(smem_u32(xs + (size_t)sx * C::A_STAGE));
              const uint64_t db = make_desc<SW>(smem_u32(bs + (size_t)sb * C::B_STAGE));
              const uint32_t d0 = tmem_base + ACC1_COL;
#pragma unroll
              for (int k = 0; k < C::KSTEPS; ++k) {
                if (k >= 1 && kb == K1_STAGES - 1) break;                   // last stage: 16 valid K columns
                const uint32_t acc = (kb | k) ? 1u : 0u;
                const uint64_t a = da + (uint64_t)((k * 32) >> 4);
                if (COLL) {
                  // same A tile back to back: Ah x {Wh_c, Wl_c} for every chunk, then Al x Wh_c
#pragma unroll
                  for (int c = 0; c < N1C; ++c) {
                    const uint64_t w = db + (uint64_t)((c * C::W1_CHUNK + k * 32) >> 4);
                    if (SPLIT == 3) {
                      if (c == 0) umma_f16_pair<1>(d0 + c * (N1CH / 2), a, w, idesc1, acc);
                      else umma_f16_pair<2>(d0 + c * (N1CH / 2), a, w, idesc1, acc);
                      if (c == N1C - 1) umma_f16_pair<3>(d0 + c * (N1CH / 2), a, w + (uint64_t)(C::B_HALF >> 4), idesc1, 1u);
                      else umma_f16_pair<2>(d0 + c * (N1CH / 2), a, w + (uint64_t)(C::B_HALF >> 4), idesc1, 1u);
                    } else {
                      if (c == 0) umma_f16_pair<1>(d0 + c * (N1CH / 2), a, w, idesc1, acc);
                      else if (c == N1C - 1) umma_f16_pair<3>(d0 + c * (N1CH / 2), a, w, idesc1, acc);
                      else umma_f16_pair<2>(d0 + c * (N1CH / 2), a, w, idesc1, acc);
                    }
                  }
                  if (SPLIT == 3) {
#pragma unroll
                    for (int c = 0; c < N1C; ++c) {
                      const uint64_t w = db + (uint64_t)((c * C::W1_CHUNK + k * 32) >> 4);
                      const uint64_t al = a + (uint64_t)(C::A_HALF >> 4);
                      if (c == 0) umma_f16_pair<1>(d0 + c * (N1CH / 2), al, w, idesc1, 1u);
                      else if (c == N1C - 1) umma_f16_pair<3>(d0 + c * (N1CH / 2), al, w, idesc1, 1u);
                      else umma_f16_pair<2>(d0 + c * (N1CH / 2), al, w, idesc1, 1u);
                    }
                  }
                } else {
#pragma unroll
                  for (int c = 0; c < N1C; ++c) {
                    const uint64_t w = db + (uint64_t)((c * C::W1_CHUNK + k * 32) >> 4);
                    umma_f16<2>(d0 + c * (N1CH / 2), a, w, idesc1, acc);
                    if (SPLIT == 3) {
                      umma_f16<2>(d0 + c * (N1CH / 2), a + (uint64_t)(C::A_HALF >> 4), w, idesc1, 1u);
                      umma_f16<2>(d0 + c * (N1CH / 2), a, w + (uint64_t)(C::B_HALF >> 4), idesc1, 1u);
                    }
                  }
                }
              }
              umma_commit_pair(&x_empty[sx], pair_mask);
              umma_commit_pair(&b_empty[sb], all_mask);
              if (kb == K1_STAGES - 1) umma_commit_pair(a1_full, pair_mask);
            }
            __syncwarp();
            if (++sb == nbr) { sb = 0; pb ^= 1; }
            if (++sx == NX) { sx = 0; px ^= 1; }
          }
          if (v == 0) { timed_wait(a3_empty, (rq & 1) ^ 1, st, 3); tc_fence_after(); }
          for (int q = 0; q < K3_STAGES; ++q) {
            timed_wait(&h_full[sh], ph, st, 4);
            timed_wait(&b_full[sb], pb, st, 5);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t da = make_desc<SW>(smem_u32(hs + (size_t)sh * C::A_STAGE));
              const uint64_t db = make_desc<SW>(smem_u32(bs + (size_t)sb * C::B_STAGE));
              const uint32_t d0 = tmem_base + ACC3_COL;
#pragma unroll
              for (int k = 0; k < C::KSTEPS; ++k) {
                const uint32_t acc = (v | q | k) ? 1u : 0u;
                const uint64_t a = da + (uint64_t)((k * 32) >> 4);
                if (COLL) {
                  static_assert(N3C == 2, "collector sequence below is written for two N chunks");
                  const uint64_t w0 = db + (uint64_t)((k * 32) >> 4), w1 = db + (uint64_t)((C::F_CHUNK + k * 32) >> 4);
                  const uint32_t e0 = d0, e1 = d0 + N3CH / 2;
                  if (SPLIT == 3) {
                    const uint64_t lo = (uint64_t)(C::B_HALF >> 4), al = a + (uint64_t)(C::A_HALF >> 4);
                    umma_f16_pair<1>(e0, a, w0, idesc3, acc);
                    umma_f16_pair<2>(e0, a, w0 + lo, idesc3, 1u);
                    umma_f16_pair<2>(e1, a, w1, idesc3, acc);
                    umma_f16_pair<3>(e1, a, w1 + lo, idesc3, 1u);
                    umma_f16_pair<1>(e0, al, w0, idesc3, 1u);
                    umma_f16_pair<3>(e1, al, w1, idesc3, 1u);
                  } else {
                    umma_f16_pair<1>(e0, a, w0, idesc3, acc);
                    umma_f16_pair<3>(e1, a, w1, idesc3, acc);
                  }
                } else {
#pragma unroll
                  for (int e = 0; e < N3C; ++e) {
                    const uint64_t w = db + (uint64_t)((e * C::F_CHUNK + k * 32) >> 4);
                    umma_f16<2>(d0 + e * (N3CH / 2), a, w, idesc3, acc);
                    if (SPLIT == 3) {
                      umma_f16<2>(d0 + e * (N3CH / 2), a + (uint64_t)(C::A_HALF >> 4), w, idesc3, 1u);
                      umma_f16<2>(d0 + e * (N3CH / 2), a, w + (uint64_t)(C::B_HALF >> 4), idesc3, 1u);
                    }
                  }
                }
              }
              umma_commit_pair(&h_empty[sh], pair_mask);
              umma_commit_pair(&b_empty[sb], all_mask);
              if (v == 1 && q == K3_STAGES - 1) umma_commit_pair(a3_full, pair_mask);
            }
            __syncwarp();
            if (++sb == nbr) { sb = 0; pb ^= 1; }
            if (++sh == NH) { sh = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp < 6) {
    // =========================== epilogue warps (2..5) ===========================
    const int sub = warp & 3;
    const int row = (sub & 1) * 32 + lane;                 // 0..63
    const int half = sub >> 1;                             // lanes 64..127 hold the upper half of each chunk
    const uint32_t tlane = tmem_base + ((uint32_t)(sub * 32) << 16);
    uint32_t av = 0, rq = 0, hi_count = 0;                 // hi_count: chunks this half has produced
    for (int itn = 0; itn < niter; ++itn, ++rq) {
      const int ray = item_of(itn);
      const bool valid = valid_of(itn);
      for (int v = 0; v < 2; ++v, ++av) {
        timed_wait(a1_full, av & 1, st, 0);
        tc_fence_after();
        uint32_t r[2][32];
        tmem_ld32(tlane + ACC1_COL, r[0]);                                    // chunk 0 in flight
#pragma unroll
        for (int cj = 0; cj < 9; ++cj) {
          const int c = cj / 3, jj = cj - c * 3;
          tmem_ld_wait();                                                      // r[cj & 1] has landed
          if (cj + 1 < 9) {                                                    // next chunk overlaps this one's math
            const int c2 = (cj + 1) / 3, j2 = (cj + 1) - c2 * 3;
            tmem_ld32(tlane + ACC1_COL + (uint32_t)(c2 * (N1CH / 2) + j2 * 32), r[(cj + 1) & 1]);
          }
          long long tc0 = st ? clock64() : 0;
          // KS = 32: one H stage per (block, lane half); KS = 64: the two lane halves of a block share a stage,
          // half h filling the K columns [32 h, 32 h + 32)
          const uint32_t hq = KS == 32 ? (hi_count + cj) * 2 + (uint32_t)half : hi_count + cj;   // global H stage index
          const int sh = hq % NH;
          const int n0 = c * N1CH + half * (N1CH / 2) + jj * 32;
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a = fmaxf(__uint_as_float(r[cj & 1][2 * i]) + sbias1[n0 + 2 * i], 0.f);
            float b = fmaxf(__uint_as_float(r[cj & 1][2 * i + 1]) + sbias1[n0 + 2 * i + 1], 0.f);
            split2<SPLIT == 3>(a, b, hi[i], lo[i]);
          }
          if (st) st[3] += (unsigned long long)(clock64() - tc0);
          timed_wait(&h_empty[sh], ((hq / NH) & 1) ^ 1, st, 1);
          tc0 = st ? clock64() : 0;
          uint8_t *dst = hs + (size_t)sh * C::A_STAGE;
#pragma unroll
          for (int c16 = 0; c16 < 4; ++c16) {
            const uint32_t off = swz_offset<SW>(row, (KS == 64 ? half * 4 : 0) + c16);
            *reinterpret_cast<uint4 *>(dst + off) = make_uint4(hi[4 * c16], hi[4 * c16 + 1], hi[4 * c16 + 2], hi[4 * c16 + 3]);
            if (SPLIT == 3) *reinterpret_cast<uint4 *>(dst + C::A_HALF + off) = make_uint4(lo[4 * c16], lo[4 * c16 + 1], lo[4 * c16 + 2], lo[4 * c16 + 3]);
          }
          if (st) { st[3] += (unsigned long long)(clock64() - tc0); tc0 = clock64(); }
          fence_async_smem();
          if (st) { st[4] += (unsigned long long)(clock64() - tc0); tc0 = clock64(); }
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&h_full[sh], leader_crank);
          if (st) st[5] += (unsigned long long)(clock64() - tc0);
        }
        hi_count += 9;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(a1_empty, leader_crank);
      }
      if (!C::A3W) {
        // acc3 drain on these warps, staged in the H ring: it is idle from a3_full (every GEMM3 MMA of this ray has
        // retired) until these same warps drain the next ray's acc1 - one barrier at the end keeps a fast warp's H
        // writes away from a slow warp's staging.  (bf16 mode: warps 14..17 do this concurrently, see below.)
        const long long td0 = st ? clock64() : 0;
        drain_acc3(std::integral_constant<int, 128>{}, ray, valid, rq, hs + sub * 4096, st);
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (st) st[6] += (unsigned long long)(clock64() - td0);
      }

    }
  } else if (C::A3W && warp >= BASE_WARPS) {
    // =========================== acc3 drain warps (14..17, bf16 mode) ===========================
    for (int itn = 0; itn < niter; ++itn)
      drain_acc3(std::integral_constant<int, 64>{}, item_of(itn), valid_of(itn), (uint32_t)itn, a3stage + (warp & 3) * A3_STAGE, nullptr);
  } else {
    // =========================== gather producers (warps 6..13) ===========================
    // Four groups of two warps; group gi owns X-ring slot gi and produces the K-stages with
    // xq % 4 == gi, so four stages are in flight per CTA.  Each thread issues all tap loads of
    // four (row, channel-group) items before consuming them (memory-level parallelism).
    const int pt = threadIdx.x - 6 * 32;                   // 0..255
    const int gi = pt >> 6, gt = pt & 63;                  // group, thread within group
    const FT *fbase[3] = {reinterpret_cast<const FT *>(p.feat[0]), reinterpret_cast<const FT *>(p.feat[1]),
                          reinterpret_cast<const FT *>(p.feat[2])};
    auto build_taps = [&](int ray_l, int buf) {
      if (pt < ROWS * 3) {
        const int rr = pt / 3, lvl = pt - rr * 3;
        const float *G = p.geom + (row_base(ray_l) + rr) * CAR_GEOM_STRIDE;
        const int h = lvl == 0 ? p.H / 4 : (lvl == 1 ? p.H / 2 : p.H);
        const int w = lvl == 0 ? p.W / 4 : (lvl == 1 ? p.W / 2 : p.W);
        TapEntry *tb = taps + buf * (ROWS * 3 * 2);
        tb[(rr * 3 + lvl) * 2 + 0] = make_taps(G[G_GX], G[G_GY], w, h, true);
        tb[(rr * 3 + lvl) * 2 + 1] = make_taps(G[G_GXC], G[G_GYC], w, h, false);
        if (lvl == 0) {
          float *th = tanhs + buf * (ROWS * 8);
#pragma unroll
          for (int i = 0; i < 3; ++i) { th[rr * 8 + i] = G[G_T0 + i]; th[rr * 8 + 4 + i] = G[G_T1 + i]; }
        }
      }
    };
    build_taps(item_of(0), 0);
    for (int it_ray = 0; it_ray < niter; ++it_ray) {
      const int ray = item_of(it_ray);
      const int scene = (p.g0 + ray / p.hpr) / p.R;
      const int buf = it_ray & 1;
      { long long tb0 = clock64();
        asm volatile("bar.sync 1, 256;" ::: "memory");     // table[buf] complete; table[buf^1] no longer read
        if (st) st[2] += (unsigned long long)(clock64() - tb0); }
      if (it_ray + 1 < niter) build_taps(item_of(it_ray + 1), buf ^ 1);
      const TapEntry *tb = taps + buf * (ROWS * 3 * 2);
      const float *th = tanhs + buf * (ROWS * 8);
      // this ray's 38 stages are numbered s = v*19 + kb; the group takes those with (xq0 + s) % 4 == gi
      const uint32_t xq0 = (uint32_t)it_ray * (2 * K1_STAGES);
      for (int sidx = (int)((gi + 4 - (xq0 & 3)) & 3); sidx < 2 * K1_STAGES; sidx += 4) {
        const uint32_t xq = xq0 + (uint32_t)sidx;
        const int v = sidx >= K1_STAGES ? 1 : 0, kb = sidx - v * K1_STAGES;
        const int oc = (v == (int)rank) ? 0 : 1;           // own line (border taps) or cross-view taps
        const int sx = (int)(xq % NX);
        const uint32_t xpar = ((xq / NX) & 1) ^ 1;            // wait for the slot only right before the first store
        uint8_t *dst = xs + (size_t)sx * C::A_STAGE;
        if (kb < K1_STAGES - 1) {
          const int ch0 = kb * KS;
          const int lvl = ch0 < 256 ? 0 : (ch0 < 512 ? 1 : 2);
          const int coff = ch0 - (lvl == 0 ? 0 : (lvl == 1 ? 256 : 512));
          const int Cc = lvl == 2 ? 64 : 256;
          const int h = lvl == 0 ? p.H / 4 : (lvl == 1 ? p.H / 2 : p.H);
          const int w = lvl == 0 ? p.W / 4 : (lvl == 1 ? p.W / 2 : p.W);
          const FT *img = fbase[lvl] + (size_t)(scene * 2 + v) * h * w * Cc + coff;
          // An item = one 16-byte piece of one sample row: CPI channels (4 fp32 / 8 bf16).  A stage is
          // 64 rows x IPR items; thread gt handles items gt + 64 i, i < IPR, with a rolling window of 4 items
          // (16 x LDG.128) in flight; the ring slot is awaited only right before the first store.
          constexpr int CPI = sizeof(FT) == 4 ? 4 : 8;
          constexpr int IPR = KS / CPI;                    // items per row = items per thread: 8/16 (fp32), 4/8 (bf16)
          uint4 x[4][4];
          float wt[4][4];
          auto load_item = [&](int slot, int i) {
            const int item = gt + 64 * i, rr = item / IPR, grp = item % IPR;
            const TapEntry t = tb[(rr * 3 + lvl) * 2 + oc];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              wt[slot][k] = t.off[k] >= 0 ? t.w[k] : 0.f;
              const int o = t.off[k] >= 0 ? t.off[k] : 0;
              x[slot][k] = __ldg(reinterpret_cast<const uint4 *>(img + (size_t)o * Cc + grp * CPI));
            }
          };
#pragma unroll
          for (int i = 0; i < 4; ++i) load_item(i, i);
          timed_wait(&x_empty[sx], xpar, st, 0);
#pragma unroll
          for (int i = 0; i < IPR; ++i) {
            const int item = gt + 64 * i, rr = item / IPR, grp = item % IPR, sl = i & 3;
            if (sizeof(FT) == 4) {
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float ww = wt[sl][k];
                acc.x = fmaf(__uint_as_float(x[sl][k].x), ww, acc.x); acc.y = fmaf(__uint_as_float(x[sl][k].y), ww, acc.y);
                acc.z = fmaf(__uint_as_float(x[sl][k].z), ww, acc.z); acc.w = fmaf(__uint_as_float(x[sl][k].w), ww, acc.w);
              }
              if (i + 4 < IPR) load_item(sl, i + 4);
              uint32_t h0, l0, h1, l1;
              split2<SPLIT == 3>(acc.x, acc.y, h0, l0);
              split2<SPLIT == 3>(acc.z, acc.w, h1, l1);
              const uint32_t off = swz_offset<SW>(rr, grp >> 1) + (uint32_t)((grp & 1) * 8);
              *reinterpret_cast<uint2 *>(dst + off) = make_uint2(h0, h1);
              if (SPLIT == 3) *reinterpret_cast<uint2 *>(dst + C::A_HALF + off) = make_uint2(l0, l1);
            } else {
              float a8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) a8[j] = 0.f;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t uu[4] = {x[sl][k].x, x[sl][k].y, x[sl][k].z, x[sl][k].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  a8[2 * j] = fmaf(__uint_as_float(uu[j] << 16), wt[sl][k], a8[2 * j]);
                  a8[2 * j + 1] = fmaf(__uint_as_float(uu[j] & 0xffff0000u), wt[sl][k], a8[2 * j + 1]);
                }
              }
              if (i + 4 < IPR) load_item(sl, i + 4);
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) split2<SPLIT == 3>(a8[2 * j], a8[2 * j + 1], hi[j], lo[j]);
              const uint32_t off = swz_offset<SW>(rr, grp);
              *reinterpret_cast<uint4 *>(dst + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              if (SPLIT == 3) *reinterpret_cast<uint4 *>(dst + C::A_HALF + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        } else {
          // last stage: [tanh(pt_v / 5) (3) | zeros]; only the first 16 K-columns are multiplied
          timed_wait(&x_empty[sx], xpar, st, 0);
          constexpr int CHK = SW / 16;                     // 16-byte chunks per row: 4 / 8
#pragma unroll
          for (int i = 0; i < CHK; ++i) {
            const int item = gt + 64 * i, rr = item / CHK, c16 = item % CHK;
            uint32_t hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
            if (c16 == 0) {
              const float *T = th + rr * 8 + v * 4;
              split2<SPLIT == 3>(T[0], T[1], hi[0], lo[0]);
              split2<SPLIT == 3>(T[2], 0.f, hi[1], lo[1]);
            }
            const uint32_t off = swz_offset<SW>(rr, c16);
            *reinterpret_cast<uint4 *>(dst + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (SPLIT == 3) *reinterpret_cast<uint4 *>(dst + C::A_HALF + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        { const long long tf0 = st ? clock64() : 0;
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&x_full[sx], leader_crank);
          if (st) st[3] += (unsigned long long)(clock64() - tf0); }
      }
    }
  }
  if (st && lane == 0) {
    const unsigned long long tot = (unsigned long long)(clock64() - t_begin);
    if (warp == 1) { for (int i = 0; i < 6; ++i) atomicAdd(p.stats + i, st[i]); atomicAdd(p.stats + 6, tot); }
    else if (warp == 2) { for (int i = 0; i < 7; ++i) atomicAdd(p.stats + 8 + i, st[i]); atomicAdd(p.stats + 15, tot); atomicAdd(p.stats + 22, st[7]); }
    else if (warp == 6) { atomicAdd(p.stats + 16, st[0]); atomicAdd(p.stats + 17, tot); atomicAdd(p.stats + 18, st[2]); atomicAdd(p.stats + 19, st[3]); }
    else if (warp == 0) { atomicAdd(p.stats + 20, st[0]); atomicAdd(p.stats + 21, tot); }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, 512);
  }
}

}  // namespace

unsigned long long *g_fused_stats = nullptr;   // device buffer [32], set by car_debug_set_fused_stats

// Launch for rays [g0,g1) of the current chunk.  Returns 0 or an error code.
int launch_fused_encode(const car_render_args &a, int g0, int g1, const float *geom, float *value,
                        uint16_t *kh_hi, uint16_t *kh_lo, cudaStream_t st) {
  const int split3 = a.precision == CAR_PREC_FP32_3XBF16;
  // stage width: 64 K-columns in bf16 mode (needs the column-permuted F), 32 in the hi+lo mode.  CAR_KS=32 forces
  // the narrow stages (A/B diagnostics).
  static int ks_env = -1;
  if (ks_env < 0) { const char *e_ = getenv("CAR_KS"); ks_env = e_ ? atoi(e_) : 0; }
  const int KS = (!split3 && ks_env != 32 && a.weights.kv_fold64.hi) ? 64 : 32;
  const car_mat &W1 = a.weights.enc1, &F = KS == 64 ? a.weights.kv_fold64 : a.weights.kv_fold;
  if (a.P % ROWS != 0 || a.P > 256 || !F.hi || F.N != N3 || F.K != 2 * N1 || W1.N != N1 || W1.K != CAR_K_ENC) {
    set_error("fused encode: unsupported configuration (P=%d)", a.P);
    return -20;
  }
  CUtensorMap t1h, t1l, tfh, tfl;
  int rc;
  if ((rc = make_tmap_bf16(&t1h, W1.hi, W1.N, W1.K, W1.K, N1CH / 2, KS))) return rc;
  if ((rc = make_tmap_bf16(&tfh, F.hi, F.N, F.K, F.K, N3CH / 2, KS))) return rc;
  if (split3) {
    if ((rc = make_tmap_bf16(&t1l, W1.lo, W1.N, W1.K, W1.K, N1CH / 2, KS))) return rc;
    if ((rc = make_tmap_bf16(&tfl, F.lo, F.N, F.K, F.K, N3CH / 2, KS))) return rc;
  } else { t1l = t1h; tfl = tfh; }
  FusedParams p;
  for (int i = 0; i < 3; ++i) p.feat[i] = a.feat[i];
  p.H = a.H; p.W = a.W; p.R = a.R; p.g0 = g0; p.g1 = g1;
  p.P = a.P; p.hpr = a.P / ROWS;
  p.geom = geom; p.bias1 = W1.bias; p.biasf = F.bias;
  p.value = value; p.kh_hi = kh_hi; p.kh_lo = kh_lo;
  p.stats = g_fused_stats;
  const int ops = split3 ? 2 : 1;
  const int a_stage = ROWS * KS * 2 * ops;
  const int b_stage = N1C * (N1CH / 2) * KS * 2 * ops;
  const int nh = KS == 64 ? Cfg<1, 64>::NH : (split3 ? Cfg<3, 32>::NH : Cfg<1, 32>::NH);
  const int nx = KS == 64 ? Cfg<1, 64>::NX : (split3 ? Cfg<3, 32>::NX : Cfg<1, 32>::NX);
  const size_t fixed = (size_t)(nx + nh) * a_stage + 2 * (ROWS * 3 * 2 * sizeof(TapEntry) + ROWS * 8 * 4) + (N1 + N3) * 4 +
                       (2 * NXMAX + 2 * NHMAX + 2 * MAXB + 4) * 8 + 16 + 1024 + (KS == 64 && Cfg<1, 64>::A3W ? 4 * A3_STAGE : 0);
  int nb = (int)((227 * 1024 - fixed) / b_stage);
  if (nb > MAXB) nb = MAXB;
  if (nb < 2) { set_error("fused encode: not enough shared memory"); return -21; }
  p.nb = nb;
  const size_t smem = fixed + (size_t)nb * b_stage;
  const int sms = sm_count();
  int pairs = sms / 2;
  if (pairs > (g1 - g0) * p.hpr) pairs = (g1 - g0) * p.hpr;
  static int cl_env = -1;
  if (cl_env < 0) { const char *e_ = getenv("CAR_CLUSTER"); cl_env = e_ ? atoi(e_) : 2; if (cl_env != 2 && cl_env != 4) cl_env = 2;   /* 4 = two pairs share multicast weight loads: correct, but the lock-step costs more than the L2 traffic it saves (measured 897 vs 522 ms) */ }
  p.cl = cl_env;
  if (p.cl == 4) { pairs &= ~1; if (pairs < 2) { p.cl = 2; pairs = (g1 - g0) * p.hpr < sms / 2 ? (g1 - g0) * p.hpr : sms / 2; } }
  cudaError_t e = cudaSuccess;
  prof_pre(CAR_ST_FUSED, st);
#define CAR_LAUNCH(S, T, K)                                                                                 \
  do {                                                                                                      \
    e = cudaFuncSetAttribute(k_fused_encode<S, T, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) {                                                                                   \
      cudaLaunchConfig_t cfg = {};                                                                            \
      cfg.gridDim = dim3(pairs * 2); cfg.blockDim = dim3(Cfg<S, K>::THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st; \
      cudaLaunchAttribute at[1];                                                                              \
      at[0].id = cudaLaunchAttributeClusterDimension;                                                         \
      at[0].val.clusterDim.x = p.cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;                  \
      cfg.attrs = at; cfg.numAttrs = 1;                                                                       \
      /* persistent kernel: every cluster must be resident at once.  Clusters of 4 do not tile the GPCs as      \
         densely as pairs do, so ask how many fit and launch exactly that many (the work split follows gridDim) */ \
      int maxc = 0;                                                                                           \
      if (cudaOccupancyMaxActiveClusters(&maxc, k_fused_encode<S, T, K>, &cfg) == cudaSuccess && maxc > 0 &&  \
          maxc * p.cl < pairs * 2) {                                                                          \
        pairs = maxc * (p.cl / 2);                                                                            \
        cfg.gridDim = dim3(pairs * 2);                                                                        \
      }                                                                                                       \
      if (getenv("CAR_FUSED_VERBOSE")) {                                                                      \
        cudaFuncAttributes fa;                                                                                \
        cudaFuncGetAttributes(&fa, k_fused_encode<S, T, K>);                                                  \
        fprintf(stderr, "fused encode: cluster %d, %d clusters fit, grid %d, block %d, regs %d, maxThreadsPerBlock %d, smem %zu\n", \
                p.cl, maxc, pairs * 2, (int)Cfg<S, K>::THREADS, fa.numRegs, fa.maxThreadsPerBlock, smem);     \
      }                                                                                                       \
      e = cudaLaunchKernelEx(&cfg, k_fused_encode<S, T, K>, t1h, t1l, tfh, tfl, p);                           \
    }                                                                                                         \
  } while (0)
  if (split3) { if (a.feat_bf16) CAR_LAUNCH(3, __nv_bfloat16, 32); else CAR_LAUNCH(3, float, 32); }
  else if (KS == 64) { if (a.feat_bf16) CAR_LAUNCH(1, __nv_bfloat16, 64); else CAR_LAUNCH(1, float, 64); }
  else { if (a.feat_bf16) CAR_LAUNCH(1, __nv_bfloat16, 32); else CAR_LAUNCH(1, float, 32); }
#undef CAR_LAUNCH
  prof_post(st);
  if (e != cudaSuccess) { set_error("fused encode: %s", cudaGetErrorString(e)); return (int)e; }
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("fused encode launch: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  return 0;
}

}  // namespace car

extern "C" int car_debug_set_fused_stats(void *dev_u64x32) {
  car::g_fused_stats = reinterpret_cast<unsigned long long *>(dev_u64x32);
  return 0;
}
