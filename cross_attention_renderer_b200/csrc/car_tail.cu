// Per-ray attention tail (P == 64 or 128): the small per-sample MLPs and both attention rounds in two
// kernels instead of five GEMM launches + two attention launches with every [rows x 128] fp32
// intermediate round-tripping through HBM.
//
//   phase A (reference models.py:491,529,532-545,573-594)
//     K  = key_map_2(relu(key_map(.)))            A operand: relu(key_map) bf16 hi/lo from the fused kernel
//     Q1 = query_embed_2(relu(query_embed(local)))
//     s  = <K,Q1>/16, joint softmax over the ray's 128 samples -> at_wt, at_wt_max, z_sum = sum a*V,
//     expected depth.  Q1 is also written out (fp32) for phase B.
//   (between the phases: encode_latent and query_repeat_embed[:, :128] per ray, M = rays)
//   phase B (models.py:552-565)
//     Q2 = query_repeat_embed_2(relu(query_repeat_embed[:,128:](local) + u_ray))
//     s2 = <Q2,Q1>/16, softmax, z = sum a2*V + 2*z_sum
//
// One CTA = one ray at a time, persistent over the chunk.  A ray is TT = P / 64 tiles of 128 sample rows
// (UMMA M = 128; TT = 1: both contexts in one tile, TT = 2: one tile per context): the MLPs and the row
// dots run per tile, the joint softmax and the V sums once per ray over all of its tiles.
//   warp 0  TMA: relu(key_map) tile and the weight K-blocks (ring)
//   warp 1  MMA issuer (tcgen05 cta_group::1, N = 128), accumulators in TMEM
//   warps 2-5: one thread per sample row: TMEM -> bias/ReLU -> bf16 hi/lo A operand of the next
//           GEMM (smem, 128B swizzle), row dots, softmax (shuffles + named barrier).
//   warps 6-13: two groups of four V-sum warps; group g takes the rays whose softmax weights the row warps post in
//           weight buffer g (alternate rays): attention-weighted V sums, at_wt / argmax / depth outputs.  The V sums
//           are bound by the latency of their loads (24 x LDG.128 in flight per lane, 11.7 k cycles per ray against
//           6.5 k for the row warps' MLP chain); two rays in flight per CTA double the bytes in flight.
#include <math.h>

#include "car_common.cuh"
#include "car_umma.cuh"

namespace car {
extern unsigned long long *g_fused_stats;
int make_tmap_bf16(CUtensorMap *tm, const uint16_t *base, int rows, int K, int ld, int box_rows, int box_k);

namespace {
using namespace ptx;

// (An L2 prefetch of the next ray's V / Q1 with cp.async.bulk.prefetch.L2 doubled the DRAM reads of both
// phases - ncu: 456 / 429 KB per ray against 291 / 220 KB algorithmic, L2 hit rate 13-20 % - and was removed.)
// TMA, MMA, 4 row warps, VG groups of 4 V-sum warps.  Two groups (448 threads, 128 registers per thread) pay in the
// single-MMA bf16 mode (tail 77 -> 72.5 ms per step); in the hi+lo mode the row threads need their 168 registers -
// at 128 their MLP chain grows from 14.1 k to 15.6 k cycles per ray and becomes the bottleneck (99 -> 109 ms) - so
// that mode keeps one group (320 threads).
template <int SPLIT> struct TailShape {
  static constexpr int VG = SPLIT == 3 ? 1 : 2;
  static constexpr int THREADS = (2 + 4 + 4 * VG) * 32;
};
constexpr int NBMAX = 5;

struct TailParams {
  car_render_args a;                  // sizes, geometry outputs (at_wt, at_wt_max, depth_ray), cams.qinv
  int g0, g1;
  const float *geom;                  // (rows,32)
  const float *value;                 // (rows,288) fp32
  float *q1;                          // per ray [128 cols][128 rows] fp32 (column-major so that the one-row-per-lane
                                      // accesses coalesce): written in phase A, read in phase B
  float *zsum;                        // (rays,288)
  uint16_t *zs_hi, *zs_lo;            // (rays,288) optional bf16 hi (+lo) copy of this phase's per-ray output: zsum (phase A, A operand
                                      // of the row-bias GEMM) / z (phase B, operand of the fused colour MLP)
  const float *rowbias;               // (rays,128)  phase B
  float *zfin;                        // (rays,288)  phase B
  const float *bias_k2, *bias_q1, *bias_q2, *bias_r2;
  int nb;
  int resident;                       // the phase's weight K-blocks (5 in phase A, 3 in phase B) all fit the B slots (bf16 mode):
                                      // loaded once per CTA and kept instead of streamed once per tile.  Removes 30 % of the
                                      // tail's L2 traffic; the kernel time did not change (72.3 ms per step: the per-ray
                                      // dependent chain bounds it)
  unsigned long long *stats;         // optional [16]: row-thread cycle accounting of CTA 0 (see scripts/tail_stalls.py)
};

template <int SPLIT> struct TCfg {
  static constexpr int OPS = SPLIT == 3 ? 2 : 1;
  static constexpr int KB_BYTES = 128 * 128;                  // one K-block (64 bf16) of a 128-row tile
  static constexpr int TILE_HALF = 2 * KB_BYTES;              // 128 x 128 bf16 (hi part)
  static constexpr int TILE = TILE_HALF * OPS;
  static constexpr int B_HALF = KB_BYTES;                     // 128 weight rows x 64 K
  static constexpr int B_STAGE = B_HALF * OPS;
};

template <bool WITH_LO>
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

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void rows_sync() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

// PHASE 0 = A, 1 = B
template <int SPLIT, int PHASE, int TT>
__global__ void __launch_bounds__(TailShape<SPLIT>::THREADS, 1)
k_tail(const __grid_constant__ CUtensorMap tm_kh_hi, const __grid_constant__ CUtensorMap tm_kh_lo,
       const __grid_constant__ CUtensorMap tm_w0_hi, const __grid_constant__ CUtensorMap tm_w0_lo,   // key2 (A only)
       const __grid_constant__ CUtensorMap tm_w1_hi, const __grid_constant__ CUtensorMap tm_w1_lo,   // K=16 layer
       const __grid_constant__ CUtensorMap tm_w2_hi, const __grid_constant__ CUtensorMap tm_w2_lo,   // 128x128 layer
       TailParams p) {
  using C = TCfg<SPLIT>;
  constexpr int VG = TailShape<SPLIT>::VG;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *at0 = smem;                                        // relu(key_map) tile (phase A)
  constexpr int Q1_BYTES = 128 * 128 * 4;                     // Q1 of one ray, [col][row] fp32 (phase B: staged by cp.async)
  const float *q1s = reinterpret_cast<const float *>(at0);
  uint8_t *at1 = at0 + (PHASE == 0 ? C::TILE : Q1_BYTES);     // local (first 32 B of each row) then the hidden tile
  uint8_t *bs = at1 + C::TILE;
  float *arow = reinterpret_cast<float *>(bs + (size_t)p.nb * C::B_STAGE);   // [2][128 TT] softmax weights
  float *part_all = arow + 256 * TT;                                                // [2 groups][4 TT][288] V sums per 32-row chunk
  float *red = part_all + 2 * 4 * TT * CAR_C_LAT;                              // [32] scratch
  float *vred_all = red + 32;                                                  // [2 groups][32] scratch of the V warps
  float *sbias = vred_all + 64;                                                // [3][128] hidden / output / key biases
  uint64_t *bars = reinterpret_cast<uint64_t *>(sbias + 3 * 128);
  uint64_t *kh_full = bars, *kh_empty = bars + 1, *loc_full = bars + 2, *hid_full = bars + 3;
  uint64_t *t_full = bars + 4, *k_full = bars + 5, *done = bars + 6;
  uint64_t *b_full = bars + 8, *b_empty = b_full + NBMAX;
  uint64_t *a_full = b_empty + NBMAX, *a_empty = a_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nrays = p.g1 - p.g0;
  constexpr int P = 64 * TT;
  // this CTA's work list: rays blockIdx.x, blockIdx.x + gridDim.x, ...; TT consecutive 128-row tiles per ray
  const int my_rays = nrays > (int)blockIdx.x ? (nrays - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int ntile = my_rays * TT;
  auto ray_of = [&](int it_) { return (int)blockIdx.x + (it_ / TT) * (int)gridDim.x; };
  auto tile_of = [&](int it_) { return ray_of(it_) * TT + (it_ % TT); };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_w1_hi);
    prefetch_tmap(&tm_w2_hi);
    mbar_init(kh_full, 1); mbar_init(kh_empty, 1); mbar_init(loc_full, 4); mbar_init(hid_full, 4);
    mbar_init(t_full, 1); mbar_init(k_full, 1); mbar_init(done, 4);
    for (int s = 0; s < p.nb; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 4); }
    fence_barrier_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 192) {
    const int c = threadIdx.x - 64;
    sbias[c] = PHASE == 0 ? p.bias_q1[c] : 0.f;
    sbias[128 + c] = PHASE == 0 ? p.bias_q2[c] : p.bias_r2[c];
    sbias[256 + c] = PHASE == 0 ? p.bias_k2[c] : 0.f;
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t ACCK = 0, ACCT = 128;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      uint32_t bq = 0;
      auto load_w = [&](const CUtensorMap *hi, const CUtensorMap *lo, int k0) {
        const int s = bq % p.nb;
        mbar_wait(&b_empty[s], ((bq / p.nb) & 1) ^ 1);
        uint8_t *st = bs + (size_t)s * C::B_STAGE;
        mbar_expect_tx(&b_full[s], (uint32_t)C::B_STAGE);
        tma_load_2d(st, hi, &b_full[s], k0, 0);
        if (SPLIT == 3) tma_load_2d(st + C::B_HALF, lo, &b_full[s], k0, 0);
        ++bq;
      };
      for (int it = 0; it < ntile; ++it) {
        if (PHASE == 0) {
          mbar_wait(kh_empty, (it & 1) ^ 1);
          mbar_expect_tx(kh_full, (uint32_t)C::TILE);
          const int r0 = tile_of(it) * 128;
          tma_load_2d(at0, &tm_kh_hi, kh_full, 0, r0);
          tma_load_2d(at0 + C::KB_BYTES, &tm_kh_hi, kh_full, 64, r0);
          if (SPLIT == 3) {
            tma_load_2d(at0 + C::TILE_HALF, &tm_kh_lo, kh_full, 0, r0);
            tma_load_2d(at0 + C::TILE_HALF + C::KB_BYTES, &tm_kh_lo, kh_full, 64, r0);
          }
        }
        if (p.resident && it > 0) continue;              // weights stay in their slots
        load_w(&tm_w1_hi, &tm_w1_lo, 0);                 // K = 16 layer: columns 16..63 are OOB zero fill
        if (PHASE == 0) {
          load_w(&tm_w0_hi, &tm_w0_lo, 0);
          load_w(&tm_w0_hi, &tm_w0_lo, 64);
        }
        load_w(&tm_w2_hi, &tm_w2_lo, 0);
        load_w(&tm_w2_hi, &tm_w2_lo, 64);
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc_bf16(128, 128);
    uint32_t bq = 0, tq = 0;
    // one 64-wide K-block: `ksteps` MMAs (x3 in the split mode) of A(base a) x B(stage) into d
    auto gemm_kb = [&](uint32_t d, uint32_t a_addr, int ksteps, bool first) {
      const int s = bq % p.nb;
      mbar_wait(&b_full[s], p.resident ? 0u : ((bq / p.nb) & 1));     // resident: completed once, never re-armed
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = make_desc<128>(a_addr);
        const uint64_t db = make_desc<128>(smem_u32(bs + (size_t)s * C::B_STAGE));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k >= ksteps) break;
          const uint32_t acc = (first && k == 0) ? 0u : 1u;
          const uint64_t a = da + (uint64_t)((k * 32) >> 4), w = db + (uint64_t)((k * 32) >> 4);
          umma_f16<1>(d, a, w, idesc, acc);
          if (SPLIT == 3) {
            umma_f16<1>(d, a + (uint64_t)(C::TILE_HALF >> 4), w, idesc, 1u);
            umma_f16<1>(d, a, w + (uint64_t)(C::B_HALF >> 4), idesc, 1u);
          }
        }
        if (!p.resident) umma_commit(&b_empty[s]);
      }
      __syncwarp();
      ++bq;
    };
    for (int it = 0; it < ntile; ++it) {
      mbar_wait(done, (it & 1) ^ 1);                     // row threads finished reading both accumulators
      tc_fence_after();
      mbar_wait(loc_full, it & 1);
      tc_fence_after();
      gemm_kb(tmem_base + ACCT, smem_u32(at1), 1, true);                    // K = 16 layer (first: the row threads drain it next)
      if (elect_one()) umma_commit(t_full);
      __syncwarp();
      ++tq;
      if (PHASE == 0) {
        mbar_wait(kh_full, it & 1);
        tc_fence_after();
        gemm_kb(tmem_base + ACCK, smem_u32(at0), 4, true);
        gemm_kb(tmem_base + ACCK, smem_u32(at0 + C::KB_BYTES), 4, false);
        if (elect_one()) { umma_commit(kh_empty); umma_commit(k_full); }
        __syncwarp();
      }
      mbar_wait(hid_full, it & 1);
      tc_fence_after();
      gemm_kb(tmem_base + ACCT, smem_u32(at1), 4, true);
      gemm_kb(tmem_base + ACCT, smem_u32(at1 + C::KB_BYTES), 4, false);
      if (elect_one()) umma_commit(t_full);
      __syncwarp();
      ++tq;
    }
  } else if (warp < 6) {
    // =========================== row threads (warps 2..5) ===========================
    const int sub = warp & 3;
    const int row = sub * 32 + lane;                     // 0..127: ctx = row >> 6, sample k = row & 63
    const uint32_t tlane = tmem_base + ((uint32_t)(sub * 32) << 16);
    const int rt = row;                                  // thread id among the 128 row threads (for column loops)
    uint32_t tq = 0;
    unsigned long long tacc[6] = {0, 0, 0, 0, 0, 0};     // 0 drain 1 score-wait 2 score 3 softmax 4 vsum 5 total
    const bool rec = p.stats && blockIdx.x == 0 && warp == 2;
    const long long tbeg = clock64();
    // local_coords of a tile are fetched one tile ahead so their latency is off the critical path
    float4 l0, l1, l2, l3;
    auto fetch_geom = [&](int tile_l) {
      const float *Gp = p.geom + ((size_t)tile_l * 128 + row) * CAR_GEOM_STRIDE;
      l0 = __ldg(reinterpret_cast<const float4 *>(Gp + G_LOCAL)); l1 = __ldg(reinterpret_cast<const float4 *>(Gp + G_LOCAL + 4));
      l2 = __ldg(reinterpret_cast<const float4 *>(Gp + G_LOCAL + 8)); l3 = __ldg(reinterpret_cast<const float4 *>(Gp + G_LOCAL + 12));
    };
    auto write_loc = [&]() {
      // local_coords (16 fp32) -> bf16 hi(+lo), first 32 bytes of this row of the at1 tile (K-block 0)
      uint32_t hi[8], lo[8];
      split2<SPLIT == 3>(l0.x, l0.y, hi[0], lo[0]); split2<SPLIT == 3>(l0.z, l0.w, hi[1], lo[1]);
      split2<SPLIT == 3>(l1.x, l1.y, hi[2], lo[2]); split2<SPLIT == 3>(l1.z, l1.w, hi[3], lo[3]);
      split2<SPLIT == 3>(l2.x, l2.y, hi[4], lo[4]); split2<SPLIT == 3>(l2.z, l2.w, hi[5], lo[5]);
      split2<SPLIT == 3>(l3.x, l3.y, hi[6], lo[6]); split2<SPLIT == 3>(l3.z, l3.w, hi[7], lo[7]);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t off = swz_offset<128>(row, c);
        *reinterpret_cast<uint4 *>(at1 + off) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
        if (SPLIT == 3) *reinterpret_cast<uint4 *>(at1 + C::TILE_HALF + off) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(loc_full);
    };
    // hidden layer: accT -> (+bias | +row bias) -> ReLU -> bf16 hi/lo -> at1 (A operand of the 128x128 layer)
    auto drain_hidden = [&](int ray_l) {
      const long long td = rec ? clock64() : 0;
      mbar_wait(t_full, tq & 1); ++tq;
      tc_fence_after();
      {
        const float4 *hb4 = reinterpret_cast<const float4 *>(PHASE == 0 ? sbias : p.rowbias + (size_t)ray_l * 128);
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          tmem_ld32(tlane + ACCT + (uint32_t)(j * 32), r);
          float4 hb[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) hb[i] = PHASE == 0 ? hb4[j * 8 + i] : __ldg(hb4 + j * 8 + i);
          tmem_ld_wait();
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float x0 = fmaxf(__uint_as_float(r[4 * i]) + hb[i].x, 0.f);
            const float x1 = fmaxf(__uint_as_float(r[4 * i + 1]) + hb[i].y, 0.f);
            const float x2 = fmaxf(__uint_as_float(r[4 * i + 2]) + hb[i].z, 0.f);
            const float x3 = fmaxf(__uint_as_float(r[4 * i + 3]) + hb[i].w, 0.f);
            split2<SPLIT == 3>(x0, x1, hi[2 * i], lo[2 * i]);
            split2<SPLIT == 3>(x2, x3, hi[2 * i + 1], lo[2 * i + 1]);
          }
          uint8_t *dst = at1 + (j >> 1) * C::KB_BYTES;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t off = swz_offset<128>(row, (j & 1) * 4 + c);
            *reinterpret_cast<uint4 *>(dst + off) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
            if (SPLIT == 3) *reinterpret_cast<uint4 *>(dst + C::TILE_HALF + off) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
          }
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(hid_full);
      }
      if (rec) tacc[0] += (unsigned long long)(clock64() - td);
    };
    // phase B: Q1 of a ray (64 KB, written by phase A) is staged into smem with per-thread 16-byte
    // cp.async copies issued one ray ahead (after the barrier that ends the previous ray's reads)
    auto stage_q1 = [&](int tile_l) {
      const char *src = reinterpret_cast<const char *>(p.q1 + (size_t)tile_l * 128 * 128);
      const uint32_t dst = smem_u32(at0);
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        const uint32_t o = (uint32_t)(k * 128 + rt) * 16u;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + o), "l"(src + o) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (ntile > 0) {
      if (PHASE == 1) stage_q1(tile_of(0));
      fetch_geom(tile_of(0));
      write_loc();
      if (ntile > 1) fetch_geom(tile_of(1));
      drain_hidden(ray_of(0));
    }
    float sc_prev = 0.f;                                 // TT == 2: score of this row in the ray's first tile
    for (int it = 0; it < ntile; ++it) {
      const int tile = tile_of(it), t = it % TT;
      // ---- scores: <K,Q1> (phase A) or <Q2,Q1> (phase B), each thread its own row ----
      float sc = 0.f;
      long long tt = rec ? clock64() : 0;
      float *q1col = p.q1 + (size_t)tile * 128 * 128 + row;     // element (col c, this row) at q1col[c * 128]
      if (PHASE == 0) { mbar_wait(k_full, it & 1); }
      else { asm volatile("cp.async.wait_group 0;" ::: "memory"); rows_sync(); }   // every thread's Q1 chunks landed
      mbar_wait(t_full, tq & 1); ++tq;
      tc_fence_after();
      if (rec) { tacc[1] += (unsigned long long)(clock64() - tt); tt = clock64(); }
      {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t rt_[32], rk[32];
          tmem_ld32(tlane + ACCT + (uint32_t)(j * 32), rt_);
          if (PHASE == 0) tmem_ld32(tlane + ACCK + (uint32_t)(j * 32), rk);
          tmem_ld_wait();
          if (PHASE == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float q[4], k4[4];
              const float4 bq4 = *reinterpret_cast<const float4 *>(sbias + 128 + j * 32 + i);
              const float4 bk4 = *reinterpret_cast<const float4 *>(sbias + 256 + j * 32 + i);
              const float bq_[4] = {bq4.x, bq4.y, bq4.z, bq4.w}, bk_[4] = {bk4.x, bk4.y, bk4.z, bk4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                q[e] = __uint_as_float(rt_[i + e]) + bq_[e];
                k4[e] = __uint_as_float(rk[i + e]) + bk_[e];
                sc = fmaf(k4[e], q[e], sc);
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) q1col[(size_t)(j * 32 + i + e) * 128] = q[e];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 br4 = *reinterpret_cast<const float4 *>(sbias + 128 + j * 32 + i);
              const float br_[4] = {br4.x, br4.y, br4.z, br4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float x = __uint_as_float(rt_[i + e]) + br_[e];
                sc = fmaf(x, q1s[(j * 32 + i + e) * 128 + row], sc);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(done);                               // accumulators and at1 may be reused
      }
      if (it + 1 < ntile) {                               // next tile's K=16 operand (at1 is free: hidden GEMM retired)
        write_loc();
        if (it + 2 < ntile) fetch_geom(tile_of(it + 2));
      }
      if (rec) { tacc[2] += (unsigned long long)(clock64() - tt); tt = clock64(); }
      sc = sc / 16.0f;
      if (TT == 2 && t == 0) {
        // first tile of the ray: keep the score, hand the pipeline to the second tile
        sc_prev = sc;
        if (PHASE == 1) { rows_sync(); stage_q1(tile_of(it + 1)); }   // all four warps are past their Q1 reads
        drain_hidden(ray_of(it + 1));
        continue;
      }
      // ---- joint softmax over the ray's 128 TT samples (this thread: one row per tile) ----
      float mx = warp_max(TT == 2 ? fmaxf(sc, sc_prev) : sc);
      if (lane == 0) red[sub] = mx;
      rows_sync();
      if (PHASE == 1 && it + 1 < ntile) stage_q1(tile_of(it + 1));   // all four warps are past their Q1 reads
      mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
      const float e = expf(sc - mx), e0 = TT == 2 ? expf(sc_prev - mx) : 0.f;
      float sm = warp_sum(TT == 2 ? e0 + e : e);
      if (lane == 0) red[4 + sub] = sm;
      rows_sync();
      sm = (red[4] + red[5]) + (red[6] + red[7]);
      const float aw = e / sm;
      {
        // post the weights for the V-sum warps (double-buffered per ray: buffer ir&1 was last read two rays ago)
        const long long tw = rec ? clock64() : 0;
        const uint32_t ir = (uint32_t)(it / TT), buf = ir & 1;
        mbar_wait(&a_empty[buf], ((ir >> 1) & 1) ^ 1);
        if (TT == 2) arow[buf * 256 + row] = e0 / sm;
        arow[buf * (128 * TT) + (TT - 1) * 128 + row] = aw;
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[buf]);
        if (rec) tacc[4] += (unsigned long long)(clock64() - tw);
      }
      if (rec) tacc[3] += (unsigned long long)(clock64() - tt);
      // next tile's hidden layer: its 128x128 GEMM then runs while the V warps sum this ray
      if (it + 1 < ntile) drain_hidden(ray_of(it + 1));
    }
    if (rec && lane == 0) {
      tacc[5] = (unsigned long long)(clock64() - tbeg);
      for (int i = 0; i < 6; ++i) atomicAdd(p.stats + i, tacc[i]);
    }
  } else {
    // =========================== V-sum warps (6..9) ===========================
    // per tile, warp vs sums rows [32 vs, 32 vs + 32): sum_i a[i] * V[i][0..288), lanes over float4 columns
    const int vg = VG == 2 ? (warp - 6) >> 2 : 0;         // V group (VG == 2: = weight buffer = parity of the CTA's ray counter)
    const int vs = (warp - 6) & 3;
    const int vt = vs * 32 + lane;
    float *part = part_all + vg * (4 * TT * CAR_C_LAT);
    float *vred = vred_all + vg * 32;
    const bool rec = p.stats && blockIdx.x == 0 && vs == 0 && vg == 0;
    unsigned long long vacc[3] = {0, 0, 0};              // 0 wait weights 1 loads+fma 2 reduce+store
    const int l2 = lane < 8 ? 64 + lane : lane;          // third float4 column only exists for lanes 0..7
    const float m2 = lane < 8 ? 1.f : 0.f;
    for (int ir = vg; ir < my_rays; ir += VG) {
      const int ray = (int)blockIdx.x + ir * (int)gridDim.x;
      const uint32_t buf = (uint32_t)ir & 1;           // VG == 2: == vg
      const int g = p.g0 + ray, scene = g / p.a.R, rr = g - scene * p.a.R;
      float w0 = 0.f, w1 = 0.f, w2 = 0.f;               // phase A: this warp's share of sum a * clamp(pt)
      long long tv = rec ? clock64() : 0;
#pragma unroll
      for (int t = 0; t < TT; ++t) {
        const int tile = ray * TT + t;
        const float *V = p.value + ((size_t)tile * 128 + vs * 32) * CAR_C_LAT;
        const float *aw = arow + buf * (128 * TT) + t * 128 + vs * 32;
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0, acc2 = acc0;
        float4 v0[2][4], v1[2][4], v2[2][4];
        auto loadb = [&](int b, int s_) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 *vr = reinterpret_cast<const float4 *>(V + (size_t)(b * 4 + i) * CAR_C_LAT);
            v0[s_][i] = __ldg(vr + lane); v1[s_][i] = __ldg(vr + 32 + lane); v2[s_][i] = __ldg(vr + l2);
          }
        };
        loadb(0, 0);
        loadb(1, 1);
        float ptx = 0.f, pty = 0.f, ptz = 0.f;          // clamp(pt) of row vt (phase A: expected depth)
        if (PHASE == 0) {
          const float *Gp = p.geom + ((size_t)tile * 128 + vt) * CAR_GEOM_STRIDE + G_PTC;
          ptx = __ldg(Gp); pty = __ldg(Gp + 1); ptz = __ldg(Gp + 2);
        }
        if (t == 0) {
          mbar_wait(&a_full[buf], ((uint32_t)ir >> 1) & 1);
          if (rec) { vacc[0] += (unsigned long long)(clock64() - tv); tv = clock64(); }
        }
        if (PHASE == 0) {
          // outputs that only need the posted weights: at_wt, per-context argmax, expected 3-D point
          const float awr = aw[lane];
          const int lr = t * 128 + vt, ctx = lr / P, kk = lr - ctx * P;   // row of the ray -> (context, sample)
          p.a.at_wt[((size_t)(scene * 2 + ctx) * p.a.R + rr) * P + kk] = awr;
          float bv = awr; int bi = kk;                   // first maximum of this 32-row chunk
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
          }
          w0 += warp_sum(awr * ptx); w1 += warp_sum(awr * pty); w2 += warp_sum(awr * ptz);   // models.py:577-582
          if (lane == 0) { vred[t * 4 + vs] = bv; reinterpret_cast<int *>(vred)[8 + t * 4 + vs] = bi; }
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const float4 a4 = *reinterpret_cast<const float4 *>(aw + b * 4);
          const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = av[i], a2 = a * m2;
            const float4 x0 = v0[b & 1][i], x1 = v1[b & 1][i], x2 = v2[b & 1][i];
            acc0.x = fmaf(a, x0.x, acc0.x); acc0.y = fmaf(a, x0.y, acc0.y); acc0.z = fmaf(a, x0.z, acc0.z); acc0.w = fmaf(a, x0.w, acc0.w);
            acc1.x = fmaf(a, x1.x, acc1.x); acc1.y = fmaf(a, x1.y, acc1.y); acc1.z = fmaf(a, x1.z, acc1.z); acc1.w = fmaf(a, x1.w, acc1.w);
            acc2.x = fmaf(a2, x2.x, acc2.x); acc2.y = fmaf(a2, x2.y, acc2.y); acc2.z = fmaf(a2, x2.z, acc2.z); acc2.w = fmaf(a2, x2.w, acc2.w);
          }
          if (b + 2 < 8) loadb(b + 2, b & 1);
        }
        float4 *pp = reinterpret_cast<float4 *>(part + (t * 4 + vs) * CAR_C_LAT);   // 32-row chunk t*4+vs of the ray
        pp[lane] = acc0; pp[32 + lane] = acc1;
        if (lane < 8) pp[64 + lane] = acc2;
      }
      if (PHASE == 0 && lane == 0) { vred[16 + vs * 3] = w0; vred[17 + vs * 3] = w1; vred[18 + vs * 3] = w2; }
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_empty[buf]);        // weights of this buffer consumed
      if (rec) { vacc[1] += (unsigned long long)(clock64() - tv); tv = clock64(); }
      if (vg == 0) asm volatile("bar.sync 3, 128;" ::: "memory"); else asm volatile("bar.sync 4, 128;" ::: "memory");
      if (PHASE == 0 && vt < 2) {
        // context c = rows [c P, (c+1) P) = chunks [2 TT c, 2 TT (c+1)) in row order: first maximum wins
        const int c = vt;
        float m0 = vred[2 * TT * c];
        int i0 = reinterpret_cast<int *>(vred)[8 + 2 * TT * c];
#pragma unroll
        for (int q = 1; q < 2 * TT; ++q) {
          const float m1 = vred[2 * TT * c + q];
          if (m1 > m0) { m0 = m1; i0 = reinterpret_cast<int *>(vred)[8 + 2 * TT * c + q]; }
        }
        p.a.at_wt_max[(size_t)(scene * 2 + c) * p.a.R + rr] = i0;
      }
      if (PHASE == 0 && vt == 2) {
        const float x = (vred[16] + vred[19]) + (vred[22] + vred[25]);
        const float y = (vred[17] + vred[20]) + (vred[23] + vred[26]);
        const float z = (vred[18] + vred[21]) + (vred[24] + vred[27]);
        const float *qi = p.a.cams.qinv + (size_t)scene * 16;
        const float zc = ((qi[8] * x + qi[9] * y) + qi[10] * z) + qi[11];
        p.a.depth_ray[(size_t)scene * p.a.R + rr] = fminf(fmaxf(zc, 0.f), 10.f);
      }
      for (int c = vt; c < CAR_C_LAT; c += 128) {
        float z0, z1;                                    // per-context sums: chunks [0, 2 TT) and [2 TT, 4 TT)
        if (TT == 1) {
          z0 = part[0 * CAR_C_LAT + c] + part[1 * CAR_C_LAT + c];
          z1 = part[2 * CAR_C_LAT + c] + part[3 * CAR_C_LAT + c];
        } else {
          z0 = (part[0 * CAR_C_LAT + c] + part[1 * CAR_C_LAT + c]) + (part[2 * CAR_C_LAT + c] + part[3 * CAR_C_LAT + c]);
          z1 = (part[4 * CAR_C_LAT + c] + part[5 * CAR_C_LAT + c]) + (part[6 * CAR_C_LAT + c] + part[7 * CAR_C_LAT + c]);
        }
        if (PHASE == 0) {
          const float zz = z0 + z1;
          p.zsum[(size_t)ray * CAR_C_LAT + c] = zz;                              // models.py:537-540
          if (p.zs_hi) {
            const __nv_bfloat16 hb = __float2bfloat16_rn(zz);
            p.zs_hi[(size_t)ray * CAR_C_LAT + c] = __bfloat16_as_ushort(hb);
            if (SPLIT == 3) p.zs_lo[(size_t)ray * CAR_C_LAT + c] = __bfloat16_as_ushort(__float2bfloat16_rn(zz - __bfloat162float(hb)));
          }
        } else {
          const float zs = p.zsum[(size_t)ray * CAR_C_LAT + c];
          const float zz = (z0 + zs) + (z1 + zs);                                // models.py:561-564
          p.zfin[(size_t)ray * CAR_C_LAT + c] = zz;
          if (p.zs_hi) {                                   // bf16 hi (+lo) copy: z operand of the fused colour MLP
            const __nv_bfloat16 hb = __float2bfloat16_rn(zz);
            p.zs_hi[(size_t)ray * CAR_C_LAT + c] = __bfloat16_as_ushort(hb);
            if (SPLIT == 3) p.zs_lo[(size_t)ray * CAR_C_LAT + c] = __bfloat16_as_ushort(__float2bfloat16_rn(zz - __bfloat162float(hb)));
          }
        }
      }
      if (vg == 0) asm volatile("bar.sync 3, 128;" ::: "memory"); else asm volatile("bar.sync 4, 128;" ::: "memory");    // part[], vred[] reused by the next ray
      if (rec) vacc[2] += (unsigned long long)(clock64() - tv);
    }
    if (rec && lane == 0)
      for (int i = 0; i < 3; ++i) atomicAdd(p.stats + 8 + i, vacc[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 256);
  }
}

}  // namespace

// phase 0: needs kh (relu(key_map) hi/lo), writes q1, zsum, at_wt, at_wt_max, depth_ray
// phase 1: needs rowbias, q1, zsum; writes zfin
int launch_tail(const car_render_args &a, int phase, int g0, int g1, const float *geom, const float *value,
                const uint16_t *kh_hi, const uint16_t *kh_lo, float *q1, float *zsum, const float *rowbias,
                float *zfin, uint16_t *zs_hi, uint16_t *zs_lo, cudaStream_t st) {
  const int split3 = a.precision == CAR_PREC_FP32_3XBF16;
  const car_weights &W = a.weights;
  if (a.P != 64 && a.P != 128) { set_error("tail kernel needs P == 64 or 128"); return -30; }
  const int TT = a.P / 64;
  const int nrays = g1 - g0;
  const car_mat &w0 = W.key2, &w1 = phase == 0 ? W.qry1 : W.rep1_loc, &w2 = phase == 0 ? W.qry2 : W.rep2;
  CUtensorMap tk_h, tk_l, t0h, t0l, t1h, t1l, t2h, t2l;
  int rc;
  auto mk = [&](CUtensorMap *th, CUtensorMap *tl, const uint16_t *hi, const uint16_t *lo, int rows, int K) {
    if ((rc = make_tmap_bf16(th, hi, rows, K, K, 128, 64))) return rc;
    if (split3) { if ((rc = make_tmap_bf16(tl, lo, rows, K, K, 128, 64))) return rc; } else *tl = *th;
    return 0;
  };
  if (phase == 0) { if (mk(&tk_h, &tk_l, kh_hi, kh_lo, nrays * 128 * TT, 128)) return rc; }
  if (mk(&t0h, &t0l, w0.hi, w0.lo, 128, 128)) return rc;
  if (mk(&t1h, &t1l, w1.hi, w1.lo, 128, 16)) return rc;
  if (mk(&t2h, &t2l, w2.hi, w2.lo, 128, 128)) return rc;
  if (phase != 0) { tk_h = t0h; tk_l = t0l; }
  TailParams p;
  p.a = a; p.g0 = g0; p.g1 = g1; p.geom = geom; p.value = value; p.q1 = q1; p.zsum = zsum;
  p.rowbias = rowbias; p.zfin = zfin; p.zs_hi = zs_hi; p.zs_lo = zs_lo;
  p.stats = g_fused_stats ? g_fused_stats + 32 + phase * 16 : nullptr;
  p.bias_k2 = W.key2.bias; p.bias_q1 = W.qry1.bias; p.bias_q2 = W.qry2.bias; p.bias_r2 = W.rep2.bias;
  const int ops = split3 ? 2 : 1;
  const size_t tile = 2 * 128 * 128 * ops, bstage = 128 * 128 * ops;
  const size_t fixed = (phase == 0 ? 2 * tile : tile + 128 * 128 * 4) + (256 * TT + 2 * 4 * TT * CAR_C_LAT + 32 + 64 + 3 * 128) * 4 + (12 + 2 * NBMAX) * 8 + 16 + 1024;
  int nb = (int)((227 * 1024 - fixed) / bstage);
  const int per_tile = phase == 0 ? 5 : 3;              // weight K-blocks a tile consumes
  if (nb > NBMAX) nb = NBMAX;
  if (nb < 2) { set_error("tail: not enough shared memory"); return -31; }
  p.resident = nb >= per_tile;
  if (p.resident) nb = per_tile;                         // slot = block index
  p.nb = nb;
  const size_t smem = fixed + (size_t)nb * bstage;
  const int sms = sm_count();
  const int grid = nrays < sms ? nrays : sms;
  cudaError_t e = cudaSuccess;
  prof_pre(CAR_ST_ATTENTION, st);
#define CAR_TAIL(S, PH, T)                                                                              \
  do {                                                                                                  \
    e = cudaFuncSetAttribute(k_tail<S, PH, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    if (e == cudaSuccess)                                                                               \
      k_tail<S, PH, T><<<grid, TailShape<S>::THREADS, smem, st>>>(tk_h, tk_l, t0h, t0l, t1h, t1l, t2h, t2l, p); \
  } while (0)
#define CAR_TAIL_T(S, PH) do { if (TT == 1) CAR_TAIL(S, PH, 1); else CAR_TAIL(S, PH, 2); } while (0)
  if (split3) { if (phase == 0) CAR_TAIL_T(3, 0); else CAR_TAIL_T(3, 1); }
  else { if (phase == 0) CAR_TAIL_T(1, 0); else CAR_TAIL_T(1, 1); }
#undef CAR_TAIL_T
#undef CAR_TAIL
  prof_post(st);
  if (e != cudaSuccess) { set_error("tail: %s", cudaGetErrorString(e)); return (int)e; }
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("tail launch: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  return 0;
}

}  // namespace car
