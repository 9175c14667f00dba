// Colour MLP phi + mask / white fill for one ray chunk in ONE persistent tcgen05 kernel
// (reference models.py:597-621 -> resnet_block_fc.py:132-168, ResnetFC with 3 blocks, d_hidden = 128):
//
//   x = lin_in(coords18);  for i in 0..2:  x += lin_z[i](z);  x += fc_1[i](relu(fc_0[i](relu(x))));  rgb = lin_out(relu(x))
//
// One CTA = tiles of 128 rays (UMMA M = 128, cta_group::1, N = 128).  x lives in a TMEM accumulator for the whole
// chain: lin_in writes it, every lin_z / fc_1 GEMM accumulates onto it (the biases are added when it is drained), so
// the fp32 state never leaves the SM.  The ten weight matrices are packed side by side along K into one
// [128][1792] matrix (car_weights::phi_pack: lin_in 64 | lin_z0 320 | fc_0 128 | fc_1 128 | lin_z1 320 | ...,
// zero padded) and stream through a TMA ring as 28 consecutive 64-wide K blocks per tile; the z operand
// (bf16 hi / lo written by the attention tail) streams through a second ring.
//   warp 0   TMA producer        warp 1   MMA issuer
//   warps 2-5  one thread per ray: coords -> bf16 operand; TMEM -> bias / ReLU -> bf16 hi(+lo) operand tile
//              (128-byte swizzle) of the next GEMM; final: lin_out (N = 3) as three dot products, mask, white fill.
// Replaces 10 k_gemm_umma launches + operand-split and SIMT lin_out launches per chunk (11.8 -> ~1.5 ms per step).
#include <math.h>

#include "car_common.cuh"
#include "car_umma.cuh"

namespace car {
int make_tmap_bf16(CUtensorMap *tm, const uint16_t *base, int rows, int K, int ld, int box_rows, int box_k);

namespace {
using namespace ptx;

constexpr int THREADS = 192;
constexpr int NBLK = 28;               // weight K blocks per tile
constexpr int ZBLK = 5;                // z K blocks per lin_z GEMM (288 -> 320)
constexpr int NB = 2, NZ = 2;          // ring depths

struct PhiParams {
  car_render_args a;
  int g0, g1;
  const uint8_t *overlap;              // (rays, 2)
  const float *b_in, *b_z[3], *b_fc0[3], *b_fc1[3];
  const float *w_out, *b_out;          // lin_out [3][128] fp32, [3]
};

template <int SPLIT> struct PCfg {
  static constexpr int OPS = SPLIT == 3 ? 2 : 1;
  static constexpr int KB_BYTES = 128 * 128;                  // one K block (64 bf16) of 128 rows
  static constexpr int TILE_HALF = 2 * KB_BYTES;              // 128 x 128 bf16
  static constexpr int TILE = TILE_HALF * OPS;
  static constexpr int STAGE = KB_BYTES * OPS;                // ring stage: one K block hi (+lo)
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

template <int SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
k_phi(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
      const __grid_constant__ CUtensorMap tm_z_hi, const __grid_constant__ CUtensorMap tm_z_lo, PhiParams p) {
  using C = PCfg<SPLIT>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *at = smem;                                         // operand tile [128 rows][128 K] hi (+lo)
  uint8_t *bs = at + C::TILE;                                 // NB weight stages
  uint8_t *zs = bs + NB * C::STAGE;                           // NZ z stages
  float *sb = reinterpret_cast<float *>(zs + NZ * C::STAGE);  // [4][128] x-bias partial sums, [3][128] fc_0 biases
  float *swo = sb + 7 * 128;                                  // lin_out [3][128] + bias [3] (+1 pad)
  uint64_t *bars = reinterpret_cast<uint64_t *>(swo + 3 * 128 + 4);
  uint64_t *b_full = bars, *b_empty = b_full + NB, *z_full = b_empty + NB, *z_empty = z_full + NZ;
  uint64_t *at_full = z_empty + NZ, *x_full = at_full + 1, *n_full = x_full + 1, *x_empty = n_full + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(x_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nrays = p.g1 - p.g0;
  const int ntiles = (nrays + 127) / 128;
  const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_w_hi);
    prefetch_tmap(&tm_z_hi);
    for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < NZ; ++s) { mbar_init(&z_full[s], 1); mbar_init(&z_empty[s], 1); }
    mbar_init(at_full, 4); mbar_init(x_full, 1); mbar_init(n_full, 1); mbar_init(x_empty, 4);
    fence_barrier_init();
  }
  if (threadIdx.x >= 64) {
    // bias of x at its four drains: b_in + b_z0 | + b_fc1_0 + b_z1 | + b_fc1_1 + b_z2 | + b_fc1_2
    const int c = threadIdx.x - 64;
    float s0 = p.b_in[c] + p.b_z[0][c];
    sb[c] = s0;
    s0 = (s0 + p.b_fc1[0][c]) + p.b_z[1][c]; sb[128 + c] = s0;
    s0 = (s0 + p.b_fc1[1][c]) + p.b_z[2][c]; sb[256 + c] = s0;
    s0 = s0 + p.b_fc1[2][c]; sb[384 + c] = s0;
    for (int i = 0; i < 3; ++i) sb[512 + i * 128 + c] = p.b_fc0[i][c];
    for (int i = 0; i < 3; ++i) swo[i * 128 + c] = p.w_out[i * 128 + c];
    if (c < 3) swo[384 + c] = p.b_out[c];
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t ACCX = 0, ACCN = 128;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      uint32_t bq = 0, zq = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int r0 = ((int)blockIdx.x + it * (int)gridDim.x) * 128;
        for (int blk = 0; blk < NBLK; ++blk) {
          // z blocks accompany the lin_z weight blocks: packed block index 1 + 9 i + kb, kb < 5
          const int rel = blk - 1, ii = rel >= 0 ? rel / 9 : -1, kb = rel >= 0 ? rel - ii * 9 : -1;
          if (rel >= 0 && kb < ZBLK) {
            const int s = zq % NZ;
            mbar_wait(&z_empty[s], ((zq / NZ) & 1) ^ 1);
            mbar_expect_tx(&z_full[s], (uint32_t)C::STAGE);
            tma_load_2d(zs + (size_t)s * C::STAGE, &tm_z_hi, &z_full[s], kb * 64, r0);
            if (SPLIT == 3) tma_load_2d(zs + (size_t)s * C::STAGE + C::KB_BYTES, &tm_z_lo, &z_full[s], kb * 64, r0);
            ++zq;
          }
          const int s = bq % NB;
          mbar_wait(&b_empty[s], ((bq / NB) & 1) ^ 1);
          mbar_expect_tx(&b_full[s], (uint32_t)C::STAGE);
          tma_load_2d(bs + (size_t)s * C::STAGE, &tm_w_hi, &b_full[s], blk * 64, 0);
          if (SPLIT == 3) tma_load_2d(bs + (size_t)s * C::STAGE + C::KB_BYTES, &tm_w_lo, &b_full[s], blk * 64, 0);
          ++bq;
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc_bf16(128, 128);
    uint32_t bq = 0, zq = 0, atq = 0;
    // one 64-wide K block: A at a_addr (lo copy a_lo bytes further), B = next weight stage
    auto gemm_kb = [&](uint32_t d, uint32_t a_addr, uint32_t a_lo, bool first, int zslot) {
      const int s = bq % NB;
      mbar_wait(&b_full[s], (bq / NB) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = make_desc<128>(a_addr);
        const uint64_t db = make_desc<128>(smem_u32(bs + (size_t)s * C::STAGE));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = (first && k == 0) ? 0u : 1u;
          const uint64_t a = da + (uint64_t)((k * 32) >> 4), w = db + (uint64_t)((k * 32) >> 4);
          umma_f16<1>(d, a, w, idesc, acc);
          if (SPLIT == 3) {
            umma_f16<1>(d, a + (uint64_t)(a_lo >> 4), w, idesc, 1u);
            umma_f16<1>(d, a, w + (uint64_t)(C::KB_BYTES >> 4), idesc, 1u);
          }
        }
        umma_commit(&b_empty[s]);
        if (zslot >= 0) umma_commit(&z_empty[zslot]);
      }
      __syncwarp();
      ++bq;
    };
    auto wait_at = [&]() { mbar_wait(at_full, atq & 1); ++atq; tc_fence_after(); };
    for (int it = 0; it < my_tiles; ++it) {
      mbar_wait(x_empty, (it & 1) ^ 1);                    // previous tile's final drain has read x
      tc_fence_after();
      wait_at();                                           // coords operand
      gemm_kb(tmem_base + ACCX, smem_u32(at), C::TILE_HALF, true, -1);          // lin_in (K block 0 of the tile)
      for (int i = 0; i < 3; ++i) {
        for (int kb = 0; kb < ZBLK; ++kb) {                // x += lin_z[i](z)
          const int s = zq % NZ;
          mbar_wait(&z_full[s], (zq / NZ) & 1);
          gemm_kb(tmem_base + ACCX, smem_u32(zs + (size_t)s * C::STAGE), C::KB_BYTES, false, s);
          ++zq;
        }
        if (elect_one()) umma_commit(x_full);
        __syncwarp();
        wait_at();                                         // relu(x) operand
        gemm_kb(tmem_base + ACCN, smem_u32(at), C::TILE_HALF, true, -1);        // fc_0
        gemm_kb(tmem_base + ACCN, smem_u32(at + C::KB_BYTES), C::TILE_HALF, false, -1);
        if (elect_one()) umma_commit(n_full);
        __syncwarp();
        wait_at();                                         // relu(net) operand
        gemm_kb(tmem_base + ACCX, smem_u32(at), C::TILE_HALF, false, -1);       // x += fc_1
        gemm_kb(tmem_base + ACCX, smem_u32(at + C::KB_BYTES), C::TILE_HALF, false, -1);
      }
      if (elect_one()) umma_commit(x_full);                // final x
      __syncwarp();
    }
  } else {
    // =========================== row threads (warps 2..5) ===========================
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(sub * 32) << 16);
    uint32_t xq = 0, nq = 0;
    // acc (TMEM columns) + bias -> ReLU -> bf16 hi(+lo) -> operand tile `at`
    auto drain_to_at = [&](uint32_t acc_col, const float *bias) {
      uint32_t r[32];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        tmem_ld32(tlane + acc_col + (uint32_t)(j * 32), r);
        tmem_ld_wait();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x0 = fmaxf(__uint_as_float(r[2 * i]) + bias[j * 32 + 2 * i], 0.f);
          const float x1 = fmaxf(__uint_as_float(r[2 * i + 1]) + bias[j * 32 + 2 * i + 1], 0.f);
          split2<SPLIT == 3>(x0, x1, hi[i], lo[i]);
        }
        uint8_t *dst = at + (j >> 1) * C::KB_BYTES;
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
      if (lane == 0) mbar_arrive(at_full);
    };
    for (int it = 0; it < my_tiles; ++it) {
      const int gl = ((int)blockIdx.x + it * (int)gridDim.x) * 128 + row;      // ray within the chunk
      const bool live = gl < nrays;
      const int g = p.g0 + (live ? gl : nrays - 1), scene = g / p.a.R, rr = g - scene * p.a.R;
      {
        // coords18 = [d0 m0 o0 | d1 m1 o1] (models.py:597-602) -> K block 0, columns 0..17, zeros up to 63
        float c[24];
#pragma unroll
        for (int q = 0; q < 24; ++q) c[q] = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float *co = p.a.coords + ((size_t)(scene * 2 + j) * p.a.R + rr) * 9;
#pragma unroll
          for (int q = 0; q < 9; ++q) c[j * 9 + q] = co[q];
        }
        // the previous tile's last use of `at` (fc_1 of block 2) retired before its final x_full, which these
        // threads waited for: `at` is free
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint32_t hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
          if (ch < 3) {
#pragma unroll
            for (int q = 0; q < 4; ++q) split2<SPLIT == 3>(c[ch * 8 + 2 * q], c[ch * 8 + 2 * q + 1], hi[q], lo[q]);
          }
          const uint32_t off = swz_offset<128>(row, ch);
          *reinterpret_cast<uint4 *>(at + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (SPLIT == 3) *reinterpret_cast<uint4 *>(at + C::TILE_HALF + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(at_full);
      }
      for (int i = 0; i < 3; ++i) {
        mbar_wait(x_full, xq & 1); ++xq;
        tc_fence_after();
        drain_to_at(ACCX, sb + i * 128);                   // relu(x)
        mbar_wait(n_full, nq & 1); ++nq;
        tc_fence_after();
        drain_to_at(ACCN, sb + 512 + i * 128);             // relu(fc_0(relu(x)))
      }
      // final: rgb = lin_out(relu(x)); valid = any context overlaps; white fill (models.py:611-616)
      mbar_wait(x_full, xq & 1); ++xq;
      tc_fence_after();
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
      {
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          tmem_ld32(tlane + ACCX + (uint32_t)(j * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int cidx = j * 32 + i;
            const float x = fmaxf(__uint_as_float(r[i]) + sb[384 + cidx], 0.f);
            o0 = fmaf(x, swo[cidx], o0); o1 = fmaf(x, swo[128 + cidx], o1); o2 = fmaf(x, swo[256 + cidx], o2);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(x_empty);
      if (live) {
        const float valid = (p.overlap[gl * 2] | p.overlap[gl * 2 + 1]) ? 1.f : 0.f;
        p.a.valid_mask[g] = valid;
        p.a.rgb[(size_t)g * 3 + 0] = (o0 + swo[384]) * valid + 1.f * (1.f - valid);
        p.a.rgb[(size_t)g * 3 + 1] = (o1 + swo[385]) * valid + 1.f * (1.f - valid);
        p.a.rgb[(size_t)g * 3 + 2] = (o2 + swo[386]) * valid + 1.f * (1.f - valid);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 256);
  }
}

}  // namespace

// z_hi / z_lo: (rays, 288) bf16 hi (+lo) of the attention output of rays [g0, g1)
int launch_phi_fused(const car_render_args &a, int g0, int g1, const uint16_t *z_hi, const uint16_t *z_lo,
                     const uint8_t *overlap, cudaStream_t st) {
  const int split3 = a.precision == CAR_PREC_FP32_3XBF16;
  const car_weights &W = a.weights;
  const car_mat &PK = W.phi_pack;
  const int nrays = g1 - g0;
  if (nrays <= 0) return 0;
  if (!PK.hi || PK.N != 128 || PK.K != NBLK * 64) { set_error("phi: weights.phi_pack must be [128][%d]", NBLK * 64); return -40; }
  CUtensorMap twh, twl, tzh, tzl;
  int rc;
  if ((rc = make_tmap_bf16(&twh, PK.hi, 128, PK.K, PK.K, 128, 64))) return rc;
  if ((rc = make_tmap_bf16(&tzh, z_hi, nrays, CAR_C_LAT, CAR_C_LAT, 128, 64))) return rc;
  if (split3) {
    if ((rc = make_tmap_bf16(&twl, PK.lo, 128, PK.K, PK.K, 128, 64))) return rc;
    if ((rc = make_tmap_bf16(&tzl, z_lo, nrays, CAR_C_LAT, CAR_C_LAT, 128, 64))) return rc;
  } else { twl = twh; tzl = tzh; }
  PhiParams p;
  p.a = a; p.g0 = g0; p.g1 = g1; p.overlap = overlap;
  p.b_in = W.phi_in.bias;
  for (int i = 0; i < 3; ++i) { p.b_z[i] = W.phi_z[i].bias; p.b_fc0[i] = W.phi_fc0[i].bias; p.b_fc1[i] = W.phi_fc1[i].bias; }
  p.w_out = W.phi_out.f32; p.b_out = W.phi_out.bias;
  const int ops = split3 ? 2 : 1;
  const size_t smem = (size_t)(2 * 128 * 128 * ops) + (size_t)(NB + NZ) * 128 * 128 * ops + (7 * 128 + 3 * 128 + 4) * 4 +
                      (2 * NB + 2 * NZ + 4) * 8 + 16 + 1024;
  const int ntiles = (nrays + 127) / 128;
  const int sms = sm_count();
  const int grid = ntiles < sms ? ntiles : sms;
  cudaError_t e;
  prof_pre(CAR_ST_PHI, st);
  if (split3) {
    e = cudaFuncSetAttribute(k_phi<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) k_phi<3><<<grid, THREADS, smem, st>>>(twh, twl, tzh, tzl, p);
  } else {
    e = cudaFuncSetAttribute(k_phi<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) k_phi<1><<<grid, THREADS, smem, st>>>(twh, twl, tzh, tzl, p);
  }
  prof_post(st);
  if (e != cudaSuccess) { set_error("phi: %s", cudaGetErrorString(e)); return (int)e; }
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("phi launch: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  return 0;
}

}  // namespace car
