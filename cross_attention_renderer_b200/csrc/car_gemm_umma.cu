// tcgen05 GEMM for the per-sample MLP stack (sm_100a only).
//
//   C[M][N] = act( A[M][K] · W[N][K]^T + bias[N] + row_bias[m / rows_per_group][N] )
//
// A and W are bf16, K-major (row-major [rows][K]).  SPLIT == 3 is the fp32-equivalent mode:
// both operands are given as hi + lo bf16 halves and every logical product is three MMAs
// accumulated in the same fp32 TMEM tile:  Ah·Wh + Al·Wh + Ah·Wl  (error ~2^-17 relative).
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D tiles (128B swizzle) into a smem ring
//   warp 1      MMA issuer (one elected lane): tcgen05.mma cta_group::1 kind::f16, M=128,
//               N=BN, K=16 per instruction; accumulators in TMEM, double buffered
//   warps 2..5  epilogue: tcgen05.ld 32x32b -> bias/ReLU -> fp32 and/or bf16 hi(+lo) stores
// Pipelines: smem full/empty mbarriers (TMA <-> MMA) and TMEM full/empty mbarriers
// (MMA <-> epilogue), so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Backward-pass forms (car_backward.cu):
//   data gradient    dA = dY · W        A operand = dY, "W" operand = W^T; the epilogue applies the ReLU
//                                       subgradient mask (a saved fp32 activation) and / or accumulates
//   weight gradient  dW += dY^T · A     both operands given transposed ([N][M] and [K][M]: the long row
//                                       dimension becomes the contraction), split along it over `splits`
//                                       work items per output tile, partial tiles added with fp32 atomics
#include <cuda.h>
#include <cuda_bf16.h>

#include "car_common.cuh"

namespace car {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // bf16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int MAX_STAGES = 8;

struct UmmaParams {
  int M, N, K, BN, stages, tmem_cols;
  const float *bias, *row_bias;
  int rows_per_group, relu;            // relu: 0 none, 1 on every output, 2 only on the bf16 operand copy
  const float *add_src;                // optional fp32 [M][ldc] added before activation (may alias out_f32)
  const float *mask;                   // optional fp32 [M][ldmask]: result = 0 where mask <= 0 (not together with add_src)
  int ldmask;
  int splits, kb_per;                  // split along K: work item = (tile, split); kb_per K blocks per split
  int atomic;                          // out_f32 += result with atomics (split-K partial tiles)
  float *out_f32;
  uint16_t *out_hi, *out_lo;
  int ldc;
};

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tm, uint64_t *bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (ignored for swizzled K-major, set to 1) |
//   [32,46) SBO >> 4 = 1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n
__device__ __forceinline__ uint32_t make_idesc(int n) {
  uint32_t d = 0;
  d |= 1u << 4;                 // c_format = F32
  d |= 1u << 7;                 // a_format = BF16
  d |= 1u << 10;                // b_format = BF16
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(BM >> 4) << 24;
  return d;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b, float &ra, float &rb) {
  __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
  ra = a - __bfloat162float(ha);
  rb = b - __bfloat162float(hb);
  return (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
}

template <int SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_gemm_umma(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
            const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
            UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages] x {A_hi, (A_lo), W_hi, (W_lo)}, then barriers
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_bytes = BM * BK * 2;                 // 16 KB
  const int w_bytes = p.BN * BK * 2;
  const int stage_bytes = (a_bytes + w_bytes) * (SPLIT == 3 ? 2 : 1);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)p.stages * stage_bytes);
  uint64_t *empty = full + MAX_STAGES;
  uint64_t *tfull = empty + MAX_STAGES;            // [2]
  uint64_t *tempty = tfull + 2;                    // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_n = p.N / p.BN;
  const int num_m = (p.M + BM - 1) / BM;
  const int num_tiles = num_m * num_n * p.splits;          // work items: split index fastest
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_a_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_w_hi)) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int acc_stride = p.tmem_cols / 2;          // columns per accumulator buffer

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int tile = t / p.splits, kb0 = (t % p.splits) * p.kb_per, kb1 = min(num_kb, kb0 + p.kb_per);
        int m0 = (tile / num_n) * BM, n0 = (tile % num_n) * p.BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t *st = smem + (size_t)stage * stage_bytes;
          mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
          tma_load_2d(st, &tm_a_hi, &full[stage], kb * BK, m0);
          if (SPLIT == 3) tma_load_2d(st + a_bytes, &tm_a_lo, &full[stage], kb * BK, m0);
          uint8_t *sw = st + (SPLIT == 3 ? 2 : 1) * a_bytes;
          tma_load_2d(sw, &tm_w_hi, &full[stage], kb * BK, n0);
          if (SPLIT == 3) tma_load_2d(sw + w_bytes, &tm_w_lo, &full[stage], kb * BK, n0);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);          // epilogue drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * acc_stride);
        const int kb0 = (t % p.splits) * p.kb_per, kb1 = min(num_kb, kb0 + p.kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sa_lo = sa + a_bytes;
          const uint32_t sw = sa + (SPLIT == 3 ? 2 : 1) * a_bytes;
          const uint32_t sw_lo = sw + w_bytes;
          int ksteps = (p.K - kb * BK + UMMA_K - 1) / UMMA_K;
          if (ksteps > BK / UMMA_K) ksteps = BK / UMMA_K;
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t koff = (uint32_t)k * UMMA_K * 2;     // bytes inside the 128-byte swizzle row
            const uint32_t first = ((kb - kb0) | k) ? 1u : 0u;
            umma_f16(tmem_d, make_desc_sw128(sa + koff), make_desc_sw128(sw + koff), idesc, first);
            if (SPLIT == 3) {
              umma_f16(tmem_d, make_desc_sw128(sa_lo + koff), make_desc_sw128(sw + koff), idesc, 1u);
              umma_f16(tmem_d, make_desc_sw128(sa + koff), make_desc_sw128(sw_lo + koff), idesc, 1u);
            }
          }
          umma_commit(&empty[stage]);                    // frees the smem slot when the MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);                        // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue (warps 2..5) =================
    const int sub = warp & 3;                            // TMEM sub-partition this warp may read
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int tile = t / p.splits;
      int m0 = (tile / num_n) * BM, n0 = (tile % num_n) * p.BN;
      mbar_wait(&tfull[acc], acc_phase);
      tcgen05_fence_after();
      const int row = m0 + sub * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * acc_stride);
      const float *rb = (p.row_bias && row < p.M) ? p.row_bias + (size_t)(row / p.rows_per_group) * p.N : nullptr;
      // 32 accumulator columns per step: one 128-byte line of fp32 (or 64 B of bf16) per lane
      // The residual / accumulate source of a 32-column block (fp32, one 128-byte line per lane) is fetched
      // in one batch BEFORE the block's math (and before the TMEM wait of the caller): consumed load by
      // load it cost one global round trip per 8 columns, 16 per tile, and dominated the small GEMMs.
      auto fetch_add = [&](float4 (&av)[8], int c0, int cnt) {
        if ((!p.add_src && !p.mask) || row >= p.M) return;
        const float4 *src = p.add_src ? reinterpret_cast<const float4 *>(p.add_src + (size_t)row * p.ldc + n0 + c0)
                                      : reinterpret_cast<const float4 *>(p.mask + (size_t)row * p.ldmask + n0 + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i * 4 < cnt) av[i] = src[i];
      };
      auto emit = [&](const uint32_t *r, const float4 (&av)[8], int c0, int cnt) {
        if (row >= p.M) return;
#pragma unroll
        for (int i0 = 0; i0 < 32; i0 += 8) {
          if (i0 >= cnt) break;
          const int n = n0 + c0 + i0;
          const size_t o = (size_t)row * p.ldc + n;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i0 + i]);
          if (p.bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4 *>(p.bias + n)), b1 = __ldg(reinterpret_cast<const float4 *>(p.bias + n + 4));
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          }
          if (rb) {
            const float4 b0 = __ldg(reinterpret_cast<const float4 *>(rb + n)), b1 = __ldg(reinterpret_cast<const float4 *>(rb + n + 4));
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          }
          if (p.add_src) {
            const float4 a0 = av[i0 / 4], a1 = av[i0 / 4 + 1];
            v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w;
            v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
          }
          if (p.relu == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (p.mask && !p.add_src) {
            const float4 a0 = av[i0 / 4], a1 = av[i0 / 4 + 1];
            const float mk[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = mk[i] > 0.f ? v[i] : 0.f;
          }
          if (p.atomic) {
#pragma unroll
            for (int i = 0; i < 8; ++i) atomicAdd(p.out_f32 + o + i, v[i]);
          } else if (p.out_f32) {
            *reinterpret_cast<float4 *>(p.out_f32 + o) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4 *>(p.out_f32 + o + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
          if (p.out_hi) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float x0 = v[2 * i], x1 = v[2 * i + 1];
              if (p.relu == 2) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }   // ReLU only on the operand copy
              float ra, rbb, d0, d1;
              hi[i] = pack_bf16x2(x0, x1, ra, rbb);
              lo[i] = pack_bf16x2(ra, rbb, d0, d1);
            }
            *reinterpret_cast<uint4 *>(p.out_hi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (p.out_lo) *reinterpret_cast<uint4 *>(p.out_lo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      };
      {
        uint32_t r0[32], r1[32];
        float4 a0[8], a1[8];
        const int n32 = p.BN / 32;
        if (n32 > 0) { tmem_ld32(taddr, r0); fetch_add(a0, 0, 32); }
        for (int j = 0; j < n32; j += 2) {
          tmem_ld_wait();
          if (j + 1 < n32) { tmem_ld32(taddr + (uint32_t)((j + 1) * 32), r1); fetch_add(a1, (j + 1) * 32, 32); }
          emit(r0, a0, j * 32, 32);
          if (j + 1 < n32) {
            tmem_ld_wait();
            if (j + 2 < n32) { tmem_ld32(taddr + (uint32_t)((j + 2) * 32), r0); fetch_add(a0, (j + 2) * 32, 32); }
            emit(r1, a1, (j + 1) * 32, 32);
          }
        }
        if (p.BN & 16) {
          uint32_t rt[16];
          tmem_ld16(taddr + (uint32_t)(n32 * 32), rt);
          fetch_add(a0, n32 * 32, 16);
          tmem_ld_wait();
          uint32_t rt32[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) rt32[i] = rt[i];
          emit(rt32, a0, n32 * 32, 16);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);          // 4 arrivals (one per epilogue warp)
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D bf16 tensor [rows][K] with row pitch ld elements; box = [box_rows][64], 128B swizzle
int make_map(CUtensorMap *tm, const uint16_t *base, int rows, int K, int ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return -10; }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d (rows=%d K=%d ld=%d box=%d)", (int)r, rows, K, ld, box_rows); return -11; }
  return 0;
}

int pick_bn(int N) {
  if (N % 192 == 0) return 192;
  if (N % 144 == 0) return 144;
  if (N % 128 == 0) return 128;
  if (N <= 256 && N % 16 == 0) return N;
  return 0;
}

}  // namespace

int launch_gemm_umma(const uint16_t *a_hi, const uint16_t *a_lo, int lda, const uint16_t *w_hi,
                     const uint16_t *w_lo, int ldw, int M, int N, int K, int split3, const GemmEpi &epi,
                     const UmmaOut &out, cudaStream_t st) {
  if (M <= 0) return 0;
  int BN = pick_bn(N);
  if (!BN || K % 16 || (lda % 8) || (ldw % 8)) { set_error("gemm_umma: unsupported shape M=%d N=%d K=%d lda=%d", M, N, K, lda); return -12; }
  if (split3 && (!a_lo || !w_lo)) { set_error("gemm_umma: split3 needs lo operands"); return -13; }
  CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
  int rc;
  if ((rc = make_map(&ta_hi, a_hi, M, K, lda, BM))) return rc;
  if ((rc = make_map(&tw_hi, w_hi, N, K, ldw, BN))) return rc;
  if (split3) {
    if ((rc = make_map(&ta_lo, a_lo, M, K, lda, BM))) return rc;
    if ((rc = make_map(&tw_lo, w_lo, N, K, ldw, BN))) return rc;
  } else { ta_lo = ta_hi; tw_lo = tw_hi; }
  UmmaParams p;
  p.M = M; p.N = N; p.K = K; p.BN = BN;
  int stage_bytes = (BM * BK * 2 + BN * BK * 2) * (split3 ? 2 : 1);
  int budget = 220 * 1024 - 1024 - 256;
  p.stages = budget / stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  if (p.stages < 2) { set_error("gemm_umma: tile too large"); return -14; }
  p.tmem_cols = BN <= 64 ? 128 : (BN <= 128 ? 256 : 512);
  p.bias = epi.bias; p.row_bias = epi.row_bias; p.rows_per_group = epi.rows_per_group > 0 ? epi.rows_per_group : 1;
  p.relu = epi.relu_out;
  p.add_src = epi.accumulate ? out.f32_add : nullptr;
  p.mask = epi.mask; p.ldmask = epi.ldmask;
  if (p.mask && p.add_src) { set_error("gemm_umma: mask and accumulate cannot be combined"); return -15; }
  if (p.mask && (epi.ldmask % 4)) { set_error("gemm_umma: ldmask %% 4"); return -15; }
  // split along K (weight gradients: few output tiles, very long contraction): about two work items per SM
  const int num_kb = (K + BK - 1) / BK;
  const int out_tiles = ((M + BM - 1) / BM) * (N / BN);
  int splits = 1;
  if (out.atomic) {
    splits = (2 * sm_count() + out_tiles - 1) / out_tiles;
    if (splits > (num_kb + 3) / 4) splits = (num_kb + 3) / 4;     // at least 4 K blocks per item
    if (splits < 1) splits = 1;
    if (!out.f32 || out.hi || epi.bias || epi.row_bias || epi.relu_out || p.mask || p.add_src) {
      set_error("gemm_umma: the atomic (split-K) form writes fp32 sums only"); return -15;
    }
  }
  p.kb_per = (num_kb + splits - 1) / splits;
  p.splits = (num_kb + p.kb_per - 1) / p.kb_per;          // no empty split
  p.atomic = out.atomic;
  p.out_f32 = out.f32; p.out_hi = out.hi; p.out_lo = out.lo; p.ldc = out.ldc;
  size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256;
  const int sms = sm_count();
  int tiles = out_tiles * p.splits;
  int grid = tiles < sms ? tiles : sms;
  cudaError_t e;
  prof_pre(-1, st);
  if (split3) {
    e = cudaFuncSetAttribute(k_gemm_umma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) k_gemm_umma<3><<<grid, NUM_THREADS, smem, st>>>(ta_hi, ta_lo, tw_hi, tw_lo, p);
  } else {
    e = cudaFuncSetAttribute(k_gemm_umma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) k_gemm_umma<1><<<grid, NUM_THREADS, smem, st>>>(ta_hi, ta_lo, tw_hi, tw_lo, p);
  }
  prof_post(st);
  if (e != cudaSuccess) { set_error("gemm_umma: %s", cudaGetErrorString(e)); return (int)e; }
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("gemm_umma launch: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  return 0;
}

}  // namespace car

extern "C" int car_gemm_umma_test(const uint16_t *a_hi, const uint16_t *a_lo, const uint16_t *w_hi,
                                  const uint16_t *w_lo, const float *bias, float *c, int M, int N, int K,
                                  int split3, int relu, void *stream) {
  car::GemmEpi e;
  e.bias = bias; e.row_bias = nullptr; e.rows_per_group = 1; e.relu_in = 0; e.relu_out = relu; e.accumulate = 0;
  car::UmmaOut o;
  o.f32 = c; o.hi = nullptr; o.lo = nullptr; o.ldc = N;
  return car::launch_gemm_umma(a_hi, a_lo, K, w_hi, w_lo, K, M, N, K, split3, e, o, (cudaStream_t)stream);
}
