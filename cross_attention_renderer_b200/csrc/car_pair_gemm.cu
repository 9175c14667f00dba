// CTA-pair tcgen05 GEMM (cta_group::2, UMMA M = 128 = 64 rows per CTA).
//
// This is the building block of the fused per-ray kernel: with 64 rows per CTA the fp32
// accumulator of an N-wide tile occupies only N/2 TMEM columns per CTA ("2x2" datapath
// layout: lanes 0-63 hold columns [0,N/2) of rows 0-63, lanes 64-127 hold columns [N/2,N)),
// so a 128-row tile can keep a 576-wide and a 416-wide accumulator resident at once
// (288 + 208 = 496 <= 512 columns), which a single CTA cannot.
//
//   C[M][N] = act(A[M][K] · W[N][K]^T + bias),  N = nch * NCH
//
// Per pair and tile of 128 rows: each CTA TMA-loads its 64 rows of A and its half of every
// NCH-row block of W (rows [c*NCH + rank*NCH/2, +NCH/2)); the leader CTA issues the MMAs for
// both; both epilogues drain their own TMEM.
#include "car_common.cuh"
#include "car_umma.cuh"

namespace car {
namespace {

using namespace ptx;

constexpr int UMMA_K = 16;
constexpr int THREADS = 192;
constexpr int MAX_STAGES = 6;

struct PairParams {
  int M, N, K, NCH, nch, stages, tmem_cols;
  const float *bias;
  int relu;
  float *out;      // [M][N] fp32
  float *dump;     // optional raw TMEM dump: [pairs*2][128 lanes][N/2] (first tile of each pair)
};

template <int SPLIT, int BK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_pair_gemm(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
            const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo, PairParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_bytes = 64 * BK * 2;                       // 64 rows x 128 B
  const int wc_bytes = (p.NCH / 2) * BK * 2;             // this CTA's half of one N-chunk
  const int w_bytes = wc_bytes * p.nch;
  const int stage_bytes = (a_bytes + w_bytes) * (SPLIT == 3 ? 2 : 1);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)p.stages * stage_bytes);
  uint64_t *empty = full + MAX_STAGES;
  uint64_t *tfull = empty + MAX_STAGES;
  uint64_t *tempty = tfull + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int num_tiles = (p.M + 127) / 128;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_a_hi);
    prefetch_tmap(&tm_w_hi);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1);
    mbar_init(tempty, 8);                                // 4 epilogue warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<2>(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (both CTAs load their halves) =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += npairs) {
        const int m0 = t * 128 + (int)rank * 64;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t *st = smem + (size_t)stage * stage_bytes;
          if (leader) mbar_expect_tx(&full[stage], (uint32_t)stage_bytes * 2);   // both CTAs' bytes
          tma_load_2d_pair(st, &tm_a_hi, &full[stage], kb * BK, m0);
          if (SPLIT == 3) tma_load_2d_pair(st + a_bytes, &tm_a_lo, &full[stage], kb * BK, m0);
          uint8_t *sw = st + (SPLIT == 3 ? 2 : 1) * a_bytes;
          for (int c = 0; c < p.nch; ++c) {
            const int n0 = c * p.NCH + (int)rank * (p.NCH / 2);
            tma_load_2d_pair(sw + (size_t)c * wc_bytes, &tm_w_hi, &full[stage], kb * BK, n0);
            if (SPLIT == 3) tma_load_2d_pair(sw + w_bytes + (size_t)c * wc_bytes, &tm_w_lo, &full[stage], kb * BK, n0);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, p.NCH);
      int stage = 0; uint32_t phase = 0, tphase = 0;
      for (int t = pair; t < num_tiles; t += npairs) {
        mbar_wait(tempty, tphase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sa_lo = sa + a_bytes;
          const uint32_t sw = sa + (SPLIT == 3 ? 2 : 1) * a_bytes;
          const uint32_t sw_lo = sw + w_bytes;
          int ksteps = (p.K - kb * BK + UMMA_K - 1) / UMMA_K;
          if (ksteps > BK / UMMA_K) ksteps = BK / UMMA_K;
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t koff = (uint32_t)k * UMMA_K * 2;
            const uint32_t acc = (kb | k) ? 1u : 0u;
            for (int c = 0; c < p.nch; ++c) {
              const uint32_t d = tmem_base + (uint32_t)(c * (p.NCH / 2));
              const uint32_t wb = sw + (uint32_t)c * wc_bytes + koff, wl = sw_lo + (uint32_t)c * wc_bytes + koff;
              umma_f16<2>(d, make_desc<BK * 2>(sa + koff), make_desc<BK * 2>(wb), idesc, acc);
              if (SPLIT == 3) {
                umma_f16<2>(d, make_desc<BK * 2>(sa_lo + koff), make_desc<BK * 2>(wb), idesc, 1u);
                umma_f16<2>(d, make_desc<BK * 2>(sa + koff), make_desc<BK * 2>(wl), idesc, 1u);
              }
            }
          }
          umma_commit_pair(&empty[stage], 0x3);          // frees the slot in both CTAs
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(tfull, 0x3);
        tphase ^= 1;
      }
    }
  } else {
    // ================= epilogue (warps 2..5 of both CTAs) =================
    const int sub = warp & 3;                            // TMEM sub-partition = lanes [32*sub, +32)
    const int row_in_cta = (sub & 1) * 32 + lane;        // rows 0..63 appear twice: lanes 0-63 and 64-127
    const int half = sub >> 1;                           // which half of each N-chunk these lanes hold
    uint32_t tphase = 0;
    bool first = true;
    for (int t = pair; t < num_tiles; t += npairs) {
      mbar_wait(tfull, tphase);
      tc_fence_after();
      const int row = t * 128 + (int)rank * 64 + row_in_cta;
      const uint32_t tlane = tmem_base + ((uint32_t)(sub * 32) << 16);
      for (int c = 0; c < p.nch; ++c) {
        for (int j0 = 0; j0 < p.NCH / 2; j0 += 8) {
          uint32_t r[8];
          tmem_ld8(tlane + (uint32_t)(c * (p.NCH / 2) + j0), r);
          tmem_ld_wait();
          if (first && p.dump) {
            float *dp = p.dump + ((size_t)(pair * 2 + rank) * 128 + sub * 32 + lane) * (p.N / 2) + c * (p.NCH / 2) + j0;
#pragma unroll
            for (int i = 0; i < 8; ++i) dp[i] = __uint_as_float(r[i]);
          }
          if (row < p.M) {
            const int n0 = c * p.NCH + half * (p.NCH / 2) + j0;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float x = __uint_as_float(r[i]);
              if (p.bias) x += __ldg(p.bias + n0 + i);
              if (p.relu) x = fmaxf(x, 0.f);
              v[i] = x;
            }
            float *o = p.out + (size_t)row * p.N + n0;
#pragma unroll
            for (int i = 0; i < 8; i += 4) *reinterpret_cast<float4 *>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        }
      }
      first = false;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty, 0);     // leader's barrier
      tphase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                        // peer smem / TMEM stay valid until both are done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, (uint32_t)p.tmem_cols);
  }
}

}  // namespace

int make_tmap_bf16(CUtensorMap *tm, const uint16_t *base, int rows, int K, int ld, int box_rows, int box_k);

}  // namespace car

extern "C" int car_gemm_pair_test(const uint16_t *a_hi, const uint16_t *a_lo, const uint16_t *w_hi,
                                  const uint16_t *w_lo, const float *bias, float *c, float *dump, int M, int N,
                                  int K, int nch, int split3, int relu, int max_pairs, int bk, void *stream) {
  using namespace car;
  const int BK = bk;
  if (BK != 64 && BK != 32) { set_error("pair gemm: bk must be 32 or 64"); return -12; }
  if (nch <= 0 || N % nch) { set_error("pair gemm: bad nch"); return -12; }
  int NCH = N / nch;
  if (NCH > 256 || NCH % 16 || K % 16 || N / 2 > 512) { set_error("pair gemm: unsupported N=%d nch=%d K=%d", N, nch, K); return -12; }
  CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
  int rc;
  if ((rc = make_tmap_bf16(&ta_hi, a_hi, M, K, K, 64, BK))) return rc;
  if ((rc = make_tmap_bf16(&tw_hi, w_hi, N, K, K, NCH / 2, BK))) return rc;
  if (split3) {
    if ((rc = make_tmap_bf16(&ta_lo, a_lo, M, K, K, 64, BK))) return rc;
    if ((rc = make_tmap_bf16(&tw_lo, w_lo, N, K, K, NCH / 2, BK))) return rc;
  } else { ta_lo = ta_hi; tw_lo = tw_hi; }
  PairParams p;
  p.M = M; p.N = N; p.K = K; p.NCH = NCH; p.nch = nch;
  int stage_bytes = (64 * BK * 2 + (NCH / 2) * BK * 2 * nch) * (split3 ? 2 : 1);
  p.stages = (220 * 1024 - 1280) / stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  if (p.stages < 2) { set_error("pair gemm: stage too large (%d B)", stage_bytes); return -14; }
  int cols = N / 2;
  p.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  p.bias = bias; p.relu = relu; p.out = c; p.dump = dump;
  size_t smem = (size_t)p.stages * stage_bytes + 1280;
  int tiles = (M + 127) / 128;
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int pairs = sms / 2;
  if (pairs > tiles) pairs = tiles;
  if (max_pairs > 0 && pairs > max_pairs) pairs = max_pairs;
  cudaError_t e;
  cudaStream_t st = (cudaStream_t)stream;
#define CAR_PG(S, B)                                                                                     \
  do {                                                                                                   \
    e = cudaFuncSetAttribute(k_pair_gemm<S, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    if (e == cudaSuccess) k_pair_gemm<S, B><<<pairs * 2, THREADS, smem, st>>>(ta_hi, ta_lo, tw_hi, tw_lo, p); \
  } while (0)
  if (split3) { if (BK == 64) CAR_PG(3, 64); else CAR_PG(3, 32); }
  else { if (BK == 64) CAR_PG(1, 64); else CAR_PG(1, 32); }
#undef CAR_PG
  if (e != cudaSuccess) { set_error("pair gemm: %s", cudaGetErrorString(e)); return (int)e; }
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("pair gemm launch: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  return 0;
}
