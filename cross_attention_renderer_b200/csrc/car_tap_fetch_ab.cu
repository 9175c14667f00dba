// A/B microbenchmark of the bilinear tap fetch (libcar_b200_test.so only; DESIGN.md "tap fetch A/B").
//
// One K-stage of the fused kernel's gather is 64 sample rows x 32 fp32 channels: four taps of 128 bytes per row
// (32 KB of taps) mixed into an 8 KB bf16 operand stage.  Two ways to bring the taps to the SM:
//   variant 0  LDG producers (what k_fused_encode does): 64 threads per stage, every thread keeps a rolling
//              window of four items = 16 x LDG.128 in flight and mixes in registers; no shared-memory staging.
//   variant 1  TMA tile::gather4: one cp.async.bulk.tensor.2d...tile::gather4 per sample row fetches its four tap
//              rows (4 x 128 B) of the [pixels][C] view into a 32 KB shared-memory staging slot (ring of nslot);
//              the consumer threads read the taps back from shared memory (LDS.128) and mix.
// Both write the mixed bf16 values (4 channels = 8 bytes per item) to global memory so results can be compared.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int ROWS = 64, KS = 32, CONSUMERS = 256;
constexpr int STAGE_BYTES = ROWS * 4 * KS * 4;                 // 32 KB of taps per stage

struct TapArgs {
  const float *map;         // [pixels][C] fp32 (one NHWC level)
  const int4 *taps;         // [rows] four pixel indices
  const float4 *wts;        // [rows] four bilinear weights
  uint2 *out;               // [rows][C / 4] four bf16 per item
  int C, n_rows;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a transfer that never completes traps instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  for (long long spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (spin > (1ll << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_gather4(void *dst, const CUtensorMap *tm, uint64_t *bar, int col, int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ uint2 mix4(const float4 (&x)[4], const float4 w) {
  const float ww[4] = {w.x, w.y, w.z, w.w};
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    acc.x = fmaf(x[k].x, ww[k], acc.x); acc.y = fmaf(x[k].y, ww[k], acc.y);
    acc.z = fmaf(x[k].z, ww[k], acc.z); acc.w = fmaf(x[k].w, ww[k], acc.w);
  }
  const uint32_t h0 = __bfloat16_as_ushort(__float2bfloat16_rn(acc.x)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(acc.y)) << 16);
  const uint32_t h1 = __bfloat16_as_ushort(__float2bfloat16_rn(acc.z)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(acc.w)) << 16);
  return make_uint2(h0, h1);
}

// ---- variant 0: LDG producers (the fused kernel's scheme: 4 groups of 64 threads, one stage per group at a time) ----
__global__ void __launch_bounds__(CONSUMERS)
k_tap_ldg(TapArgs a) {
  const int gi = threadIdx.x >> 6, gt = threadIdx.x & 63;
  const int chunks = a.C / KS, n_stages = (a.n_rows / ROWS) * chunks;
  for (int s = blockIdx.x * 4 + gi; s < n_stages; s += gridDim.x * 4) {
    const int row0 = (s / chunks) * ROWS, ch0 = (s % chunks) * KS;
    float4 x[4][4], wt[4];
    auto load_item = [&](int slot, int i) {
      const int item = gt + 64 * i, rr = item >> 3, grp = item & 7;
      const int4 t = __ldg(a.taps + row0 + rr);
      wt[slot] = __ldg(a.wts + row0 + rr);
      const float *base = a.map + ch0 + grp * 4;
      x[slot][0] = __ldg(reinterpret_cast<const float4 *>(base + (size_t)t.x * a.C));
      x[slot][1] = __ldg(reinterpret_cast<const float4 *>(base + (size_t)t.y * a.C));
      x[slot][2] = __ldg(reinterpret_cast<const float4 *>(base + (size_t)t.z * a.C));
      x[slot][3] = __ldg(reinterpret_cast<const float4 *>(base + (size_t)t.w * a.C));
    };
#pragma unroll
    for (int i = 0; i < 4; ++i) load_item(i, i);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int item = gt + 64 * i, rr = item >> 3, grp = item & 7, sl = i & 3;
      const uint2 v = mix4(x[sl], wt[sl]);
      if (i + 4 < 8) load_item(sl, i + 4);
      a.out[(size_t)(row0 + rr) * (a.C / 4) + (ch0 >> 2) + grp] = v;
    }
  }
}

// ---- variant 1: TMA tile::gather4 into a shared-memory ring, 256 consumer threads + one issuing warp ----
__global__ void __launch_bounds__(CONSUMERS + 64)
k_tap_gather4(const __grid_constant__ CUtensorMap tm, TapArgs a, int nslot, int issue_warps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)nslot * STAGE_BYTES);
  uint64_t *empty = full + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = a.C / KS, n_stages = (a.n_rows / ROWS) * chunks;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nslot; ++s) { mbar_init(&full[s], issue_warps); mbar_init(&empty[s], CONSUMERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp >= CONSUMERS / 32) {
    // issuing warp(s): every lane issues the gather4 of its sample rows (one warp: two rows per lane; two warps: one)
    const int iw = warp - CONSUMERS / 32, per = ROWS / (32 * issue_warps);
    int slot = 0; uint32_t phase = 0;
    for (int s = blockIdx.x; s < n_stages; s += gridDim.x) {
      const int row0 = (s / chunks) * ROWS, ch0 = (s % chunks) * KS;
      if (lane == 0) { mbar_wait(&empty[slot], phase ^ 1); mbar_expect_tx(&full[slot], STAGE_BYTES / issue_warps); }
      __syncwarp();
      uint8_t *dst = smem + (size_t)slot * STAGE_BYTES;
      for (int j = 0; j < per; ++j) {
        const int rr = (iw * per + j) * 32 + lane;
        const int4 t = __ldg(a.taps + row0 + rr);
        tma_gather4(dst + rr * 512, &tm, &full[slot], ch0, t.x, t.y, t.z, t.w);
      }
      if (++slot == nslot) { slot = 0; phase ^= 1; }
    }
  } else {
    int slot = 0; uint32_t phase = 0;
    for (int s = blockIdx.x; s < n_stages; s += gridDim.x) {
      const int row0 = (s / chunks) * ROWS, ch0 = (s % chunks) * KS;
      float4 wt[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) wt[i] = __ldg(a.wts + row0 + ((threadIdx.x + CONSUMERS * i) >> 3));
      mbar_wait(&full[slot], phase);
      const uint8_t *src = smem + (size_t)slot * STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int item = threadIdx.x + CONSUMERS * i, rr = item >> 3, grp = item & 7;
        float4 x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = *reinterpret_cast<const float4 *>(src + rr * 512 + k * 128 + grp * 16);
        a.out[(size_t)(row0 + rr) * (a.C / 4) + (ch0 >> 2) + grp] = mix4(x, wt[i]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
      if (++slot == nslot) { slot = 0; phase ^= 1; }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

}  // namespace

// Returns 0 and the average milliseconds per launch in *ms; negative / CUDA error code otherwise.
extern "C" int car_tap_fetch_ab(const float *map, int pixels, int C, const int *taps, const float *wts, void *out, int n_rows,
                                int variant, int box_rows, int nslot, int ctas_per_sm, int iters, float *ms, void *stream,
                                int issue_warps) {
  if (!map || !taps || !wts || !out || !ms || C % KS || n_rows % ROWS || iters < 1) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  TapArgs a;
  a.map = map; a.taps = reinterpret_cast<const int4 *>(taps); a.wts = reinterpret_cast<const float4 *>(wts);
  a.out = reinterpret_cast<uint2 *>(out); a.C = C; a.n_rows = n_rows;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaError_t e = cudaSuccess;
  if (variant == 0) {
    const int grid = sms * (ctas_per_sm > 0 ? ctas_per_sm : 2);
    k_tap_ldg<<<grid, CONSUMERS, 0, st>>>(a);                                   // warm-up
    cudaEventRecord(e0, st);
    for (int i = 0; i < iters; ++i) k_tap_ldg<<<grid, CONSUMERS, 0, st>>>(a);
    cudaEventRecord(e1, st);
  } else {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return -10;
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)pixels};
    cuuint64_t strides[1] = {(cuuint64_t)C * 4};
    cuuint32_t box[2] = {(cuuint32_t)KS, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeTiledFn>(ptr)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(map), dims, strides, box, estr,
                                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -11;
    if (nslot < 1 || nslot > 6 || (issue_warps != 1 && issue_warps != 2)) return -1;
    const size_t smem = (size_t)nslot * STAGE_BYTES + 1024 + 256;
    if ((e = cudaFuncSetAttribute(k_tap_gather4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
    const int grid = sms * (ctas_per_sm > 0 ? ctas_per_sm : 1);
    const int threads = CONSUMERS + 32 * issue_warps;
    k_tap_gather4<<<grid, threads, smem, st>>>(tm, a, nslot, issue_warps);
    cudaEventRecord(e0, st);
    for (int i = 0; i < iters; ++i) k_tap_gather4<<<grid, threads, smem, st>>>(tm, a, nslot, issue_warps);
    cudaEventRecord(e1, st);
  }
  e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  float t = 0.f;
  if (e == cudaSuccess) cudaEventElapsedTime(&t, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms = t / iters;
  return (int)e;
}
