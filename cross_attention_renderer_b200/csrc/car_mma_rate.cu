// Micro-benchmark: back-to-back tcgen05.mma issue from resident smem operands (no loads in the
// loop).  Reports cycles per MMA for a given cta_group / M / N / swizzle, to size the fused kernel.
#include "car_common.cuh"
#include "car_umma.cuh"

namespace car {
namespace {
using namespace ptx;

template <int CG, int SW>
__global__ void __launch_bounds__(128, 1)
k_mma_rate(int M, int N, int iters, int nops, unsigned long long *out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t wbar[4];
  __shared__ uint32_t slot;
  // operands: A tile at 0 (up to 128 rows x SW bytes = 16 KB), B tile at 32 KB (up to 256 rows)
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  if (warp == 0 && lane == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&wbar[i], 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<CG>(&slot, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (nops < 0) {
    // multi-warp issue test: -nops warps each stream `iters` MMAs (N <= 128) into their own accumulator
    const int nw = -nops;
    if (warp < nw && lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc_bf16(M, N);
      const uint32_t sa = smem_u32(smem) + warp * 8192, sb = smem_u32(smem + 32 * 1024) + warp * 16384;
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t koff = (uint32_t)(i & (SW / 32 - 1)) * 32;
        umma_f16<CG>(tmem + (uint32_t)(warp * 128), make_desc<SW>(sa + koff), make_desc<SW>(sb + koff), idesc, 1u);
      }
      if (CG == 2) umma_commit_pair(&wbar[warp], 0x1); else umma_commit(&wbar[warp]);
      mbar_wait(&wbar[warp], 0);
      long long t1 = clock64();
      if (blockIdx.x == 0) out[warp] = (unsigned long long)(t1 - t0);
    }
  } else
  if (warp == 0 && lane == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(M, N);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32 * 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      for (int o = 0; o < nops; ++o) {
        const uint32_t koff = (uint32_t)((i + o) & (SW / 32 - 1)) * 32;
        const uint32_t d = tmem + (uint32_t)((o % 2) * 256);
        umma_f16<CG>(d, make_desc<SW>(sa + koff), make_desc<SW>(sb + koff), idesc, 1u);
      }
    }
    if (CG == 2) umma_commit_pair(&bar, 0x1); else umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<CG>(tmem, 512); }
}
}  // namespace
}  // namespace car

extern "C" int car_mma_rate_test(int cg, int M, int N, int sw, int iters, int nops, int ctas, void *out_u64, void *stream) {
  using namespace car;
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = 97 * 1024 + 1024;
  cudaError_t e = cudaSuccess;
  unsigned long long *out = (unsigned long long *)out_u64;
#define RATE(CG, SW)                                                                                   \
  do {                                                                                                 \
    e = cudaFuncSetAttribute(k_mma_rate<CG, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) break;                                                                       \
    if (CG == 2) {                                                                                     \
      cudaLaunchConfig_t cfg = {};                                                                     \
      cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = st; \
      cudaLaunchAttribute at[1];                                                                       \
      at[0].id = cudaLaunchAttributeClusterDimension;                                                  \
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;              \
      cfg.attrs = at; cfg.numAttrs = 1;                                                                \
      e = cudaLaunchKernelEx(&cfg, k_mma_rate<CG, SW>, M, N, iters, nops, out);                        \
    } else {                                                                                           \
      k_mma_rate<CG, SW><<<ctas, 128, smem, st>>>(M, N, iters, nops, out);                             \
    }                                                                                                  \
  } while (0)
  if (cg == 1) { if (sw == 128) RATE(1, 128); else RATE(1, 64); }
  else { if (sw == 128) RATE(2, 128); else RATE(2, 64); }
#undef RATE
  if (e != cudaSuccess) { set_error("mma rate: %s", cudaGetErrorString(e)); return (int)e; }
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("mma rate launch: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}
