/*
 * car_b200_test.h - entry points of libcar_b200_test.so: micro-benchmarks and building-block test
 * kernels used while sizing the fused per-ray kernel (scripts/mma_rate.py, tests/test_gpu_gemm.py).
 * They are NOT part of the product library (libcar_b200.so, include/car_b200.h) and replace nothing
 * in the reference.
 */
#ifndef CAR_B200_TEST_H
#define CAR_B200_TEST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Micro-benchmark: cycles for iters*nops back-to-back tcgen05.mma (M x N x 16, bf16) from resident smem. */
int car_mma_rate_test(int cg, int M, int N, int sw, int iters, int nops, int ctas, void *out_u64, void *stream);

/* A/B of the bilinear tap fetch (csrc/car_tap_fetch_ab.cu): variant 0 = LDG producers with a rolling window (the
 * fused kernel's scheme), variant 1 = TMA tile::gather4 of the four tap rows into a shared-memory ring of `nslot`
 * 32 KB stages (tensor-map box = 1 row x 32 columns; a 4-row box is rejected by the hardware).  map: [pixels][C] fp32; taps: [n_rows][4] pixel indices; wts: [n_rows][4]; out: [n_rows][C] bf16.
 * Returns 0 and the milliseconds per launch (average of `iters` launches after one warm-up).                     */
int car_tap_fetch_ab(const float *map, int pixels, int C, const int *taps, const float *wts, void *out, int n_rows,
                     int variant, int box_rows, int nslot, int ctas_per_sm, int iters, float *ms, void *stream,
                     int issue_warps /* 1 or 2 warps issuing the gather4 instructions */);

/* CTA-pair (cta_group::2) tcgen05 GEMM, the building block of the fused per-ray kernel, exported
 * for tests: C[M][N] = A·W^T (+bias); N is processed as `nch` MMA chunks; `dump` (optional)
 * receives the raw TMEM image [pairs*2][128 lanes][N/2] of each pair's first tile. */
int car_gemm_pair_test(const uint16_t *a_hi, const uint16_t *a_lo, const uint16_t *w_hi,
                       const uint16_t *w_lo, const float *bias, float *c, float *dump, int M, int N,
                       int K, int nch, int split3, int relu, int max_pairs, int bk /*32|64*/, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CAR_B200_TEST_H */
