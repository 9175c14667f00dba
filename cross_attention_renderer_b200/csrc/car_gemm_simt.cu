// Exact-fp32 SIMT GEMM:  C[M][N] (+)= act( relu_in(A)[M][K] · W[N][K]^T + bias + row_bias ).
// Used by CAR_PREC_FP32_SIMT (the reference-arithmetic mode) and, in every mode, for the
// per-ray colour MLP phi (reference resnet_block_fc.py:132-168) whose M is only the ray count.
//
// 128x64 output tile, BK = 16, 256 threads, 8x4 register block per thread.
#include "car_common.cuh"

namespace car {
namespace {

constexpr int BM = 128, BN = 64, BK = 16, PAD = 4;

__global__ void __launch_bounds__(256)
k_gemm_simt(const float *__restrict__ A, int lda, const float *__restrict__ W, int ldw,
            float *__restrict__ C, int ldc, int M, int N, int K, GemmEpi epi) {
  __shared__ float As[BK][BM + PAD];
  __shared__ float Ws[BK][BN + PAD];
  int tid = threadIdx.x;
  int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  int tx = tid & 15, ty = tid >> 4;          // tx -> 4 columns, ty -> 8 rows
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  int a_row = tid >> 2, a_kq = (tid & 3) * 4;     // 64 rows x 4 k-quads per pass, 2 passes
  int w_row = tid >> 2, w_kq = (tid & 3) * 4;     // 64 rows x 4 k-quads

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      int r = a_row + pass * 64;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M) v = *reinterpret_cast<const float4 *>(A + (size_t)(m0 + r) * lda + k0 + a_kq);
      if (epi.relu_in) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      As[a_kq + 0][r] = v.x; As[a_kq + 1][r] = v.y; As[a_kq + 2][r] = v.z; As[a_kq + 3][r] = v.w;
    }
    {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + w_row < N) v = *reinterpret_cast<const float4 *>(W + (size_t)(n0 + w_row) * ldw + k0 + w_kq);
      Ws[w_kq + 0][w_row] = v.x; Ws[w_kq + 1][w_row] = v.y; Ws[w_kq + 2][w_row] = v.z; Ws[w_kq + 3][w_row] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4 *>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4 *>(&As[k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4 *>(&Ws[k][tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + ty * 8 + i;
    if (m >= M) continue;
    const float *rb = epi.row_bias ? epi.row_bias + (size_t)(m / epi.rows_per_group) * N : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (epi.bias) v += epi.bias[n];
      if (rb) v += rb[n];
      if (epi.relu_out) v = fmaxf(v, 0.f);
      float *c = C + (size_t)m * ldc + n;
      if (epi.accumulate) v = *c + v;
      *c = v;
    }
  }
}

}  // namespace

void launch_gemm_simt(const float *A, int lda, const float *W, int ldw, float *C, int ldc, int M,
                      int N, int K, const GemmEpi &epi, cudaStream_t st) {
  if (M <= 0 || N <= 0) return;
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
  prof_pre(-1, st);
  k_gemm_simt<<<grid, 256, 0, st>>>(A, lda, W, ldw, C, ldc, M, N, K, epi);
  prof_post(st);
  count_launch();
}

}  // namespace car
