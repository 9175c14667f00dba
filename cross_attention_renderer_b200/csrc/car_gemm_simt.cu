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
      if (epi.mask && !(epi.mask[(size_t)m * epi.ldmask + n] > 0.f)) v = 0.f;
      float *c = C + (size_t)m * ldc + n;
      if (epi.accumulate) v = *c + v;
      *c = v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Weight gradient:  dW[N][K] += dY[M][N]^T · act(A[M][K]),  db[N] += column sums of dY.
// 64x64 output tile per CTA, the M (row) range is split over gridDim.z; partial tiles are
// combined with fp32 atomics.  Rows are read row-major (coalesced along N / K).
// ---------------------------------------------------------------------------------------
constexpr int WG_BN = 64, WG_BK = 64, WG_BM = 16;

__device__ __forceinline__ float4 load4_guard(const float *p, int valid) {
  if (valid >= 4 && (((uintptr_t)p) & 15) == 0) return *reinterpret_cast<const float4 *>(p);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid > 0) v.x = p[0];
  if (valid > 1) v.y = p[1];
  if (valid > 2) v.z = p[2];
  if (valid > 3) v.w = p[3];
  return v;
}

__global__ void __launch_bounds__(256)
k_wgrad_simt(const float *__restrict__ dY, int ldy, const float *__restrict__ A, int lda,
             float *__restrict__ dW, int ldw, float *__restrict__ db, int M, int N, int K, int relu_a,
             int m_per_cta) {
  __shared__ __align__(16) float Ys[WG_BM][WG_BN];
  __shared__ __align__(16) float As[WG_BM][WG_BK];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * WG_BN, k0 = blockIdx.y * WG_BK;
  const int m_begin = blockIdx.z * m_per_cta;
  const int m_end = min(M, m_begin + m_per_cta);
  const int tx = tid & 15, ty = tid >> 4;          // tx -> 4 k columns, ty -> 4 n rows
  const int lrow = tid >> 4, lcol = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;
  const bool do_bias = db != nullptr && blockIdx.y == 0 && tid < WG_BN;
  for (int m0 = m_begin; m0 < m_end; m0 += WG_BM) {
    int r = m0 + lrow;
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f), av = y;
    if (r < m_end) {
      y = load4_guard(dY + (size_t)r * ldy + n0 + lcol, N - (n0 + lcol));
      av = load4_guard(A + (size_t)r * lda + k0 + lcol, K - (k0 + lcol));
      if (relu_a) { av.x = fmaxf(av.x, 0.f); av.y = fmaxf(av.y, 0.f); av.z = fmaxf(av.z, 0.f); av.w = fmaxf(av.w, 0.f); }
    }
    *reinterpret_cast<float4 *>(&Ys[lrow][lcol]) = y;
    *reinterpret_cast<float4 *>(&As[lrow][lcol]) = av;
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < WG_BM; ++rr) {
      float4 yv = *reinterpret_cast<const float4 *>(&Ys[rr][ty * 4]);
      float4 aa = *reinterpret_cast<const float4 *>(&As[rr][tx * 4]);
      float yy[4] = {yv.x, yv.y, yv.z, yv.w}, a4[4] = {aa.x, aa.y, aa.z, aa.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(yy[i], a4[j], acc[i][j]);
    }
    if (do_bias) {
#pragma unroll
      for (int rr = 0; rr < WG_BM; ++rr) bsum += Ys[rr][tid];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(dW + (size_t)n * ldw + k, acc[i][j]);
    }
  }
  if (do_bias && n0 + tid < N) atomicAdd(db + n0 + tid, bsum);
}

__global__ void k_transpose(const float *__restrict__ src, int rows, int cols, int ld_src,
                            float *__restrict__ dst) {
  __shared__ float tile[32][33];
  int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(size_t)r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) dst[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

}  // namespace

void launch_wgrad_simt(const float *dY, int ldy, const float *A, int lda, float *dW, int ldw, float *db,
                       int M, int N, int K, int relu_a, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0 || !dW) return;
  int nt = (N + WG_BN - 1) / WG_BN, kt = (K + WG_BK - 1) / WG_BK;
  int split = (148 * 4 + nt * kt - 1) / (nt * kt);
  int max_split = (M + 63) / 64;
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  int m_per_cta = ((M + split - 1) / split + WG_BM - 1) / WG_BM * WG_BM;
  split = (M + m_per_cta - 1) / m_per_cta;
  dim3 grid(nt, kt, split);
  prof_pre(-1, st);
  k_wgrad_simt<<<grid, 256, 0, st>>>(dY, ldy, A, lda, dW, ldw, db, M, N, K, relu_a, m_per_cta);
  prof_post(st);
  count_launch();
}

void launch_transpose(const float *src, int rows, int cols, int ld_src, float *dst, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  prof_pre(-1, st);
  k_transpose<<<grid, block, 0, st>>>(src, rows, cols, ld_src, dst);
  prof_post(st);
  count_launch();
}

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
