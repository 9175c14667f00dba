// Per-ray attention rounds, expected-depth output, colour-MLP input assembly and the
// final mask / white fill.
//
// Reference arithmetic restated here:
//   round 1   models.py:532-545   scores <K,Q1>/16, joint softmax over 2P, weighted V sum
//   round 2   models.py:555-565   scores <Q2,Q1>/16 (Q1 = coords_embed, not K), residual
//   depth     models.py:573-594   argmax, attention-weighted clamp(pt), inverse(query c2w)
//   mask      models.py:615-621
#include <math.h>

#include "car_common.cuh"

namespace car {
namespace {

constexpr int ATT_THREADS = 128;
constexpr int MAX_ROWS = 768;      // n * P: 2 * 256 on the n_view = 2 path, 3 * 256 on the general one

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

// scores[i] = <x_i, y_i> / 16 for the 2P rows of one ray, then joint softmax (in smem).
__device__ void scores_softmax(const float *__restrict__ x, const float *__restrict__ y, int rows,
                               float *sc, float *red) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < rows; i += ATT_THREADS / 32) {
    float4 a = __ldg(reinterpret_cast<const float4 *>(x + (size_t)i * 128) + lane);
    float4 b = __ldg(reinterpret_cast<const float4 *>(y + (size_t)i * 128) + lane);
    float p = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    p = warp_sum(p);
    if (lane == 0) sc[i] = p / 16.0f;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < rows; i += ATT_THREADS) mx = fmaxf(mx, sc[i]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sm = 0.f;
  for (int i = threadIdx.x; i < rows; i += ATT_THREADS) {
    float e = expf(sc[i] - mx);
    sc[i] = e;
    sm += e;
  }
  sm = warp_sum(sm);
  if (lane == 0) red[warp] = sm;
  __syncthreads();
  sm = (red[0] + red[1]) + (red[2] + red[3]);
  for (int i = threadIdx.x; i < rows; i += ATT_THREADS) sc[i] = sc[i] / sm;
  __syncthreads();
}

__global__ void __launch_bounds__(ATT_THREADS)
k_attention1(car_render_args a, int g0, const float *__restrict__ key,
             const float *__restrict__ q1, const float *__restrict__ value,
             const float *__restrict__ geom, float *__restrict__ zsum) {
  __shared__ float sc[MAX_ROWS];
  __shared__ float red[8];
  int gl = blockIdx.x, g = g0 + gl;
  int s = g / a.R, r = g - s * a.R;
  int P = a.P, rows = 2 * P;
  size_t row0 = (size_t)gl * rows;
  scores_softmax(key + row0 * 128, q1 + row0 * 128, rows, sc, red);
  // at_wt (b*2,R,P) and per-context argmax (first maximum)
  for (int i = threadIdx.x; i < rows; i += ATT_THREADS) {
    int j = i / P, k = i - j * P;
    a.at_wt[((size_t)(s * 2 + j) * a.R + r) * P + k] = sc[i];
  }
  if (threadIdx.x < 2) {
    int j = threadIdx.x;
    float best = sc[j * P];
    int bi = 0;
    for (int k = 1; k < P; ++k)
      if (sc[j * P + k] > best) { best = sc[j * P + k]; bi = k; }
    a.at_wt_max[(size_t)(s * 2 + j) * a.R + r] = bi;
  }
  // z_sum = sum_ctx sum_k a * V   (models.py:537-540)
  for (int c = threadIdx.x; c < CAR_C_LAT; c += ATT_THREADS) {
    const float *v = value + row0 * CAR_C_LAT + c;
    float z0 = 0.f, z1 = 0.f;
    for (int k = 0; k < P; ++k) z0 += sc[k] * v[(size_t)k * CAR_C_LAT];
    for (int k = 0; k < P; ++k) z1 += sc[P + k] * v[(size_t)(P + k) * CAR_C_LAT];
    zsum[(size_t)gl * CAR_C_LAT + c] = z0 + z1;
  }
  // expected 3-D point -> depth in the query camera (models.py:577-590, geometry.py:395-406)
  if (threadIdx.x >= 32 && threadIdx.x < 38) {
    int t = threadIdx.x - 32, j = t / 3, comp = t - j * 3;
    const float *G = geom + (row0 + (size_t)j * P) * CAR_GEOM_STRIDE + G_PTC + comp;
    float w = 0.f;
    for (int k = 0; k < P; ++k) w += sc[j * P + k] * G[(size_t)k * CAR_GEOM_STRIDE];
    red[2 + t] = w;                       // red[2..7] (red[0..1] no longer needed)
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float x = red[2] + red[5], y = red[3] + red[6], z = red[4] + red[7];
    const float *qi = a.cams.qinv + (size_t)s * 16;
    float zc = ((qi[8] * x + qi[9] * y) + qi[10] * z) + qi[11];
    a.depth_ray[(size_t)s * a.R + r] = fminf(fmaxf(zc, 0.f), 10.f);
  }
}

__global__ void __launch_bounds__(ATT_THREADS)
k_attention2(car_render_args a, int g0, const float *__restrict__ q2, const float *__restrict__ q1,
             const float *__restrict__ value, const float *__restrict__ zsum,
             float *__restrict__ zfin, float *__restrict__ att2) {
  __shared__ float sc[MAX_ROWS];
  __shared__ float red[8];
  int gl = blockIdx.x;
  int P = a.P, rows = 2 * P;
  size_t row0 = (size_t)gl * rows;
  scores_softmax(q2 + row0 * 128, q1 + row0 * 128, rows, sc, red);
  if (att2)                                     // training mode: round-2 weights for the backward pass
    for (int i = threadIdx.x; i < rows; i += ATT_THREADS) att2[row0 + i] = sc[i];
  // per ctx: sum_k a2*V + z_sum, then summed over ctx (models.py:561-564)
  for (int c = threadIdx.x; c < CAR_C_LAT; c += ATT_THREADS) {
    const float *v = value + row0 * CAR_C_LAT + c;
    float z0 = 0.f, z1 = 0.f;
    for (int k = 0; k < P; ++k) z0 += sc[k] * v[(size_t)k * CAR_C_LAT];
    for (int k = 0; k < P; ++k) z1 += sc[P + k] * v[(size_t)(P + k) * CAR_C_LAT];
    float zs = zsum[(size_t)gl * CAR_C_LAT + c];
    zfin[(size_t)gl * CAR_C_LAT + c] = (z0 + zs) + (z1 + zs);
  }
}

// coords18 padded to 32: [d0 m0 o0 | d1 m1 o1 | 0...] (models.py:597-602)
__global__ void k_phi_prep(car_render_args a, int g0, int g1, float *__restrict__ c18) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int n = (g1 - g0) * 32;
  if (idx >= n) return;
  int gl = idx >> 5, c = idx & 31;
  int g = g0 + gl, s = g / a.R, r = g - s * a.R;
  float v = 0.f;
  if (c < 18) {
    int j = c / 9, q = c - j * 9;
    v = a.coords[((size_t)(s * 2 + j) * a.R + r) * 9 + q];
  }
  c18[idx] = v;
}

__global__ void k_finalize(car_render_args a, int g0, int g1, const float *__restrict__ rgb3,
                           const uint8_t *__restrict__ overlap) {
  int gl = blockIdx.x * blockDim.x + threadIdx.x;
  if (gl >= g1 - g0) return;
  int g = g0 + gl;
  float valid = (overlap[gl * 2] | overlap[gl * 2 + 1]) ? 1.f : 0.f;
  a.valid_mask[g] = valid;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    a.rgb[(size_t)g * 3 + c] = rgb3[(size_t)gl * 3 + c] * valid + 1.f * (1.f - valid);
}

// ---- general branches: n contexts, latent width L (car_general_args) ------------------------------
__global__ void __launch_bounds__(ATT_THREADS)
k_attention1_g(car_general_args a, int n, int L, int g0, const float *__restrict__ key, const float *__restrict__ q1,
               const float *__restrict__ value, const float *__restrict__ geom, float *__restrict__ zsum) {
  __shared__ float sc[MAX_ROWS];
  __shared__ float red[16];
  int gl = blockIdx.x, g = g0 + gl;
  int s = g / a.R, r = g - s * a.R;
  int P = a.P, rows = n * P;
  size_t row0 = (size_t)gl * rows;
  scores_softmax(key + row0 * 128, q1 + row0 * 128, rows, sc, red);
  for (int i = threadIdx.x; i < rows; i += ATT_THREADS) {
    int j = i / P, k = i - j * P;
    a.at_wt[((size_t)(s * n + j) * a.R + r) * P + k] = sc[i];
  }
  if (threadIdx.x < n) {                                  // per-context argmax, first maximum (models.py:574)
    int j = threadIdx.x;
    float best = sc[j * P];
    int bi = 0;
    for (int k = 1; k < P; ++k)
      if (sc[j * P + k] > best) { best = sc[j * P + k]; bi = k; }
    a.at_wt_max[(size_t)(s * n + j) * a.R + r] = bi;
  }
  // z_sum = sum over contexts of the per-context weighted sums, contexts added in order (models.py:537-540)
  for (int c = threadIdx.x; c < L; c += ATT_THREADS) {
    const float *v = value + row0 * L + c;
    float z = 0.f;
    for (int j = 0; j < n; ++j) {
      float zj = 0.f;
      for (int k = 0; k < P; ++k) zj += sc[j * P + k] * v[(size_t)(j * P + k) * L];
      z = j == 0 ? zj : z + zj;
    }
    zsum[(size_t)gl * L + c] = z;
  }
  // expected 3-D point: per context, summed over contexts, then the query camera's inverse (models.py:577-590)
  if (threadIdx.x >= 32 && threadIdx.x < 32 + 3 * n) {
    int t = threadIdx.x - 32, j = t / 3, comp = t - j * 3;
    const float *G = geom + (row0 + (size_t)j * P) * CAR_GG_STRIDE + GG_PTC + comp;
    float w = 0.f;
    for (int k = 0; k < P; ++k) w += sc[j * P + k] * G[(size_t)k * CAR_GG_STRIDE];
    red[4 + t] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float x = red[4], y = red[5], z = red[6];
    for (int j = 1; j < n; ++j) { x += red[4 + 3 * j]; y += red[5 + 3 * j]; z += red[6 + 3 * j]; }
    const float *qi = a.cams.qinv + (size_t)s * 16;
    float zc = ((qi[8] * x + qi[9] * y) + qi[10] * z) + qi[11];
    a.depth_ray[(size_t)s * a.R + r] = fminf(fmaxf(zc, 0.f), 10.f);
  }
}

__global__ void __launch_bounds__(ATT_THREADS)
k_attention2_g(car_general_args a, int n, int L, int g0, const float *__restrict__ q2, const float *__restrict__ q1,
               const float *__restrict__ value, const float *__restrict__ zsum, float *__restrict__ zfin) {
  __shared__ float sc[MAX_ROWS];
  __shared__ float red[16];
  int gl = blockIdx.x;
  int P = a.P, rows = n * P;
  size_t row0 = (size_t)gl * rows;
  scores_softmax(q2 + row0 * 128, q1 + row0 * 128, rows, sc, red);
  // per context: sum_k a2 * V + z_sum, then summed over the contexts (models.py:561-564)
  for (int c = threadIdx.x; c < L; c += ATT_THREADS) {
    const float *v = value + row0 * L + c;
    const float zs = zsum[(size_t)gl * L + c];
    float z = 0.f;
    for (int j = 0; j < n; ++j) {
      float zj = 0.f;
      for (int k = 0; k < P; ++k) zj += sc[j * P + k] * v[(size_t)(j * P + k) * L];
      z = j == 0 ? (zj + zs) : z + (zj + zs);
    }
    zfin[(size_t)gl * L + c] = z;
  }
}

// coords (9 per context) padded to 32: [d_0 m_0 o_0 | d_1 m_1 o_1 | ... | 0]  (models.py:597-602)
__global__ void k_phi_prep_g(car_general_args a, int g0, int g1, float *__restrict__ c32) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (g1 - g0) * 32) return;
  int gl = idx >> 5, c = idx & 31;
  int g = g0 + gl, s = g / a.R, r = g - s * a.R;
  float v = 0.f;
  if (c < 9 * a.n_view) {
    int j = c / 9, q = c - j * 9;
    v = a.coords[((size_t)(s * a.n_view + j) * a.R + r) * 9 + q];
  }
  c32[idx] = v;
}

__global__ void k_finalize_g(car_general_args a, int g0, int g1, const float *__restrict__ rgb3,
                             const uint8_t *__restrict__ overlap) {
  int gl = blockIdx.x * blockDim.x + threadIdx.x;
  if (gl >= g1 - g0) return;
  int g = g0 + gl;
  int any = 0;
  for (int j = 0; j < a.n_view; ++j) any |= overlap[gl * a.n_view + j];
  float valid = any ? 1.f : 0.f;
  a.valid_mask[g] = valid;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    a.rgb[(size_t)g * 3 + c] = rgb3[(size_t)gl * 3 + c] * valid + 1.f * (1.f - valid);
}

}  // namespace

void launch_attention1_general(const car_general_args &a, const GenShape &gs, int g0, int g1, const float *key, const float *q1,
                               const float *value, const float *geom, float *zsum, cudaStream_t st) {
  if (g1 <= g0) return;
  prof_pre(CAR_ST_ATTENTION, st);
  k_attention1_g<<<g1 - g0, ATT_THREADS, 0, st>>>(a, gs.n, gs.L, g0, key, q1, value, geom, zsum);
  prof_post(st);
  count_launch();
}

void launch_attention2_general(const car_general_args &a, const GenShape &gs, int g0, int g1, const float *q2, const float *q1,
                               const float *value, const float *zsum, float *zfin, cudaStream_t st) {
  if (g1 <= g0) return;
  prof_pre(CAR_ST_ATTENTION, st);
  k_attention2_g<<<g1 - g0, ATT_THREADS, 0, st>>>(a, gs.n, gs.L, g0, q2, q1, value, zsum, zfin);
  prof_post(st);
  count_launch();
}

void launch_phi_prep_general(const car_general_args &a, int g0, int g1, float *c32, cudaStream_t st) {
  int n = (g1 - g0) * 32;
  if (n <= 0) return;
  prof_pre(CAR_ST_PHI, st);
  k_phi_prep_g<<<(n + 255) / 256, 256, 0, st>>>(a, g0, g1, c32);
  prof_post(st);
  count_launch();
}

void launch_finalize_general(const car_general_args &a, int g0, int g1, const float *rgb3, const uint8_t *overlap,
                             cudaStream_t st) {
  int n = g1 - g0;
  if (n <= 0) return;
  prof_pre(CAR_ST_PHI, st);
  k_finalize_g<<<(n + 127) / 128, 128, 0, st>>>(a, g0, g1, rgb3, overlap);
  prof_post(st);
  count_launch();
}

void launch_attention1(const car_render_args &a, int g0, int g1, const float *key, const float *q1,
                       const float *value, const float *geom, float *zsum, float *,
                       cudaStream_t st) {
  if (g1 <= g0) return;
  prof_pre(CAR_ST_ATTENTION, st);
  k_attention1<<<g1 - g0, ATT_THREADS, 0, st>>>(a, g0, key, q1, value, geom, zsum);
  prof_post(st);
  count_launch();
}

void launch_attention2(const car_render_args &a, int g0, int g1, const float *q2, const float *q1,
                       const float *value, const float *zsum, float *zfin, float *att2, cudaStream_t st) {
  if (g1 <= g0) return;
  prof_pre(CAR_ST_ATTENTION, st);
  k_attention2<<<g1 - g0, ATT_THREADS, 0, st>>>(a, g0, q2, q1, value, zsum, zfin, att2);
  prof_post(st);
  count_launch();
}

void launch_phi_prep(const car_render_args &a, int g0, int g1, float *c18, cudaStream_t st) {
  int n = (g1 - g0) * 32;
  if (n <= 0) return;
  prof_pre(CAR_ST_PHI, st);
  k_phi_prep<<<(n + 255) / 256, 256, 0, st>>>(a, g0, g1, c18);
  prof_post(st);
  count_launch();
}

void launch_finalize(const car_render_args &a, int g0, int g1, const float *rgb3,
                     const uint8_t *overlap, cudaStream_t st) {
  int n = g1 - g0;
  if (n <= 0) return;
  prof_pre(CAR_ST_PHI, st);
  k_finalize<<<(n + 127) / 128, 128, 0, st>>>(a, g0, g1, rgb3, overlap);
  prof_post(st);
  count_launch();
}

}  // namespace car
