// Backward pass of the per-ray rendering path (car_render_backward, include/car_b200.h).
//
// Reference: autograd of CrossAttentionRenderer.forward (models.py:278-621) as driven by
// train_loss.backward() (training.py:125).  The forward (car_render_forward with train = 1)
// left every activation of the ray range in its workspace; this file walks the graph backwards:
//
//   rgb / depth cotangents -> colour MLP phi (resnet_block_fc.py:132-168)
//     -> attention round 2 (models.py:555-565) -> query_repeat_embed(_2), encode_latent
//     -> attention round 1 (:532-545) + expected depth (:577-594)
//     -> query_embed(_2), key_map(_2), latent_value -> query_encode_latent(_2) (both views)
//     -> scatter-add into the feature maps (grid_sample backward, :278,317).
//
// ReLU subgradients follow torch (grad * (out > 0)); softmax backward is
// ds = a * (da - sum(a * da)); clamp backward passes min <= x <= max.  Nothing upstream of
// the sample coordinates has a gradient (pt / depth detached, models.py:327-328,516).
//
// Matrix products: exact fp32 (k_gemm_simt for the data gradients with transposed weights, k_wgrad_simt for the
// weight gradients), or - CAR_PREC_FP32_3XBF16 - the per-sample layers on tcgen05 (k_gemm_umma, hi + lo bf16
// operands: data gradients with the ReLU mask in the epilogue, weight gradients split along the row dimension);
// the per-ray layers (M = rays) stay exact fp32 either way.
#include <math.h>
#include <string.h>

#include <cuda_bf16.h>

#include "car_common.cuh"

namespace car {
namespace {

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

constexpr int BT = 128;            // threads of the per-ray kernels
constexpr int MAX_ROWS = 512;      // 2 * P, P <= 256

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ float block_sum(float v, float *red) {     // BT threads; red[4]
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return (red[0] + red[1]) + (red[2] + red[3]);
}

// da[i] = <V_i, dz> for the 2P rows of one ray (warp per row)
__device__ void value_dots(const float *__restrict__ v, const float *dz, int rows, float *da) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < rows; i += BT / 32) {
    const float *r = v + (size_t)i * CAR_C_LAT;
    float p = 0.f;
    for (int c = lane; c < CAR_C_LAT; c += 32) p = fmaf(__ldg(r + c), dz[c], p);
    p = warp_sum(p);
    if (lane == 0) da[i] = p;
  }
}

// ---- attention round 2 backward (models.py:555-565) -------------------------------------
//   zfin = sum_rows a2 * V + 2 * zsum ;  s2 = <Q2, Q1> / 16
// out: dQ2 = ds2 * Q1, dQ1 = ds2 * Q2, dV = a2 * dzf, d_zsum = 2 * dzf
__global__ void __launch_bounds__(BT)
k_attn2_bwd(int P, const float *__restrict__ d_zfin, const float *__restrict__ value,
            const float *__restrict__ att2, const float *__restrict__ q1, const float *__restrict__ q2,
            float *__restrict__ dq2, float *__restrict__ dq1, float *__restrict__ dv,
            float *__restrict__ d_zsum) {
  __shared__ __align__(16) float dz[CAR_C_LAT];
  __shared__ float da[MAX_ROWS], aw[MAX_ROWS];
  __shared__ float red[4];
  int gl = blockIdx.x, rows = 2 * P;
  size_t row0 = (size_t)gl * rows;
  for (int c = threadIdx.x; c < CAR_C_LAT; c += BT) dz[c] = d_zfin[(size_t)gl * CAR_C_LAT + c];
  for (int i = threadIdx.x; i < rows; i += BT) aw[i] = att2[row0 + i];
  __syncthreads();
  value_dots(value + row0 * CAR_C_LAT, dz, rows, da);
  __syncthreads();
  float part = 0.f;
  for (int i = threadIdx.x; i < rows; i += BT) part += aw[i] * da[i];
  float dot = block_sum(part, red);
  for (int i = threadIdx.x; i < rows; i += BT) da[i] = aw[i] * (da[i] - dot) / 16.0f;     // ds2
  __syncthreads();
  for (int idx = threadIdx.x; idx < rows * 32; idx += BT) {
    int i = idx >> 5, c4 = idx & 31;
    float ds = da[i];
    float4 a = __ldg(reinterpret_cast<const float4 *>(q1 + (row0 + i) * 128) + c4);
    float4 b = __ldg(reinterpret_cast<const float4 *>(q2 + (row0 + i) * 128) + c4);
    reinterpret_cast<float4 *>(dq2 + (row0 + i) * 128)[c4] = make_float4(ds * a.x, ds * a.y, ds * a.z, ds * a.w);
    reinterpret_cast<float4 *>(dq1 + (row0 + i) * 128)[c4] = make_float4(ds * b.x, ds * b.y, ds * b.z, ds * b.w);
  }
  for (int idx = threadIdx.x; idx < rows * (CAR_C_LAT / 4); idx += BT) {
    int i = idx / (CAR_C_LAT / 4), c4 = idx - i * (CAR_C_LAT / 4);
    float w = aw[i];
    float4 d = *reinterpret_cast<const float4 *>(dz + c4 * 4);
    reinterpret_cast<float4 *>(dv + (row0 + i) * CAR_C_LAT)[c4] = make_float4(w * d.x, w * d.y, w * d.z, w * d.w);
  }
  for (int c = threadIdx.x; c < CAR_C_LAT; c += BT) d_zsum[(size_t)gl * CAR_C_LAT + c] = 2.0f * dz[c];
}

// ---- attention round 1 + expected depth backward (models.py:532-545,577-594) ------------
//   zsum = sum_rows a1 * V ; s1 = <K, Q1> / 16 ; depth = clamp(qinv[2,:3] . sum a1 * clamp(pt) + qinv[2,3], 0, 10)
// out: dK = ds1 * Q1, dQ1 += ds1 * K, dV += a1 * d_zsum
__global__ void __launch_bounds__(BT)
k_attn1_bwd(car_render_args a, int g0, const float *__restrict__ d_zsum, const float *__restrict__ d_depth,
            const float *__restrict__ value, const float *__restrict__ key, const float *__restrict__ q1,
            const float *__restrict__ geom, float *__restrict__ dk, float *__restrict__ dq1,
            float *__restrict__ dv) {
  __shared__ __align__(16) float dz[CAR_C_LAT];
  __shared__ float da[MAX_ROWS], aw[MAX_ROWS];
  __shared__ float red[4];
  __shared__ float w3[8];
  int gl = blockIdx.x, g = g0 + gl;
  int s = g / a.R, r = g - s * a.R;
  int P = a.P, rows = 2 * P;
  size_t row0 = (size_t)gl * rows;
  for (int c = threadIdx.x; c < CAR_C_LAT; c += BT) dz[c] = d_zsum[(size_t)gl * CAR_C_LAT + c];
  for (int i = threadIdx.x; i < rows; i += BT) {
    int j = i / P, k = i - j * P;
    aw[i] = a.at_wt[((size_t)(s * 2 + j) * a.R + r) * P + k];
  }
  __syncthreads();
  value_dots(value + row0 * CAR_C_LAT, dz, rows, da);
  // depth term: recompute the pre-clamp camera-space z exactly like k_attention1
  float dzc = 0.f;
  const float *qi = a.cams.qinv + (size_t)s * 16;
  if (d_depth) {
    if (threadIdx.x < 6) {
      int t = threadIdx.x, j = t / 3, comp = t - j * 3;
      const float *G = geom + (row0 + (size_t)j * P) * CAR_GEOM_STRIDE + G_PTC + comp;
      float w = 0.f;
      for (int k = 0; k < P; ++k) w += aw[j * P + k] * G[(size_t)k * CAR_GEOM_STRIDE];
      w3[t] = w;
    }
    __syncthreads();
    float x = w3[0] + w3[3], y = w3[1] + w3[4], z = w3[2] + w3[5];
    float zc = ((qi[8] * x + qi[9] * y) + qi[10] * z) + qi[11];
    dzc = (zc >= 0.f && zc <= 10.f) ? d_depth[(size_t)s * a.R + r] : 0.f;
  }
  __syncthreads();
  if (dzc != 0.f) {
    for (int i = threadIdx.x; i < rows; i += BT) {
      const float *G = geom + (row0 + i) * CAR_GEOM_STRIDE + G_PTC;
      da[i] += dzc * ((qi[8] * G[0] + qi[9] * G[1]) + qi[10] * G[2]);
    }
  }
  __syncthreads();
  float part = 0.f;
  for (int i = threadIdx.x; i < rows; i += BT) part += aw[i] * da[i];
  float dot = block_sum(part, red);
  for (int i = threadIdx.x; i < rows; i += BT) da[i] = aw[i] * (da[i] - dot) / 16.0f;     // ds1
  __syncthreads();
  for (int idx = threadIdx.x; idx < rows * 32; idx += BT) {
    int i = idx >> 5, c4 = idx & 31;
    float ds = da[i];
    float4 q = __ldg(reinterpret_cast<const float4 *>(q1 + (row0 + i) * 128) + c4);
    float4 k = __ldg(reinterpret_cast<const float4 *>(key + (row0 + i) * 128) + c4);
    reinterpret_cast<float4 *>(dk + (row0 + i) * 128)[c4] = make_float4(ds * q.x, ds * q.y, ds * q.z, ds * q.w);
    float4 *p = reinterpret_cast<float4 *>(dq1 + (row0 + i) * 128) + c4;
    float4 o = *p;
    o.x += ds * k.x; o.y += ds * k.y; o.z += ds * k.z; o.w += ds * k.w;
    *p = o;
  }
  for (int idx = threadIdx.x; idx < rows * (CAR_C_LAT / 4); idx += BT) {
    int i = idx / (CAR_C_LAT / 4), c4 = idx - i * (CAR_C_LAT / 4);
    float w = aw[i];
    float4 d = *reinterpret_cast<const float4 *>(dz + c4 * 4);
    float4 *p = reinterpret_cast<float4 *>(dv + (row0 + i) * CAR_C_LAT) + c4;
    float4 o = *p;
    o.x += w * d.x; o.y += w * d.y; o.z += w * d.z; o.w += w * d.w;
    *p = o;
  }
}

// per-ray sum of a [rows][128] gradient over the ray's 2P rows (d of the per-ray row bias)
__global__ void __launch_bounds__(128)
k_raysum128(const float *__restrict__ src, int rows_per_ray, float *__restrict__ dst) {
  int gl = blockIdx.x, c = threadIdx.x;
  const float *p = src + (size_t)gl * rows_per_ray * 128 + c;
  float s0 = 0.f, s1 = 0.f;
  int i = 0;
  for (; i + 1 < rows_per_ray; i += 2) { s0 += p[(size_t)i * 128]; s1 += p[(size_t)(i + 1) * 128]; }
  if (i < rows_per_ray) s0 += p[(size_t)i * 128];
  dst[(size_t)gl * 128 + c] = s0 + s1;
}

// d_rgb3 = d_rgb * valid (models.py:615-616);  dx = (d_rgb3 · W_out) * (x3 > 0)  (lin_out after ReLU)
__global__ void __launch_bounds__(128)
k_rgb_bwd(car_render_args a, int g0, int nr, const float *__restrict__ d_rgb, const uint8_t *__restrict__ overlap,
          const float *__restrict__ w_out /*[3][128]*/, const float *__restrict__ x3,
          float *__restrict__ d_rgb3 /*[nr][4]*/, float *__restrict__ dx /*[nr][128]*/) {
  int gl = blockIdx.x, c = threadIdx.x;
  if (gl >= nr) return;
  int g = g0 + gl;
  float valid = (overlap[gl * 2] | overlap[gl * 2 + 1]) ? 1.f : 0.f;
  float d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) d[k] = d_rgb ? d_rgb[(size_t)g * 3 + k] * valid : 0.f;
  if (c < 4) d_rgb3[(size_t)gl * 4 + c] = c < 3 ? d[c] : 0.f;
  float v = (d[0] * w_out[c] + d[1] * w_out[128 + c]) + d[2] * w_out[256 + c];
  dx[(size_t)gl * 128 + c] = x3[(size_t)gl * 128 + c] > 0.f ? v : 0.f;
}

// fp32 rows -> the bf16 hi + lo operand copies the tcgen05 gradient GEMMs read: row-major ([M][W], tight; data
// gradient) and / or transposed ([W][M]; weight gradient, the row dimension becomes the contraction), plus the
// column sums (bias gradient).  64 rows x 32 columns per CTA.
__global__ void __launch_bounds__(256)
k_split_ops(const float *__restrict__ src, int ld, int M, int W, uint16_t *__restrict__ hi, uint16_t *__restrict__ lo,
            uint16_t *__restrict__ hiT, uint16_t *__restrict__ loT, float *__restrict__ db) {
  __shared__ uint32_t tile[32][65];                 // [column][row]: hi | lo << 16
  __shared__ float csum[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 32, c = c0 + tx;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = ty + 8 * i, r = r0 + rl;
    const bool in = r < M && c < W;
    const float v = in ? src[(size_t)r * ld + c] : 0.f;
    s += v;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    const uint32_t hb = __bfloat16_as_ushort(h), lb = __bfloat16_as_ushort(l);
    tile[tx][rl] = hb | (lb << 16);
    if (hi && in) { hi[(size_t)r * W + c] = (uint16_t)hb; lo[(size_t)r * W + c] = (uint16_t)lb; }
  }
  if (db) csum[ty][tx] = s;
  __syncthreads();
  if (db && ty == 0 && c < W) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += csum[j][tx];
    atomicAdd(db + c, t);
  }
  if (hiT) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cl = ty + 8 * j, cc = c0 + cl;
      if (cc >= W) continue;
      const uint32_t p0 = tile[cl][2 * tx], p1 = tile[cl][2 * tx + 1];
      const int r = r0 + 2 * tx;
      const size_t o = (size_t)cc * M + r;
      if (r + 1 < M) {                              // M is even: 4-byte aligned pair
        *reinterpret_cast<uint32_t *>(hiT + o) = (p0 & 0xffffu) | (p1 << 16);
        *reinterpret_cast<uint32_t *>(loT + o) = (p0 >> 16) | (p1 & 0xffff0000u);
      } else if (r < M) {
        hiT[o] = (uint16_t)(p0 & 0xffffu); loT[o] = (uint16_t)(p0 >> 16);
      }
    }
  }
}

void launch_split_ops(const float *src, int ld, int M, int W, uint16_t *hi, uint16_t *lo, uint16_t *hiT, uint16_t *loT,
                      float *db, cudaStream_t st) {
  if (M <= 0 || W <= 0) return;
  dim3 grid((M + 63) / 64, (W + 31) / 32);
  prof_pre(-1, st);
  k_split_ops<<<grid, 256, 0, st>>>(src, ld, M, W, hi, lo, hiT, loT, db);
  prof_post(st);
  count_launch();
}

// ---- backward workspace ------------------------------------------------------------------
struct BwWs {
  // transposed weights [K][N]
  float *enc1T, *enc2T, *valueT, *key1T, *key2T, *qry2T, *rep2T, *rep1gT, *enclatT;
  float *phizT[3], *fc0T[3], *fc1T[3];
  // recomputed colour-MLP activations, per ray
  float *xs[3], *nets[3], *x3, *xrun;
  // per-ray gradients
  float *d_rgb3, *dx, *dnet, *d_zfin, *d_zsum, *drowbias, *dg;
  // per-sample gradients
  float *ds, *dq1, *dhid, *dv, *dinterp, *dh1, *dxin;
  // tensor-core precision: bf16 hi / lo operand copies of the current gradient (row-major and transposed), of the
  // current saved activation (transposed), of local_coords (transposed, used twice) and of the transposed weights
  uint16_t *dy_hi, *dy_lo, *dyT_hi, *dyT_lo, *aT_hi, *aT_lo, *locT_hi, *locT_lo;
  uint16_t *enc1T_hi, *enc1T_lo, *enc2T_hi, *enc2T_lo, *valueT_hi, *valueT_lo, *key1T_hi, *key1T_lo;
  uint16_t *key2T_hi, *key2T_lo, *qry2T_hi, *qry2T_lo, *rep2T_hi, *rep2T_lo;
  size_t bytes;
};

BwWs carve_bw(char *base, int precision, int P, int rays) {
  BwWs w;
  memset(&w, 0, sizeof(w));
  size_t off = 0;
  auto take = [&](size_t n) { char *p = base ? base + off : nullptr; off += align_up(n * 4); return (float *)p; };
  size_t rows = (size_t)rays * 2 * P;
  w.enc1T = take((size_t)CAR_C_FEAT * CAR_C_FEAT);      // only the 576 feature columns get a data gradient
  w.enc2T = take((size_t)CAR_C_FEAT * CAR_C_LAT);
  w.valueT = take((size_t)CAR_C_FEAT * CAR_C_LAT);
  w.key1T = take((size_t)CAR_C_FEAT * 128);
  w.key2T = take(128 * 128); w.qry2T = take(128 * 128); w.rep2T = take(128 * 128); w.rep1gT = take(128 * 128);
  w.enclatT = take((size_t)CAR_C_LAT * 128);
  for (int i = 0; i < 3; ++i) { w.phizT[i] = take((size_t)CAR_C_LAT * 128); w.fc0T[i] = take(128 * 128); w.fc1T[i] = take(128 * 128); }
  for (int i = 0; i < 3; ++i) { w.xs[i] = take((size_t)rays * 128); w.nets[i] = take((size_t)rays * 128); }
  w.x3 = take((size_t)rays * 128); w.xrun = take((size_t)rays * 128);
  w.d_rgb3 = take((size_t)rays * 4); w.dx = take((size_t)rays * 128); w.dnet = take((size_t)rays * 128);
  w.d_zfin = take((size_t)rays * CAR_C_LAT); w.d_zsum = take((size_t)rays * CAR_C_LAT);
  w.drowbias = take((size_t)rays * 128); w.dg = take((size_t)rays * 128);
  w.ds = take(rows * 128); w.dq1 = take(rows * 128); w.dhid = take(rows * 128);
  w.dv = take(rows * CAR_C_LAT); w.dinterp = take(rows * CAR_C_FEAT);
  w.dh1 = take(rows * 2 * CAR_C_FEAT); w.dxin = take(rows * 2 * CAR_C_FEAT);
  if (precision == CAR_PREC_FP32_3XBF16) {
    auto take16 = [&](size_t n) { return (uint16_t *)take((n + 1) / 2); };
    const size_t big = rows * 2 * CAR_C_FEAT;
    w.dy_hi = take16(big); w.dy_lo = take16(big); w.dyT_hi = take16(big); w.dyT_lo = take16(big);
    w.aT_hi = take16(rows * 2 * CAR_K_ENC); w.aT_lo = take16(rows * 2 * CAR_K_ENC);
    w.locT_hi = take16(rows * 16); w.locT_lo = take16(rows * 16);
    const size_t F = CAR_C_FEAT, L = CAR_C_LAT;
    w.enc1T_hi = take16(F * F); w.enc1T_lo = take16(F * F);
    w.enc2T_hi = take16(F * L); w.enc2T_lo = take16(F * L);
    w.valueT_hi = take16(F * L); w.valueT_lo = take16(F * L);
    w.key1T_hi = take16(F * 128); w.key1T_lo = take16(F * 128);
    w.key2T_hi = take16(128 * 128); w.key2T_lo = take16(128 * 128);
    w.qry2T_hi = take16(128 * 128); w.qry2T_lo = take16(128 * 128);
    w.rep2T_hi = take16(128 * 128); w.rep2T_lo = take16(128 * 128);
  }
  w.bytes = off;
  return w;
}

GemmEpi plain() { GemmEpi e; e.bias = nullptr; e.row_bias = nullptr; e.rows_per_group = 1; e.relu_in = 0; e.relu_out = 0; e.accumulate = 0; return e; }
GemmEpi masked(const float *mask, int ld, int accumulate = 0) { GemmEpi e = plain(); e.mask = mask; e.ldmask = ld; e.accumulate = accumulate; return e; }
GemmEpi accum() { GemmEpi e = plain(); e.accumulate = 1; return e; }

// data gradient:  dA[M][K] = dY[M][N] · W[N][K]   computed as gemm(dY, W^T) with WT = [K][N]
void dgrad(const float *dY, int ldy, const float *WT, int N, int K, float *dA, int lda, int M, const GemmEpi &e,
           cudaStream_t st) {
  StageScope ss(CAR_ST_BWD_DGRAD);
  launch_gemm_simt(dY, ldy, WT, N, dA, lda, M, K, N, e, st);
}

void wgrad(const float *dY, int ldy, const float *A, int lda, const car_mat_grad &g, int ldw, int M, int N, int K,
           int relu_a, cudaStream_t st) {
  if (!g.w) return;
  StageScope ss(CAR_ST_BWD_WGRAD);
  launch_wgrad_simt(dY, ldy, A, lda, g.w, ldw, g.bias, M, N, K, relu_a, st);
}

}  // namespace
}  // namespace car

using namespace car;

extern "C" {

size_t car_backward_workspace_bytes(int precision, int P, int rays) { return carve_bw(nullptr, precision, P, rays).bytes; }

int car_unpack_features(const float *nhwc, float *nchw, int bn, int C, int h, int w, void *stream) {
  if (!nchw || !nhwc || bn <= 0 || C <= 0 || h <= 0 || w <= 0) { set_error("car_unpack_features: bad argument"); return -1; }
  launch_unpack_features(nhwc, nchw, bn, C, h, w, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("car_unpack_features: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

int car_render_backward(const car_backward_args *pb) {
  count_launch(-car_last_launch_count());
  if (!pb || !pb->fwd) { set_error("car_render_backward: null args"); return -1; }
  const car_backward_args &b = *pb;
  const car_render_args &a = *b.fwd;
  if (b.abi_version != CAR_ABI_VERSION || a.abi_version != CAR_ABI_VERSION) { set_error("ABI version mismatch"); return -2; }
  if (!a.train || a.precision == CAR_PREC_BF16) { set_error("car_render_backward needs the arguments of a train=1 forward (CAR_PREC_FP32_SIMT or CAR_PREC_FP32_3XBF16)"); return -12; }
  if (b.precision != CAR_PREC_FP32_SIMT && b.precision != CAR_PREC_FP32_3XBF16) { set_error("car_render_backward: precision %d (0 or 1)", b.precision); return -5; }
  const bool tc = b.precision == CAR_PREC_FP32_3XBF16;
  if (!b.d_rgb && !b.d_depth_ray) { set_error("car_render_backward: no cotangent given"); return -6; }
  if (!b.workspace) { set_error("car_render_backward: null workspace"); return -6; }
  const int g0 = a.ray_begin, g1 = a.ray_end, nr = g1 - g0;
  if (nr <= 0) return 0;
  if (carve_bw(nullptr, b.precision, a.P, nr).bytes > b.workspace_bytes) { set_error("backward workspace too small: %zu bytes", b.workspace_bytes); return -8; }
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) { set_error("no CUDA device (there is no CPU fallback)"); return -7; }
  const bool any_feat = b.d_feat[0] || b.d_feat[1] || b.d_feat[2];
  if (any_feat && !(b.d_feat[0] && b.d_feat[1] && b.d_feat[2])) { set_error("d_feat: give all three levels or none"); return -6; }

  const Workspace f = carve((char *)a.workspace, a.precision, a.P, nr, 0, 1);     // saved activations
  const BwWs w = carve_bw((char *)b.workspace, b.precision, a.P, nr);
  const car_weights &W = a.weights;
  const car_weight_grads &G = b.grads;
  cudaStream_t st = (cudaStream_t)b.stream;
  const int P = a.P, rows = nr * 2 * P, rows2 = rows * 2;
  const int L = CAR_C_LAT, F = CAR_C_FEAT;
  StageScope sc(CAR_ST_BACKWARD);

  // transposed weights for the data gradients
  launch_transpose(W.enc1.f32, F, F, CAR_K_ENC, w.enc1T, st);           // [576 k][576 n]  (feature columns only)
  launch_transpose(W.enc2.f32, L, F, F, w.enc2T, st);                   // [576][288]
  launch_transpose(W.value.f32, L, F, F, w.valueT, st);
  launch_transpose(W.key1.f32, 128, F, F, w.key1T, st);                 // [576][128]
  launch_transpose(W.key2.f32, 128, 128, 128, w.key2T, st);
  launch_transpose(W.qry2.f32, 128, 128, 128, w.qry2T, st);
  launch_transpose(W.rep2.f32, 128, 128, 128, w.rep2T, st);
  launch_transpose(W.rep1_g.f32, 128, 128, 128, w.rep1gT, st);
  launch_transpose(W.enc_lat.f32, 128, L, L, w.enclatT, st);            // [288][128]
  for (int i = 0; i < 3; ++i) {
    launch_transpose(W.phi_z[i].f32, 128, L, L, w.phizT[i], st);
    launch_transpose(W.phi_fc0[i].f32, 128, 128, 128, w.fc0T[i], st);
    launch_transpose(W.phi_fc1[i].f32, 128, 128, 128, w.fc1T[i], st);
  }

  if (tc) {   // operand copies of the transposed weights the per-sample data gradients multiply by
    launch_split_rows(w.enc1T, F, w.enc1T_hi, w.enc1T_lo, F, F, st);
    launch_split_rows(w.enc2T, L, w.enc2T_hi, w.enc2T_lo, F, L, st);
    launch_split_rows(w.valueT, L, w.valueT_hi, w.valueT_lo, F, L, st);
    launch_split_rows(w.key1T, 128, w.key1T_hi, w.key1T_lo, F, 128, st);
    launch_split_rows(w.key2T, 128, w.key2T_hi, w.key2T_lo, 128, 128, st);
    launch_split_rows(w.qry2T, 128, w.qry2T_hi, w.qry2T_lo, 128, 128, st);
    launch_split_rows(w.rep2T, 128, w.rep2T_hi, w.rep2T_lo, 128, 128, st);
  }
  // The per-sample layers (M = rows or 2*rows) in the tensor-core precision: every gradient GEMM runs on tcgen05
  // (car_gemm_umma.cu) with hi + lo bf16 operands.  ops(): operand copies of the gradient dY that the layer's GEMMs read
  // (+ its bias gradient); act(): transposed copy of the saved activation; wg(): dW += dY^T·A; dg(): dA = dY·W.
  int rc = 0;
  auto ops = [&](const float *dY, int ld, int M, int N, bool rowmajor, const car_mat_grad &g) {
    StageScope ss(CAR_ST_BWD_OPS);
    launch_split_ops(dY, ld, M, N, rowmajor ? w.dy_hi : nullptr, rowmajor ? w.dy_lo : nullptr, g.w ? w.dyT_hi : nullptr,
                     g.w ? w.dyT_lo : nullptr, g.w ? g.bias : nullptr, st);
  };
  auto act = [&](const float *A, int ld, int M, int K) {
    StageScope ss(CAR_ST_BWD_OPS);
    launch_split_ops(A, ld, M, K, nullptr, nullptr, w.aT_hi, w.aT_lo, nullptr, st);
  };
  auto wg = [&](const uint16_t *aT_hi, const uint16_t *aT_lo, const car_mat_grad &g, int ldw, int M, int N, int K) {
    if (!g.w || rc) return;
    StageScope ss(CAR_ST_BWD_WGRAD);
    for (int k0 = 0; k0 < K && !rc; k0 += CAR_C_FEAT) {      // K = 592: the 576 feature columns, then the 16 others
      const int kn = K - k0 < CAR_C_FEAT ? K - k0 : CAR_C_FEAT;
      UmmaOut o; o.f32 = g.w + k0; o.hi = nullptr; o.lo = nullptr; o.ldc = ldw; o.atomic = 1;
      rc = launch_gemm_umma(w.dyT_hi, w.dyT_lo, M, aT_hi + (size_t)k0 * M, aT_lo + (size_t)k0 * M, M, N, kn, M, 1, plain(), o, st);
    }
  };
  auto dg = [&](const uint16_t *wT_hi, const uint16_t *wT_lo, int N, int K, float *dA, int lda, int M, const GemmEpi &e) {
    if (rc) return;
    StageScope ss(CAR_ST_BWD_DGRAD);
    UmmaOut o; o.f32 = dA; o.f32_add = e.accumulate ? dA : nullptr; o.hi = nullptr; o.lo = nullptr; o.ldc = lda;
    rc = launch_gemm_umma(w.dy_hi, w.dy_lo, N, wT_hi, wT_lo, N, M, K, N, 1, e, o, st);
  };

  // ---- colour MLP: recompute the layer inputs (resnet_block_fc.py:146-166), then go back ----
  {
    GemmEpi e = plain();
    e.bias = W.phi_in.bias;
    launch_gemm_simt(f.c18, 32, W.phi_in.f32, 32, w.xrun, 128, nr, 128, 32, e, st);
    for (int i = 0; i < 3; ++i) {
      GemmEpi ez = plain(); ez.bias = W.phi_z[i].bias; ez.accumulate = 1;
      launch_gemm_simt(f.zfin, L, W.phi_z[i].f32, L, w.xrun, 128, nr, 128, L, ez, st);
      cudaMemcpyAsync(w.xs[i], w.xrun, (size_t)nr * 128 * 4, cudaMemcpyDeviceToDevice, st);
      GemmEpi e0 = plain(); e0.bias = W.phi_fc0[i].bias; e0.relu_in = 1;
      launch_gemm_simt(w.xrun, 128, W.phi_fc0[i].f32, 128, w.nets[i], 128, nr, 128, 128, e0, st);
      GemmEpi e1 = plain(); e1.bias = W.phi_fc1[i].bias; e1.relu_in = 1; e1.accumulate = 1;
      launch_gemm_simt(w.nets[i], 128, W.phi_fc1[i].f32, 128, w.xrun, 128, nr, 128, 128, e1, st);
    }
    cudaMemcpyAsync(w.x3, w.xrun, (size_t)nr * 128 * 4, cudaMemcpyDeviceToDevice, st);
  }
  prof_pre(-1, st);
  k_rgb_bwd<<<nr, 128, 0, st>>>(a, g0, nr, b.d_rgb, f.overlap, W.phi_out.f32, w.x3, w.d_rgb3, w.dx);
  prof_post(st);
  count_launch();
  wgrad(w.d_rgb3, 4, w.x3, 128, G.phi_out, 128, nr, 3, 128, 1, st);
  for (int i = 2; i >= 0; --i) {
    // x_out = x' + fc_1(relu(net)),  net = fc_0(relu(x')),  x' = x_in + lin_z[i](z)
    wgrad(w.dx, 128, w.nets[i], 128, G.phi_fc1[i], 128, nr, 128, 128, 1, st);
    dgrad(w.dx, 128, w.fc1T[i], 128, 128, w.dnet, 128, nr, masked(w.nets[i], 128), st);
    wgrad(w.dnet, 128, w.xs[i], 128, G.phi_fc0[i], 128, nr, 128, 128, 1, st);
    dgrad(w.dnet, 128, w.fc0T[i], 128, 128, w.dx, 128, nr, masked(w.xs[i], 128, 1), st);
    wgrad(w.dx, 128, f.zfin, L, G.phi_z[i], L, nr, 128, L, 0, st);
    dgrad(w.dx, 128, w.phizT[i], 128, L, w.d_zfin, L, nr, i == 2 ? plain() : accum(), st);
  }
  wgrad(w.dx, 128, f.c18, 32, G.phi_in, 32, nr, 128, 32, 0, st);

  // ---- attention round 2 and the repeat-query MLP (models.py:548-565) ----
  prof_pre(-1, st);
  k_attn2_bwd<<<nr, BT, 0, st>>>(P, w.d_zfin, f.value, f.att2, f.q1, f.q2, w.ds, w.dq1, w.dv, w.d_zsum);
  prof_post(st);
  count_launch();
  if (tc) {
    ops(w.ds, 128, rows, 128, true, G.rep2);
    act(f.hid_r, 128, rows, 128);
    wg(w.aT_hi, w.aT_lo, G.rep2, 128, rows, 128, 128);
    dg(w.rep2T_hi, w.rep2T_lo, 128, 128, w.dhid, 128, rows, masked(f.hid_r, 128));
    ops(w.dhid, 128, rows, 128, false, G.rep1_loc);
    launch_split_ops(f.geom + G_LOCAL, CAR_GEOM_STRIDE, rows, 16, nullptr, nullptr, w.locT_hi, w.locT_lo, nullptr, st);
    wg(w.locT_hi, w.locT_lo, G.rep1_loc, 16, rows, 128, 16);
    if (rc) return rc;
  } else {
    wgrad(w.ds, 128, f.hid_r, 128, G.rep2, 128, rows, 128, 128, 0, st);
    dgrad(w.ds, 128, w.rep2T, 128, 128, w.dhid, 128, rows, masked(f.hid_r, 128), st);
    wgrad(w.dhid, 128, f.geom + G_LOCAL, CAR_GEOM_STRIDE, G.rep1_loc, 16, rows, 128, 16, 0, st);
  }
  prof_pre(-1, st);
  k_raysum128<<<nr, 128, 0, st>>>(w.dhid, 2 * P, w.drowbias);
  prof_post(st);
  count_launch();
  wgrad(w.drowbias, 128, f.g, 128, G.rep1_g, 128, nr, 128, 128, 0, st);
  dgrad(w.drowbias, 128, w.rep1gT, 128, 128, w.dg, 128, nr, plain(), st);
  wgrad(w.dg, 128, f.zsum, L, G.enc_lat, L, nr, 128, L, 0, st);
  dgrad(w.dg, 128, w.enclatT, 128, L, w.d_zsum, L, nr, accum(), st);

  // ---- attention round 1 + depth (models.py:532-545,577-594) ----
  prof_pre(-1, st);
  k_attn1_bwd<<<nr, BT, 0, st>>>(a, g0, w.d_zsum, b.d_depth_ray, f.value, f.key, f.q1, f.geom, w.ds, w.dq1, w.dv);
  prof_post(st);
  count_launch();
  if (tc) {
    // geometric query (:529)
    ops(w.dq1, 128, rows, 128, true, G.qry2);
    act(f.hid_q, 128, rows, 128);
    wg(w.aT_hi, w.aT_lo, G.qry2, 128, rows, 128, 128);
    dg(w.qry2T_hi, w.qry2T_lo, 128, 128, w.dhid, 128, rows, masked(f.hid_q, 128));
    ops(w.dhid, 128, rows, 128, false, G.qry1);
    wg(w.locT_hi, w.locT_lo, G.qry1, 16, rows, 128, 16);
    // key (:491)
    ops(w.ds, 128, rows, 128, true, G.key2);
    act(f.hid, 128, rows, 128);
    wg(w.aT_hi, w.aT_lo, G.key2, 128, rows, 128, 128);
    dg(w.key2T_hi, w.key2T_lo, 128, 128, w.dhid, 128, rows, masked(f.hid, 128));
    ops(w.dhid, 128, rows, 128, true, G.key1);
    act(f.interp, F, rows, F);                                       // interp^T serves key_map and latent_value
    wg(w.aT_hi, w.aT_lo, G.key1, F, rows, 128, F);
    dg(w.key1T_hi, w.key1T_lo, 128, F, w.dinterp, F, rows, plain());
    ops(w.dv, L, rows, L, true, G.value);
    wg(w.aT_hi, w.aT_lo, G.value, F, rows, L, F);
    dg(w.valueT_hi, w.valueT_lo, L, F, w.dinterp, F, rows, accum());
    // per-view encoder MLP (models.py:333-342)
    ops(w.dinterp, L, rows2, L, true, G.enc2);
    act(f.h1, F, rows2, F);
    wg(w.aT_hi, w.aT_lo, G.enc2, F, rows2, L, F);
    dg(w.enc2T_hi, w.enc2T_lo, L, F, w.dh1, F, rows2, masked(f.h1, F));
    ops(w.dh1, F, rows2, F, any_feat, G.enc1);
    act(f.x, CAR_K_ENC, rows2, CAR_K_ENC);
    wg(w.aT_hi, w.aT_lo, G.enc1, CAR_K_ENC, rows2, F, CAR_K_ENC);
    if (any_feat) {
      dg(w.enc1T_hi, w.enc1T_lo, F, F, w.dxin, F, rows2, plain());
      if (!rc) { StageScope ss(CAR_ST_BWD_SCATTER); launch_gather_backward(a, g0, g1, f.geom, w.dxin, b.d_feat, st); }
    }
    if (rc) return rc;
  } else {
    // geometric query  Q1 = query_embed_2(relu(query_embed(local)))   (:529)
    wgrad(w.dq1, 128, f.hid_q, 128, G.qry2, 128, rows, 128, 128, 0, st);
    dgrad(w.dq1, 128, w.qry2T, 128, 128, w.dhid, 128, rows, masked(f.hid_q, 128), st);
    wgrad(w.dhid, 128, f.geom + G_LOCAL, CAR_GEOM_STRIDE, G.qry1, 16, rows, 128, 16, 0, st);
    // key  K = key_map_2(relu(key_map(interp)))   (:491)
    wgrad(w.ds, 128, f.hid, 128, G.key2, 128, rows, 128, 128, 0, st);
    dgrad(w.ds, 128, w.key2T, 128, 128, w.dhid, 128, rows, masked(f.hid, 128), st);
    wgrad(w.dhid, 128, f.interp, F, G.key1, F, rows, 128, F, 0, st);
    wgrad(w.dv, L, f.interp, F, G.value, F, rows, L, F, 0, st);
    dgrad(w.dhid, 128, w.key1T, 128, F, w.dinterp, F, rows, plain(), st);
    dgrad(w.dv, L, w.valueT, L, F, w.dinterp, F, rows, accum(), st);

    // ---- per-view encoder MLP (models.py:333-342): rows*2 view-rows of 288 / 576 / 592 ----
    wgrad(w.dinterp, L, f.h1, F, G.enc2, F, rows2, L, F, 0, st);
    dgrad(w.dinterp, L, w.enc2T, L, F, w.dh1, F, rows2, masked(f.h1, F), st);
    wgrad(w.dh1, F, f.x, CAR_K_ENC, G.enc1, CAR_K_ENC, rows2, F, CAR_K_ENC, 0, st);
    if (any_feat) {
      dgrad(w.dh1, F, w.enc1T, F, F, w.dxin, F, rows2, plain(), st);
      { StageScope ss(CAR_ST_BWD_SCATTER); launch_gather_backward(a, g0, g1, f.geom, w.dxin, b.d_feat, st); }
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("backward kernel launch failed: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

}  // extern "C"
