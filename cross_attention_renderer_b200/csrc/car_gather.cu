// Feature-map packing (NCHW -> NHWC) and the dual bilinear gather that builds the
// per-sample encoder inputs.
//
// Reference arithmetic restated here (PyTorch CUDA grid_sample formulas,
// ATen/native/cuda/GridSampler.cuh:23-31,56-59,139-168, as used at reference
// models.py:278 (padding 'border') and models.py:317 (padding 'zeros')).
#include <cuda_bf16.h>
#include <math.h>

#include "car_common.cuh"

namespace car {
namespace {

// ---------------------------------------------------------------------------
// NCHW fp32 -> NHWC {fp32,bf16}: 32x32 smem-tiled transpose of the (C, h*w)
// matrix of every image; coalesced on both sides.
// ---------------------------------------------------------------------------
template <bool BF16>
__global__ void k_pack_features(const float *__restrict__ src, void *__restrict__ dst, int C,
                                int hw) {
  __shared__ float tile[32][33];
  int img = blockIdx.z;
  int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float *s = src + (size_t)img * C * hw;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < hw) ? s[(size_t)c * hw + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, c = c0 + threadIdx.x;
    if (p < hw && c < C) {
      float v = tile[threadIdx.x][i];
      size_t o = ((size_t)img * hw + p) * C + c;
      if (BF16) reinterpret_cast<__nv_bfloat16 *>(dst)[o] = __float2bfloat16_rn(v);
      else reinterpret_cast<float *>(dst)[o] = v;
    }
  }
}

// ---------------------------------------------------------------------------
// Bilinear tap set-up.
// ---------------------------------------------------------------------------
struct Taps {
  int off[4];      // texel offsets (y*w+x), -1 if the tap is out of bounds
  float w[4];      // nw, ne, sw, se
};

__device__ __forceinline__ float downgrade(float x) {
  // safe_downgrade_to_int_range (GridSampler.cuh:139-147)
  if (x > 2147483646.0f || x < -2147483648.0f || !isfinite(x)) return -100.0f;
  return x;
}

__device__ __forceinline__ Taps make_taps(float gx, float gy, int w, int h, bool border) {
  // separately rounded ops (no FMA contraction) so the integer taps are a pure function of
  // the float sample coordinate, identical to the oracle's unfused evaluation
  float ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)w), 1.f), 2.f);
  float iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)h), 1.f), 2.f);
  if (border) {
    ix = fminf((float)(w - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(h - 1), fmaxf(iy, 0.f));
  }
  ix = downgrade(ix);
  iy = downgrade(iy);
  float fx0 = floorf(ix), fy0 = floorf(iy);
  int x0 = (int)fx0, y0 = (int)fy0;
  float wx1 = __fsub_rn(ix, fx0), wx0 = __fsub_rn(__fadd_rn(fx0, 1.f), ix);
  float wy1 = __fsub_rn(iy, fy0), wy0 = __fsub_rn(__fadd_rn(fy0, 1.f), iy);
  Taps t;
  t.w[0] = wx0 * wy0; t.w[1] = wx1 * wy0; t.w[2] = wx0 * wy1; t.w[3] = wx1 * wy1;
  bool xin0 = x0 >= 0 && x0 < w, xin1 = x0 + 1 >= 0 && x0 + 1 < w;
  bool yin0 = y0 >= 0 && y0 < h, yin1 = y0 + 1 >= 0 && y0 + 1 < h;
  t.off[0] = (xin0 && yin0) ? y0 * w + x0 : -1;
  t.off[1] = (xin1 && yin0) ? y0 * w + x0 + 1 : -1;
  t.off[2] = (xin0 && yin1) ? (y0 + 1) * w + x0 : -1;
  t.off[3] = (xin1 && yin1) ? (y0 + 1) * w + x0 + 1 : -1;
  return t;
}

__device__ __forceinline__ float4 ld4(const float *p) {
  return __ldg(reinterpret_cast<const float4 *>(p));
}
__device__ __forceinline__ float4 ld4(const __nv_bfloat16 *p) {
  uint2 u = __ldg(reinterpret_cast<const uint2 *>(p));
  float4 f;
  f.x = __uint_as_float(u.x << 16);
  f.y = __uint_as_float(u.x & 0xffff0000u);
  f.z = __uint_as_float(u.y << 16);
  f.w = __uint_as_float(u.y & 0xffff0000u);
  return f;
}

// out += val * weight in tap order nw, ne, sw, se (same accumulation as
// grid_sampler_2d_kernel).
template <typename T>
__device__ __forceinline__ float4 lerp4(const T *__restrict__ img, int C, int c, const Taps &t) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (t.off[k] >= 0) {
      float4 v = ld4(img + (size_t)t.off[k] * C + c);
      acc.x = fmaf(v.x, t.w[k], acc.x);
      acc.y = fmaf(v.y, t.w[k], acc.y);
      acc.z = fmaf(v.z, t.w[k], acc.z);
      acc.w = fmaf(v.w, t.w[k], acc.w);
    }
  }
  return acc;
}

__device__ __forceinline__ void split_bf16(float v, uint16_t &hi, uint16_t &lo) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  float r = v - __bfloat162float(h);
  __nv_bfloat16 l = __float2bfloat16_rn(r);
  hi = __bfloat16_as_ushort(h);
  lo = __bfloat16_as_ushort(l);
}

template <typename OUT>
__device__ __forceinline__ void store4(OUT *of, uint16_t *oh, uint16_t *ol, size_t o, float4 v);

template <>
__device__ __forceinline__ void store4<float>(float *of, uint16_t *, uint16_t *, size_t o,
                                              float4 v) {
  *reinterpret_cast<float4 *>(of + o) = v;
}
template <>
__device__ __forceinline__ void store4<uint16_t>(uint16_t *, uint16_t *oh, uint16_t *ol, size_t o,
                                                 float4 v) {
  uint16_t h[4], l[4];
  split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]);
  split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
  uint2 ph = make_uint2(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16));
  *reinterpret_cast<uint2 *>(oh + o) = ph;
  if (ol) {
    uint2 pl = make_uint2(l[0] | ((uint32_t)l[1] << 16), l[2] | ((uint32_t)l[3] << 16));
    *reinterpret_cast<uint2 *>(ol + o) = pl;
  }
}

// One warp per sample row.  X[row][view][592]:
//   [0,256) level 0, [256,512) level 1, [512,576) level 2, [576,579) tanh(pt_view/5), rest 0.
// For a context-j row, the primary (border) gather reads view j and fills slot j; the
// cross-view (zeros) gather reads view 1-j at the reprojected point and fills slot 1-j
// (reference models.py:330-342).
template <typename FT, typename OUT>
__global__ void __launch_bounds__(256)
k_gather(car_render_args a, int g0, int g1, const float *__restrict__ geom, OUT *out_f,
         uint16_t *out_hi, uint16_t *out_lo) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  long nrows = (long)(g1 - g0) * 2 * a.P;
  if (warp >= nrows) return;
  int rj = warp / a.P;
  int j = rj & 1;
  int g = g0 + (rj >> 1);
  int s = g / a.R;
  const float *G = geom + (size_t)warp * CAR_GEOM_STRIDE;
  float gx = G[G_GX], gy = G[G_GY], gxc = G[G_GXC], gyc = G[G_GYC];
  size_t row_base = (size_t)warp * 2 * CAR_K_ENC;
  size_t own_base = row_base + (size_t)j * CAR_K_ENC;
  size_t oth_base = row_base + (size_t)(1 - j) * CAR_K_ENC;
  int chan0 = 0;
#pragma unroll
  for (int lvl = 0; lvl < 3; ++lvl) {
    int C = lvl == 2 ? 64 : 256;
    int h = lvl == 0 ? a.H / 4 : (lvl == 1 ? a.H / 2 : a.H);
    int w = lvl == 0 ? a.W / 4 : (lvl == 1 ? a.W / 2 : a.W);
    const FT *base = reinterpret_cast<const FT *>(a.feat[lvl]);
    const FT *own = base + (size_t)(s * 2 + j) * h * w * C;
    const FT *oth = base + (size_t)(s * 2 + (1 - j)) * h * w * C;
    Taps to = make_taps(gx, gy, w, h, true);
    Taps tc = make_taps(gxc, gyc, w, h, false);
    for (int c = lane * 4; c < C; c += 128) {
      float4 vo = lerp4(own, C, c, to);
      float4 vc = lerp4(oth, C, c, tc);
      store4<OUT>(out_f, out_hi, out_lo, own_base + chan0 + c, vo);
      store4<OUT>(out_f, out_hi, out_lo, oth_base + chan0 + c, vc);
    }
    chan0 += C;
  }
  // tail: tanh(pt_view/5) then zero pad; 16 columns = 4 lanes x float4 per view
  if (lane < 8) {
    int v = lane >> 2, q = lane & 3;
    const float *T = G + (v == 0 ? G_T0 : G_T1);
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q == 0) { t.x = T[0]; t.y = T[1]; t.z = T[2]; }
    store4<OUT>(out_f, out_hi, out_lo, row_base + (size_t)v * CAR_K_ENC + 576 + q * 4, t);
  }
}

// General branches (car_b200.h, car_general_args): one warp per sample row, fp32 output.
//   parts == 1, xw == 592 (n_view = 1): [own 576 | tanh(pt/5) 3 | tanh(pt/100) 3 | 0]
//   parts == 1, xw == 576 (no_latent_concat): [own 576]
//   parts == 2 (n_view = 2):  slot j = own view (border taps), slot 1-j = the other view at the re-projected point
//   parts == 3 (n_view = 3):  part 0 = own view; parts 1, 2 = the other views in ascending order at GG_C0 / GG_C1
// each part followed by its tanh triple and zero padding to 592 (models.py:330-342, 436-446, 483-485).
__global__ void __launch_bounds__(256)
k_gather_g(car_general_args a, GenShape gs, int g0, int g1, const float *__restrict__ geom, float *__restrict__ out) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  const int n = gs.n;
  long nrows = (long)(g1 - g0) * n * a.P;
  if (warp >= nrows) return;
  int rj = warp / a.P;
  int j = rj % n;
  int g = g0 + rj / n;
  int s = g / a.R;
  const float *G = geom + (size_t)warp * CAR_GG_STRIDE;
  float *row = out + (size_t)warp * gs.parts * gs.xw;
  for (int part = 0; part < gs.parts; ++part) {
    // which view is sampled, where, with which padding, and into which slot
    int view = j, slot = part;
    float gx = G[GG_GX], gy = G[GG_GY];
    bool border = true;
    if (gs.parts == 2) {
      if (part == 1) { view = 1 - j; gx = G[GG_C0]; gy = G[GG_C0 + 1]; border = false; }
      slot = view;                                      // channel halves are ordered (view 0, view 1)
    } else if (gs.parts == 3 && part > 0) {
      int idx = 0;
      for (int jj = 0; jj < 3; ++jj) { if (jj == j) continue; if (++idx == part) { view = jj; break; } }
      gx = G[part == 1 ? GG_C0 : GG_C1]; gy = G[(part == 1 ? GG_C0 : GG_C1) + 1]; border = false;
    }
    float *dst = row + (size_t)slot * gs.xw;
    int chan0 = 0;
#pragma unroll
    for (int lvl = 0; lvl < 3; ++lvl) {
      int C = lvl == 2 ? 64 : 256;
      int h = lvl == 0 ? a.H / 4 : (lvl == 1 ? a.H / 2 : a.H);
      int w = lvl == 0 ? a.W / 4 : (lvl == 1 ? a.W / 2 : a.W);
      const float *img = a.feat[lvl] + (size_t)(s * n + view) * h * w * C;
      Taps t = make_taps(gx, gy, w, h, border);
      for (int c = lane * 4; c < C; c += 128)
        *reinterpret_cast<float4 *>(dst + chan0 + c) = lerp4(img, C, c, t);
      chan0 += C;
    }
    if (gs.xw > CAR_C_FEAT && lane < 4) {
      // tail columns 576..591: tanh triple(s), zero padding
      float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n == 1) {
        if (lane == 0) t4 = make_float4(G[GG_T], G[GG_T + 1], G[GG_T + 2], G[GG_T + 3]);
        if (lane == 1) t4 = make_float4(G[GG_T + 4], G[GG_T + 5], 0.f, 0.f);
      } else if (lane == 0) {
        // n_view = 2: the triple of the VIEW in this slot (GG_T + 3*view); n_view = 3: the triple of this part
        const float *T = G + GG_T + 3 * (gs.parts == 2 ? slot : part);
        t4 = make_float4(T[0], T[1], T[2], 0.f);
      }
      *reinterpret_cast<float4 *>(dst + CAR_C_FEAT + lane * 4) = t4;
    }
  }
}

// ---------------------------------------------------------------------------
// Backward of the two gathers (grid_sample backward w.r.t. the input maps,
// ATen/native/cuda/GridSampler.cuh grid_sampler_2d_backward: safe_add_2d of
// gOut * tap weight).  One warp per sample row; d(X)[row][view][576] is
// scattered with 16-byte vector atomics into the NHWC gradient maps.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void scatter4(float *__restrict__ img, int C, int c, const Taps &t, float4 g) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (t.off[k] >= 0) {
      float4 v = make_float4(g.x * t.w[k], g.y * t.w[k], g.z * t.w[k], g.w * t.w[k]);
      atomicAdd(reinterpret_cast<float4 *>(img + (size_t)t.off[k] * C + c), v);
    }
  }
}

__global__ void __launch_bounds__(256)
k_gather_backward(car_render_args a, int g0, int g1, const float *__restrict__ geom,
                  const float *__restrict__ dx, float *d0, float *d1, float *d2) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  long nrows = (long)(g1 - g0) * 2 * a.P;
  if (warp >= nrows) return;
  int rj = warp / a.P;
  int j = rj & 1;
  int g = g0 + (rj >> 1);
  int s = g / a.R;
  const float *G = geom + (size_t)warp * CAR_GEOM_STRIDE;
  float gx = G[G_GX], gy = G[G_GY], gxc = G[G_GXC], gyc = G[G_GYC];
  const float *dx_own = dx + ((size_t)warp * 2 + j) * CAR_C_FEAT;
  const float *dx_oth = dx + ((size_t)warp * 2 + (1 - j)) * CAR_C_FEAT;
  float *dl[3] = {d0, d1, d2};
  int chan0 = 0;
#pragma unroll
  for (int lvl = 0; lvl < 3; ++lvl) {
    int C = lvl == 2 ? 64 : 256;
    int h = lvl == 0 ? a.H / 4 : (lvl == 1 ? a.H / 2 : a.H);
    int w = lvl == 0 ? a.W / 4 : (lvl == 1 ? a.W / 2 : a.W);
    float *own = dl[lvl] + (size_t)(s * 2 + j) * h * w * C;
    float *oth = dl[lvl] + (size_t)(s * 2 + (1 - j)) * h * w * C;
    Taps to = make_taps(gx, gy, w, h, true);
    Taps tc = make_taps(gxc, gyc, w, h, false);
    for (int c = lane * 4; c < C; c += 128) {
      float4 go = __ldg(reinterpret_cast<const float4 *>(dx_own + chan0 + c));
      float4 gc = __ldg(reinterpret_cast<const float4 *>(dx_oth + chan0 + c));
      scatter4(own, C, c, to, go);
      scatter4(oth, C, c, tc, gc);
    }
    chan0 += C;
  }
}

// NHWC fp32 -> NCHW fp32 (inverse of k_pack_features)
__global__ void k_unpack_features(const float *__restrict__ src, float *__restrict__ dst, int C, int hw) {
  __shared__ float tile[32][33];
  int img = blockIdx.z;
  int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < hw && c < C) ? src[((size_t)img * hw + p) * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < hw) dst[((size_t)img * C + c) * hw + p] = tile[threadIdx.x][i];
  }
}

__global__ void k_split_rows(const float *__restrict__ src, int src_stride, uint16_t *__restrict__ hi,
                             uint16_t *__restrict__ lo, long n, int width) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long r = i / width;
  int c = (int)(i - r * width);
  uint16_t h, l;
  split_bf16(src[r * src_stride + c], h, l);
  hi[i] = h;
  if (lo) lo[i] = l;
}

}  // namespace

void launch_split_rows(const float *src, int src_stride, uint16_t *hi, uint16_t *lo, int rows,
                       int width, cudaStream_t st) {
  long n = (long)rows * width;
  if (n <= 0) return;
  prof_pre(CAR_ST_GEMM_SMALL, st);
  k_split_rows<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, src_stride, hi, lo, n, width);
  prof_post(st);
  count_launch();
}

void launch_pack_features(const float *nchw, void *nhwc, int bn, int C, int h, int w, int bf16,
                          cudaStream_t st) {
  int hw = h * w;
  dim3 grid((hw + 31) / 32, (C + 31) / 32, bn), block(32, 8);
  prof_pre(CAR_ST_PACK, st);
  if (bf16) k_pack_features<true><<<grid, block, 0, st>>>(nchw, nhwc, C, hw);
  else k_pack_features<false><<<grid, block, 0, st>>>(nchw, nhwc, C, hw);
  prof_post(st);
  count_launch();
}

void launch_gather_backward(const car_render_args &a, int g0, int g1, const float *geom, const float *dx,
                            float *const d_feat[3], cudaStream_t st) {
  long nrows = (long)(g1 - g0) * 2 * a.P;
  if (nrows <= 0) return;
  unsigned blocks = (unsigned)((nrows * 32 + 255) / 256);
  prof_pre(-1, st);                                   // tagged by the caller (CAR_ST_BWD_SCATTER)
  k_gather_backward<<<blocks, 256, 0, st>>>(a, g0, g1, geom, dx, d_feat[0], d_feat[1], d_feat[2]);
  prof_post(st);
  count_launch();
}

void launch_unpack_features(const float *nhwc, float *nchw, int bn, int C, int h, int w, cudaStream_t st) {
  int hw = h * w;
  dim3 grid((hw + 31) / 32, (C + 31) / 32, bn), block(32, 8);
  prof_pre(CAR_ST_PACK, st);
  k_unpack_features<<<grid, block, 0, st>>>(nhwc, nchw, C, hw);
  prof_post(st);
  count_launch();
}

void launch_gather_general(const car_general_args &a, const GenShape &gs, int g0, int g1, const float *geom, float *x,
                           cudaStream_t st) {
  long nrows = (long)(g1 - g0) * gs.n * a.P;
  if (nrows <= 0) return;
  unsigned blocks = (unsigned)((nrows * 32 + 255) / 256);
  prof_pre(CAR_ST_GATHER, st);
  k_gather_g<<<blocks, 256, 0, st>>>(a, gs, g0, g1, geom, x);
  prof_post(st);
  count_launch();
}

void launch_gather(const car_render_args &a, int g0, int g1, const float *geom, float *out_f32,
                   uint16_t *out_hi, uint16_t *out_lo, cudaStream_t st) {
  long nrows = (long)(g1 - g0) * 2 * a.P;
  if (nrows <= 0) return;
  unsigned blocks = (unsigned)((nrows * 32 + 255) / 256);
  prof_pre(CAR_ST_GATHER, st);
  if (a.feat_bf16) {
    if (out_f32) k_gather<__nv_bfloat16, float><<<blocks, 256, 0, st>>>(a, g0, g1, geom, out_f32, nullptr, nullptr);
    else k_gather<__nv_bfloat16, uint16_t><<<blocks, 256, 0, st>>>(a, g0, g1, geom, nullptr, out_hi, out_lo);
  } else {
    if (out_f32) k_gather<float, float><<<blocks, 256, 0, st>>>(a, g0, g1, geom, out_f32, nullptr, nullptr);
    else k_gather<float, uint16_t><<<blocks, 256, 0, st>>>(a, g0, g1, geom, nullptr, out_hi, out_lo);
  }
  prof_post(st);
  count_launch();
}

}  // namespace car
