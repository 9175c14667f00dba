// Ray set-up, epipolar clipping and per-sample geometry.
//
// THIS FILE IS COMPILED WITH -fmad=false: every fp32/fp64 operation below is a
// separately rounded IEEE operation in the order written, so that the float
// sample coordinates (pixel_val) are bit-identical to the fixed-order oracle
// (oracle/car_oracle.py stages A.1-A.3) on any device.
//
// Reference arithmetic restated here:
//   ray through pixel / Plücker   geometry.py:236-245,353-371,409-433
//   epipolar segment              epipolar.py:74-162,175-253 ; models.py:226-258
//   line samples                  models.py:261,271-275
//   fp64 triangulation            geometry.py:98-162
//   cross-view reprojection       models.py:30-39,285-325 ; geometry.py:374-393 ; utils/util.py:16-19
//   geometric query feature       models.py:494-528 ; geometry.py:313-324
#include <math.h>

#include "car_common.cuh"

namespace car {
namespace {

struct V3 { float x, y, z; };
struct D3 { double x, y, z; };

__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
  V3 r;
  r.x = a.y * b.z - a.z * b.y;
  r.y = a.z * b.x - a.x * b.z;
  r.z = a.x * b.y - a.y * b.x;
  return r;
}
__device__ __forceinline__ D3 cross3d(D3 a, D3 b) {
  D3 r;
  r.x = a.y * b.z - a.z * b.y;
  r.y = a.z * b.x - a.x * b.z;
  r.z = a.x * b.y - a.y * b.x;
  return r;
}
__device__ __forceinline__ float norm3(float x, float y, float z) {
  return sqrtf((x * x + y * y) + z * z);
}

// Unit direction of the ray through pixel (u,v) and its moment, expressed in the
// frame the 4x4 M (row-major, rows 0..2 used) maps into.  K=4 dot order:
// ((M0*x + M1*y) + M2) + M3.
__device__ __forceinline__ void ray_through_pixel(float u, float v, float fx, float fy, float cx,
                                                  float cy, const float *__restrict__ M, V3 &d,
                                                  V3 &m) {
  float xl = (u - cx) / fx;
  float yl = (v - cy) / fy;
  float p0 = ((M[0] * xl + M[1] * yl) + M[2]) + M[3];
  float p1 = ((M[4] * xl + M[5] * yl) + M[6]) + M[7];
  float p2 = ((M[8] * xl + M[9] * yl) + M[10]) + M[11];
  V3 o = {M[3], M[7], M[11]};
  float dx = p0 - o.x, dy = p1 - o.y, dz = p2 - o.z;
  float nrm = fmaxf(norm3(dx, dy, dz), 1e-12f);
  d.x = dx / nrm; d.y = dy / nrm; d.z = dz / nrm;
  m = cross3(o, d);
}

__device__ __forceinline__ bool in_bounds(float x, float y) {
  return (x >= -1e-6f) && (y >= -1e-6f) && (x <= 1.000001f) && (y <= 1.000001f);
}

struct Edge { float t, x, y; bool valid; };

__device__ __forceinline__ Edge frame_edge(int dim, float val, const float *Kn, V3 o, V3 d) {
  // Kn = {k00,k01,k02,k10,k11,k12} of the H-normalised intrinsics
  float fs = dim == 0 ? Kn[0] : Kn[4];
  float fo = dim == 0 ? Kn[4] : Kn[0];
  float cs = dim == 0 ? Kn[2] : Kn[5];
  float co = dim == 0 ? Kn[5] : Kn[2];
  float os = dim == 0 ? o.x : o.y, oo = dim == 0 ? o.y : o.x;
  float ds = dim == 0 ? d.x : d.y, dd = dim == 0 ? d.y : d.x;
  float c = (val - cs) / fs;
  Edge e;
  e.t = (c * o.z - os) / (ds - c * d.z);
  float num = fo * (oo * (c * d.z - ds) + dd * (os - c * o.z));
  float den = d.z * os - ds * o.z;
  float other = co + num / den;
  e.x = dim == 0 ? val : other;
  e.y = dim == 0 ? other : val;
  float zz = o.z + e.t * d.z;
  e.valid = in_bounds(e.x, e.y) && (zz > -1e-6f);
  return e;
}

__device__ __forceinline__ void project_norm(V3 p, const float *Kn, float &x, float &y) {
  float den = p.z + 1e-8f;
  float qx = p.x / den, qy = p.y / den, qz = p.z / den;
  x = (Kn[0] * qx + Kn[1] * qy) + Kn[2] * qz;
  y = (Kn[3] * qx + Kn[4] * qy) + Kn[5] * qz;
}

__device__ __forceinline__ float to_grid(float c) {
  float g = (c - 0.5f) * 2.0f;
  return isfinite(g) ? g : 0.0f;
}

// Clip the ray (origin o, unit direction d, context frame) to the context image: epipolar.py:175-253 with
// extrinsics = identity and H-normalised intrinsics (models.py:226-258).  seg = segment end points in grid
// coordinates after the NaN / Inf scrub; ov = overlaps_image.
__device__ __forceinline__ void epipolar_clip(const float *__restrict__ K, float Hf, V3 o, V3 d, RaySeg &sg, bool &ov) {
  float Kn[6] = {K[0] / Hf, K[1] / Hf, K[2] / Hf, K[4] / Hf, K[5] / Hf, K[6] / Hf};
  Edge e[4];
  e[0] = frame_edge(0, 0.0f, Kn, o, d);
  e[1] = frame_edge(0, 1.0f, Kn, o, d);
  e[2] = frame_edge(1, 0.0f, Kn, o, d);
  e[3] = frame_edge(1, 1.0f, Kn, o, d);
  // min / max over the 4 edges; invalid -> +-inf; strict compare => first index wins ties
  const float inf = INFINITY;
  float bt = e[0].valid ? e[0].t : inf;
  Edge fmin = e[0];
#pragma unroll
  for (int i = 1; i < 4; ++i) {
    float ti = e[i].valid ? e[i].t : inf;
    if (ti < bt) { bt = ti; fmin = e[i]; }
  }
  bt = e[0].valid ? e[0].t : -inf;
  Edge fmax = e[0];
#pragma unroll
  for (int i = 1; i < 4; ++i) {
    float ti = e[i].valid ? e[i].t : -inf;
    if (ti > bt) { bt = ti; fmax = e[i]; }
  }
  bool depth_zero = o.z < 1e-6f;
  bool at_cam = norm3(o.x, o.y, o.z) < 1e-6f;
  V3 p0 = at_cam ? d : o;
  float x0, y0, xi, yi;
  project_norm(p0, Kn, x0, y0);
  bool v0 = in_bounds(x0, y0) && (p0.z > -1e-6f);
  v0 = v0 && !(depth_zero && !at_cam);
  project_norm(d, Kn, xi, yi);
  bool vi = in_bounds(xi, yi) && (d.z > -1e-6f);
  float minx = v0 ? x0 : fmin.x, miny = v0 ? y0 : fmin.y;
  float maxx = vi ? xi : fmax.x, maxy = vi ? yi : fmax.y;
  bool minv = v0 || fmin.valid, maxv = vi || fmax.valid;
  sg.sx = to_grid(minx); sg.sy = to_grid(miny);
  sg.ex = to_grid(maxx); sg.ey = to_grid(maxy);
  ov = minv && maxv;
}

__global__ void k_ray_setup(car_render_args a, int g0, int g1, RaySeg *__restrict__ seg,
                            uint8_t *__restrict__ overlap) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int n = (g1 - g0) * 2;
  if (idx >= n) return;
  int g = g0 + idx / 2, j = idx & 1;
  int s = g / a.R, r = g - s * a.R;
  const float *Q = a.cams.Q + (size_t)(s * 2 + j) * 16;
  const float *Kq = a.cams.Kq + (size_t)s * 16;
  const float *K = a.cams.K + (size_t)(s * 2 + j) * 16;
  float u = a.uv[((size_t)s * a.R + r) * 2 + 0];
  float v = a.uv[((size_t)s * a.R + r) * 2 + 1];
  V3 d, m;
  ray_through_pixel(u, v, Kq[0], Kq[5], Kq[2], Kq[6], Q, d, m);
  V3 o = {Q[3], Q[7], Q[11]};
  float *co = a.coords + ((size_t)(s * 2 + j) * a.R + r) * 9;
  co[0] = d.x; co[1] = d.y; co[2] = d.z;
  co[3] = m.x; co[4] = m.y; co[5] = m.z;
  co[6] = o.x; co[7] = o.y; co[8] = o.z;

  RaySeg sg;
  bool ov;
  epipolar_clip(K, (float)a.H, o, d, sg, ov);
  seg[idx] = sg;
  overlap[idx] = ov ? 1 : 0;
}

__device__ __forceinline__ V3 xform_point(const float *__restrict__ T, V3 p) {
  V3 r;
  r.x = ((T[0] * p.x + T[1] * p.y) + T[2] * p.z) + T[3];
  r.y = ((T[4] * p.x + T[5] * p.y) + T[6] * p.z) + T[7];
  r.z = ((T[8] * p.x + T[9] * p.y) + T[10] * p.z) + T[11];
  return r;
}

__device__ __forceinline__ float nan_to_num0(float x) {   // torch.nan_to_num(x, 0)
  if (isnan(x)) return 0.0f;
  if (isinf(x)) return x > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return x;
}

__global__ void k_sample_geometry(car_render_args a, int g0, int g1,
                                  const RaySeg *__restrict__ seg, float *__restrict__ geom) {
  // The 128-byte records of a block are contiguous in `geom`: they are assembled in shared memory
  // (row stride 33 floats: conflict-free for both access patterns) and written out with coalesced
  // 16-byte stores; 32 scalar stores per thread at a 128-byte lane stride made the kernel LSU-bound.
  __shared__ float srec[128][33];
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long n = (long)(g1 - g0) * 2 * a.P;
  const bool live = idx < n;
  if (!live) idx = n - 1;                                   // compute a valid sample, store nothing
  int k = (int)(idx % a.P);
  int rj = (int)(idx / a.P);
  int j = rj & 1;
  int g = g0 + (rj >> 1);
  int s = g / a.R, r = g - s * a.R;
  RaySeg sg = seg[rj];
  float iv = a.interval[k];
  float gx = sg.sx + (sg.ex - sg.sx) * iv;
  float gy = sg.sy + (sg.ey - sg.sy) * iv;
  float *pv = a.pixel_val + (((size_t)(s * 2 + j) * a.R + r) * a.P + k) * 2;
  if (live) *reinterpret_cast<float2 *>(pv) = make_float2(gx, gy);

  const float *co = a.coords + ((size_t)(s * 2 + j) * a.R + r) * 9;
  V3 d = {co[0], co[1], co[2]}, m = {co[3], co[4], co[5]}, o = {co[6], co[7], co[8]};
  const float *K = a.cams.K + (size_t)(s * 2 + j) * 16;
  float fx = K[0], fy = K[5], cx = K[2], cy = K[6];
  float px = (gx + 1.0f) / 2.0f * (float)(a.W - 1);
  float py = (gy + 1.0f) / 2.0f * (float)(a.H - 1);
  V3 l2f, m2f;
  ray_through_pixel(px, py, fx, fy, cx, cy, a.cams.Cself + (size_t)(s * 2 + j) * 16, l2f, m2f);
  // fp64 closest point on the query ray (geometry.py:132-162)
  D3 l1 = {(double)d.x, (double)d.y, (double)d.z};
  D3 m1 = {(double)m.x, (double)m.y, (double)m.z};
  D3 l2 = {(double)l2f.x, (double)l2f.y, (double)l2f.z};
  D3 m2 = {(double)m2f.x, (double)m2f.y, (double)m2f.z};
  D3 nn = cross3d(l1, l2);
  D3 aa = cross3d(l2, nn);
  D3 t1 = cross3d(m1, aa);
  double sdot = (m2.x * nn.x + m2.y * nn.y) + m2.z * nn.z;
  double nrm = sqrt((nn.x * nn.x + nn.y * nn.y) + nn.z * nn.z);
  double cd = nrm * nrm + 1e-12;
  double p1x = (-t1.x + sdot * l1.x) / cd;
  double p1y = (-t1.y + sdot * l1.y) / cd;
  double p1z = (-t1.z + sdot * l1.z) / cd;
  V3 pt;
  pt.x = isfinite(p1x) ? (float)p1x : 0.0f;
  pt.y = isfinite(p1y) ? (float)p1y : 0.0f;
  pt.z = isfinite(p1z) ? (float)p1z : 0.0f;

  const float *Rel = a.cams.Rel + (size_t)s * 64;      // [k][j][4][4]
  V3 pv0 = xform_point(Rel + (0 * 2 + j) * 16, pt);
  V3 pv1 = xform_point(Rel + (1 * 2 + j) * 16, pt);
  V3 oth = j == 0 ? pv1 : pv0;
  const float *Ko = a.cams.K + (size_t)(s * 2 + (1 - j)) * 16;
  float xp = Ko[0] * oth.x / (oth.z + 1e-12f) + Ko[2];
  float yp = Ko[5] * oth.y / (oth.z + 1e-12f) + Ko[6];
  xp = isfinite(xp) ? xp : 1e10f;
  yp = isfinite(yp) ? yp : 1e10f;
  float gxc = (xp / (float)(a.W - 1)) * 2.0f - 1.0f;
  float gyc = (yp / (float)(a.H - 1)) * 2.0f - 1.0f;

  float *G = srec[threadIdx.x];
  G[G_GX] = gx; G[G_GY] = gy; G[G_GXC] = gxc; G[G_GYC] = gyc;
  G[G_T0 + 0] = tanhf(nan_to_num0(pv0.x) / 5.0f);
  G[G_T0 + 1] = tanhf(nan_to_num0(pv0.y) / 5.0f);
  G[G_T0 + 2] = tanhf(nan_to_num0(pv0.z) / 5.0f);
  G[G_T1 + 0] = tanhf(nan_to_num0(pv1.x) / 5.0f);
  G[G_T1 + 1] = tanhf(nan_to_num0(pv1.y) / 5.0f);
  G[G_T1 + 2] = tanhf(nan_to_num0(pv1.z) / 5.0f);
  // local_coords = [cam_rays, 0,0,0, ray_dir, tanh(depth/{1,10,100,1000}), ray origin]
  float rx = (px - cx) / fx, ry = (py - cy) / fy;
  float rn = fmaxf(norm3(rx, ry, 1.0f), 1e-12f);
  float *L = G + G_LOCAL;
  L[0] = rx / rn; L[1] = ry / rn; L[2] = 1.0f / rn;
  L[3] = 0.f; L[4] = 0.f; L[5] = 0.f;
  L[6] = d.x; L[7] = d.y; L[8] = d.z;
  float depth = norm3(pt.x - o.x, pt.y - o.y, pt.z - o.z);
  if (!isfinite(depth)) depth = 1000000.0f;
  L[9] = tanhf(depth);
  L[10] = tanhf(depth / 10.0f);
  L[11] = tanhf(depth / 100.0f);
  L[12] = tanhf(depth / 1000.0f);
  L[13] = o.x; L[14] = o.y; L[15] = o.z;
  G[G_PTC + 0] = fminf(fmaxf(pt.x, -100.0f), 100.0f);
  G[G_PTC + 1] = fminf(fmaxf(pt.y, -100.0f), 100.0f);
  G[G_PTC + 2] = fminf(fmaxf(pt.z, -100.0f), 100.0f);
  G[13] = 0.f; G[14] = 0.f; G[15] = 0.f;
  __syncthreads();
  const long rec0 = (long)blockIdx.x * blockDim.x;          // first record of this block
  float4 *out = reinterpret_cast<float4 *>(geom + (size_t)rec0 * CAR_GEOM_STRIDE);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int i4 = threadIdx.x + q * 128;                   // float4 index within the block's 16 KB
    const int rec = i4 >> 3, e = (i4 & 7) * 4;
    if (rec0 + rec < n) out[i4] = make_float4(srec[rec][e], srec[rec][e + 1], srec[rec][e + 2], srec[rec][e + 3]);
  }
}

// ---------------------------------------------------------------------------------------
// General branches (n_view = 1 / 3, no_sample, no_latent_concat): same arithmetic, n contexts.
// ---------------------------------------------------------------------------------------
// geometry.project (geometry.py:374-393) + util.normalize_for_grid_sample (utils/util.py:16-19)
__device__ __forceinline__ void project_to_grid(V3 p, const float *__restrict__ Kj, float Wm1, float Hm1, float &gx, float &gy) {
  float xp = Kj[0] * p.x / (p.z + 1e-12f) + Kj[2];
  float yp = Kj[5] * p.y / (p.z + 1e-12f) + Kj[6];
  xp = isfinite(xp) ? xp : 1e10f;
  yp = isfinite(yp) ? yp : 1e10f;
  gx = (xp / Wm1) * 2.0f - 1.0f;
  gy = (yp / Hm1) * 2.0f - 1.0f;
}

// One thread per (ray, context): Plücker coordinates, and either the clipped epipolar segment or
// (CAR_FLAG_NO_SAMPLE, geometry.py:165-187) the projections of the ray's points at the depths in
// `interval`, written straight to pixel_val, with valid = any sample strictly inside (-1, 1)^2.
__global__ void k_ray_setup_g(car_general_args a, int g0, int g1, RaySeg *__restrict__ seg,
                              uint8_t *__restrict__ overlap) {
  const int n = a.n_view;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (g1 - g0) * n) return;
  int g = g0 + idx / n, j = idx % n;
  int s = g / a.R, r = g - s * a.R;
  const float *Q = a.cams.Q + (size_t)(s * n + j) * 16;
  const float *Kq = a.cams.Kq + (size_t)s * 16;
  const float *K = a.cams.K + (size_t)(s * n + j) * 16;
  float u = a.uv[((size_t)s * a.R + r) * 2 + 0];
  float v = a.uv[((size_t)s * a.R + r) * 2 + 1];
  V3 d, m;
  ray_through_pixel(u, v, Kq[0], Kq[5], Kq[2], Kq[6], Q, d, m);
  V3 o = {Q[3], Q[7], Q[11]};
  float *co = a.coords + ((size_t)(s * n + j) * a.R + r) * 9;
  co[0] = d.x; co[1] = d.y; co[2] = d.z;
  co[3] = m.x; co[4] = m.y; co[5] = m.z;
  co[6] = o.x; co[7] = o.y; co[8] = o.z;
  RaySeg sg = {0.f, 0.f, 0.f, 0.f};
  bool ov = false;
  if (a.flags & CAR_FLAG_NO_SAMPLE) {
    float *pv = a.pixel_val + (((size_t)(s * n + j) * a.R + r) * a.P) * 2;
    const float Wm1 = (float)(a.W - 1), Hm1 = (float)(a.H - 1);
    for (int k = 0; k < a.P; ++k) {
      const float t = a.interval[k];
      V3 pt = {o.x + t * d.x, o.y + t * d.y, o.z + t * d.z};
      float gx, gy;
      project_to_grid(pt, K, Wm1, Hm1, gx, gy);
      pv[2 * k] = gx; pv[2 * k + 1] = gy;
      ov = ov || (gx < 1.0f && gx > -1.0f && gy < 1.0f && gy > -1.0f);
    }
  } else {
    epipolar_clip(K, (float)a.H, o, d, sg, ov);
  }
  seg[idx] = sg;
  overlap[idx] = ov ? 1 : 0;
}

// One thread per (ray, sample index) handling all n contexts (n_view = 3 needs the triangulated points of
// the other contexts' samples, models.py:354-423).  Writes the 48-float records (car_common.cuh GG_*).
__global__ void k_sample_geometry_g(car_general_args a, int g0, int g1, const RaySeg *__restrict__ seg,
                                    float *__restrict__ geom) {
  const int n = a.n_view;
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)(g1 - g0) * a.P) return;
  const int k = (int)(idx % a.P);
  const int gl = (int)(idx / a.P);
  const int g = g0 + gl;
  const int s = g / a.R, r = g - s * a.R;
  const bool nosample = (a.flags & CAR_FLAG_NO_SAMPLE) != 0;
  const bool concat = !(a.flags & CAR_FLAG_NO_LATENT_CONCAT);
  V3 pt[3];
  float gxs[3], gys[3];
  for (int j = 0; j < n; ++j) {
    float gx, gy;
    float *pv = a.pixel_val + (((size_t)(s * n + j) * a.R + r) * a.P + k) * 2;
    if (nosample) { gx = pv[0]; gy = pv[1]; }
    else {
      const RaySeg sg = seg[gl * n + j];
      const float iv = a.interval[k];
      gx = sg.sx + (sg.ex - sg.sx) * iv;
      gy = sg.sy + (sg.ey - sg.sy) * iv;
      *reinterpret_cast<float2 *>(pv) = make_float2(gx, gy);
    }
    gxs[j] = gx; gys[j] = gy;
    const float *co = a.coords + ((size_t)(s * n + j) * a.R + r) * 9;
    V3 d = {co[0], co[1], co[2]}, m = {co[3], co[4], co[5]}, o = {co[6], co[7], co[8]};
    const float *K = a.cams.K + (size_t)(s * n + j) * 16;
    const float fx = K[0], fy = K[5], cx = K[2], cy = K[6];
    const float px = (gx + 1.0f) / 2.0f * (float)(a.W - 1);
    const float py = (gy + 1.0f) / 2.0f * (float)(a.H - 1);
    V3 l2f, m2f;
    ray_through_pixel(px, py, fx, fy, cx, cy, a.cams.Cself + (size_t)(s * n + j) * 16, l2f, m2f);
    D3 l1 = {(double)d.x, (double)d.y, (double)d.z};
    D3 m1 = {(double)m.x, (double)m.y, (double)m.z};
    D3 l2 = {(double)l2f.x, (double)l2f.y, (double)l2f.z};
    D3 m2 = {(double)m2f.x, (double)m2f.y, (double)m2f.z};
    D3 nn = cross3d(l1, l2);
    D3 aa = cross3d(l2, nn);
    D3 t1 = cross3d(m1, aa);
    double sdot = (m2.x * nn.x + m2.y * nn.y) + m2.z * nn.z;
    double nrm = sqrt((nn.x * nn.x + nn.y * nn.y) + nn.z * nn.z);
    double cd = nrm * nrm + 1e-12;
    double p1x = (-t1.x + sdot * l1.x) / cd;
    double p1y = (-t1.y + sdot * l1.y) / cd;
    double p1z = (-t1.z + sdot * l1.z) / cd;
    pt[j].x = isfinite(p1x) ? (float)p1x : 0.0f;
    pt[j].y = isfinite(p1y) ? (float)p1y : 0.0f;
    pt[j].z = isfinite(p1z) ? (float)p1z : 0.0f;
    float *G = geom + ((size_t)(gl * n + j) * a.P + k) * CAR_GG_STRIDE;
    G[GG_GX] = gx; G[GG_GY] = gy;
    // local_coords = [cam_rays, 0,0,0, ray_dir, tanh(depth/{1,10,100,1000}), ray origin]  (models.py:494-528)
    const float rx = (px - cx) / fx, ry = (py - cy) / fy;
    const float rn = fmaxf(norm3(rx, ry, 1.0f), 1e-12f);
    float *L = G + GG_LOCAL;
    L[0] = rx / rn; L[1] = ry / rn; L[2] = 1.0f / rn;
    L[3] = 0.f; L[4] = 0.f; L[5] = 0.f;
    L[6] = d.x; L[7] = d.y; L[8] = d.z;
    float depth = norm3(pt[j].x - o.x, pt[j].y - o.y, pt[j].z - o.z);
    if (!isfinite(depth)) depth = 1000000.0f;
    L[9] = tanhf(depth);
    L[10] = tanhf(depth / 10.0f);
    L[11] = tanhf(depth / 100.0f);
    L[12] = tanhf(depth / 1000.0f);
    L[13] = o.x; L[14] = o.y; L[15] = o.z;
    G[GG_PTC + 0] = fminf(fmaxf(pt[j].x, -100.0f), 100.0f);
    G[GG_PTC + 1] = fminf(fmaxf(pt[j].y, -100.0f), 100.0f);
    G[GG_PTC + 2] = fminf(fmaxf(pt[j].z, -100.0f), 100.0f);
    for (int q = 2; q < GG_PTC; ++q) G[q] = 0.f;
    for (int q = GG_PTC + 3; q < GG_LOCAL; ++q) G[q] = 0.f;
  }
  const float Wm1 = (float)(a.W - 1), Hm1 = (float)(a.H - 1);
  for (int j = 0; j < n; ++j) {
    float *G = geom + ((size_t)(gl * n + j) * a.P + k) * CAR_GG_STRIDE;
    if (n == 1) {
      // models.py:483: [tanh(pt/5), tanh(pt/100)] (pt already scrubbed, geometry.py:126-127)
      G[GG_T + 0] = tanhf(pt[0].x / 5.0f); G[GG_T + 1] = tanhf(pt[0].y / 5.0f); G[GG_T + 2] = tanhf(pt[0].z / 5.0f);
      G[GG_T + 3] = tanhf(pt[0].x / 100.0f); G[GG_T + 4] = tanhf(pt[0].y / 100.0f); G[GG_T + 5] = tanhf(pt[0].z / 100.0f);
    } else if (!concat) {
      // raw features only
    } else if (n == 2) {
      // models.py:285-342: parts in (view 0, view 1) order; the other view's features at the re-projected point
      const float *Rel = a.cams.Rel + (size_t)s * 64;               // [k][j][4][4]
      V3 pv0 = xform_point(Rel + (0 * 2 + j) * 16, pt[j]);
      V3 pv1 = xform_point(Rel + (1 * 2 + j) * 16, pt[j]);
      V3 oth = j == 0 ? pv1 : pv0;
      project_to_grid(oth, a.cams.K + (size_t)(s * 2 + (1 - j)) * 16, Wm1, Hm1, G[GG_C0], G[GG_C0 + 1]);
      G[GG_T + 0] = tanhf(nan_to_num0(pv0.x) / 5.0f); G[GG_T + 1] = tanhf(nan_to_num0(pv0.y) / 5.0f); G[GG_T + 2] = tanhf(nan_to_num0(pv0.z) / 5.0f);
      G[GG_T + 3] = tanhf(nan_to_num0(pv1.x) / 5.0f); G[GG_T + 4] = tanhf(nan_to_num0(pv1.y) / 5.0f); G[GG_T + 5] = tanhf(nan_to_num0(pv1.z) / 5.0f);
    } else {
      // n_view = 3 (models.py:345-475): row of context a = j.  part 0: own features, tanh(ptv[a][a] / 5); parts 1, 2:
      // the other contexts jj in ascending order: view jj's maps at project(ptv[a][jj], K_jj) where
      // ptv[a][jj] = Rel[a][jj] . pt[jj] (the same (ray, sample) of context jj's OWN line), and tanh(ptv[a][jj] / 5)
      const float *Rel = a.cams.Rel + (size_t)s * 9 * 16 + (size_t)j * 3 * 16;   // Rel[s][a = j][.]
      V3 own = xform_point(Rel + j * 16, pt[j]);
      G[GG_T + 0] = tanhf(nan_to_num0(own.x) / 5.0f); G[GG_T + 1] = tanhf(nan_to_num0(own.y) / 5.0f); G[GG_T + 2] = tanhf(nan_to_num0(own.z) / 5.0f);
      int part = 1;
      for (int jj = 0; jj < 3; ++jj) {
        if (jj == j) continue;
        V3 q = xform_point(Rel + jj * 16, pt[jj]);
        float *cc = G + (part == 1 ? GG_C0 : GG_C1);
        project_to_grid(q, a.cams.K + (size_t)(s * 3 + jj) * 16, Wm1, Hm1, cc[0], cc[1]);
        float *T = G + GG_T + 3 * part;
        T[0] = tanhf(nan_to_num0(q.x) / 5.0f); T[1] = tanhf(nan_to_num0(q.y) / 5.0f); T[2] = tanhf(nan_to_num0(q.z) / 5.0f);
        ++part;
      }
    }
  }
}

}  // namespace

void launch_ray_setup_general(const car_general_args &a, int g0, int g1, RaySeg *seg, uint8_t *overlap, cudaStream_t st) {
  int n = (g1 - g0) * a.n_view;
  if (n <= 0) return;
  prof_pre(CAR_ST_RAYSETUP, st);
  k_ray_setup_g<<<(n + 127) / 128, 128, 0, st>>>(a, g0, g1, seg, overlap);
  prof_post(st);
  count_launch();
}

void launch_sample_geometry_general(const car_general_args &a, int g0, int g1, const RaySeg *seg, float *geom, cudaStream_t st) {
  long n = (long)(g1 - g0) * a.P;
  if (n <= 0) return;
  prof_pre(CAR_ST_SAMPLE_GEOM, st);
  k_sample_geometry_g<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(a, g0, g1, seg, geom);
  prof_post(st);
  count_launch();
}

void launch_ray_setup(const car_render_args &a, int g0, int g1, RaySeg *seg, uint8_t *overlap,
                      cudaStream_t st) {
  int n = (g1 - g0) * 2;
  if (n <= 0) return;
  prof_pre(CAR_ST_RAYSETUP, st);
  k_ray_setup<<<(n + 127) / 128, 128, 0, st>>>(a, g0, g1, seg, overlap);
  prof_post(st);
  count_launch();
}

void launch_sample_geometry(const car_render_args &a, int g0, int g1, const RaySeg *seg,
                            float *geom, cudaStream_t st) {
  long n = (long)(g1 - g0) * 2 * a.P;
  if (n <= 0) return;
  prof_pre(CAR_ST_SAMPLE_GEOM, st);
  k_sample_geometry<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(a, g0, g1, seg, geom);
  prof_post(st);
  count_launch();
}

}  // namespace car
