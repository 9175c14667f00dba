// The other forward branches of the reference behind the C ABI (include/car_b200.h, car_general_args):
//   n_view = 1   (models.py:478-485)      n_view = 3 (models.py:345-475)
//   n_view = 2 with no_sample (models.py:219-222 -> geometry.py:165-187) or no_latent_concat (models.py:476-477)
// followed by the shared tail (models.py:487-621).  Same stages as car_render_forward with a different gather
// fan-out: general geometry kernels (car_geometry.cu), the general gather (car_gather.cu), exact-fp32 GEMMs
// (car_gemm_simt.cu) and n-context attention kernels (car_attention.cu).  The n_view = 2 / default-flag hot
// path stays in car_api.cu with its fused tcgen05 kernels.
#include <string.h>

#include "car_common.cuh"

namespace car {
namespace {

struct GenWs {
  RaySeg *seg; uint8_t *overlap;
  float *geom, *x, *h1, *interp, *value, *hid, *key, *q1, *q2;
  float *zsum, *g, *rowbias, *zfin, *c32, *px, *pnet, *rgb3;
  // tensor-core precision: bf16 hi / lo operand copies
  uint16_t *x_hi, *x_lo, *h1_hi, *h1_lo, *in_hi, *in_lo, *hid_hi, *hid_lo, *loc_hi, *loc_lo;
  size_t bytes;
};

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

GenWs carve(char *base, const GenShape &gs, int P, int chunk) {
  GenWs w;
  memset(&w, 0, sizeof(w));
  size_t off = 0;
  const size_t rows = (size_t)chunk * gs.n * P;
  auto take = [&](size_t bytes) { char *p = base ? base + off : nullptr; off += align_up(bytes); return p; };
  const bool enc = gs.parts > 1;                       // per-sample encoder (query_encode_latent + _2)
  const bool merge = gs.parts == 1 && gs.xw == CAR_K_ENC;   // n_view = 1: update_val_merge
  w.seg = (RaySeg *)take((size_t)chunk * gs.n * sizeof(RaySeg));
  w.overlap = (uint8_t *)take((size_t)chunk * gs.n);
  w.geom = (float *)take(rows * CAR_GG_STRIDE * 4);
  w.x = (float *)take(rows * gs.parts * gs.xw * 4);
  if (enc) w.h1 = (float *)take(rows * gs.parts * CAR_C_FEAT * 4);
  w.interp = (enc || merge) ? (float *)take(rows * gs.ci * 4) : w.x;    // no_latent_concat: the raw features
  w.value = (float *)take(rows * gs.L * 4);
  w.hid = (float *)take(rows * 128 * 4);
  w.key = (float *)take(rows * 128 * 4);
  w.q1 = (float *)take(rows * 128 * 4);
  w.q2 = (float *)take(rows * 128 * 4);
  w.zsum = (float *)take((size_t)chunk * gs.L * 4);
  w.g = (float *)take((size_t)chunk * 128 * 4);
  w.rowbias = (float *)take((size_t)chunk * 128 * 4);
  w.zfin = (float *)take((size_t)chunk * gs.L * 4);
  w.c32 = (float *)take((size_t)chunk * 32 * 4);
  w.px = (float *)take((size_t)chunk * 128 * 4);
  w.pnet = (float *)take((size_t)chunk * 128 * 4);
  w.rgb3 = (float *)take((size_t)chunk * 4 * 4);
  // operand copies of the tensor-core precision (allocated unconditionally: one workspace size per configuration)
  w.x_hi = (uint16_t *)take(rows * gs.parts * gs.xw * 2);
  w.x_lo = (uint16_t *)take(rows * gs.parts * gs.xw * 2);
  if (enc) { w.h1_hi = (uint16_t *)take(rows * gs.parts * CAR_C_FEAT * 2); w.h1_lo = (uint16_t *)take(rows * gs.parts * CAR_C_FEAT * 2); }
  if (enc || merge) { w.in_hi = (uint16_t *)take(rows * gs.ci * 2); w.in_lo = (uint16_t *)take(rows * gs.ci * 2); }
  else { w.in_hi = w.x_hi; w.in_lo = w.x_lo; }
  w.hid_hi = (uint16_t *)take(rows * 128 * 2);
  w.hid_lo = (uint16_t *)take(rows * 128 * 2);
  w.loc_hi = (uint16_t *)take(rows * 16 * 2);
  w.loc_lo = (uint16_t *)take(rows * 16 * 2);
  w.bytes = off;
  return w;
}

GemmEpi epi(const float *bias, int relu_out, int relu_in = 0, int accumulate = 0, const float *row_bias = nullptr,
            int rows_per_group = 1) {
  GemmEpi e;
  e.bias = bias; e.row_bias = row_bias; e.rows_per_group = rows_per_group;
  e.relu_in = relu_in; e.relu_out = relu_out; e.accumulate = accumulate;
  return e;
}

int check_mat(const car_mat &m, int N, int K, const char *name, bool tc = false) {
  if (tc && (!m.hi || !m.lo)) { set_error("general path: weight %s needs bf16 hi / lo copies for the tensor-core precision", name); return -11; }
  if (!m.f32 || m.N != N || m.K != K) { set_error("general path: weight %s must be fp32 [%d][%d] (got N=%d K=%d)", name, N, K, m.N, m.K); return -11; }
  return 0;
}

void gemm(const float *A, int lda, const car_mat &m, float *C, int ldc, int M, const GemmEpi &e, cudaStream_t st) {
  launch_gemm_simt(A, lda, m.f32, m.K, C, ldc, M, m.N, m.K, e, st);
}

}  // namespace
}  // namespace car

using namespace car;

extern "C" {

int car_general_default_chunk_rays(int n_view, int flags, int P) {
  if (n_view < 1 || n_view > 3 || P < 1) return 1;
  long c = (1L << 17) / ((long)n_view * P);            // ~131 k sample rows per chunk (up to ~21 KB of fp32 activations each)
  return (int)(c < 1 ? 1 : c);
}

size_t car_general_workspace_bytes(int n_view, int flags, int P, int chunk_rays) {
  if (n_view < 1 || n_view > 3 || P < 1 || chunk_rays < 1) return 0;
  return carve(nullptr, gen_shape(n_view, flags), P, chunk_rays).bytes;
}

int car_render_forward_general(const car_general_args *pa) {
  reset_launch_count();
  if (!pa) { set_error("null args"); return -1; }
  const car_general_args &a = *pa;
  if (a.abi_version != CAR_ABI_VERSION) { set_error("ABI version mismatch: got %d want %d", a.abi_version, CAR_ABI_VERSION); return -2; }
  if (a.n_view < 1 || a.n_view > 3 || (a.flags & ~3) || (a.n_view != 2 && a.flags)) {
    set_error("general path: n_view %d / flags %d (flags apply to n_view = 2; n_view in 1..3)", a.n_view, a.flags); return -3;
  }
  if (a.b <= 0 || a.R <= 0 || a.P < 2 || a.P > 256 || a.H < 4 || a.W < 4 || (a.H & 3) || (a.W & 3)) {
    set_error("bad sizes b=%d R=%d P=%d H=%d W=%d", a.b, a.R, a.P, a.H, a.W); return -3;
  }
  const long total = (long)a.b * a.R;
  if (a.ray_begin < 0 || a.ray_end > total || a.ray_begin > a.ray_end) { set_error("bad ray range [%d,%d) of %ld", a.ray_begin, a.ray_end, total); return -4; }
  if (!a.feat[0] || !a.feat[1] || !a.feat[2] || !a.uv || !a.interval || !a.rgb || !a.valid_mask || !a.depth_ray || !a.at_wt ||
      !a.at_wt_max || !a.pixel_val || !a.coords || !a.workspace || !a.cams.Q || !a.cams.Cself || !a.cams.Rel || !a.cams.qinv ||
      !a.cams.K || !a.cams.Kq) { set_error("null pointer in car_general_args"); return -6; }
  if (a.ray_begin == a.ray_end) return 0;
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) { set_error("no CUDA device (there is no CPU fallback)"); return -7; }
  const GenShape gs = gen_shape(a.n_view, a.flags);
  const car_general_weights &W = a.weights;
  const bool enc = gs.parts > 1, merge = gs.parts == 1 && gs.xw == CAR_K_ENC;
  int rc;
  if (a.precision != CAR_PREC_FP32_SIMT && a.precision != CAR_PREC_FP32_3XBF16) { set_error("general path: precision %d (0 or 1)", a.precision); return -5; }
  const bool tc = a.precision == CAR_PREC_FP32_3XBF16;
  if (enc && ((rc = check_mat(W.enc1, CAR_C_FEAT, CAR_K_ENC, "enc1", tc)) || (rc = check_mat(W.enc2, CAR_C_LAT, CAR_C_FEAT, "enc2", tc)))) return rc;
  if (merge && (rc = check_mat(W.merge, CAR_C_FEAT, CAR_K_ENC, "merge", tc))) return rc;
  if ((rc = check_mat(W.value, gs.L, gs.ci, "value", tc)) || (rc = check_mat(W.key1, 128, gs.ci, "key1", tc)) ||
      (rc = check_mat(W.key2, 128, 128, "key2", tc)) || (rc = check_mat(W.qry1, 128, 16, "qry1", tc)) || (rc = check_mat(W.qry2, 128, 128, "qry2", tc)) ||
      (rc = check_mat(W.rep1_loc, 128, 16, "rep1_loc", tc)) || (rc = check_mat(W.rep1_g, 128, 128, "rep1_g")) ||
      (rc = check_mat(W.rep2, 128, 128, "rep2", tc)) || (rc = check_mat(W.enc_lat, 128, gs.L, "enc_lat")) ||
      (rc = check_mat(W.phi_in, 128, 32, "phi_in")) || (rc = check_mat(W.phi_out, 3, 128, "phi_out"))) return rc;
  for (int i = 0; i < 3; ++i)
    if ((rc = check_mat(W.phi_z[i], 128, gs.L, "phi_z")) || (rc = check_mat(W.phi_fc0[i], 128, 128, "phi_fc0")) ||
        (rc = check_mat(W.phi_fc1[i], 128, 128, "phi_fc1"))) return rc;

  int chunk = a.chunk_rays > 0 ? a.chunk_rays : car_general_default_chunk_rays(a.n_view, a.flags, a.P);
  const int span = a.ray_end - a.ray_begin;
  if (chunk > span) chunk = span;
  while (a.chunk_rays == 0 && chunk > 1 && carve(nullptr, gs, a.P, chunk).bytes > a.workspace_bytes) chunk = (chunk + 1) / 2;
  if (carve(nullptr, gs, a.P, chunk).bytes > a.workspace_bytes) { set_error("workspace too small: %zu bytes", a.workspace_bytes); return -8; }
  const GenWs w = carve((char *)a.workspace, gs, a.P, chunk);
  cudaStream_t st = (cudaStream_t)a.stream;

  for (int g0 = a.ray_begin; g0 < a.ray_end; g0 += chunk) {
    const int g1 = g0 + chunk < a.ray_end ? g0 + chunk : a.ray_end;
    const int nr = g1 - g0;
    const int rows = nr * gs.n * a.P;
    launch_ray_setup_general(a, g0, g1, w.seg, w.overlap, st);
    launch_sample_geometry_general(a, g0, g1, w.seg, w.geom, st);
    launch_gather_general(a, gs, g0, g1, w.geom, w.x, st);
    // per-sample feature stage
    if (tc) {
      // tcgen05 GEMMs (car_gemm_umma.cu), operands as bf16 hi + lo between the layers, fp32 where attention reads them
      auto out_bf = [&](uint16_t *hi, uint16_t *lo, int ldc, float *f32 = nullptr) { UmmaOut o; o.f32 = f32; o.hi = hi; o.lo = lo; o.ldc = ldc; return o; };
      auto out_f = [&](float *f32, int ldc) { UmmaOut o; o.f32 = f32; o.hi = nullptr; o.lo = nullptr; o.ldc = ldc; return o; };
      auto mm = [&](const uint16_t *ah, const uint16_t *al, int lda, const car_mat &m, int M, const GemmEpi &e, const UmmaOut &o) {
        return launch_gemm_umma(ah, al, lda, m.hi, m.lo, m.K, M, m.N, m.K, 1, e, o, st);
      };
      launch_split_rows(w.x, gs.xw, w.x_hi, w.x_lo, rows * gs.parts, gs.xw, st);
      launch_split_rows(w.geom + GG_LOCAL, CAR_GG_STRIDE, w.loc_hi, w.loc_lo, rows, 16, st);
      if (enc) {
        { StageScope sc(CAR_ST_GEMM_ENC1);
          if ((rc = mm(w.x_hi, w.x_lo, CAR_K_ENC, W.enc1, rows * gs.parts, epi(W.enc1.bias, 1), out_bf(w.h1_hi, w.h1_lo, CAR_C_FEAT)))) return rc; }
        { StageScope sc(CAR_ST_GEMM_ENC2);
          if ((rc = mm(w.h1_hi, w.h1_lo, CAR_C_FEAT, W.enc2, rows * gs.parts, epi(W.enc2.bias, 0),
                       out_bf(w.in_hi, w.in_lo, CAR_C_LAT, a.debug_interp ? w.interp : nullptr)))) return rc; }
      } else if (merge) {
        StageScope sc(CAR_ST_GEMM_ENC1);
        if ((rc = mm(w.x_hi, w.x_lo, CAR_K_ENC, W.merge, rows, epi(W.merge.bias, 0),
                     out_bf(w.in_hi, w.in_lo, CAR_C_FEAT, a.debug_interp ? w.interp : nullptr)))) return rc;
      }
      if (a.debug_interp)
        cudaMemcpyAsync(a.debug_interp + (size_t)(g0 - a.ray_begin) * gs.n * a.P * gs.ci, w.interp, (size_t)rows * gs.ci * 4,
                        cudaMemcpyDeviceToDevice, st);
      { StageScope sc(CAR_ST_GEMM_KV);
        if ((rc = mm(w.in_hi, w.in_lo, gs.ci, W.value, rows, epi(W.value.bias, 0), out_f(w.value, gs.L)))) return rc;
        if ((rc = mm(w.in_hi, w.in_lo, gs.ci, W.key1, rows, epi(W.key1.bias, 1), out_bf(w.hid_hi, w.hid_lo, 128)))) return rc; }
      if ((rc = mm(w.hid_hi, w.hid_lo, 128, W.key2, rows, epi(W.key2.bias, 0), out_f(w.key, 128)))) return rc;
      if ((rc = mm(w.loc_hi, w.loc_lo, 16, W.qry1, rows, epi(W.qry1.bias, 1), out_bf(w.hid_hi, w.hid_lo, 128)))) return rc;
      if ((rc = mm(w.hid_hi, w.hid_lo, 128, W.qry2, rows, epi(W.qry2.bias, 0), out_f(w.q1, 128)))) return rc;
    } else {
    if (enc) {
      // query_encode_latent (+ReLU) and query_encode_latent_2 on every part (models.py:333-342, 436-446): M = rows * parts;
      // the parts of a row land side by side = the part-major order of weights.value / key1
      { StageScope sc(CAR_ST_GEMM_ENC1);
        gemm(w.x, CAR_K_ENC, W.enc1, w.h1, CAR_C_FEAT, rows * gs.parts, epi(W.enc1.bias, 1), st); }
      { StageScope sc(CAR_ST_GEMM_ENC2);
        gemm(w.h1, CAR_C_FEAT, W.enc2, w.interp, CAR_C_LAT, rows * gs.parts, epi(W.enc2.bias, 0), st); }
    } else if (merge) {
      StageScope sc(CAR_ST_GEMM_ENC1);
      gemm(w.x, CAR_K_ENC, W.merge, w.interp, CAR_C_FEAT, rows, epi(W.merge.bias, 0), st);      // models.py:484-485 (no ReLU)
    }
    if (a.debug_interp)
      cudaMemcpyAsync(a.debug_interp + (size_t)(g0 - a.ray_begin) * gs.n * a.P * gs.ci, w.interp, (size_t)rows * gs.ci * 4,
                      cudaMemcpyDeviceToDevice, st);
    { StageScope sc(CAR_ST_GEMM_KV);
      gemm(w.interp, gs.ci, W.value, w.value, gs.L, rows, epi(W.value.bias, 0), st);              // models.py:487
      gemm(w.interp, gs.ci, W.key1, w.hid, 128, rows, epi(W.key1.bias, 1), st); }                 // :491
    gemm(w.hid, 128, W.key2, w.key, 128, rows, epi(W.key2.bias, 0), st);
    gemm(w.geom + GG_LOCAL, CAR_GG_STRIDE, W.qry1, w.hid, 128, rows, epi(W.qry1.bias, 1), st);    // :529
    gemm(w.hid, 128, W.qry2, w.q1, 128, rows, epi(W.qry2.bias, 0), st);
    }
    launch_attention1_general(a, gs, g0, g1, w.key, w.q1, w.value, w.geom, w.zsum, st);           // :532-545, 573-594
    gemm(w.zsum, gs.L, W.enc_lat, w.g, 128, nr, epi(W.enc_lat.bias, 0), st);                      // :548
    gemm(w.g, 128, W.rep1_g, w.rowbias, 128, nr, epi(W.rep1_g.bias, 0), st);
    if (tc) {
      UmmaOut oh; oh.f32 = nullptr; oh.hi = w.hid_hi; oh.lo = w.hid_lo; oh.ldc = 128;
      UmmaOut oq; oq.f32 = w.q2; oq.hi = nullptr; oq.lo = nullptr; oq.ldc = 128;
      if ((rc = launch_gemm_umma(w.loc_hi, w.loc_lo, 16, W.rep1_loc.hi, W.rep1_loc.lo, 16, rows, 128, 16, 1,
                                 epi(nullptr, 1, 0, 0, w.rowbias, gs.n * a.P), oh, st))) return rc;
      if ((rc = launch_gemm_umma(w.hid_hi, w.hid_lo, 128, W.rep2.hi, W.rep2.lo, 128, rows, 128, 128, 1, epi(W.rep2.bias, 0), oq, st))) return rc;
    } else {
      gemm(w.geom + GG_LOCAL, CAR_GG_STRIDE, W.rep1_loc, w.hid, 128, rows, epi(nullptr, 1, 0, 0, w.rowbias, gs.n * a.P), st);
      gemm(w.hid, 128, W.rep2, w.q2, 128, rows, epi(W.rep2.bias, 0), st);
    }
    launch_attention2_general(a, gs, g0, g1, w.q2, w.q1, w.value, w.zsum, w.zfin, st);            // :555-565
    if (a.debug_zfinal)
      cudaMemcpyAsync(a.debug_zfinal + (size_t)(g0 - a.ray_begin) * gs.L, w.zfin, (size_t)nr * gs.L * 4, cudaMemcpyDeviceToDevice, st);
    // colour MLP (resnet_block_fc.py:132-168)
    { StageScope sc(CAR_ST_PHI);
      launch_phi_prep_general(a, g0, g1, w.c32, st);
      gemm(w.c32, 32, W.phi_in, w.px, 128, nr, epi(W.phi_in.bias, 0), st);
      for (int i = 0; i < 3; ++i) {
        gemm(w.zfin, gs.L, W.phi_z[i], w.px, 128, nr, epi(W.phi_z[i].bias, 0, 0, 1), st);
        gemm(w.px, 128, W.phi_fc0[i], w.pnet, 128, nr, epi(W.phi_fc0[i].bias, 0, 1, 0), st);
        gemm(w.pnet, 128, W.phi_fc1[i], w.px, 128, nr, epi(W.phi_fc1[i].bias, 0, 1, 1), st);
      }
      gemm(w.px, 128, W.phi_out, w.rgb3, 3, nr, epi(W.phi_out.bias, 0, 1, 0), st);
      launch_finalize_general(a, g0, g1, w.rgb3, w.overlap, st); }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch failed: %s", cudaGetErrorString(e)); return (int)e; }
  }
  return 0;
}

}  // extern "C"
