// C-ABI entry points (include/car_b200.h) and the host-side stage scheduler.
//
// car_render_forward walks the flattened (scene-major) ray range in chunks and, per chunk,
// enqueues the stage kernels on the caller's stream.  It replaces the body of
// CrossAttentionRenderer.forward for n_view = 2 (reference models.py:206-621) after the
// 4x4 pose algebra (models.py:207-211), which the Python host keeps in torch.
#include <stdlib.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "car_common.cuh"

namespace car {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
void reset_launch_count() { g_launches = 0; }
int sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }
  if (!cache[dev]) cudaDeviceGetAttribute(&cache[dev], cudaDevAttrMultiProcessorCount, dev);
  return cache[dev];
}

// ---- optional per-stage event timing -------------------------------------------------
struct ProfRec { cudaEvent_t a, b; int stage; };
// process-wide (not thread_local): autograd runs car_render_backward on its own thread and its launches belong to the
// profile the calling thread opened; profiling is a single-caller diagnostic
static bool g_prof_on = false;
static std::vector<ProfRec> *g_prof = nullptr;
static thread_local int g_stage = CAR_ST_GEMM_SMALL;
void set_stage(int s) { g_stage = s; }
int cur_stage() { return g_stage; }
void prof_pre(int stage, cudaStream_t st) {
  if (!g_prof_on) return;
  ProfRec r;
  r.stage = stage < 0 ? g_stage : stage;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  g_prof->push_back(r);
}
void prof_post(cudaStream_t st) {
  if (!g_prof_on) return;
  cudaEventRecord(g_prof->back().b, st);
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// use_fused: bit 0 fused gather+encode, bit 1 fused attention tail (effective only for tensor-core
// precisions with P == 64); the unfused activations are then never allocated.
// train: keep every activation the backward pass reads in its own fp32 buffer (exact-fp32 path, or the unfused
// tensor-core path, which adds the bf16 operand copies between its layers).
Workspace carve(char *base, int precision, int P, int chunk, int use_fused, int train) {
  if (train) use_fused = 0;
  const bool fused = precision != CAR_PREC_FP32_SIMT && (use_fused & 1) && P % 64 == 0;
  const bool tail = fused && (use_fused & 2) && (P == 64 || P == 128);
  Workspace w;
  memset(&w, 0, sizeof(w));
  size_t off = 0;
  size_t rows = (size_t)chunk * 2 * P;
  auto take = [&](size_t bytes) { char *p = base ? base + off : nullptr; off += align_up(bytes); return p; };
  w.seg = (RaySeg *)take((size_t)chunk * 2 * sizeof(RaySeg));
  w.overlap = (uint8_t *)take((size_t)chunk * 2);
  w.geom = (float *)take(rows * CAR_GEOM_STRIDE * 4);
  w.value = (float *)take(rows * CAR_C_LAT * 4);
  w.q1 = (float *)take(rows * 128 * 4);
  if (!tail) {
    w.key = (float *)take(rows * 128 * 4);
    w.q2 = (float *)take(rows * 128 * 4);
  }
  if (precision == CAR_PREC_FP32_SIMT || train) {
    w.x = (float *)take(rows * 2 * CAR_K_ENC * 4);
    w.h1 = (float *)take(rows * 2 * CAR_C_FEAT * 4);
    w.interp = (float *)take(rows * CAR_C_FEAT * 4);
    w.hid = (float *)take(rows * 128 * 4);
    if (train) {
      w.hid_q = (float *)take(rows * 128 * 4);
      w.hid_r = (float *)take(rows * 128 * 4);
      w.att2 = (float *)take(rows * 4);
    }
  }
  if (precision != CAR_PREC_FP32_SIMT) {
    bool lo = precision == CAR_PREC_FP32_3XBF16;
    w.hid_hi = (uint16_t *)take(rows * 128 * 2);
    if (lo) w.hid_lo = (uint16_t *)take(rows * 128 * 2);
    if (!tail) {
      w.loc_hi = (uint16_t *)take(rows * 16 * 2);
      if (lo) w.loc_lo = (uint16_t *)take(rows * 16 * 2);
    }
    if (!fused) {
      w.x_hi = (uint16_t *)take(rows * 2 * CAR_K_ENC * 2);
      w.h1_hi = (uint16_t *)take(rows * 2 * CAR_C_FEAT * 2);
      w.in_hi = (uint16_t *)take(rows * CAR_C_FEAT * 2);
      if (lo) {
        w.x_lo = (uint16_t *)take(rows * 2 * CAR_K_ENC * 2);
        w.h1_lo = (uint16_t *)take(rows * 2 * CAR_C_FEAT * 2);
        w.in_lo = (uint16_t *)take(rows * CAR_C_FEAT * 2);
      }
      if (!train) w.interp = (float *)take(rows * CAR_C_FEAT * 4);   // only filled when debug.interp is set
    }
  }
  w.zsum = (float *)take((size_t)chunk * CAR_C_LAT * 4);
  w.g = (float *)take((size_t)chunk * 128 * 4);
  w.rowbias = (float *)take((size_t)chunk * 128 * 4);
  w.zfin = (float *)take((size_t)chunk * CAR_C_LAT * 4);
  w.c18 = (float *)take((size_t)chunk * 32 * 4);
  w.px = (float *)take((size_t)chunk * 128 * 4);
  w.pnet = (float *)take((size_t)chunk * 128 * 4);
  w.rgb3 = (float *)take((size_t)chunk * 4 * 4);
  w.pr_hi = (uint16_t *)take((size_t)chunk * 576 * 2);
  w.pr_lo = (uint16_t *)take((size_t)chunk * 576 * 2);
  w.bytes = off;
  return w;
}

namespace {

GemmEpi epi(const float *bias, int relu_out, int relu_in = 0, int accumulate = 0,
            const float *row_bias = nullptr, int rows_per_group = 1) {
  GemmEpi e;
  e.bias = bias; e.row_bias = row_bias; e.rows_per_group = rows_per_group;
  e.relu_in = relu_in; e.relu_out = relu_out; e.accumulate = accumulate;
  return e;
}

void gemm(const float *A, int lda, const car_mat &m, float *C, int ldc, int M, const GemmEpi &e,
          cudaStream_t st) {
  launch_gemm_simt(A, lda, m.f32, m.K, C, ldc, M, m.N, m.K, e, st);
}

void dump(float *dst, const float *src, size_t row_off, size_t rows, size_t width, cudaStream_t st) {
  if (!dst) return;
  cudaMemcpyAsync(dst + row_off * width, src, rows * width * sizeof(float), cudaMemcpyDeviceToDevice, st);
}

// ---- per-sample stage, exact fp32 -------------------------------------------------------
void sample_stage_simt(const car_render_args &a, const Workspace &w, int g0, int g1, cudaStream_t st) {
  const car_weights &W = a.weights;
  int nr = g1 - g0;
  int rows = nr * 2 * a.P;
  launch_gather(a, g0, g1, w.geom, w.x, nullptr, nullptr, st);
  // A.7 encoder MLP on both views of every sample (models.py:333-342): M = rows*2
  { StageScope sc(CAR_ST_GEMM_ENC1);
    gemm(w.x, CAR_K_ENC, W.enc1, w.h1, CAR_C_FEAT, rows * 2, epi(W.enc1.bias, 1), st); }
  { StageScope sc(CAR_ST_GEMM_ENC2);
    gemm(w.h1, CAR_C_FEAT, W.enc2, w.interp, CAR_C_LAT, rows * 2, epi(W.enc2.bias, 0), st); }
  // A.8 value / key / geometric query (models.py:487-529)
  { StageScope sc(CAR_ST_GEMM_KV);
    gemm(w.interp, CAR_C_FEAT, W.value, w.value, CAR_C_LAT, rows, epi(W.value.bias, 0), st);
    gemm(w.interp, CAR_C_FEAT, W.key1, w.hid, 128, rows, epi(W.key1.bias, 1), st); }
  // training mode keeps the three 128-wide hidden layers apart (the backward pass reads them)
  float *hid_q = w.hid_q ? w.hid_q : w.hid, *hid_r = w.hid_r ? w.hid_r : w.hid;
  gemm(w.hid, 128, W.key2, w.key, 128, rows, epi(W.key2.bias, 0), st);
  gemm(w.geom + G_LOCAL, CAR_GEOM_STRIDE, W.qry1, hid_q, 128, rows, epi(W.qry1.bias, 1), st);
  gemm(hid_q, 128, W.qry2, w.q1, 128, rows, epi(W.qry2.bias, 0), st);
  // A.9 round 1
  launch_attention1(a, g0, g1, w.key, w.q1, w.value, w.geom, w.zsum, nullptr, st);
  // A.10 round 2: per-ray part of query_repeat_embed becomes a row bias (models.py:548-553)
  gemm(w.zsum, CAR_C_LAT, W.enc_lat, w.g, 128, nr, epi(W.enc_lat.bias, 0), st);
  gemm(w.g, 128, W.rep1_g, w.rowbias, 128, nr, epi(W.rep1_g.bias, 0), st);
  gemm(w.geom + G_LOCAL, CAR_GEOM_STRIDE, W.rep1_loc, hid_r, 128, rows,
       epi(nullptr, 1, 0, 0, w.rowbias, 2 * a.P), st);
  gemm(hid_r, 128, W.rep2, w.q2, 128, rows, epi(W.rep2.bias, 0), st);
  launch_attention2(a, g0, g1, w.q2, w.q1, w.value, w.zsum, w.zfin, w.att2, st);
}

// ---- per-sample stage, tcgen05 ------------------------------------------------------------
int sample_stage_umma(const car_render_args &a, const Workspace &w, int g0, int g1, cudaStream_t st);
int sample_stage_train_umma(const car_render_args &a, const Workspace &w, int g0, int g1, cudaStream_t st);
int phi_stage_umma(const car_render_args &a, const Workspace &w, int g0, int g1, cudaStream_t st);

// ---- per-ray colour MLP (resnet_block_fc.py:132-168), always exact fp32 ------------------
void phi_stage(const car_render_args &a, const Workspace &w, int g0, int g1, cudaStream_t st) {
  const car_weights &W = a.weights;
  int nr = g1 - g0;
  StageScope sc(CAR_ST_PHI);
  launch_phi_prep(a, g0, g1, w.c18, st);
  gemm(w.c18, 32, W.phi_in, w.px, 128, nr, epi(W.phi_in.bias, 0), st);
  for (int i = 0; i < 3; ++i) {
    gemm(w.zfin, CAR_C_LAT, W.phi_z[i], w.px, 128, nr, epi(W.phi_z[i].bias, 0, 0, 1), st);
    gemm(w.px, 128, W.phi_fc0[i], w.pnet, 128, nr, epi(W.phi_fc0[i].bias, 0, 1, 0), st);
    gemm(w.pnet, 128, W.phi_fc1[i], w.px, 128, nr, epi(W.phi_fc1[i].bias, 0, 1, 1), st);
  }
  gemm(w.px, 128, W.phi_out, w.rgb3, 3, nr, epi(W.phi_out.bias, 0, 1, 0), st);
  launch_finalize(a, g0, g1, w.rgb3, w.overlap, st);
}

}  // namespace
}  // namespace car

using namespace car;

extern "C" {

int car_version(void) { return CAR_ABI_VERSION; }
const char *car_last_error(void) { return g_err; }
int car_last_launch_count(void) { return g_launches; }

int car_profile_begin(void) {
  if (!g_prof) g_prof = new std::vector<ProfRec>();
  for (auto &r : *g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof->clear();
  g_prof_on = true;
  return 0;
}

int car_profile_end(float *ms, int *launches, int n) {
  if (!g_prof_on || !g_prof) { set_error("car_profile_end without car_profile_begin"); return -1; }
  g_prof_on = false;
  for (int i = 0; i < n; ++i) { if (ms) ms[i] = 0.f; if (launches) launches[i] = 0; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { set_error("car_profile_end: %s", cudaGetErrorString(e)); return (int)e; }
  for (auto &r : *g_prof) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    if (r.stage >= 0 && r.stage < n) { if (ms) ms[r.stage] += t; if (launches) launches[r.stage] += 1; }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof->clear();
  return 0;
}

size_t car_features_bytes(int bn, int H, int W, int level, int bf16) {
  int C = level == 2 ? 64 : 256;
  int h = level == 0 ? H / 4 : (level == 1 ? H / 2 : H);
  int w = level == 0 ? W / 4 : (level == 1 ? W / 2 : W);
  return (size_t)bn * h * w * C * (bf16 ? 2 : 4);
}

int car_pack_features(const float *nchw, void *nhwc, int bn, int C, int h, int w, int bf16,
                      void *stream) {
  if (!nchw || !nhwc || bn <= 0 || C <= 0 || h <= 0 || w <= 0) {
    set_error("car_pack_features: bad argument");
    return -1;
  }
  launch_pack_features(nchw, nhwc, bn, C, h, w, bf16, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("car_pack_features: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

int car_default_chunk_rays(int precision, int P, int use_fused) {
  const bool fused = precision != CAR_PREC_FP32_SIMT && (use_fused & 1) && P % 64 == 0;
  long rows_target = fused ? (1 << 21) : (1 << 19);   // sample rows per chunk (fused path keeps ~2.4 KB per row)
  long c = rows_target / (2 * (long)P);
  if (c < 1) c = 1;
  return (int)c;
}

size_t car_workspace_bytes(int precision, int P, int chunk_rays, int use_fused) {
  return carve(nullptr, precision, P, chunk_rays, use_fused, 0).bytes;
}

size_t car_train_workspace_bytes(int precision, int P, int rays) {
  return carve(nullptr, precision, P, rays, 0, 1).bytes;
}

int car_render_forward(const car_render_args *pa) {
  g_launches = 0;
  if (!pa) { set_error("null args"); return -1; }
  const car_render_args &a = *pa;
  if (a.abi_version != CAR_ABI_VERSION) { set_error("ABI version mismatch: got %d want %d", a.abi_version, CAR_ABI_VERSION); return -2; }
  if (a.b <= 0 || a.R <= 0 || a.P < 2 || a.P > 256 || a.H < 4 || a.W < 4 || (a.H & 3) || (a.W & 3)) {
    set_error("bad sizes b=%d R=%d P=%d H=%d W=%d", a.b, a.R, a.P, a.H, a.W); return -3;
  }
  long total = (long)a.b * a.R;
  if (a.ray_begin < 0 || a.ray_end > total || a.ray_begin > a.ray_end) { set_error("bad ray range [%d,%d) of %ld", a.ray_begin, a.ray_end, total); return -4; }
  if (a.precision < 0 || a.precision > 2) { set_error("bad precision %d", a.precision); return -5; }
  if (!a.feat[0] || !a.feat[1] || !a.feat[2] || !a.uv || !a.interval || !a.rgb || !a.valid_mask ||
      !a.depth_ray || !a.at_wt || !a.at_wt_max || !a.pixel_val || !a.coords || !a.workspace ||
      !a.cams.Q || !a.cams.Cself || !a.cams.Rel || !a.cams.qinv || !a.cams.K || !a.cams.Kq) {
    set_error("null pointer in car_render_args"); return -6;
  }
  if (a.ray_begin == a.ray_end) return 0;
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) { set_error("no CUDA device (there is no CPU fallback)"); return -7; }
  // largest chunk that fits the given workspace
  if ((a.debug.interp && (a.use_fused & 1)) || ((a.debug.key || a.debug.q2) && (a.use_fused & 2))) {
    set_error("debug taps interp/key/q2 need the unfused stages: clear the matching use_fused bits"); return -10;
  }
  if (a.chunk_rays < 0) { set_error("bad chunk_rays %d", a.chunk_rays); return -3; }
  int chunk = a.chunk_rays > 0 ? a.chunk_rays : car_default_chunk_rays(a.precision, a.P, a.use_fused);
  int span = a.ray_end - a.ray_begin;
  if (chunk > span) chunk = span;
  const int use_fused = a.train ? 0 : a.use_fused;
  if (a.train) {
    // training mode: the whole ray range is one chunk and its activations stay in the workspace
    // for car_render_backward
    if (a.precision == CAR_PREC_BF16) { set_error("train=1 needs precision CAR_PREC_FP32_SIMT or CAR_PREC_FP32_3XBF16"); return -12; }
    if (a.feat_bf16) { set_error("train=1 needs fp32 feature maps"); return -12; }
    chunk = span;
    if (carve(nullptr, a.precision, a.P, chunk, 0, 1).bytes > a.workspace_bytes) {
      set_error("train=1: workspace must hold the whole ray range (car_train_workspace_bytes), got %zu bytes", a.workspace_bytes);
      return -8;
    }
  }
  // an explicit chunk is honoured exactly; the default is halved until it fits the given workspace
  while (a.chunk_rays == 0 && chunk > 1 && carve(nullptr, a.precision, a.P, chunk, use_fused, a.train).bytes > a.workspace_bytes) chunk = (chunk + 1) / 2;
  if (carve(nullptr, a.precision, a.P, chunk, use_fused, a.train).bytes > a.workspace_bytes) { set_error("workspace too small: %zu bytes", a.workspace_bytes); return -8; }
  Workspace w = carve((char *)a.workspace, a.precision, a.P, chunk, use_fused, a.train);
  cudaStream_t st = (cudaStream_t)a.stream;

  for (int g0 = a.ray_begin; g0 < a.ray_end; g0 += chunk) {
    int g1 = g0 + chunk < a.ray_end ? g0 + chunk : a.ray_end;
    size_t row_off = (size_t)(g0 - a.ray_begin) * 2 * a.P;
    size_t rows = (size_t)(g1 - g0) * 2 * a.P;
    launch_ray_setup(a, g0, g1, w.seg, w.overlap, st);
    launch_sample_geometry(a, g0, g1, w.seg, w.geom, st);
    if (a.precision == CAR_PREC_FP32_SIMT) {
      sample_stage_simt(a, w, g0, g1, st);
      dump(a.debug.x, w.x, row_off, rows, 2 * CAR_K_ENC, st);
    } else if (a.train) {
      int rc = sample_stage_train_umma(a, w, g0, g1, st);
      if (rc) return rc;
    } else {
      int rc = sample_stage_umma(a, w, g0, g1, st);
      if (rc) return rc;
    }
    dump(a.debug.geom, w.geom, row_off, rows, CAR_GEOM_STRIDE, st);
    dump(a.debug.interp, w.interp, row_off, rows, CAR_C_FEAT, st);
    dump(a.debug.value, w.value, row_off, rows, CAR_C_LAT, st);
    dump(a.debug.key, w.key, row_off, rows, 128, st);
    dump(a.debug.q1, w.q1, row_off, rows, 128, st);
    dump(a.debug.q2, w.q2, row_off, rows, 128, st);
    dump(a.debug.zfinal, w.zfin, (size_t)(g0 - a.ray_begin), (size_t)(g1 - g0), CAR_C_LAT, st);
    if (a.precision == CAR_PREC_FP32_SIMT || a.train) phi_stage(a, w, g0, g1, st);     // per-ray layers: exact fp32 in training
    else { int rc = phi_stage_umma(a, w, g0, g1, st); if (rc) return rc; }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch failed: %s", cudaGetErrorString(e)); return (int)e; }
  }
  return 0;
}

}  // extern "C"

namespace car {
namespace {
// Colour MLP phi on tcgen05 (M = rays): x is kept in fp32 (w.px); every layer's A operand is the
// bf16 hi(+lo) copy emitted by the previous GEMM's epilogue (ReLU applied to the copy only).
int phi_stage_umma(const car_render_args &a, const Workspace &w, int g0, int g1, cudaStream_t st) {
  const car_weights &W = a.weights;
  const int split3 = a.precision == CAR_PREC_FP32_3XBF16;
  const int nr = g1 - g0;
  StageScope sc(CAR_ST_PHI);
  if (W.phi_pack.hi) {
    // one persistent kernel (car_phi.cu); its z operand is already in pr_hi / pr_lo when the fused tail produced it
    const bool tail = (a.use_fused & 1) && (a.use_fused & 2) && (a.P == 64 || a.P == 128);
    if (!tail) launch_split_rows(w.zfin, CAR_C_LAT, w.pr_hi, split3 ? w.pr_lo : nullptr, nr, CAR_C_LAT, st);
    return launch_phi_fused(a, g0, g1, w.pr_hi, split3 ? w.pr_lo : nullptr, w.overlap, st);
  }
  uint16_t *c_hi = w.pr_hi, *z_hi = c_hi + (size_t)nr * 32, *x_hi = z_hi + (size_t)nr * 288, *n_hi = x_hi + (size_t)nr * 128;
  uint16_t *c_lo = w.pr_lo, *z_lo = c_lo + (size_t)nr * 32, *x_lo = z_lo + (size_t)nr * 288, *n_lo = x_lo + (size_t)nr * 128;
  if (!split3) c_lo = z_lo = x_lo = n_lo = nullptr;
  launch_phi_prep(a, g0, g1, w.c18, st);
  launch_split_rows(w.c18, 32, c_hi, c_lo, nr, 32, st);
  launch_split_rows(w.zfin, CAR_C_LAT, z_hi, z_lo, nr, CAR_C_LAT, st);
  int rc;
  auto mm = [&](const uint16_t *ah, const uint16_t *al, int lda, const car_mat &m, const GemmEpi &e, const UmmaOut &o) {
    return launch_gemm_umma(ah, al, lda, m.hi, m.lo, m.K, nr, m.N, m.K, split3, e, o, st);
  };
  auto out = [&](float *f32, bool add, uint16_t *hi, uint16_t *lo) {
    UmmaOut o; o.f32 = f32; o.f32_add = add ? f32 : nullptr; o.hi = hi; o.lo = lo; o.ldc = 128; return o; };
  if ((rc = mm(c_hi, c_lo, 32, W.phi_in, epi(W.phi_in.bias, 0), out(w.px, false, nullptr, nullptr)))) return rc;
  for (int i = 0; i < 3; ++i) {
    // x += lin_z[i](z); operand copy = relu(x)
    if ((rc = mm(z_hi, z_lo, CAR_C_LAT, W.phi_z[i], epi(W.phi_z[i].bias, 2, 0, 1), out(w.px, true, x_hi, x_lo)))) return rc;
    // relu(fc_0(relu(x)))
    if ((rc = mm(x_hi, x_lo, 128, W.phi_fc0[i], epi(W.phi_fc0[i].bias, 1), out(nullptr, false, n_hi, n_lo)))) return rc;
    // x += fc_1(relu(net))
    if ((rc = mm(n_hi, n_lo, 128, W.phi_fc1[i], epi(W.phi_fc1[i].bias, 0, 0, 1), out(w.px, true, nullptr, nullptr)))) return rc;
  }
  gemm(w.px, 128, W.phi_out, w.rgb3, 3, nr, epi(W.phi_out.bias, 0, 1, 0), st);       // N = 3: exact fp32
  launch_finalize(a, g0, g1, w.rgb3, w.overlap, st);
  return 0;
}

// Training forward in the tensor-core precision: the dataflow of sample_stage_simt, every per-sample GEMM on tcgen05
// with hi + lo bf16 operands; each epilogue writes the fp32 activation car_render_backward reads AND the operand copy
// of the next layer.  The per-ray layers (M = rays) stay exact fp32.
int sample_stage_train_umma(const car_render_args &a, const Workspace &w, int g0, int g1, cudaStream_t st) {
  const car_weights &W = a.weights;
  const int nr = g1 - g0, rows = nr * 2 * a.P;
  int rc;
  auto out = [&](float *f32, int ldc, uint16_t *hi = nullptr, uint16_t *lo = nullptr) {
    UmmaOut o; o.f32 = f32; o.hi = hi; o.lo = lo; o.ldc = ldc; return o; };
  auto mm = [&](const uint16_t *ah, const uint16_t *al, int lda, const car_mat &m, int M, const GemmEpi &e, const UmmaOut &o) {
    return launch_gemm_umma(ah, al, lda, m.hi, m.lo, m.K, M, m.N, m.K, 1, e, o, st);
  };
  launch_gather(a, g0, g1, w.geom, w.x, nullptr, nullptr, st);
  launch_split_rows(w.x, CAR_K_ENC, w.x_hi, w.x_lo, rows * 2, CAR_K_ENC, st);
  launch_split_rows(w.geom + G_LOCAL, CAR_GEOM_STRIDE, w.loc_hi, w.loc_lo, rows, 16, st);
  { StageScope sc(CAR_ST_GEMM_ENC1);
    if ((rc = mm(w.x_hi, w.x_lo, CAR_K_ENC, W.enc1, rows * 2, epi(W.enc1.bias, 1), out(w.h1, CAR_C_FEAT, w.h1_hi, w.h1_lo)))) return rc; }
  { StageScope sc(CAR_ST_GEMM_ENC2);
    if ((rc = mm(w.h1_hi, w.h1_lo, CAR_C_FEAT, W.enc2, rows * 2, epi(W.enc2.bias, 0), out(w.interp, CAR_C_LAT, w.in_hi, w.in_lo)))) return rc; }
  { StageScope sc(CAR_ST_GEMM_KV);
    if ((rc = mm(w.in_hi, w.in_lo, CAR_C_FEAT, W.value, rows, epi(W.value.bias, 0), out(w.value, CAR_C_LAT)))) return rc;
    if ((rc = mm(w.in_hi, w.in_lo, CAR_C_FEAT, W.key1, rows, epi(W.key1.bias, 1), out(w.hid, 128, w.hid_hi, w.hid_lo)))) return rc; }
  if ((rc = mm(w.hid_hi, w.hid_lo, 128, W.key2, rows, epi(W.key2.bias, 0), out(w.key, 128)))) return rc;
  if ((rc = mm(w.loc_hi, w.loc_lo, 16, W.qry1, rows, epi(W.qry1.bias, 1), out(w.hid_q, 128, w.hid_hi, w.hid_lo)))) return rc;
  if ((rc = mm(w.hid_hi, w.hid_lo, 128, W.qry2, rows, epi(W.qry2.bias, 0), out(w.q1, 128)))) return rc;
  launch_attention1(a, g0, g1, w.key, w.q1, w.value, w.geom, w.zsum, nullptr, st);
  gemm(w.zsum, CAR_C_LAT, W.enc_lat, w.g, 128, nr, epi(W.enc_lat.bias, 0), st);
  gemm(w.g, 128, W.rep1_g, w.rowbias, 128, nr, epi(W.rep1_g.bias, 0), st);
  if ((rc = mm(w.loc_hi, w.loc_lo, 16, W.rep1_loc, rows, epi(nullptr, 1, 0, 0, w.rowbias, 2 * a.P),
               out(w.hid_r, 128, w.hid_hi, w.hid_lo)))) return rc;
  if ((rc = mm(w.hid_hi, w.hid_lo, 128, W.rep2, rows, epi(W.rep2.bias, 0), out(w.q2, 128)))) return rc;
  launch_attention2(a, g0, g1, w.q2, w.q1, w.value, w.zsum, w.zfin, w.att2, st);
  return 0;
}

// Tensor-core per-sample stage: same dataflow as sample_stage_simt with every GEMM on
// tcgen05 (car_gemm_umma.cu), operands kept as bf16 hi(+lo) between stages.
int sample_stage_umma(const car_render_args &a, const Workspace &w, int g0, int g1, cudaStream_t st) {
  const car_weights &W = a.weights;
  const int split3 = a.precision == CAR_PREC_FP32_3XBF16;
  const int nr = g1 - g0;
  const int rows = nr * 2 * a.P;
  int rc;
  auto out_bf = [&](uint16_t *hi, uint16_t *lo, int ldc, float *f32 = nullptr) {
    UmmaOut o; o.f32 = f32; o.hi = hi; o.lo = split3 ? lo : nullptr; o.ldc = ldc; return o; };
  auto out_f = [&](float *f32, int ldc) { UmmaOut o; o.f32 = f32; o.hi = nullptr; o.lo = nullptr; o.ldc = ldc; return o; };
  auto mm = [&](const uint16_t *ah, const uint16_t *al, int lda, const car_mat &m, int M, const GemmEpi &e,
                const UmmaOut &o) {
    return launch_gemm_umma(ah, al, lda, m.hi, m.lo, m.K, M, m.N, m.K, split3, e, o, st);
  };
  const bool fused = (a.use_fused & 1) && a.P % 64 == 0;
  const bool tail = fused && (a.use_fused & 2) && (a.P == 64 || a.P == 128);
  if (fused && !W.kv_fold.hi) { set_error("use_fused needs weights.kv_fold"); return -11; }
  if (!tail) launch_split_rows(w.geom + G_LOCAL, CAR_GEOM_STRIDE, w.loc_hi, split3 ? w.loc_lo : nullptr, rows, 16, st);
  if (fused) {
    // gather + enc1 + (enc2 ∘ [value; key1]) in one CTA-pair kernel: V and relu(key1) per sample
    if ((rc = launch_fused_encode(a, g0, g1, w.geom, w.value, w.hid_hi, split3 ? w.hid_lo : nullptr, st))) return rc;
    if (tail) {
      // per-ray tail: K, Q1, attention round 1 | per-ray 288->128->128 | Q2, attention round 2
      // zsum also leaves phase A as bf16 hi(+lo) (in the colour MLP's operand scratch, free until phi runs): the
      // per-ray row bias is then ONE tcgen05 GEMM with rep1_g o enc_lat folded (was two exact-fp32 SIMT GEMMs: 6.2 ms/step)
      const bool fold = W.rowb_fold.hi != nullptr;
      if ((rc = launch_tail(a, 0, g0, g1, w.geom, w.value, w.hid_hi, w.hid_lo, w.q1, w.zsum, nullptr, nullptr,
                            fold ? w.pr_hi : nullptr, fold && split3 ? w.pr_lo : nullptr, st))) return rc;
      if (fold) {
        if ((rc = mm(w.pr_hi, w.pr_lo, CAR_C_LAT, W.rowb_fold, nr, epi(W.rowb_fold.bias, 0), out_f(w.rowbias, 128)))) return rc;
      } else {
        gemm(w.zsum, CAR_C_LAT, W.enc_lat, w.g, 128, nr, epi(W.enc_lat.bias, 0), st);
        gemm(w.g, 128, W.rep1_g, w.rowbias, 128, nr, epi(W.rep1_g.bias, 0), st);
      }
      // phase B leaves z as fp32 (debug taps) and as bf16 hi(+lo): the z operand of the fused colour MLP
      const bool zpk = W.phi_pack.hi != nullptr;
      if ((rc = launch_tail(a, 1, g0, g1, w.geom, w.value, nullptr, nullptr, w.q1, w.zsum, w.rowbias, w.zfin,
                            zpk ? w.pr_hi : nullptr, zpk && split3 ? w.pr_lo : nullptr, st))) return rc;
      return 0;
    }
  } else {
    launch_gather(a, g0, g1, w.geom, nullptr, w.x_hi, w.x_lo, st);
    // A.7 encoder MLP on both views of every sample: M = rows*2
    { StageScope sc(CAR_ST_GEMM_ENC1);
      if ((rc = mm(w.x_hi, w.x_lo, CAR_K_ENC, W.enc1, rows * 2, epi(W.enc1.bias, 1), out_bf(w.h1_hi, w.h1_lo, CAR_C_FEAT)))) return rc; }
    { StageScope sc(CAR_ST_GEMM_ENC2);
      if ((rc = mm(w.h1_hi, w.h1_lo, CAR_C_FEAT, W.enc2, rows * 2, epi(W.enc2.bias, 0),
                   out_bf(w.in_hi, w.in_lo, CAR_C_LAT, a.debug.interp ? w.interp : nullptr)))) return rc; }
    // A.8 value / key / geometric query
    { StageScope sc(CAR_ST_GEMM_KV);
      if ((rc = mm(w.in_hi, w.in_lo, CAR_C_FEAT, W.value, rows, epi(W.value.bias, 0), out_f(w.value, CAR_C_LAT)))) return rc;
      if ((rc = mm(w.in_hi, w.in_lo, CAR_C_FEAT, W.key1, rows, epi(W.key1.bias, 1), out_bf(w.hid_hi, w.hid_lo, 128)))) return rc; }
  }
  if ((rc = mm(w.hid_hi, w.hid_lo, 128, W.key2, rows, epi(W.key2.bias, 0), out_f(w.key, 128)))) return rc;
  if ((rc = mm(w.loc_hi, w.loc_lo, 16, W.qry1, rows, epi(W.qry1.bias, 1), out_bf(w.hid_hi, w.hid_lo, 128)))) return rc;
  if ((rc = mm(w.hid_hi, w.hid_lo, 128, W.qry2, rows, epi(W.qry2.bias, 0), out_f(w.q1, 128)))) return rc;
  launch_attention1(a, g0, g1, w.key, w.q1, w.value, w.geom, w.zsum, nullptr, st);
  // per-ray 288->128->128 (M = rays): exact fp32
  gemm(w.zsum, CAR_C_LAT, W.enc_lat, w.g, 128, nr, epi(W.enc_lat.bias, 0), st);
  gemm(w.g, 128, W.rep1_g, w.rowbias, 128, nr, epi(W.rep1_g.bias, 0), st);
  if ((rc = mm(w.loc_hi, w.loc_lo, 16, W.rep1_loc, rows, epi(nullptr, 1, 0, 0, w.rowbias, 2 * a.P),
               out_bf(w.hid_hi, w.hid_lo, 128)))) return rc;
  if ((rc = mm(w.hid_hi, w.hid_lo, 128, W.rep2, rows, epi(W.rep2.bias, 0), out_f(w.q2, 128)))) return rc;
  launch_attention2(a, g0, g1, w.q2, w.q1, w.value, w.zsum, w.zfin, nullptr, st);
  return 0;
}
}  // namespace
}  // namespace car
