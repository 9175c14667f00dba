// Shared declarations of the B200 per-ray renderer kernels (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "car_b200.h"

#define CAR_GEOM_STRIDE 32   // floats per sample row in the geometry record

// Geometry record layout (one row per epipolar sample), see car_debug::geom
enum {
  G_GX = 0, G_GY = 1,        // primary grid coords (pixel_val)
  G_GXC = 2, G_GYC = 3,      // cross-view grid coords
  G_T0 = 4,                  // tanh(pt in view-0 frame / 5) [3]
  G_T1 = 7,                  // tanh(pt in view-1 frame / 5) [3]
  G_PTC = 10,                // clamp(pt, -100, 100) [3], then 3 floats of padding
  G_LOCAL = 16,              // local_coords [16] (16-byte aligned: it is a GEMM A operand)
};

// Per (ray, ctx) record produced by the ray set-up kernel.
struct RaySeg {
  float sx, sy, ex, ey;      // segment start / end in grid coords (after scrub)
};

namespace car {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
void reset_launch_count();
int sm_count();                     // multiprocessors of the CURRENT device (cached per device)
// Profiling hooks (car_profile_begin/end): call around a kernel launch.
void prof_pre(int stage, cudaStream_t st);
void prof_post(cudaStream_t st);
void set_stage(int stage);          // stage id applied to the following launches
int cur_stage();
struct StageScope {                 // RAII: tag launches in a scope
  int prev;
  explicit StageScope(int s) : prev(cur_stage()) { set_stage(s); }
  ~StageScope() { set_stage(prev); }
};

// ---- car_geometry.cu (compiled with -fmad=false) -------------------------
void launch_ray_setup(const car_render_args &a, int g0, int g1, RaySeg *seg, uint8_t *overlap,
                      cudaStream_t st);
void launch_sample_geometry(const car_render_args &a, int g0, int g1, const RaySeg *seg,
                            float *geom, cudaStream_t st);

// ---- car_gather.cu -------------------------------------------------------
void launch_pack_features(const float *nchw, void *nhwc, int bn, int C, int h, int w, int bf16,
                          cudaStream_t st);
// Builds encoder inputs X[row][view][592] from the packed maps.
//   out_f32 != null: fp32 rows; else bf16 hi (+lo if out_lo != null)
void launch_gather(const car_render_args &a, int g0, int g1, const float *geom, float *out_f32,
                   uint16_t *out_hi, uint16_t *out_lo, cudaStream_t st);

// fp32 rows (row stride src_stride) -> contiguous bf16 hi (+lo) rows of `width` columns
void launch_split_rows(const float *src, int src_stride, uint16_t *hi, uint16_t *lo, int rows,
                       int width, cudaStream_t st);

// ---- car_gemm_simt.cu ------------------------------------------------------
struct GemmEpi {
  const float *bias;        // [N] or null
  const float *row_bias;    // [M / rows_per_group][N] or null (per-ray bias)
  int rows_per_group;
  int relu_in;              // apply relu to A on load
  int relu_out;             // tcgen05 GEMM: 2 = ReLU only on the bf16 operand copy, fp32 output stays linear
  int accumulate;           // C += result
  const float *mask = nullptr;   // backward: result *= (mask[m][n] > 0) before accumulate (ReLU subgradient)
  int ldmask = 0;
};
void launch_gemm_simt(const float *A, int lda, const float *W, int ldw, float *C, int ldc, int M,
                      int N, int K, const GemmEpi &epi, cudaStream_t st);

// dW[N][K] += dY[M][N]^T · act(A[M][K]),  db[N] += column sums of dY  (exact fp32, atomics)
void launch_wgrad_simt(const float *dY, int ldy, const float *A, int lda, float *dW, int ldw, float *db,
                       int M, int N, int K, int relu_a, cudaStream_t st);
// dst[C][R] = src[R][C]^T   (small weight matrices)
void launch_transpose(const float *src, int rows, int cols, int ld_src, float *dst, cudaStream_t st);

// ---- car_attention.cu ------------------------------------------------------
void launch_attention1(const car_render_args &a, int g0, int g1, const float *key, const float *q1,
                       const float *value, const float *geom, float *zsum, float *rowbias,
                       cudaStream_t st);
void launch_attention2(const car_render_args &a, int g0, int g1, const float *q2, const float *q1,
                       const float *value, const float *zsum, float *zfin, float *att2, cudaStream_t st);
void launch_phi_prep(const car_render_args &a, int g0, int g1, float *c18, cudaStream_t st);
void launch_finalize(const car_render_args &a, int g0, int g1, const float *rgb3,
                     const uint8_t *overlap, cudaStream_t st);

// ---- car_gemm_umma.cu ------------------------------------------------------
struct UmmaOut {
  float *f32;               // [M][ldc] fp32 output or null
  const float *f32_add = nullptr;   // with GemmEpi::accumulate: fp32 [M][ldc] added to the product (may alias f32)
  uint16_t *hi, *lo;        // [M][ldc] bf16 split output or null
  int ldc;
  int atomic = 0;           // weight-gradient form: f32 += product, contraction split over CTAs, fp32 atomics
};
int launch_gemm_umma(const uint16_t *a_hi, const uint16_t *a_lo, int lda, const uint16_t *w_hi,
                     const uint16_t *w_lo, int ldw, int M, int N, int K, int split3,
                     const GemmEpi &epi, const UmmaOut &out, cudaStream_t st);

// ---- car_fused.cu -----------------------------------------------------------------------
int launch_fused_encode(const car_render_args &a, int g0, int g1, const float *geom, float *value,
                        uint16_t *kh_hi, uint16_t *kh_lo, cudaStream_t st);

// ---- car_phi.cu: colour MLP + mask / white fill in one persistent tcgen05 kernel -----------
int launch_phi_fused(const car_render_args &a, int g0, int g1, const uint16_t *z_hi, const uint16_t *z_lo,
                     const uint8_t *overlap, cudaStream_t st);

// ---- car_tail.cu ------------------------------------------------------------------------
int launch_tail(const car_render_args &a, int phase, int g0, int g1, const float *geom, const float *value,
                const uint16_t *kh_hi, const uint16_t *kh_lo, float *q1, float *zsum, const float *rowbias,
                float *zfin, uint16_t *zs_hi, uint16_t *zs_lo, cudaStream_t st);


// ---- general branches (n_view 1 / 3, no_sample, no_latent_concat): car_general.cu ----------
#define CAR_GG_STRIDE 48     // floats per sample row in the general geometry record
enum {
  GG_GX = 0, GG_GY = 1,      // primary grid coords
  GG_C0 = 2, GG_C1 = 4,      // cross-view grid coords of encoder parts 1 and 2 (x, y each)
  GG_T = 6,                  // tanh triples of parts 0, 1, 2 (n_view = 1: tanh(pt/5), tanh(pt/100))
  GG_PTC = 15,               // clamp(pt, -100, 100) [3]
  GG_LOCAL = 32,             // local_coords [16]
};
struct GenShape {             // derived sizes of a general-branch call
  int n, flags, parts, xw, ci, L;   // contexts; flags; encoder parts per row; X width per part; interp width; latent width
};
inline GenShape gen_shape(int n_view, int flags) {
  GenShape g;
  g.n = n_view; g.flags = flags;
  const bool noconcat = (flags & CAR_FLAG_NO_LATENT_CONCAT) != 0;
  g.parts = noconcat ? 1 : (n_view == 1 ? 1 : n_view);
  g.xw = noconcat ? CAR_C_FEAT : CAR_K_ENC;
  g.ci = noconcat ? CAR_C_FEAT : (n_view == 1 ? CAR_C_FEAT : CAR_C_LAT * n_view);
  g.L = (noconcat || n_view == 1) ? CAR_C_FEAT : CAR_C_LAT;
  return g;
}
void launch_ray_setup_general(const car_general_args &a, int g0, int g1, RaySeg *seg, uint8_t *overlap, cudaStream_t st);
void launch_sample_geometry_general(const car_general_args &a, int g0, int g1, const RaySeg *seg, float *geom, cudaStream_t st);
void launch_gather_general(const car_general_args &a, const GenShape &gs, int g0, int g1, const float *geom, float *x, cudaStream_t st);
void launch_attention1_general(const car_general_args &a, const GenShape &gs, int g0, int g1, const float *key, const float *q1,
                               const float *value, const float *geom, float *zsum, cudaStream_t st);
void launch_attention2_general(const car_general_args &a, const GenShape &gs, int g0, int g1, const float *q2, const float *q1,
                               const float *value, const float *zsum, float *zfin, cudaStream_t st);
void launch_phi_prep_general(const car_general_args &a, int g0, int g1, float *c32, cudaStream_t st);
void launch_finalize_general(const car_general_args &a, int g0, int g1, const float *rgb3, const uint8_t *overlap, cudaStream_t st);

// ---- car_api.cu: workspace carve-up for one ray chunk ------------------------------------
struct Workspace {
  RaySeg *seg; uint8_t *overlap;
  float *geom;
  // fp32 SIMT path
  float *x, *h1, *interp, *value, *hid, *key, *q1, *q2;
  // training mode (car_render_args::train): activations a backward pass needs and the
  // inference path overwrites
  float *hid_q, *hid_r, *att2;
  // tensor-core path (bf16 hi/lo operand copies)
  uint16_t *x_hi, *x_lo, *h1_hi, *h1_lo, *in_hi, *in_lo, *hid_hi, *hid_lo, *loc_hi, *loc_lo;
  // per ray
  float *zsum, *g, *rowbias, *zfin, *c18, *px, *pnet, *rgb3;
  uint16_t *pr_hi, *pr_lo;     // per-ray bf16 operand copies for the tensor-core phi: [c18 32 | zfin 288 | relu(x) 128 | relu(net) 128]
  size_t bytes;
};
Workspace carve(char *base, int precision, int P, int chunk, int use_fused, int train);

// ---- car_backward.cu / car_gather.cu ------------------------------------------------------
// scatter-add of d(encoder input)[row][view][576] into the packed NHWC feature-map gradients
void launch_gather_backward(const car_render_args &a, int g0, int g1, const float *geom, const float *dx,
                            float *const d_feat[3], cudaStream_t st);
void launch_unpack_features(const float *nhwc, float *nchw, int bn, int C, int h, int w, cudaStream_t st);

}  // namespace car
