/*
 * car_b200.h — C ABI of the B200-native per-ray rendering path.
 *
 * This is the drop-in boundary for the hot path of
 * CrossAttentionRenderer.forward(input, z=z) (reference models.py:190-626):
 * the reference has no native code, so these entry points are what a binding
 * for that path binds (ctypes stub shown in INTEGRATION.md; the in-repo host
 * side is cross_attention_renderer_b200/models.py).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless noted;
 *     nothing is allocated or freed inside the library;
 *   - every call enqueues work on the given cudaStream_t (passed as void*)
 *     and returns without synchronising; calls are re-entrant per stream;
 *   - return value: 0 = ok, <0 = argument error, >0 = cudaError_t of a failed
 *     launch; car_last_error() returns a thread-local message.
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails.
 */
#ifndef CAR_B200_H
#define CAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAR_ABI_VERSION 7

/* Arithmetic of the per-sample MLP GEMMs (everything else is fp32/fp64). */
enum car_precision {
  CAR_PREC_FP32_SIMT = 0,   /* fp32 FFMA, exact fp32 accumulate (reference arithmetic)        */
  CAR_PREC_FP32_3XBF16 = 1, /* tcgen05 kind::f16, operands split hi+lo bf16, 3 MMAs / product */
  CAR_PREC_BF16 = 2         /* tcgen05 kind::f16, single bf16 MMA, fp32 accumulate            */
};

enum { CAR_C_FEAT = 576, CAR_C_LAT = 288, CAR_C_HID = 128, CAR_C_LOCAL = 16,
       CAR_K_ENC = 592 /* 576 features + 3 tanh(pt/5) + 13 zero pad (multiple of 16) */ };

int car_version(void);
const char *car_last_error(void);

/* ------------------------------------------------------------------------
 * Feature maps.  The encoder returns z = [z1 (bn,256,H/4,W/4), z2 (bn,256,
 * H/2,W/2), z3 (bn,64,H,W)] NCHW fp32 (reference models.py:178-188).  The
 * gather wants channel-contiguous texels: one packed buffer per level,
 * layout [bn][h][w][C], fp32 or bf16.  Replaces the NCHW reads of
 * F.grid_sample at models.py:278,317.
 * ---------------------------------------------------------------------- */
size_t car_features_bytes(int bn, int H, int W, int level /*0,1,2*/, int bf16);
int car_pack_features(const float *nchw, void *nhwc, int bn, int C, int h, int w,
                      int bf16, void *stream);

/* ------------------------------------------------------------------------
 * Weights, pre-packed by the host (cross_attention_renderer_b200/packing.py):
 * every matrix is [N][K] row-major ("K-major"), K zero-padded to the stated
 * value.  *_lo pointers are only read for CAR_PREC_FP32_3XBF16; for
 * CAR_PREC_FP32_SIMT the f32 pointers are read and hi/lo ignored.
 * Names follow the reference's state_dict (models.py:102-145).
 * ---------------------------------------------------------------------- */
typedef struct car_mat {
  const float *f32;        /* [N][K] fp32                                   */
  const uint16_t *hi;      /* [N][K] bf16  (round-to-nearest of f32)         */
  const uint16_t *lo;      /* [N][K] bf16  (round-to-nearest of f32 - hi)    */
  const float *bias;       /* [N] fp32                                      */
  int32_t N, K;
} car_mat;

typedef struct car_weights {
  car_mat enc1;      /* query_encode_latent      N=576 K=592 (579 padded)            */
  car_mat enc2;      /* query_encode_latent_2    N=288 K=576                         */
  car_mat value;     /* latent_value             N=288 K=576                         */
  car_mat key1;      /* key_map                  N=128 K=576                         */
  car_mat key2;      /* key_map_2                N=128 K=128                         */
  car_mat qry1;      /* query_embed              N=128 K=16                          */
  car_mat qry2;      /* query_embed_2            N=128 K=128                         */
  car_mat rep1_loc;  /* query_repeat_embed[:,128:144]  N=128 K=16 (bias unused)      */
  car_mat rep1_g;    /* query_repeat_embed[:,0:128]    N=128 K=128 (+ its bias)      */
  car_mat rep2;      /* query_repeat_embed_2     N=128 K=128                         */
  car_mat enc_lat;   /* encode_latent (Conv1d)   N=128 K=288                         */
  car_mat phi_in;    /* phi.lin_in               N=128 K=32  (18 padded)             */
  car_mat phi_z[3];  /* phi.lin_z.i folded: W[:, :288] + W[:, 288:]   N=128 K=288    */
  car_mat phi_fc0[3];/* phi.blocks.i.fc_0        N=128 K=128                         */
  car_mat phi_fc1[3];/* phi.blocks.i.fc_1        N=128 K=128                         */
  car_mat phi_out;   /* phi.lin_out              N=3   K=128                         */
  /* [latent_value ; key_map] composed with query_encode_latent_2 (no non-linearity lies
   * between them, models.py:333-344,487-491):  F[:, v*576:(v+1)*576] = W3[:, v*288:(v+1)*288] @ W2,
   * bias = W3 @ [b2;b2] + [b_value; b_key].  Rows 0..287 -> V, rows 288..415 -> key_map pre-ReLU.
   * Used by the fused kernel (tensor-core precisions, P == 64).  N=416 K=1152. */
  car_mat kv_fold;
  /* kv_fold with its K columns permuted into the order in which the fused kernel's epilogue produces the
   * hidden activations when it runs 64-wide K stages (bf16 mode): for view v, accumulator block q = 0..8
   * (c = q / 3, j = q % 3), lane half h, i < 32:  column v*576 + q*64 + h*32 + i  holds kv_fold column
   * v*576 + c*192 + h*96 + j*32 + i.  Same bias.  hi == NULL: the kernel falls back to 32-wide stages. */
  car_mat kv_fold64;
  /* query_repeat_embed[:, :128] composed with encode_latent (Conv1d 288 -> 128, no non-linearity between them,
   * models.py:548-553): the per-ray row bias of the round-2 query MLP in ONE 128 x 288 matrix,
   * W = rep1_g @ enc_lat, bias = rep1_g @ b_enc_lat + b_query_repeat_embed.  Used by the fused-tail path. */
  car_mat rowb_fold;
  /* The ten colour-MLP matrices side by side along K (zero padded to 64-column blocks) for the fused phi kernel:
   * [lin_in 64 | lin_z0 320 | fc_0[0] 128 | fc_1[0] 128 | lin_z1 320 | fc_0[1] 128 | fc_1[1] 128 | lin_z2 320 | ...]:
   * N=128 K=1792.  Biases are read from phi_in / phi_z / phi_fc0 / phi_fc1; lin_out from phi_out.f32. */
  car_mat phi_pack;
} car_weights;

/* ------------------------------------------------------------------------
 * Cameras after the 4x4 pose algebra the reference does with torch.inverse /
 * matmul (models.py:207-211,285-286; geometry.py:404).  All row-major fp32.
 * ---------------------------------------------------------------------- */
typedef struct car_cameras {
  const float *Q;      /* (b,2,4,4)   inv(ctx c2w) @ query c2w                       */
  const float *Cself;  /* (b,2,4,4)   inv(ctx c2w) @ ctx c2w                         */
  const float *Rel;    /* (b,2,2,4,4) Rel[b][k][j] = inv(ctx_k c2w) @ ctx_j c2w       */
  const float *qinv;   /* (b,4,4)     inv(query c2w)                                 */
  const float *K;      /* (b,2,4,4)   context intrinsics (pixels)                    */
  const float *Kq;     /* (b,4,4)     query intrinsics (pixels)                      */
} car_cameras;

/* Optional per-stage dumps for parity tests (any pointer may be NULL).
 * Row index = ((scene*R + ray)*2 + ctx)*P + sample, restricted to the call's
 * ray range; "view" = which context image the features came from. */
typedef struct car_debug {
  float *geom;      /* (rows,32): gx,gy,gxc,gyc, tanh(pt_v0/5)[3], tanh(pt_v1/5)[3], clamp(pt)[3], pad[3], local[16] */
  float *x;         /* (rows,2,592) encoder inputs per view (fp32 modes only)                                      */
  float *interp;    /* (rows,576)   [enc(view0) | enc(view1)]                                                      */
  float *value;     /* (rows,288)                                                                                 */
  float *key;       /* (rows,128)                                                                                 */
  float *q1;        /* (rows,128)                                                                                 */
  float *q2;        /* (rows,128)                                                                                 */
  float *zfinal;    /* (rays,288)                                                                                 */
} car_debug;

typedef struct car_render_args {
  int32_t abi_version;            /* CAR_ABI_VERSION                                         */
  int32_t precision;              /* enum car_precision                                      */
  int32_t b, R, P, H, W;          /* scenes, rays/scene, samples/line, context image size    */
  int32_t ray_begin, ray_end;     /* this call renders flattened rays [ray_begin, ray_end)
                                     of the b*R (scene-major) rays: the multi-GPU shard     */
  int32_t feat_bf16;              /* packed features are bf16 (else fp32)                    */
  const void *feat[3];            /* packed NHWC levels, (b*2, h_l, w_l, C_l)                */
  car_weights weights;
  car_cameras cams;
  const float *uv;                /* (b,R,2) target pixel (x,y)  (models.py:198)             */
  const float *interval;          /* (P) torch.linspace(0,1,P)   (models.py:261)             */
  /* outputs: full-size tensors, only the ray range is written (models.py:217,570-621) */
  float *rgb;                     /* (b,1,R,3)                                               */
  float *valid_mask;              /* (b,R,1)                                                 */
  float *depth_ray;               /* (b,R,1)                                                 */
  float *at_wt;                   /* (b*2,R,P)  round-1 attention weights                    */
  int64_t *at_wt_max;             /* (b*2,R,1)                                               */
  float *pixel_val;               /* (b*2,R,P,2)                                             */
  float *coords;                  /* (b*2,R,9)                                               */
  void *workspace;                /* >= car_workspace_bytes(...)                             */
  size_t workspace_bytes;
  car_debug debug;                /* all-NULL in production                                  */
  void *stream;                   /* cudaStream_t                                            */
  int32_t use_fused;              /* bit 0: fused gather+encode kernel, bit 1: fused per-ray
                                     attention tail (encode: P % 64 == 0; tail: P == 64 or 128; else the unfused path).
                                     debug.interp needs bit 0 clear, debug.key/q2 bit 1 clear.   */
  int32_t train;                  /* 1: training-mode forward (reference training.py:92): the whole
                                     ray range is processed as one chunk on the unfused path and
                                     every activation stays in `workspace` (>= car_train_workspace_bytes)
                                     for car_render_backward.  CAR_PREC_FP32_SIMT (exact fp32) or
                                     CAR_PREC_FP32_3XBF16 (per-sample GEMMs of forward AND backward on
                                     tcgen05, hi + lo bf16 operands); fp32 maps.               */
  int32_t chunk_rays;             /* rays per workspace chunk; 0 = car_default_chunk_rays().  The library
                                     uses exactly this chunk (clipped to the ray range) and fails with -8
                                     if `workspace_bytes` < car_workspace_bytes(precision, P, chunk, use_fused) */
} car_render_args;

/* Rays are processed in chunks of `chunk_rays`; workspace scales with the chunk. */
size_t car_workspace_bytes(int precision, int P, int chunk_rays, int use_fused);
int car_default_chunk_rays(int precision, int P, int use_fused);
int car_render_forward(const car_render_args *args);

/* ------------------------------------------------------------------------
 * Backward pass of the path (reference: autograd of models.py:278-621 driven by
 * train_loss.backward(), training.py:125).  Gradients flow from the cotangents of
 * out['rgb'] (image loss, loss_functions.py:74-80) and out['depth_ray'] (depth
 * regulariser, :113-123) to the renderer weights and to the three feature maps; nothing
 * upstream of the sample coordinates carries a gradient (pt / depth are detached,
 * models.py:327-328,516, and the cameras are data).
 *
 * Weight gradients use the packed layout of car_weights: [N][K] fp32, K padded; bias [N],
 * and are ACCUMULATED into (+=): the caller zeroes them, which is also how several
 * micro-batches are summed.  The host maps them back to state_dict shapes
 * (cross_attention_renderer_b200/packing.py::unpack_grads).  kv_fold has no gradient of
 * its own (training runs the unfused matrices).
 * ---------------------------------------------------------------------- */
typedef struct car_mat_grad {
  float *w;       /* [N][K] fp32, same shape as car_mat.f32; NULL skips */
  float *bias;    /* [N] or NULL                                        */
} car_mat_grad;

typedef struct car_weight_grads {
  car_mat_grad enc1, enc2, value, key1, key2, qry1, qry2, rep1_loc, rep1_g, rep2, enc_lat, phi_in;
  car_mat_grad phi_z[3], phi_fc0[3], phi_fc1[3];
  car_mat_grad phi_out;
} car_weight_grads;

typedef struct car_backward_args {
  int32_t abi_version;
  const car_render_args *fwd;     /* HOST pointer: the arguments of the train=1 forward whose
                                     workspace still holds the activations (same ray range)    */
  const float *d_rgb;             /* (b,1,R,3) cotangent of out['rgb'], or NULL                 */
  const float *d_depth_ray;       /* (b,R,1)   cotangent of out['depth_ray'], or NULL           */
  car_weight_grads grads;         /* accumulated into                                           */
  float *d_feat[3];               /* packed NHWC fp32 (b*2,h_l,w_l,C_l) feature-map gradients,
                                     accumulated into (scatter-add of the bilinear taps =
                                     grid_sample backward); all NULL skips them                  */
  void *workspace;                /* >= car_backward_workspace_bytes(precision, P, rays)        */
  size_t workspace_bytes;
  void *stream;
  int32_t precision;              /* arithmetic of the gradient GEMMs of the per-sample layers:
                                     CAR_PREC_FP32_SIMT (exact fp32) or CAR_PREC_FP32_3XBF16 (tcgen05,
                                     hi + lo bf16 operands).  Independent of fwd->precision.    */
} car_backward_args;

size_t car_train_workspace_bytes(int precision, int P, int rays);
size_t car_backward_workspace_bytes(int precision, int P, int rays);
int car_render_backward(const car_backward_args *args);
/* NHWC fp32 gradient buffer -> the NCHW layout of the encoder output (inverse of car_pack_features). */
int car_unpack_features(const float *nhwc, float *nchw, int bn, int C, int h, int w, void *stream);

/* ------------------------------------------------------------------------
 * The other forward branches of the reference (models.py:219-222, 345-485):
 *   n_view = 1   one context view, features + [tanh(pt/5), tanh(pt/100)] through update_val_merge (:478-485)
 *   n_view = 3   three context views, two cross-view gathers per sample, 864-wide interleaved encode (:345-475)
 *   n_view = 2 with CAR_FLAG_NO_SAMPLE          volumetric line instead of the clipped epipolar segment
 *                                               (geometry.py:165-187)
 *   n_view = 2 with CAR_FLAG_NO_LATENT_CONCAT   raw 576-channel features, no per-sample encoder (:476-477)
 * Same stages as car_render_forward with a different gather fan-out, unfused: exact-fp32 kernels, or
 * (precision = CAR_PREC_FP32_3XBF16) the per-sample GEMMs on tcgen05.  Rows are ((scene*R + ray)*n_view + ctx)*P + sample; outputs have b*n_view leading
 * dimensions where the n_view = 2 path has b*2.
 * ---------------------------------------------------------------------- */
enum { CAR_FLAG_NO_SAMPLE = 1, CAR_FLAG_NO_LATENT_CONCAT = 2 };

typedef struct car_general_weights {
  car_mat enc1, enc2;   /* query_encode_latent(_2): N=576 K=592 / N=288 K=576 (n_view >= 2 with the latent concat) */
  car_mat merge;        /* update_val_merge, n_view = 1: N=576 K=592 (582 padded)                                  */
  car_mat value, key1;  /* latent_value / key_map: K = 576 (n_view 1, no concat) | 576 (n_view 2) | 864 (n_view 3,
                           columns re-ordered part-major: column k*288 + c holds the reference's column 3c + k)    */
  car_mat key2, qry1, qry2, rep1_loc, rep1_g, rep2;
  car_mat enc_lat;      /* N=128 K=L (L = 288, or 576 for n_view = 1 / no concat)                                  */
  car_mat phi_in;       /* N=128 K=32 (9*n_view padded)                                                           */
  car_mat phi_z[3];     /* lin_z summed over its n_view column blocks: N=128 K=L                                  */
  car_mat phi_fc0[3], phi_fc1[3];
  car_mat phi_out;
} car_general_weights;

typedef struct car_general_args {
  int32_t abi_version;
  int32_t n_view, flags;
  int32_t b, R, P, H, W;
  int32_t ray_begin, ray_end;
  const float *feat[3];           /* packed NHWC fp32 levels, (b*n_view, h_l, w_l, C_l)               */
  car_general_weights weights;
  car_cameras cams;               /* Q, Cself, K: (b,n,4,4); Rel: (b,n,n,4,4); qinv, Kq: (b,4,4)       */
  const float *uv;                /* (b,R,2)                                                          */
  const float *interval;          /* (P): linspace(0,1,P), or the depths linspace(0.1,10,P) with CAR_FLAG_NO_SAMPLE */
  float *rgb;                     /* (b,1,R,3)   */
  float *valid_mask;              /* (b,R,1)     */
  float *depth_ray;               /* (b,R,1)     */
  float *at_wt;                   /* (b*n,R,P)   */
  int64_t *at_wt_max;             /* (b*n,R,1)   */
  float *pixel_val;               /* (b*n,R,P,2) */
  float *coords;                  /* (b*n,R,9)   */
  void *workspace;                /* >= car_general_workspace_bytes(n_view, flags, P, chunk)          */
  size_t workspace_bytes;
  void *stream;
  int32_t chunk_rays;             /* 0 = car_general_default_chunk_rays                               */
  float *debug_interp;            /* optional (rows, Ci) dump of the per-sample features fed to latent_value / key_map */
  float *debug_zfinal;            /* optional (rays, L)                                               */
  int32_t precision;              /* CAR_PREC_FP32_SIMT (exact fp32) or CAR_PREC_FP32_3XBF16: the per-sample GEMMs on
                                     tcgen05 with hi+lo bf16 operands (car_mat::hi / lo); the per-ray layers stay fp32  */
} car_general_args;

size_t car_general_workspace_bytes(int n_view, int flags, int P, int chunk_rays);   /* sized for either precision */
int car_general_default_chunk_rays(int n_view, int flags, int P);
int car_render_forward_general(const car_general_args *args);

/* ------------------------------------------------------------------------
 * Optimiser step of the training loop (reference training.py:124-136: average_gradients, clip_grad_norm_(1.0),
 * torch.optim.Adam.step with betas (0.99, 0.999), train_realestate10k.py:86) as ONE kernel over a flat fp32
 * buffer holding every parameter (host side: cross_attention_renderer_b200/optim.py::FlatAdam).  Arithmetic =
 * torch.optim.Adam (amsgrad off); `grad_scale` (device scalar, may be NULL) multiplies the gradient first: the
 * clip coefficient and / or 1 / world_size.  `step` counts from 1.
 * ---------------------------------------------------------------------- */
int car_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, size_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int step, const float *grad_scale,
                  void *stream);

/* Number of kernels the last car_render_forward / car_render_backward on this thread launched. */
int car_last_launch_count(void);

/* Per-stage device timing for bench.py: between begin and end every kernel launch of this
 * thread is bracketed by CUDA events recorded on the launching stream (no host sync while
 * recording).  car_profile_end synchronises, adds the event durations per stage into
 * ms[0..n) / launches[0..n) (HOST pointers) and stops recording. */
enum car_stage { CAR_ST_RAYSETUP = 0, CAR_ST_SAMPLE_GEOM, CAR_ST_GATHER, CAR_ST_GEMM_ENC1,
                 CAR_ST_GEMM_ENC2, CAR_ST_GEMM_KV, CAR_ST_GEMM_SMALL, CAR_ST_ATTENTION,
                 CAR_ST_PHI, CAR_ST_PACK, CAR_ST_FUSED, CAR_ST_BACKWARD,
                 CAR_ST_BWD_DGRAD, CAR_ST_BWD_WGRAD, CAR_ST_BWD_OPS, CAR_ST_BWD_SCATTER, CAR_ST_COUNT };  /* FUSED = gather+enc1+enc2+kv in one kernel */
int car_profile_begin(void);
int car_profile_end(float *ms, int *launches, int n);

/* Stand-alone tcgen05 GEMM used by the tensor-core precisions, exported for
 * tests: C[M][N] = A[M][K] · W[N][K]^T (+bias) with A given as bf16 hi (+lo). */
int car_gemm_umma_test(const uint16_t *a_hi, const uint16_t *a_lo, const uint16_t *w_hi,
                       const uint16_t *w_lo, const float *bias, float *c, int M, int N, int K,
                       int split3, int relu, void *stream);

/* Diagnostics: device buffer of 32 uint64 that the fused kernel fills with per-role barrier-wait
 * cycle counts of pair 0 (layout documented in car_fused.cu); NULL disables. */
int car_debug_set_fused_stats(void *dev_u64x32);

#ifdef __cplusplus
}
#endif
#endif /* CAR_B200_H */
