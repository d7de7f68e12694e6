/*
 * nvfi_b200 — C ABI of the B200-native NVFi render + velocity-advection hot path.
 *
 * The reference (vLAR-group/NVFi) has NO native interface: its hot path is pure
 * Python/PyTorch behind `models/__init__.py`.  This header is the
 * drop-in boundary a native replacement exports: plain C structs holding device
 * pointers and scalars, `extern "C"` entry points taking a `cudaStream_t` (passed as
 * `void*`), no torch types, no global state.  Each entry point cites the reference
 * function(s) it replaces.  `nvfi_b200/_lib.py` binds it with ctypes;
 * `nvfi_b200/models/` mirrors the reference's Python surface on top.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in `_host`.
 *   - All arrays are dense, row-major, float32 unless noted.
 *   - Every function returns 0 on success, a negative NVFI_E* code on argument
 *     errors, or a positive cudaError_t value if a CUDA call failed.
 *   - Functions only enqueue work on `stream`; they never synchronise, except the
 *     `*_host` variants which copy results back and synchronise the stream.
 *   - Factor planes are consumed in a packed channels-last layout (H, W, R) produced
 *     by nvfi_pack_plane from the reference's NCHW parameter (1, R, H, W)
 *     (models/tensorf_keyframe.py:143-149); MLP weights in a transposed, zero-padded
 *     layout produced by nvfi_pack_linear from nn.Linear's (out, in).
 */
#ifndef NVFI_B200_H
#define NVFI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NVFI_ABI_VERSION 12

/* error codes */
#define NVFI_OK 0
#define NVFI_EINVAL (-1)      /* bad argument (null pointer, negative size) */
#define NVFI_EUNSUPPORTED (-2) /* shape/config outside what the kernels implement */

/* enums */
#define NVFI_ACT_SOFTPLUS 0 /* fea2denseAct, models/tensorf_keyframe.py:320-325 */
#define NVFI_ACT_RELU 1
#define NVFI_ACT_RELU_ABS 2

#define NVFI_SHADING_MLP_PE 0 /* models/tensorf_base.py:67-98 */
#define NVFI_SHADING_SH 1     /* models/tensorf_model_utils.py:292-296 */

/* arithmetic of the velocity-MLP GEMMs (NvfiField.mlp_mode; there is no process-wide setting) */
#define NVFI_MLP_DEFAULT 0   /* = NVFI_MLP_F16X3 */
#define NVFI_MLP_FP32_SIMT 1 /* FP32 FMA tile GEMM (verification path) */
#define NVFI_MLP_TF32X3 2    /* tcgen05 TF32 tensor cores, 3-term split, activations in tensor memory */
#define NVFI_MLP_TF32 3      /* tcgen05 single TF32 pass: ~1e-3 relative */
#define NVFI_MLP_F16X3 4     /* tcgen05 FP16 tensor cores, 2-way operand split (hi + lo, 22 mantissa bits),
                                3 MMAs per GEMM: FP32-grade (the product path) */

#define NVFI_GATE_AABB 0 /* VelocityAABB,    models/velocity_field.py:21-33 */
#define NVFI_GATE_SUR 1  /* VelocityAABBSur, models/velocity_field.py:36-51 */

#define NVFI_HIDDEN 128      /* width of every hidden layer on the path */
#define NVFI_VEL_IN 28       /* PositionEncoder(3) on xyzt, models/base_network.py:42-54 */
#define NVFI_VEL_LAYERS 6    /* models/velocity_field.py:58-67 */
#define NVFI_MAX_MASK_LAYERS 8
#define NVFI_MAX_MASK_DIM 32

/* One packed dense layer: wt is (k_pad, n_pad) row-major = W^T zero padded,
 * bias is (n_pad) zero padded (or NULL).  Produced by nvfi_pack_linear. */
typedef struct NvfiLinear {
  const float* wt;
  const float* bias;
  const float* w_rows; /* (128, k_pad) row-major = W zero padded, for the input-gradient GEMM
                          of the backward pass; NULL when the layer is never differentiated */
  const float* umma;   /* tensor-core image produced by nvfi_pack_linear_umma (velocity nets):
                          per 32-wide K block the TF32 "hi" slab [umma_rows][32] followed by the
                          "lo" (residual) slab, both K-major with the 128-byte swizzle; NULL when
                          the layer only runs on the FP32 SIMT path */
  const float* ummaT;  /* the same kind of image of W^T (rows = input features: 128, or 32 for the
                          28-wide first layer; K = 128 outputs), for the input-gradient GEMMs of
                          the tensor-core backward pass; NULL when never differentiated */
  const void* himg;    /* FP16-split tensor-core image produced by nvfi_pack_linear_h: per 64-wide K block
                          the FP16 "hi" slab [rows][64] followed by the "lo" (residual) slab, K-major rows
                          of 128 bytes with the 128-byte swizzle; rows = 128 (16 for the narrow head) */
  const void* himgT;   /* the same kind of image of W^T (rows = input features: 128, or 32 for the 28-wide
                          first layer; for the head: rows = 128 inputs, K = its outputs padded to 64) */
  int32_t in_dim, out_dim; /* logical sizes */
  int32_t k_pad, n_pad;    /* padded sizes: k_pad % 32 == 0, n_pad % 4 == 0 */
  int32_t umma_rows;       /* rows (N) of the image: 128 for hidden layers, 16 for a narrow head */
  int32_t ummaT_rows;
} NvfiLinear;

/* Everything the render path reads.  Mirrors the attributes of
 * TensorVMKeyframeTimeKplane (models/tensorf_keyframe.py:38-114) and TensorBase
 * (models/tensorf_base.py:134-183) that the hot path touches. */
typedef struct NvfiField {
  /* geometry (models/tensorf_base.py:214-227, 241-242) */
  float aabb_min[3], aabb_max[3];
  float inv_aabb[3]; /* 2 / (aabb_max - aabb_min), FP32 as computed by torch */
  int32_t grid[3];   /* gridSize (x, y, z) */
  int32_t num_keyframes;
  float tmax;
  float time_scale; /* tmax / (K - 1), or 1 */
  float dt_max;     /* 0.5 * tmax / (K - 1), or 1  (models/tensorf_keyframe.py:577) */
  float near, far;
  float step_size;
  int32_t n_samples;
  /* density / appearance decode */
  float density_shift, distance_scale, weight_thres;
  int32_t fea2dense_act;
  int32_t shading_mode;
  int32_t pos_pe, view_pe;
  int32_t rd, ra, app_dim; /* components per plane, appearance feature width */
  /* packed planes (H, W, R).  space k: H = grid[m1], W = grid[m0] with
   * matModeSpace = [[0,1],[0,2],[1,2]]; time k: H = K, W = grid[n0] with
   * matModeTime = [[2,3],[1,3],[0,3]]  (models/tensorf_keyframe.py:39-40) */
  const float* dplane_space[3];
  const float* dplane_time[3];
  const float* aplane_space[3];
  const float* aplane_time[3];
  NvfiLinear basis_mat;     /* Ra -> app_dim, no bias (models/tensorf_keyframe.py:129-131) */
  NvfiLinear render_mlp[3]; /* MLPRender_PE: in -> 128 -> 128 -> 3 */
  /* velocity field (models/velocity_field.py:54-98) */
  int32_t use_vel;
  NvfiLinear vel_net[NVFI_VEL_LAYERS]; /* SiLU weight net */
  NvfiLinear acc_net[NVFI_VEL_LAYERS]; /* ReLU twin net (PDE loss only) */
  int32_t vel_gate;
  float gate_lo[3], gate_hi[3]; /* velocity is zero outside [lo, hi] (normalised coords) */
  /* eval-only empty-space mask (models/tensorf_model_utils.py:417-442); NULL = none */
  const uint8_t* alpha_volume; /* (Gz, Gy, Gx) 0/1 */
  int32_t alpha_grid[3];       /* (Gx, Gy, Gz) */
  /* optional mask field (models/mask_field.py:34-83 as built at test_segm_render.py:75-80) */
  int32_t mask_layers; /* 0 = none; else number of Linear layers incl. the head */
  int32_t mask_dim;
  NvfiLinear mask_net[NVFI_MAX_MASK_LAYERS];
  /* arithmetic of the velocity-MLP GEMMs for every call that takes this field: NVFI_MLP_* */
  int32_t mlp_mode;
} NvfiField;

/* Per-call render arguments: one time for all rays (a render call in the reference
 * takes a scalar t, models/tensorf_keyframe.py:613-639). */
typedef struct NvfiRenderArgs {
  int64_t n_rays;
  const float* rays_o;  /* (n_rays, 3) */
  const float* rays_d;  /* (n_rays, 3), NOT normalised (models/camera.py:112-133) */
  const float* jitter;  /* (n_rays) stratified offsets u in [0,1), or NULL (eval) */
  int32_t ray_chunk;    /* reference chunk size (models/renderer.py:29); the
                           inside-box predicate of sample_ray is evaluated per chunk */
  const uint8_t* chunk_bg; /* (n_chunks) 1 = composite on white for this chunk, or NULL.
                              Encodes `white_bg or (training and rand < .5)`
                              (models/tensorf_keyframe.py:740). */
  int32_t white_bg;     /* used when chunk_bg == NULL */
  int32_t training;     /* 0: eval (alpha-mask skip active), 1: train */
  float t;              /* query time */
  float base_time;      /* keyframe time the samples are advected to (host: FP32 torch semantics) */
  float t_norm_base;    /* normalize_time_coord(base_time) */
  int32_t advect;       /* 0 when isclose(t, base) or !use_vel: no advection */
} NvfiRenderArgs;

/* Scratch + outputs of the forward pass; also what the backward pass re-reads. */
typedef struct NvfiRenderBuffers {
  /* per ray outputs */
  float* rgb_map;   /* (n_rays, 3) */
  float* depth_map; /* (n_rays) */
  float* acc_map;   /* (n_rays) */
  float* weights;   /* (n_rays, S) */
  float* mask_map;  /* (n_rays, mask_dim or 3), zero-filled by the caller */
  /* per sample scratch */
  float* x_adv;     /* (n_rays, S, 3) advected normalised positions of valid samples */
  uint8_t* valid;   /* (n_rays, S) */
  float* rgb;       /* (n_rays, S, 3): written only where weights > weight_thres */
  float* sigma;     /* (n_rays, S) density (saved for backward), may be NULL in eval */
  /* small */
  uint8_t* chunk_inside; /* (n_chunks) */
  int32_t* counters;     /* >= 16 ints (8-byte aligned), zeroed by the callee.  After
                            nvfi_render_backward, the int64 at [8] / [10] holds the number of
                            samples back-propagated through the appearance / velocity nets */
  int64_t* stats;        /* >= 4: [in-box samples, advected samples (those in front of each ray's termination),
                            app samples, samples sent through the velocity MLP (advected minus those outside the
                            velocity gate, which do not move); 0 for the non-product arithmetic modes], or NULL */
  float* x_mid;          /* optional (n_rays, S, 3): RK2 midpoint of the last advection step of every valid
                            sample, saved by a training forward so that the backward pass need not re-evaluate
                            the first velocity evaluation to find it (used when the call has ONE step); NULL =
                            recompute */
  /* Optional pair (both or neither), n_rays entries each: EARLY RAY TERMINATION.  With them, an advecting
   * render in NVFI_MLP_F16X3 mode marches in depth waves; a ray whose FP32 transmittance has underflowed to
   * exactly 0 at the end of a wave is not evaluated further — its remaining samples cannot change any output
   * or gradient (weights = T * alpha = 0), the reference merely computes them anyway
   * (models/tensorf_keyframe.py:641-755 has no termination).  ray_T: transmittance carried between waves;
   * ray_term[ray]: number of leading samples that were evaluated (S if the ray never terminated).  Samples
   * >= ray_term keep their `valid` flag (it is geometric) and get weights = sigma = 0; x_adv / x_mid are not
   * written there. */
  float* ray_T;
  int32_t* ray_term;
} NvfiRenderBuffers;

/* Upstream gradients and gradient accumulators for the backward pass.  Plane
 * gradients are accumulated (red.add) in the packed (H, W, R) layout — convert back with
 * nvfi_unpack_plane; linear-layer gradients come out in the packed (k_pad, n_pad) layout
 * of NvfiLinear.wt — convert back with nvfi_unpack_linear.  Every g_* accumulator must be
 * zero-initialised by the caller; scratch buffers need no initialisation. */
typedef struct NvfiRenderGrads {
  const float* g_rgb;     /* (n_rays, 3) or NULL */
  const float* g_depth;   /* (n_rays) or NULL */
  const float* g_acc;     /* (n_rays) or NULL */
  const float* g_weights; /* (n_rays, S) or NULL */
  float* g_dplane_space[3];
  float* g_dplane_time[3];
  float* g_aplane_space[3];
  float* g_aplane_time[3];
  float* g_basis_mat;      /* packed (k_pad, n_pad) */
  float* g_render_w[3];    /* packed */
  float* g_render_b[3];    /* (n_pad) */
  float* g_vel_w[NVFI_VEL_LAYERS]; /* packed */
  float* g_vel_b[NVFI_VEL_LAYERS];
  float* g_x_adv;          /* scratch (n_rays, S, 3) */
  float* g_sigma;          /* scratch (n_rays, S) */
  float* g_rgb_eff;        /* scratch (n_rays, 3): g_rgb after the clamp mask */
  float* workspace;        /* scratch: per-CTA activation stash + weight-gradient partials */
  int64_t workspace_bytes; /* >= nvfi_backward_workspace_bytes() */
} NvfiRenderGrads;

/* Destinations of nvfi_unpack_render_grads, in the PARAMETER layouts of the reference module (what
 * autograd hands to the optimiser): planes (1, R, H, W), nn.Linear weights (out, in), biases (out).
 * A NULL pointer skips that tensor. */
typedef struct NvfiParamGrads {
  float* dplane_space[3];
  float* dplane_time[3];
  float* aplane_space[3];
  float* aplane_time[3];
  float* basis_mat;
  float* render_w[3];
  float* render_b[3];
  float* vel_w[NVFI_VEL_LAYERS];
  float* vel_b[NVFI_VEL_LAYERS];
} NvfiParamGrads;

/* Gradient accumulators of nvfi_pde_loss, in the packed layouts of NvfiLinear.wt / bias
 * (convert with nvfi_unpack_linear).  All g_* must be zero-initialised by the caller. */
typedef struct NvfiPdeGrads {
  float* g_vel_w[NVFI_VEL_LAYERS]; /* weight_net   (SiLU net: through value AND Jacobian) */
  float* g_vel_b[NVFI_VEL_LAYERS];
  float* g_acc_w[NVFI_VEL_LAYERS]; /* a_weight_net (ReLU twin: through the value of a only) */
  float* g_acc_b[NVFI_VEL_LAYERS];
  float* g_acc_pts;                /* scratch (n, 3): dL/da per point */
  float* workspace;                /* scratch, >= nvfi_backward_workspace_bytes() */
  int64_t workspace_bytes;
} NvfiPdeGrads;

/* Destinations of nvfi_unpack_pde_grads: nn.Linear layouts (out, in) / (out) of both weight nets. */
typedef struct NvfiPdeParamGrads {
  float* vel_w[NVFI_VEL_LAYERS];
  float* vel_b[NVFI_VEL_LAYERS];
  float* acc_w[NVFI_VEL_LAYERS];
  float* acc_b[NVFI_VEL_LAYERS];
} NvfiPdeParamGrads;

/* ---- library info --------------------------------------------------------------- */
int nvfi_abi_version(void);
/* Bytes of `workspace` scratch nvfi_render_backward needs on the current device. */
int64_t nvfi_backward_workspace_bytes(void);

/* ---- instrumentation ---------------------------------------------------------------
 * Not part of the reference surface.  nvfi_launch_count: kernels this library has launched
 * since it was loaded (bench.py's `gpu_launches`).  With profiling enabled every launch is
 * bracketed by two CUDA events on its stream; nvfi_profile_read synchronises the device and
 * returns per-kernel totals (bench.py's live `roofline` timing — no profiler attached). */
typedef struct NvfiProfileEntry {
  char name[48];
  double ms;        /* summed device time of the launches */
  int64_t launches;
} NvfiProfileEntry;
int64_t nvfi_launch_count(void);
int nvfi_profile_enable(int on);
/* Fills up to `cap` entries; returns the number filled (< 0 on error).  reset != 0 clears. */
int nvfi_profile_read(NvfiProfileEntry* out, int cap, int reset);


/* ---- layout packing ----------------------------------------------------------------
 * Replaces nothing in the reference (it reads NCHW through F.grid_sample,
 * models/tensorf_keyframe.py:259-264); the packed layout makes one bilinear corner a
 * single contiguous R-vector. */
int nvfi_pack_plane(const float* src_nchw, float* dst_hwc, int r, int h, int w, void* stream);
int nvfi_unpack_plane(const float* src_hwc, float* dst_nchw, int r, int h, int w, void* stream);
/* nn.Linear (out,in) [+ bias] -> padded transposed (k_pad, n_pad) [+ (n_pad)] and back
 * (gradients). */
int nvfi_pack_linear(const float* w, const float* b, float* wt, float* bias_out, int out_dim,
                     int in_dim, int k_pad, int n_pad, void* stream);
int nvfi_unpack_linear(const float* wt, const float* bias_in, float* w, float* b, int out_dim,
                       int in_dim, int k_pad, int n_pad, void* stream);

/* nn.Linear (out,in) -> tensor-core image (see NvfiLinear.umma): dst holds
 * (k_pad / 32) * 2 * n_rows * 32 floats. */
int nvfi_pack_linear_umma(const float* w, float* dst, int out_dim, int in_dim, int n_rows, int k_pad,
                          void* stream);

/* nn.Linear (out,in) -> FP16-split tensor-core image (see NvfiLinear.himg).  transposed == 0: rows = outputs
 * (n_rows >= out_dim), K = inputs; transposed != 0: rows = inputs (n_rows >= in_dim), K = outputs.  K is padded
 * to k_pad (a multiple of 64); dst holds (k_pad / 64) * n_rows * 256 bytes. */
int nvfi_pack_linear_h(const float* w, void* dst, int out_dim, int in_dim, int n_rows, int k_pad,
                       int transposed, void* stream);

/* ---- rays ----------------------------------------------------------------------------
 * Camera.get_ray_bundle for selected pixels (models/camera.py:112-138):
 * pixel_ids (n) int64 flat indices (row * W + col), or NULL for all H*W pixels. */
int nvfi_raygen(const float* pose4x4, int h, int w, float focal, const int64_t* pixel_ids,
                int64_t n, float* rays_o, float* rays_d, void* stream);

/* ---- forward render -------------------------------------------------------------------
 * TensorVMKeyframeTimeKplane.forward + render_pts over all chunks of a Renderer.forward
 * call (models/renderer.py:22-56, models/tensorf_keyframe.py:613-755). */
int nvfi_render_forward(const NvfiField* field, const NvfiRenderArgs* args,
                        const NvfiRenderBuffers* buf, void* stream);

/* Backward of nvfi_render_forward (autograd of the same reference functions). */
int nvfi_render_backward(const NvfiField* field, const NvfiRenderArgs* args,
                         const NvfiRenderBuffers* buf, const NvfiRenderGrads* grads,
                         void* stream);

/* All packed gradient accumulators of a backward pass -> parameter layouts, in TWO launches (one
 * batched tiled transpose for the 12 planes, one batched un-padding transpose for the linear layers)
 * instead of one launch per tensor: at the shipped 2 048-ray training batch the per-tensor launches
 * were a tenth of the iteration. */
int nvfi_unpack_render_grads(const NvfiField* field, const NvfiRenderGrads* grads,
                             const NvfiParamGrads* out, void* stream);

/* The 24 packed gradient accumulators of nvfi_pde_loss -> nn.Linear layouts in one launch. */
int nvfi_unpack_pde_grads(const NvfiField* field, const NvfiPdeGrads* grads, const NvfiPdeParamGrads* out,
                          void* stream);

/* Host-buffer variant of the eval render (the end-to-end entry point): rays and
 * jitter come from HOST memory, rgb/depth/acc are copied back to HOST memory; the
 * device scratch in `buf` is still caller-provided.  Synchronises `stream`. */
int nvfi_render_forward_host(const NvfiField* field, const NvfiRenderArgs* args_devptrs_ignored,
                             const float* rays_o_host, const float* rays_d_host,
                             const float* jitter_host, float* dev_rays_o, float* dev_rays_d,
                             float* dev_jitter, const NvfiRenderBuffers* buf,
                             float* rgb_host, float* depth_host, float* acc_host, void* stream);

/* ---- field queries (called directly by train_segm.py:138-166, models/nvfi.py:50-64) -- */
/* integrate_pos (models/tensorf_keyframe.py:575-611): per-point t and base. x (n,3)
 * normalised; out (n,3). */
int nvfi_integrate_pos(const NvfiField* field, const float* x, const float* t, const float* base,
                       int64_t n, float* out, int32_t* counters, void* stream);
/* compute_densityfeature (models/tensorf_keyframe.py:233-272): xyzt (n,4) normalised. */
int nvfi_density_feature(const NvfiField* field, const float* xyzt, int64_t n, float* feat,
                         void* stream);
/* compute_densityfeature followed by feature2density in one pass: sigma (n). */
int nvfi_density_sigma(const NvfiField* field, const float* xyzt, int64_t n, float* sigma,
                       void* stream);
/* feature2density (models/tensorf_keyframe.py:312-325). */
int nvfi_feature2density(const NvfiField* field, const float* feat, int64_t n, float* sigma,
                         void* stream);
/* compute_appfeature (models/tensorf_keyframe.py:274-310): out (n, app_dim). */
int nvfi_app_feature(const NvfiField* field, const float* xyzt, int64_t n, float* feat,
                     int32_t* counters, void* stream);
/* VelBasis.forward / get_vel and the gated VelocityAABB(.Sur) (models/velocity_field.py):
 * out (n, 6) = [v, a] when full != 0, else (n, 3) gated velocity. */
int nvfi_velocity(const NvfiField* field, const float* xyzt, int64_t n, int32_t full, float* out,
                  int32_t* counters, void* stream);

/* ---- PDE loss (NVFi.get_vel_loss, models/nvfi.py:69-84) ----------------------------------
 * On the n occupied points xyzt (n,4) (normalised x, raw t; the occupancy filter of
 * models/nvfi.py:50-64 is nvfi_integrate_pos + nvfi_density_sigma on the host side):
 *   J = d VelBasis.forward / d(x,y,z,t)  (rows 0-2),  div = tr J[:, :3],
 *   transport = J[:, :3] v + J[:, 3] - a,
 *   loss = 5 mean(div^2) + 0.1 mean(transport^2).
 * `va` (n,6) = VelBasis.forward(xyzt) from nvfi_velocity(full = 1).  loss_sums receives
 * [sum div^2, sum transport^2] (double[2], device); the loss is 5 s0 / n + 0.1 s1 / (3 n).
 * With `grads` != NULL the gradients of that loss w.r.t. both nets are accumulated
 * (hand-written reverse pass of the forward-mode Jacobian, second order through SiLU).
 * grads->workspace is needed in both cases (per-CTA activation stash). */
int nvfi_pde_loss(const NvfiField* field, const float* xyzt, const float* va, int64_t n,
                  double* loss_sums, const NvfiPdeGrads* grads, int32_t want_grad,
                  int32_t* counters, void* stream);

/* ---- plane regularisers (one streaming pass: loss term + gradient) --------------------------
 * TVLoss.forward (utils/tensorf_utils.py:139-158) of one NCHW plane (1, C, H, W), as called by
 * TV_loss_density / TV_loss_app (models/tensorf_keyframe.py:205-231; `time_plane` = the t=True
 * variant, which weights the H differences by 3):
 *   *loss_accum += scale * 2 * (tfac * sum (x[y+1] - x[y])^2 / (C (H-1) W) + sum (x[.,x+1] - x[.,x])^2 / (C H (W-1)))
 * and, when grad != NULL, grad (same shape, overwritten) = d(that term)/dx.  `scale` carries
 * TVLoss_weight and the 1e-2 factor of the callers.  loss_accum: device double. */
int nvfi_tv_loss(const float* plane, int C, int H, int W, int time_plane, float scale,
                 double* loss_accum, float* grad, void* stream);
/* density_L1 (models/tensorf_keyframe.py:188-203): *loss_accum += scale * mean |x - offset| over n
 * elements (offset 0 for space planes, 1 for time planes); grad (n, overwritten) optional.
 * plane and grad must be 16-byte aligned. */
int nvfi_l1_loss(const float* plane, int64_t n, float offset, float scale, double* loss_accum,
                 float* grad, void* stream);

/* ---- segmentation trainer (next-row f4 of SURVEY.md section 8; not on the render path) --------------
 * K nearest neighbours of every query point among `points`, per batch element: replaces
 * pytorch3d.ops.knn_points as called by smooth_loss (utils/seg_loss.py:78-90).  query (batch, n_query, 3),
 * points (batch, n_points, 3); dist (batch, n_query, k) SQUARED L2 distances in ascending order, idx
 * (batch, n_query, k) int64 (-1 where fewer than k points exist).  k in {1, 2, 4, 8, 16}. */
int nvfi_knn_points(const float* query, const float* points, int32_t batch, int32_t n_query,
                    int32_t n_points, int32_t k, float* dist, int64_t* idx, void* stream);

/* Development aid: (tag, clock64) pairs recorded by CTA 0 of the tensor-core backward at its phase
 * boundaries into dev_buf (cap int64 entries); NULL disables. */
int nvfi_debug_timeline(long long* dev_buf, int cap);

#ifdef __cplusplus
}
#endif
#endif /* NVFI_B200_H */
