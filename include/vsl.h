/* vsl.h -- C ABI of the B200-native view-synthesis-loss library (libvsl.so).
 *
 * The reference (wrlife/tf_depth_estimation) has no FFI: its "operator interface" for this path is a set
 * of Python functions over framework tensors.  Each entry point below is what a binding for one of those
 * functions would call; the citation gives the reference function it replaces (paths relative to the
 * reference checkout).  INTEGRATION.md shows the reference-side stub for each.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous float32, layouts as in the reference (NHWC images,
 *     [B,H,W] depth, [B,H,W,2] coords with x first, row-major 3x3 / 4x4 matrices);
 *   - the caller owns every buffer, including the workspace (`ws`, size from the matching *_ws_bytes());
 *     the library allocates nothing, frees nothing and keeps no global state (one explicit exception: the peer
 *     arenas of the data-parallel optimiser step, vsl_peer_alloc / vsl_peer_free at the end of this file);
 *   - all work is enqueued on `stream` (a cudaStream_t); no call synchronises;
 *   - return value: 0 = OK, < 0 = VSL_E_* argument error (nothing enqueued), > 0 = a raw cudaError_t;
 *     no C++ exception crosses the boundary;
 *   - nullable outputs are skipped when NULL;
 *   - gather indices are exact integers for any size (the reference computes them in float32 and breaks
 *     beyond 2^24 elements, utils.py:273-294).
 */
#ifndef VSL_H_
#define VSL_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vsl_stream_t; /* cudaStream_t */

#define VSL_VERSION 100

enum { VSL_POSE_EULER = 0, VSL_POSE_ANGLEAXIS = 1, VSL_POSE_MATRIX = 2 };

enum {
  VSL_OK = 0,
  VSL_E_NULL = -1,      /* a required pointer is NULL            */
  VSL_E_SHAPE = -2,     /* non-positive or unsupported dimension */
  VSL_E_FORMAT = -3,    /* unknown pose format / mask mode       */
  VSL_E_ALIGN = -4,     /* pointer not 4-byte (or required 16-byte) aligned */
  VSL_E_UNSUPPORTED = -5
};

enum { VSL_MASK_NONE = 0, VSL_MASK_EXP = 1, VSL_MASK_CONST = 2 };
enum { VSL_IMG_F32 = 0, VSL_IMG_U8_255 = 1, VSL_IMG_U8_255_CENTRED = 2, VSL_IMG_U8_RAW = 3 };

#define VSL_MAX_SCALES 6
#define VSL_MAX_VIEWS 4

int vsl_version(void);
const char* vsl_strerror(int code);

/* ---- pose_vec2mat(vec[, format])  utils.py:79-98, utils_lr.py:106-149 (euler2mat utils.py:26-75,
 *      axis_angle_to_rotation_matrix utils_lr.py:77-103).  format: EULER or ANGLEAXIS. */
int vsl_pose_vec2mat_fwd(const float* vec /*[B,6]*/, int B, int format, float* mat /*[B,4,4]*/,
                         vsl_stream_t stream);
int vsl_pose_vec2mat_bwd(const float* vec, const float* g_mat /*[B,4,4]*/, int B, int format,
                         float* g_vec /*[B,6]*/, vsl_stream_t stream);

/* ---- projective_inverse_warp(img, depth, pose, intrinsics[, format])  utils.py:168-199,
 *      utils_lr.py:222-256 (meshgrid :142-166, pixel2cam :100-119, cam2pixel :121-140, bilinear_sampler
 *      :219-308 fused).  pose: [B,6] for EULER/ANGLEAXIS, [B,4,4] for MATRIX. */
size_t vsl_warp_ws_bytes(int B, int H, int W);
int vsl_warp_fwd(const float* img /*[B,H,W,C]*/, const float* depth /*[B,H,W]*/, const float* pose,
                 const float* K /*[B,3,3]*/, int B, int H, int W, int C, int format,
                 float* out_img /*[B,H,W,C] nullable*/, float* coords /*[B,H,W,2] nullable*/,
                 float* wmask /*[B,H,W,1] nullable*/, float* src_depth /*[B,H,W,1] nullable*/,
                 float* pose_mat /*[B,4,4] nullable*/, void* ws, vsl_stream_t stream);
/* Backward of the above for upstream gradients of (out_img, coords, wmask, src_depth, pose_mat), each
 * nullable.  g_img is accumulated with atomics (summation order not deterministic) and zeroed here first;
 * g_depth and g_pose are deterministic.  g_pose has the shape of `pose`. */
int vsl_warp_bwd(const float* img, const float* depth, const float* pose, const float* K, int B, int H,
                 int W, int C, int format, const float* g_out_img, const float* g_coords,
                 const float* g_wmask, const float* g_src_depth, const float* g_pose_mat,
                 float* g_img /*nullable*/, float* g_depth /*[B,H,W] nullable*/, float* g_pose /*nullable*/,
                 void* ws, vsl_stream_t stream);

/* ---- bilinear_sampler(imgs, coords)  utils.py:219-308; also the core of optflow_warp (utils.py:201-217)
 *      and consistent_depth_loss (utils_lr.py:369-458).  With flow != NULL the sampled coordinates are
 *      meshgrid + flow (flowx, flowy as two [B,Ht,Wt,1] planes) instead of `coords`. */
int vsl_bilinear_fwd(const float* imgs /*[B,Hs,Ws,C]*/, const float* coords /*[B,Ht,Wt,2] or NULL*/,
                     const float* flowx, const float* flowy, int B, int Hs, int Ws, int C, int Ht, int Wt,
                     float* out /*[B,Ht,Wt,C]*/, float* wmask /*[B,Ht,Wt,1] nullable*/,
                     float* coords_out /*[B,Ht,Wt,2] nullable*/, vsl_stream_t stream);
int vsl_bilinear_bwd(const float* imgs, const float* coords, const float* flowx, const float* flowy, int B,
                     int Hs, int Ws, int C, int Ht, int Wt, const float* g_out, const float* g_wmask,
                     float* g_imgs /*nullable, atomics*/, float* g_coords /*[B,Ht,Wt,2] nullable*/,
                     vsl_stream_t stream);

/* ---- consistent_depth_loss(src_depth, pred_src_depth, coords)  utils_lr.py:369-458 in one pass:
 *      err = |pred_src_depth - bilinear(src_depth, coords)|, [B,Ht,Wt,1], no reduction (as the reference).
 *      Backward for an upstream g_err: g_pred and g_coords deterministic, g_src_depth by atomics (zeroed here);
 *      each nullable. */
int vsl_consist_fwd(const float* src_depth /*[B,Hs,Ws,1]*/, const float* pred /*[B,Ht,Wt,1]*/,
                    const float* coords /*[B,Ht,Wt,2]*/, int B, int Hs, int Ws, int Ht, int Wt, float* err,
                    vsl_stream_t stream);
int vsl_consist_bwd(const float* src_depth, const float* pred, const float* coords, int B, int Hs, int Ws, int Ht,
                    int Wt, const float* g_err, float* g_src_depth, float* g_pred, float* g_coords,
                    vsl_stream_t stream);

/* ---- depth_optflow(coords)  utils.py:321-338: flow = coords - meshgrid. */
int vsl_depth_optflow(const float* coords /*[B,H,W,2]*/, int B, int H, int W, float* flowx, float* flowy,
                      vsl_stream_t stream);

/* ---- the geometry building blocks as ops of their own (the warp entry points inline them).  Planar [B,C,H,W]
 *      layouts as in the reference.
 *      meshgrid(batch, h, w, is_homogeneous)               utils.py:142-166  -> [B, 3|2, H, W]
 *      pixel2cam(depth, pixel_coords, K, is_homogeneous)   utils.py:100-119  -> [B, 4|3, H, W]
 *      cam2pixel(cam_coords, proj)                         utils.py:121-140, utils_lr.py:172-194
 *                                                          -> coords [B,H,W,2], z_u [B,H,W,1] (nullable)
 *      axis_angle_to_rotation_matrix(axis, angle)          utils_lr.py:77-103 -> [B,3,3]  */
int vsl_meshgrid(int B, int H, int W, int homogeneous, float* out, vsl_stream_t stream);
int vsl_pixel2cam_fwd(const float* depth /*[B,H,W]*/, const float* pixel_coords /*[B,3,H,W]*/,
                      const float* K /*[B,3,3]*/, int B, int H, int W, int homogeneous, float* cam,
                      vsl_stream_t stream);
int vsl_pixel2cam_bwd(const float* pixel_coords, const float* K, const float* g_cam, int B, int H, int W,
                      int homogeneous, float* g_depth /*[B,H,W]*/, vsl_stream_t stream);
int vsl_cam2pixel_fwd(const float* cam /*[B,4,H,W]*/, const float* proj /*[B,4,4]*/, int B, int H, int W,
                      float* coords, float* z, vsl_stream_t stream);
int vsl_cam2pixel_bwd(const float* cam, const float* proj, const float* g_coords /*nullable*/,
                      const float* g_z /*nullable*/, int B, int H, int W, float* g_cam /*[B,4,H,W] nullable*/,
                      float* g_proj /*[B,4,4] nullable*/, vsl_stream_t stream);
int vsl_axis_angle_fwd(const float* axis /*[B,3]*/, const float* angle /*[B]*/, int B, float* R /*[B,3,3]*/,
                       vsl_stream_t stream);
int vsl_axis_angle_bwd(const float* axis, const float* angle, const float* g_R, int B, float* g_axis,
                       float* g_angle, vsl_stream_t stream);

/* ---- compute_smooth_loss(pred_disp)  my_losses.py:27-36.  x: [B,H,W,C].  With inverse != 0 the loss is
 *      taken on 1/x (train_depth_then_cam_lr.py:217) and g_x is chained through the reciprocal.
 *      loss: one device float.  g_loss: device float (upstream), NULL means 1. */
size_t vsl_smooth_ws_bytes(int B, int H, int W, int C);
int vsl_smooth_fwd(const float* x, int B, int H, int W, int C, int inverse, float* loss, void* ws,
                   vsl_stream_t stream);
int vsl_smooth_bwd(const float* x, int B, int H, int W, int C, int inverse, const float* g_loss, float* g_x,
                   vsl_stream_t stream);

/* ---- compute_exp_reg_loss(pred, ref=[0,1])  my_losses.py:14-23,39-43: mean softmax cross-entropy of
 *      2-channel logits [N,2] against the constant label [0,1]. */
size_t vsl_expreg_ws_bytes(long long N);
int vsl_expreg_fwd(const float* logits, long long N, float* loss, void* ws, vsl_stream_t stream);
int vsl_expreg_bwd(const float* logits, long long N, const float* g_loss, float* g_logits,
                   vsl_stream_t stream);

/* ---- tf.image.resize_area pyramid (integer factors), e.g. train_depth_then_cam_lr.py:227-232.
 *      Writes levels 1..S-1 (H>>s x W>>s block means of level 0) of an NHWC image in one pass.
 *      H and W must be divisible by 2^(S-1). */
int vsl_pyramid(const float* img /*[B,H,W,C]*/, int B, int H, int W, int C, int S,
                float* const* levels /*host array of S-1 device pointers*/, vsl_stream_t stream);

/* ---- the fused multi-scale view-synthesis loss: the per-scale loop of train.py:107-135 with the
 *      explainability mask of train_depth_then_cam_lr.py:297-328 (or a constant validity mask,
 *      train_optflow_combine.py:176,187-188), forward AND backward in one pass over the data.
 *
 *      losses[5] = (pixel, smooth, exp, consist, their sum); gradients are those of losses[4] * loss_scale.  */
typedef struct {
  int B, H, W;           /* level-0 size; level s is (H>>s) x (W>>s) */
  int S, V;              /* scales (<= VSL_MAX_SCALES), source views (<= VSL_MAX_VIEWS) */
  int pose_format;       /* VSL_POSE_*: poses are [B,V,6] or [B,V,4,4] */
  int mask_mode;         /* VSL_MASK_* */
  int pixel_scale_norm;  /* data_weight / 2^s (train.py:135) or data_weight (train_depth_then_cam_lr.py:310) */
  int depth_is_inverse;  /* warp depth = 1/x (train.py:128) or x */
  int smooth_on_inverse; /* smoothness on 1/x (train_depth_then_cam_lr.py:217) or on x (train.py:108) */
  float data_weight, smooth_weight, explain_reg_weight;
  float loss_scale;      /* upstream gradient of the summed loss, folded into every gradient */
  int exact_coords;      /* 1: reference rounding sequence for coordinates / softmax / blend (bit-identical
                            sample positions to the oracle for matrix poses); 0: FMA + MUFU fast path (an even
                            number of views runs the view-paired packed-fp32x2 kernel); 2: fast path, scalar
                            kernel for any number of views (A/B comparisons) */
  int want_src_grad;     /* 1: also produce d/d(source images) into g_srcs (atomic scatter: the one output whose
                            summation order is not deterministic); needs exact_coords == 0 */
  int x_is_logit;        /* 1: x_pyr holds the disparity head's PRE-activation output; the kernel applies
                            disp = disp_scale * sigmoid(x) + disp_min (nets_optflow_depth.py:8-9,143-144) on load and
                            its derivative on store, so the head's elementwise ops and their backward disappear */
  float disp_scale, disp_min;
  int img_format;        /* VSL_IMG_*: what tgt / srcs hold.  F32: float32 (vsl_loss_fwd_bwd).  U8_*: the loader's uint8
                            (vsl_loss_fwd_bwd_u8), converted on load exactly as the reference's loaders do:
                            U8_255 (float)u8 / 255.0 (imageselect_Dataloader.py:86-93), U8_255_CENTRED ... - 0.5
                            (imageselect_Dataloader_optflow_dim11.py:128), U8_RAW the value itself
                            (imageselect_Dataloader_optflow.py:129) */
  float consist_weight;  /* > 0 (vsl_loss_consist_fwd_bwd only): the left-right depth-consistency term of
                            train_depth_then_cam_lr.py:336-340 rides on the same gather --
                            sum_s consist_weight * mean(|z_v - bilinear(source depth_v_s, coords_v)| * mask_v), z_v the
                            projected depth of the warp (utils_lr.py:172-194), the source depth map fetched by
                            consistent_depth_loss (utils_lr.py:369-458).  Needs exact_coords != 1, no want_src_grad,
                            no x_is_logit */
  float ssim_weight;     /* EXTENSION (SSIM is absent from the reference; BASELINE.json's north_star names it): a in
                            (0, 1] turns the photometric term of every scale and view into data_weight_s * [(1 - a) *
                            mean(|warp - tgt| * m) + a * mean_{B,H-2,W-2,3}(D * m_centre)], D = clip((1 - SSIM_3x3(warp,
                            tgt)) / 2, 0, 1) over VALID 3 x 3 windows (C1 = 0.01^2, C2 = 0.03^2), m the mask at the
                            window's centre: one more launch between the fused kernel and the finalize (loss_ssim_kernel:
                            warped + target tiles with their halo in shared memory, value and gradient, no
                            full-resolution intermediate).  Needs exact_coords != 1, no want_src_grad, no x_is_logit,
                            no consist_weight.  0 = off */
  void* ev_main_begin;   /* optional cudaEvent_t pair recorded on `stream` immediately around the fused  */
  void* ev_main_end;     /* loss kernel (launch 3 of the step) so a caller can time it in situ; NULL = off */
} VslLossDesc;

size_t vsl_loss_ws_bytes(const VslLossDesc* d);
/* Where the prep launch leaves its products inside the workspace (tests and tools): byte offsets, out[s] = RGB level
 * s of the target [B,Hs,Ws,3] (level 0: -1 unless the images are uint8 -- the float32 target is read in place),
 * out[S + v * S + s] = level s of source view v as zero-bordered RGBA [B,Hs+4,Ws+4,4].  out: S + V * S entries. */
int vsl_loss_ws_layout(const VslLossDesc* d, long long* out);
int vsl_loss_fwd_bwd(const VslLossDesc* d,
                     const float* tgt /*[B,H,W,3]*/, const float* const* srcs /*host array V x [B,H,W,3]*/,
                     const float* const* x_pyr /*host array S x [B,Hs,Ws,1]*/,
                     const float* poses, const float* K_pyr /*[B,S,3,3]*/,
                     const float* const* logits_pyr /*S x [B,Hs,Ws,2V], MASK_EXP*/,
                     const float* const* mask_pyr /*S x [B,Hs,Ws,1], MASK_CONST*/,
                     float* losses /*device [5]: pixel, smooth, exp, consist, their sum*/, float* const* g_x_pyr /*S x [B,Hs,Ws,1]*/,
                     float* g_poses /*same shape as poses*/, float* const* g_logits_pyr /*S, MASK_EXP*/,
                     float* const* g_srcs /*host array V x [B,H,W,3]; NULL unless want_src_grad*/,
                     void* ws, vsl_stream_t stream);
/* The same step fed with the images as the reference's input pipeline holds them before its conversion to float
 * (uint8 [B,H,W,3], imageselect_Dataloader.py:86-93): a quarter of the bytes to move and to read, results
 * bit-identical to vsl_loss_fwd_bwd on the converted images.  d->img_format selects the conversion (VSL_IMG_U8_*);
 * no gradient w.r.t. the images. */
int vsl_loss_fwd_bwd_u8(const VslLossDesc* d,
                        const unsigned char* tgt /*[B,H,W,3]*/, const unsigned char* const* srcs /*host array V*/,
                        const float* const* x_pyr, const float* poses, const float* K_pyr,
                        const float* const* logits_pyr, const float* const* mask_pyr,
                        float* losses, float* const* g_x_pyr, float* g_poses, float* const* g_logits_pyr,
                        void* ws, vsl_stream_t stream);

/* The step with the depth-consistency term (d->consist_weight > 0): src_x_pyr[v * S + s] is source view v's own
 * network output at scale s, [B,Hs,Ws,1] (its depth map is x or 1/x as d->depth_is_inverse says, like the
 * target's); the prep launch lays it into the fourth channel of that view's RGBA level, so the consistency fetch
 * is part of the 16-byte gathers the photometric term issues anyway.  g_src_x_pyr (nullable, same layout): the
 * gradient w.r.t. src_x_pyr (atomic scatter, like g_srcs).  losses[3] = the consistency term. */
int vsl_loss_consist_fwd_bwd(const VslLossDesc* d, const float* tgt, const float* const* srcs,
                             const float* const* x_pyr, const float* const* src_x_pyr /*host array V*S*/,
                             const float* poses, const float* K_pyr, const float* const* logits_pyr,
                             const float* const* mask_pyr, float* losses, float* const* g_x_pyr,
                             float* const* g_src_x_pyr /*host array V*S, nullable*/, float* g_poses,
                             float* const* g_logits_pyr, void* ws, vsl_stream_t stream);

/* ---- the flow-and-depth loss of the DeMoN-pair family: the per-scale loop of train_optflow_combine.py:138-240
 *      (optflow_warp utils.py:201-217 + depth_optflow utils.py:321-338 on the same sampler core),
 *      forward AND backward in one pass.  Per scale s, every weight / 2^s:
 *        smooth   smooth_weight * (compute_smooth_loss(pred_depth_s) + ...(flow_x_s) + ...(flow_y_s))      :142-150
 *        depth    depth_weight * mean |label_s - pred_depth_s|                                             :163-164
 *        pixel    data_weight * (mean(|warp(right_s; 1/pred_depth_s, proj, K_s) - left_s| * wmask3)
 *                              + mean(|optflow_warp(right_s, flow_x_s, flow_y_s) - left_s| * wmask3))      :169-197
 *        optflow  optflow_weight * (mean |flow_x_s - fx*| + mean |flow_y_s - fy*|),
 *                 (fx*, fy*) = depth_optflow(coords of the warp by 1/label_s)                              :204-210
 *      wmask3 = the validity mask of that ground-truth-depth warp on three channels (no gradient: data only);
 *      left_s / right_s / label_s = resize_area levels.  pred_depth is INVERSE depth (the warp uses 1/x).
 *      losses[5] = (depth, smooth, optflow, pixel, their sum = total_loss :240); gradients are those of
 *      losses[4] * loss_scale w.r.t. the three prediction pyramids.  proj: the loader's [B,4,4] target-to-source
 *      matrix (tgt2src_projs[:,0], :173).  H, W divisible by 2^(S-1); coarsest level at least 3 x 3. */
typedef struct {
  int B, H, W, S;
  float smooth_weight, depth_weight, data_weight, optflow_weight;
  float loss_scale;
} VslFlowLossDesc;

size_t vsl_flow_loss_ws_bytes(const VslFlowLossDesc* d);
int vsl_flow_loss_fwd_bwd(const VslFlowLossDesc* d, const float* left /*[B,H,W,3]*/, const float* right /*[B,H,W,3]*/,
                          const float* label /*[B,H,W,1]*/, const float* const* depth_pyr /*host array S x [B,Hs,Ws,1]*/,
                          const float* const* flowx_pyr, const float* const* flowy_pyr, const float* proj /*[B,4,4]*/,
                          const float* K_pyr /*[B,S,3,3]*/, float* losses /*device [5]*/, float* const* g_depth_pyr,
                          float* const* g_flowx_pyr, float* const* g_flowy_pyr, void* ws, vsl_stream_t stream);

/* ---- the image path of the DeMoN-pair loader on the device (imageselect_Dataloader_optflow.py:120-133, :218-236):
 *      strip = the decoded JPEG, uint8 [B,h,w,3], two frames side by side -> tf.image.resize_images(strip, [H, 2 W])
 *      (bilinear, TF 1.x ResizeBilinear semantics: no align_corners, no half-pixel centres) -> tf.to_float ->
 *      unpack_image_sequence: tgt = columns [0, W), src = columns [W, 2 W), float32 [B,H,W,3] in the 0..255 range the
 *      script feeds (its `/ 255.0` is commented out, :129). */
int vsl_unpack_strip(const unsigned char* strip, int B, int h, int w, int H, int W, float* tgt, float* src,
                     vsl_stream_t stream);

/* ---- upstream gradient of the summed loss: dst[0..n) = src[0..n) * (*num / *den) (num, den: device floats; den
 *      NULL means 1).  What TF autodiff does with the incoming gradient of `total_loss` in the reference
 *      (slim.learning.create_train_op, train_depth_then_cam_lr.py:417), applied to the whole gradient arena of a
 *      fused step in one launch; in place with a factor of exactly 1 the kernel returns without touching memory.
 *      dst, src 16-byte aligned. */
int vsl_scale(float* dst, const float* src, long long n, const float* num, const float* den, vsl_stream_t stream);

/* ---- EXTENSIONS, not in the reference (BASELINE.json's north_star names them): off unless a
 *      caller asks.  Oracle: oracle/vsl_oracle.py ssim_dissimilarity / edge_aware_smooth_loss; parity unpinned.
 *
 *      SSIM dissimilarity clip((1 - SSIM(x, y)) / 2, 0, 1) with 3x3 VALID average pools, C1 = 0.01^2,
 *      C2 = 0.03^2.  x, y: [B,H,W,C], C <= 4.  map: [B,H-2,W-2,C] (nullable); loss: device float = mean(map)
 *      (nullable; needs ws).  Backward: upstream = g_map (per element, nullable) plus, when mean_path != 0,
 *      g_loss[0] / count (g_loss NULL means 1).  Deterministic (gather form). */
size_t vsl_ssim_ws_bytes(int B, int H, int W, int C);
int vsl_ssim_fwd(const float* x, const float* y, int B, int H, int W, int C, float* map, float* loss, void* ws,
                 vsl_stream_t stream);
int vsl_ssim_bwd(const float* x, const float* y, int B, int H, int W, int C, const float* g_map,
                 const float* g_loss, int mean_path, float* g_x /*nullable*/, float* g_y /*nullable*/,
                 vsl_stream_t stream);
/*      Edge-aware first-order smoothness mean(|d_x disp| exp(-mean_c |d_x img|)) + the same along y.
 *      disp: [B,H,W,1], img: [B,H,W,C], C <= 4.  g_img nullable. */
size_t vsl_edge_smooth_ws_bytes(int B, int H, int W);
int vsl_edge_smooth_fwd(const float* disp, const float* img, int B, int H, int W, int C, float* loss, void* ws,
                        vsl_stream_t stream);
int vsl_edge_smooth_bwd(const float* disp, const float* img, int B, int H, int W, int C, const float* g_loss,
                        float* g_disp, float* g_img, vsl_stream_t stream);

/* ---- the step after the path: tf.train.AdamOptimizer(lr, beta1) as the training scripts apply it
 *      (train_depth_then_cam_lr.py:413, train.py:148), TensorFlow's ApplyAdam arithmetic, over one flat range of
 *      n parameters (param, grad, m, v cut at the same offset out of 16-byte aligned arenas).  step = t >= 1;
 *      grad is multiplied by grad_scale first (1, or 1/world to average summed rank gradients). */
int vsl_adam_step(float* param, const float* grad, float* m, float* v, long long n, float lr, float beta1,
                  float beta2, float eps, int step, float grad_scale, vsl_stream_t stream);

/*      The same step data-parallel, as ONE kernel over NVLink peer memory (reduce-scatter + Adam + all-gather):
 *      peer_grads[r] / peer_params[r] (host arrays of `world` device pointers, 16-byte aligned) are rank r's flat
 *      gradient / parameter arenas as mapped into this process (this rank's own at index `rank`).  This rank sums
 *      the gradients of all ranks over its shard [lo, hi) (multiples of 4 floats) in rank order, updates the shard
 *      with moments m_shard / v_shard (hi - lo floats each) and stores the new parameters into every rank's arena.
 *      Bracket it with vsl_peer_barrier (before: all gradients written; after: all parameters landed).
 *      vsl_peer_barrier: peer_flags[r] = rank r's flag array (>= 16 unsigned, zero-initialised) as mapped here;
 *      epoch must grow by 1 per call.  The wait is bounded in wall-clock time: timeout_ms (0 = the default of 2
 *      minutes).  If a peer does not arrive, *timed_out (int, nullable; device memory or PINNED HOST memory, so the
 *      host can poll it without synchronising) is set to 1 -- and a set *timed_out turns every later
 *      vsl_dp_adam_step / vsl_dp_step given the same pointer into a no-op (no update, no peer store): a step is never
 *      computed from a late peer's half-written gradients; the caller must treat the flag as a fatal error.
 *      vsl_dp_step: barrier -> fused step -> barrier as ONE call with the barrier epoch and Adam's step count t held in
 *      `state` (device int[4] = epoch, t, timed-out mirror, -; zero-initialised, owned by the caller; the kernels advance it), so the sequence takes no
 *      per-step host argument and can be captured into / replayed from a CUDA graph. */
/*      Peer arenas -- the one place the library allocates: memory other processes map must come straight from
 *      cudaMalloc (an IPC handle names a whole allocation).  vsl_peer_alloc zero-fills; the owner frees with
 *      vsl_peer_free after every peer has closed its mapping.  vsl_ipc_open maps a peer's arena for kernels of the
 *      CURRENT device (cudaIpcMemLazyEnablePeerAccess); handle64 is the 64-byte cudaIpcMemHandle_t. */
int vsl_peer_alloc(size_t bytes, void** ptr);
int vsl_peer_free(void* ptr);
int vsl_ipc_get_handle(void* ptr, unsigned char* handle64);
int vsl_ipc_open(const unsigned char* handle64, void** ptr);
int vsl_ipc_close(void* ptr);
int vsl_peer_barrier(unsigned* const* peer_flags, int rank, int world, unsigned epoch, int* timed_out,
                     long long timeout_ms, vsl_stream_t stream);
int vsl_dp_adam_step(const float* const* peer_grads, float* const* peer_params, int rank, int world, float* m_shard,
                     float* v_shard, long long lo, long long hi, float lr, float beta1, float beta2, float eps,
                     int step, float grad_scale, const int* timed_out /*nullable*/, vsl_stream_t stream);
int vsl_dp_step(unsigned* const* peer_flags, const float* const* peer_grads, float* const* peer_params, int rank,
                int world, float* m_shard, float* v_shard, long long lo, long long hi, float lr, float beta1, float beta2,
                float eps, float grad_scale, int* state, int* timed_out /*nullable*/, long long timeout_ms,
                vsl_stream_t stream);

/*      vsl_dp_step through the NVSwitch's multicast engine (NVLS): mc_grads / mc_params are the MULTICAST addresses of
 *      the gradient / parameter arenas (every rank's arena bound to one multicast object at the same offset, e.g.
 *      torch.distributed._symmetric_memory's multicast_ptr), own_params this rank's own parameter arena.  The shard's
 *      gradient sum is ONE multimem.ld_reduce per 16 bytes (added inside the switch), the new parameters leave as ONE
 *      multimem.st (replicated by the switch): about half the link traffic of the peer-to-peer form.  Same barriers,
 *      state block and failure behaviour as vsl_dp_step; the in-switch sum has its own fixed association order
 *      (deterministic, replicas bit-identical; not bit-identical to the rank-ordered sum).  world >= 2. */
int vsl_dp_step_mc(unsigned* const* peer_flags, const float* mc_grads, float* mc_params, const float* own_params, int rank,
                   int world, float* m_shard, float* v_shard, long long lo, long long hi, float lr, float beta1, float beta2,
                   float eps, float grad_scale, int* state, int* timed_out, long long timeout_ms, vsl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VSL_H_ */
