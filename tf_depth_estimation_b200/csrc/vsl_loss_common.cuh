// Definitions shared by the two fused loss kernels (vsl_loss.cu: general; vsl_loss_pair.cu: view-paired fast path).
#pragma once
#include "vsl_common.cuh"

namespace vsl {

#ifndef VSL_RH
#define VSL_RH 32
#endif
// 4 independent warps per block, 4 blocks per SM (128 registers x 512 threads fill the register file): measured
// best of {1x16, 2x8, 4x4, 8x2, 16x1}; block granularity only matters for how evenly the tail drains
#ifndef VSL_FUSED_WARPS
#define VSL_FUSED_WARPS 4
#endif
#ifndef VSL_FUSED_MIN_BLOCKS
#define VSL_FUSED_MIN_BLOCKS 4
#endif

// one source view in the fast arithmetic needs 96 registers: a fifth block per SM (20 warps) without a spill
#ifndef VSL_V1_MIN_BLOCKS
#define VSL_V1_MIN_BLOCKS VSL_FUSED_MIN_BLOCKS
#endif

constexpr int kRH = VSL_RH;                     // tile rows per warp
constexpr int kWarps = VSL_FUSED_WARPS;         // independent warps per block
constexpr int kThreads = 32 * kWarps;
constexpr int kHalo = 2;
constexpr int kQH = kRH + 2 * kHalo;            // rows of x held per tile
constexpr int kQS = 32 + 2 * kHalo;             // 36 columns of x
constexpr int kOH = kRH + kHalo;                // owner rows: tile + 2 above
constexpr int kPad = 2;                         // zero border (pixels) of the RGBA source levels
constexpr float kMagic = 12582912.0f;           // 1.5 * 2^23
constexpr unsigned kMagicBits = 0x4B400000u;

struct LossParams {
  int B, H, W, S, V;
  int mask_mode, depth_is_inverse, smooth_on_inverse;
  int x_is_logit;                  // x = pre-activation of the disparity head: disp = disp_scale * sigmoid(x) + disp_min
  float disp_scale, disp_min;
  const float* tgt[VSL_MAX_SCALES];                    // RGB
  const float4* src[VSL_MAX_VIEWS][VSL_MAX_SCALES];    // zero-bordered RGBA [B][Hs+4][Ws+4]
  float4* gsrc[VSL_MAX_VIEWS][VSL_MAX_SCALES];         // d/d(those levels), same layout (only with DSRC)
  const float* x[VSL_MAX_SCALES];
  const float* logits[VSL_MAX_SCALES];
  const float* mask[VSL_MAX_SCALES];
  float* g_x[VSL_MAX_SCALES];
  float* g_logits[VSL_MAX_SCALES];
  const Xform* xf;                 // [S][V][B]
  const XformQ* xq;                // [S][V][B], folded transforms of the view-paired kernel
  float* partials;                 // [n_items][NT]
  float* gsd[VSL_MAX_VIEWS][VSL_MAX_SCALES];           // consistency term: d/d(source depth), zero-bordered float planes (nullable)
  float ccon[VSL_MAX_SCALES];      // loss_scale * consist_weight / (B Hs Ws); 0 without the consistency term
  float cpix[VSL_MAX_SCALES];      // loss_scale * data_weight_s / (B Hs Ws 3)
  float cexp[VSL_MAX_SCALES];      // loss_scale * explain_reg_weight / (B Hs Ws)
  // loss_scale * smooth_weight / 2^s / count_k, one flat array per second difference (a per-scale constant the
  // compiler re-loads inside the row loop should be ONE indexed constant-bank load)
  float cxx[VSL_MAX_SCALES], cxy[VSL_MAX_SCALES], cyx[VSL_MAX_SCALES], cyy[VSL_MAX_SCALES];
  int item_begin[VSL_MAX_SCALES + 1];
  int strips[VSL_MAX_SCALES], bands[VSL_MAX_SCALES];
  int lg_vec4[VSL_MAX_SCALES];     // logits / g_logits of this scale are 16-byte aligned (and V is even)
  int x_vec2[VSL_MAX_SCALES];      // x of this scale is 8-byte aligned and its width even: the tile fill moves pairs
  float wstep[VSL_MAX_SCALES], hstep[VSL_MAX_SCALES];  // meshgrid linspace steps 2/(W-1), 2/(H-1) in fp32
  // per-scale constants precomputed on the host: under register pressure the compiler re-derives loop invariants
  // inside the row loop, and re-loading one word from the constant bank is all that should cost
  int Hs[VSL_MAX_SCALES], Ws[VSL_MAX_SCALES];
  int stride4[VSL_MAX_SCALES];     // float4 per row of a zero-bordered RGBA level
  int plane4[VSL_MAX_SCALES];      // float4 per image of it
  int coff[VSL_MAX_SCALES];        // (kPad * stride4 + kPad) - magic bias * (stride4 + 1), wrapping: see tap_issue
  float Wf[VSL_MAX_SCALES], Hf[VSL_MAX_SCALES];
};

// partial sums per tile: kLossSlots loss sums (pixel, smooth, exp, consist) + per view 12 dP entries
constexpr int kLossSlots = 4;
template <int V> struct NT { static constexpr int value = kLossSlots + 12 * V; };

// One warp's slice of dynamic shared memory (floats).  What the unified L1 does not give to shared memory is the
// cache the gathers run in, and the carve-out moves in steps (..., 100, 132, 164, 196 KB): the fast-arithmetic
// kernels keep 4 blocks x (4 warps x total x 4 B + 1 KB) just inside the 100 KB step (156 KB of L1; at 164 KB the
// kernel is 1.5 us slower, at 196 KB 4 us slower).
template <int V, bool EXACT> struct WarpSmem {
  static constexpr int qt = 0;                          // [kQH][kQS]  x or 1/x with a 2-pixel halo
  static constexpr int ha = qt + kQH * kQS;             // [kOH][2]    cxx * sign(dx2) of the 2 columns left of the tile
  static constexpr int hc = ha + kOH * 2;               // [kOH][2]    cxy*sign(dxdy) + cyx*sign(dydx), same columns
  // [kRH][3] per tile row (fast arithmetic): grid row coordinate; weight of the yy second difference (0 in the last
  // two image rows, where it does not exist); 1 / 0 = the mixed difference exists / does not (last image row)
  static constexpr int rt = hc + kOH * 2;
  // [kRH][32] x itself where the tile holds 1/x (smoothness on 1/x, warp on x): the exact mode only -- the fast
  // arithmetic takes the reciprocal of the tile value again (2 ulp)
  static constexpr int xc = rt + (EXACT ? 0 : 3 * kRH);
  static constexpr int total = (xc + (EXACT ? kRH * 32 : 0) + 3) / 4 * 4;
  static constexpr size_t block_bytes = sizeof(float) * total * kWarps;
  static_assert(EXACT || VSL_FUSED_WARPS != 4 || VSL_FUSED_MIN_BLOCKS != 4 || VSL_RH != 32 ||
                    (sizeof(float) * total * 4 + 1024) * 4 <= 100 * 1024,
                "shared memory of the fast fused kernels crosses the 100 KB carve-out");
};

VSL_DEV float signed_by(float c, float v) {  // c * sign(v), sign(0) = 0
  return (v == 0.f) ? 0.f : copysignf(c, v);
}
VSL_DEV float rcp_fast(float a) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
VSL_DEV float ex2_fast(float a) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
VSL_DEV float lg2_fast(float a) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
// sign(e) in {-1, 0, 1} with two FMA-pipe instructions: sat(e * 2^100 + 0.5) is 0, 0.5 or 1
VSL_DEV float sign_fast(float e) {
  return fmaf(__saturatef(fmaf(e, 1.2676506e30f, 0.5f)), 2.0f, -1.0f);
}

// The four second differences owned by the element at q (their top-left corner), my_losses.py:27-36:
// weighted signs for the gradient (a: xx, b: yy, c: xy + yx) and the weighted |.| sum for the loss.
// gx, gy: image coordinates of the element as unsigned (a difference exists iff its support is inside).
template <bool EXACT>
VSL_DEV void owner_signs(const float* q, unsigned gx, unsigned gy, int H, int W, float cxx, float cxy, float cyx,
                         float cyy, float& a, float& b, float& c, float& sm) {
  const float q00 = q[0], q01 = q[1], q02 = q[2], q10 = q[kQS], q11 = q[kQS + 1], q20 = q[2 * kQS];
  const float dx0 = __fsub_rn(q01, q00), dy0 = __fsub_rn(q10, q00);
  float dxx = __fsub_rn(__fsub_rn(q02, q01), dx0);
  float dyy = __fsub_rn(__fsub_rn(q20, q10), dy0);
  float dxy = __fsub_rn(__fsub_rn(q11, q10), dx0);  // d/dy of dx
  // unsigned compares fold the >= 0 tests; H, W >= 3 is guaranteed by check_desc
  if (!(gy < (unsigned)H && gx < (unsigned)(W - 2))) dxx = 0.f;
  if (!(gx < (unsigned)W && gy < (unsigned)(H - 2))) dyy = 0.f;
  const bool mixed = gx < (unsigned)(W - 1) && gy < (unsigned)(H - 1);
  if (!mixed) dxy = 0.f;
  if (EXACT) {
    float dyx = __fsub_rn(__fsub_rn(q11, q01), dy0);  // d/dx of dy
    if (!mixed) dyx = 0.f;
    a = signed_by(cxx, dxx);
    b = signed_by(cyy, dyy);
    c = signed_by(cxy, dxy) + signed_by(cyx, dyx);
    sm = cxx * fabsf(dxx) + cyy * fabsf(dyy) + cxy * fabsf(dxy) + cyx * fabsf(dyx);
  } else {
    // d/dy of dx and d/dx of dy are the same number up to rounding: the fast path evaluates it once
    const float cm = cxy + cyx;
    a = cxx * sign_fast(dxx);
    b = cyy * sign_fast(dyy);
    c = cm * sign_fast(dxy);
    sm = cxx * fabsf(dxx) + cyy * fabsf(dyy) + cm * fabsf(dxy);
  }
}

// The view-paired kernel lives in its own translation unit (vsl_loss_pair.cu); V must be even.
int launch_fused_pair(int V, const LossParams& P, cudaStream_t st);
// The SSIM part of the photometric term (vsl_loss_ssim.cu): runs after the fused kernel, before the finalize.
int launch_ssim_term(const VslLossDesc* d, const LossParams& P, cudaStream_t st);

}  // namespace vsl
