// libvsl: the fused multi-scale view-synthesis loss, forward AND backward in one pass.
//
// What the reference does per training step (train.py:107-135, train_depth_then_cam_lr.py:211-328) is a
// Python loop over S scales x V source views of ~150 TF ops each, then autodiff of all of it.  Here the
// whole thing is three launches:
//
//   1. loss_prep_kernel     resize_area levels 1..S-1 of the target (RGB) and, for every source view, ALL
//                           levels 0..S-1 re-laid as zero-bordered RGBA (16 bytes per pixel, 2 pixels of
//                           zeros all round); K_s^-1 and P = K4_s . T_v per (scale, view, batch) ride along.
//   2. loss_fused_kernel    every scale and view in ONE grid.  A WARP owns a tile of 32 columns x kRH rows
//                           of one image at one scale and never talks to another warp (no block barrier):
//        it stages x (+2 halo) in its slice of shared memory, then walks down the rows, lane = column.  Per
//        pixel: the four second differences the element owns (smoothness loss) and the gradient of the 10
//        stencils it is part of (own column carried in registers, left neighbours by shuffle);
//        back-projection; per view projection, bilinear gather, L1 against the target, explainability /
//        validity mask, softmax cross-entropy regulariser -- and, because the loss is a weighted sum of means
//        whose upstream gradient is known (loss_scale), d/dx and d/dlogits are written in the same pass and
//        dP = sum du (x) [cam;1] is accumulated in registers and reduced once per tile with a shuffle
//        butterfly.  No full-resolution intermediate is ever written.  The row loop is software-pipelined:
//        the gathers and streamed operands of row r+1 are in flight while row r is computed.
//   3. loss_finalize_kernel fixed-order reduction of the tile partials (deterministic), dT = K4^T dP summed
//                           over scales, pose chain rule, the three loss scalars.
//
// Why the zero-bordered RGBA copy of the sources: the kernel is instruction-bound and the gather is its
// most expensive part.  With packed RGB a corner is three scalar loads at a 12-byte stride (12 loads per
// pixel-view) plus the sampler's mask logic; with one pixel = one aligned float4 it is 4 loads.  The
// border implements the sampler's zero padding (utils.py:266-270) in the DATA: a corner outside the image
// reads zeros, so neither the value nor d/dx, d/dy needs the four "corner == clipped corner" masks, and
// coordinates are simply clamped to [-2, size] (where everything in reach is zero) before the floor.
//
// floor + float->int without the quarter-rate conversion pipe: t = x (+, round-down) 1.5*2^23 has
// floor(x) in its low mantissa bits, so floor(x) = t - 1.5*2^23 and the integer is a bit-cast away.
#include <algorithm>
#include <type_traits>

#include "vsl_common.cuh"
#include "vsl_prep.cuh"
#include "vsl_loss_common.cuh"

namespace vsl {

// timing experiment (wrong results): VSL_EXP_TIMELINE makes every warp record %globaltimer at the stages of its
// tile; the stamps replace the tile's partial sums (profiles/run_fused.py reads them back from the workspace)
#ifdef VSL_EXP_TIMELINE
#define VSL_STAMP(k) do { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); stamps[k] = t_; } while (0)
#else
#define VSL_STAMP(k) do { } while (0)
#endif

// What one (pixel, view) has in flight between issuing its gathers and consuming them.
struct Tap {
  float4 A, B, C, D;             // corners (x0,y0) (x1,y0) (x0,y1) (x1,y1)
  float wx0, wx1, wy0, wy1;      // (x1 - x), (x - x0), (y1 - y), (y - y0)
  float qx, qy, rz;              // projected coordinates (unclamped) and 1 / (z + eps)
  float zu;                      // z itself: the source-frame depth the consistency term compares (fast arithmetic)
  int off;                       // float4 offset of corner A inside the level (kept for the d/d(source) scatter)
};

// Projection, footprint and the four 16-byte gathers of one view.  p: the 12 floats of P (warp-uniform).
// EXACT: u = P [c; 1] in the reference's rounding sequence (c = ray * depth).  Fast: the projection folded per image,
// u_i = d * (a_i + gy * qy_i) + p[4 i + 3] with Q = P[:, :3] K^-1 (XformQ): a_i = Q[i][0] gx + Q[i][2] is a lane
// constant (the column is fixed), qy_i = Q[i][1] warp-uniform -- 6 FMAs instead of 3 + 9.
template <bool EXACT>
VSL_DEV void tap_issue(Tap& t, const float (&p)[12], float c0, float c1, float c2, const float (&a)[3],
                       const float (&qy)[3], float d, float gy, const float4* __restrict__ src, int stride4, int coff,
                       float Wf, float Hf) {
  if (EXACT) {
    const Proj q = project(p, c0, c1, c2);
    t.qx = q.x; t.qy = q.y; t.rz = 1.0f / q.zp;
  } else {
    const float u0 = fmaf(d, fmaf(gy, qy[0], a[0]), p[3]);
    const float u1 = fmaf(d, fmaf(gy, qy[1], a[1]), p[7]);
    const float u2 = fmaf(d, fmaf(gy, qy[2], a[2]), p[11]);
    t.rz = rcp_fast(u2 + kEpsZ);
    t.qx = u0 * t.rz; t.qy = u1 * t.rz;
    t.zu = u2;
  }
  // beyond [-2, size] all four corners are border zeros; clamping there changes neither value nor gradient
  const float xc = fminf(fmaxf(t.qx, -2.0f), Wf), yc = fminf(fmaxf(t.qy, -2.0f), Hf);
  const float tx = __fadd_rd(xc, kMagic), ty = __fadd_rd(yc, kMagic);
  const float fx = __fsub_rn(tx, kMagic), fy = __fsub_rn(ty, kMagic);   // floor, exactly
  t.wx1 = __fsub_rn(xc, fx); t.wy1 = __fsub_rn(yc, fy);
  if (EXACT) {   // the fast path forms 1 - w at the point of use instead of carrying two more registers
    t.wx0 = __fsub_rn(__fadd_rn(fx, 1.0f), xc); t.wy0 = __fsub_rn(__fadd_rn(fy, 1.0f), yc);
  }
  const int off = (int)(__float_as_uint(ty) * (unsigned)stride4 + __float_as_uint(tx) + (unsigned)coff);  // wraps to the true offset
  t.off = off;
#ifdef VSL_EXP_L1GATHER   // timing experiment (wrong results): every gather hits the same line
  const float4* __restrict__ g = src + (off & 1);
#else
  const float4* __restrict__ g = src + off;
#endif
  t.A = __ldg(g); t.B = __ldg(g + 1);
  t.C = __ldg(g + stride4); t.D = __ldg(g + stride4 + 1);
}

// EXACT = true : coordinates, softmax and the warped value follow the reference's rounding sequence
//                (bit-identical sample positions to the oracle for matrix poses).
// EXACT = false: the same algebra with FMA contraction, MUFU reciprocal / exp / log and the closed form
//                d(depth) = -<du, t> / depth; differs from EXACT by a few ulp per quantity.
// DSRC = true additionally scatters d/d(source levels) with 16-byte reductions (red.global.add.v4.f32) into
// gradient levels of the same zero-bordered RGBA layout; loss_fold_src_grad_kernel folds them back to level 0.
// CONS = true (fast arithmetic only) adds the left-right depth-consistency term of train_depth_then_cam_lr.py:336-340
// from the SAME gather: the source view's depth map rides in the fourth channel of its RGBA levels (written by the
// prep launch), so consistent_depth_loss' bilinear fetch (utils_lr.py:369-458) is one more blend of values that are
// already in registers: |z - sampled| * mask, its gradient w.r.t. the coordinates (through the sampler) and w.r.t. z
// join d/du, the mask's share joins d/dlogits, and d/d(source depth) is scattered with 4 float reductions into a
// zero-bordered plane of the same geometry (loss_crop_src_depth_grad_kernel returns its interior).
template <int V, bool EXACT, bool DSRC, bool CONS = false>
// the exact mode (IEEE divisions, expf / logf, unfused sequences) needs ~170 registers: one block less per SM
// beats spilling
__global__ void __launch_bounds__(kThreads, (V == 1 && !EXACT && !CONS && !DSRC ? VSL_V1_MIN_BLOCKS
                                           : V <= 2 ? (EXACT || CONS ? VSL_FUSED_MIN_BLOCKS - 1 : VSL_FUSED_MIN_BLOCKS)
                                                      : VSL_FUSED_MIN_BLOCKS / 2))
loss_fused_kernel(const LossParams P) {
  static_assert(!CONS || (!EXACT && !DSRC), "the consistency term rides on the fast arithmetic");
  constexpr int N = NT<V>::value;
  using L = WarpSmem<V, EXACT>;
  extern __shared__ float4 smem4[];
  // the shuffle tells the compiler `warp` (and everything derived from it: tile, scale, sizes, base pointers) is
  // warp-uniform, so those values live in uniform registers instead of 128-per-thread vector registers
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int tile = blockIdx.x * kWarps + warp;
  if (tile >= P.item_begin[P.S]) return;   // warps are independent: no block barrier anywhere below
#ifdef VSL_EXP_TIMELINE
  unsigned long long stamps[8];
#endif
  VSL_STAMP(0);
  float* wsm = reinterpret_cast<float*>(smem4) + warp * L::total;
  float* qt = wsm + L::qt;
  float* sha = wsm + L::ha;
  float* shc = wsm + L::hc;
  float* sxc = wsm + L::xc;
  // the row table through a per-thread address: a load the compiler believes to be warp-uniform is followed by a
  // register -> uniform-register move that waits for it; as an ordinary per-lane value the row constants are operands
  const float* rtv = reinterpret_cast<const float*>(smem4) + (threadIdx.x >> 5) * L::total + L::rt;

  // ---- which tile
  int s = 0;
  while (s + 1 < P.S && tile >= P.item_begin[s + 1]) ++s;
  const int strips = P.strips[s], per_b = strips * P.bands[s];
  const int rem = tile - P.item_begin[s];
  const int b = rem / per_b, r2 = rem - b * per_b;
  const int band = r2 / strips, strip = r2 - band * strips;
  const int H = P.Hs[s], W = P.Ws[s];
  const int y_base = band * kRH, x_base = strip * 32;
  const int rows = min(kRH, H - y_base);
  const int x = x_base + lane;
  const bool act = x < W;
  const int pix0 = b * H * W;                  // first pixel of this image inside the level (32-bit: check_desc)
  const bool use_lg = P.mask_mode == VSL_MASK_EXP;

  // ---- 1. the x tile (+halo), zero outside the image, and this image's transforms
  const float* __restrict__ xs = P.x[s] + pix0;
  {
    // 4-byte cp.async with a zero source size outside the image: every element of the tile is in flight at
    // once and no register holds it on the way
    const unsigned qt_s = (unsigned)__cvta_generic_to_shared(qt);
    if (P.x_vec2[s] != 0) {
      // 8-byte pieces: the tile starts 2 columns left of a multiple of 32 and the level's width is even, so a pair
      // (2 p, 2 p + 1) of tile columns is 8-byte aligned and lies inside or outside the image as a whole
      constexpr int kPairs = kQS / 2;
      int ty = 0, tp = lane;
      while (tp >= kPairs) { tp -= kPairs; ++ty; }
#pragma unroll 4
      for (int i = lane; i < kQH * kPairs; i += 32) {
        const int gy = y_base - kHalo + ty, gx = x_base - kHalo + 2 * tp;
        const bool in = (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
        const float* src = xs + (in ? gy * W + gx : 0);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(qt_s + 8u * i), "l"(src), "r"(in ? 8 : 0)
                     : "memory");
        tp += 32 - kPairs;
        ++ty;
        if (tp >= kPairs) { tp -= kPairs; ++ty; }
      }
    } else {
      int ty = 0, tc = lane;                     // (row, column) of flat element i, advanced without division
#pragma unroll 4
      for (int i = lane; i < kQH * kQS; i += 32) {
        const int gy = y_base - kHalo + ty, gx = x_base - kHalo + tc;
        const bool in = (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
        const float* src = xs + (in ? gy * W + gx : 0);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(qt_s + 4u * i), "l"(src), "r"(in ? 4 : 0)
                     : "memory");
        tc += 32;
        if (tc >= kQS) { tc -= kQS; ++ty; }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (!EXACT)   // the tile's row table (fp32 linspace row coordinate of utils.py:153-159, smoothness row weights)
    {
      float* rt = wsm + L::rt + 3 * lane;
      rt[0] = grid_coord(y_base + lane, H, P.hstep[s]);
      rt[1] = y_base + lane < H - 2 ? P.cyy[s] : 0.f;
      rt[2] = y_base + lane < H - 1 ? 1.f : 0.f;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (P.x_is_logit != 0) {          // disparity head on load (nets_optflow_depth.py:143-144); outside stays 0
      int ty = 0, tc = lane;
      for (int i = lane; i < kQH * kQS; i += 32) {
        const int gy = y_base - kHalo + ty, gx = x_base - kHalo + tc;
        if ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) {
          const float v = qt[i];
          const float sg = EXACT ? 1.0f / (1.0f + expf(-v)) : rcp_fast(1.0f + ex2_fast(-1.4426950408889634f * v));
          qt[i] = fmaf(P.disp_scale, sg, P.disp_min);
        }
        tc += 32;
        if (tc >= kQS) { tc -= kQS; ++ty; }
      }
      __syncwarp();
    }
    if (P.smooth_on_inverse != 0) {   // the smoothness term lives on 1/x (train_depth_then_cam_lr.py:217)
      if (EXACT && !P.depth_is_inverse) {   // ... while the warp wants x itself: keep the tile's centre rows
        for (int r = 0; r < kRH; ++r) sxc[r * 32 + lane] = qt[(r + kHalo) * kQS + lane + kHalo];
      }
      for (int i = lane; i < kQH * kQS; i += 32) {
        const float v = qt[i];
        qt[i] = v != 0.f ? __fdiv_rn(1.0f, v) : 0.f;   // outside the image the tile stays 0
      }
    }
  }
  __syncwarp();

  float pix_sum = 0.f, exp_sum = 0.f, sm_sum = 0.f, con_sum = 0.f;
  const float ccon = CONS ? P.ccon[s] : 0.f;
  float se_prod[V];                 // fast arithmetic: running product of the softmax denominators, per view
#pragma unroll
  for (int v = 0; v < V; ++v) se_prod[v] = 1.f;

  VSL_STAMP(1);
  // ---- 2. smoothness of the two columns LEFT of the tile: the weighted signs of their xx and xy/yx second
  // differences, which the gradient of columns 0 and 1 needs.  Everything else of the smoothness term is
  // evaluated in the row loop, lane = column: own column in registers, the two to the left by shuffle.
  const float cxx = P.cxx[s], cxy = P.cxy[s], cyx = P.cyx[s], cyy = P.cyy[s];
  for (int i = lane; i < kOH * 2; i += 32) {
    const int oy = i >> 1, ox = i & 1;
    float a, bb, c, sm;
    owner_signs<EXACT>(qt + oy * kQS + ox, (unsigned)(x_base - kHalo + ox), (unsigned)(y_base - kHalo + oy), H, W, cxx,
                       cxy, cyx, cyy, a, bb, c, sm);
    sha[i] = a;
    shc[i] = c;
  }
  __syncwarp();

  // ---- 3. the pixels: lane = column, walking down the rows of the tile.
  // Software pipeline: the streamed operands (target, logits, mask) of row r+1 are requested at the top of
  // row r, and the gathers of a view are re-issued for the NEXT row (or the view two ahead) as soon as the
  // current ones have been consumed, so every global load has about half a row of arithmetic to land.
  // Lanes beyond the image edge (ragged last strip) compute on the last column and contribute nothing.
  const int xl = min(x, W - 1);
  const float gx = grid_coord(xl, W, P.wstep[s]);
  const float hstep = P.hstep[s];
  const float Wf = P.Wf[s], Hf = P.Hf[s];
  const int stride4 = P.stride4[s];
  // corner offset inside the level = b * plane + (iy + kPad) * stride4 + (ix + kPad), with iy, ix still carrying
  // the magic bias (P.coff removes it, wrapping)
  const int coff = P.coff[s] + b * P.plane4[s];
  const float cpix = P.cpix[s], cexp = P.cexp[s];
  // fast arithmetic: image-border conditions of the smoothness term are data, not predicates -- a lane's weights are
  // zero where the second difference its column owns does not exist (the row conditions come from the row table)
  const float kxx_l = x < W - 2 ? cxx : 0.f;
  const float kxy_l = x < W - 1 ? cxy + cyx : 0.f;

  // Everything above read only the caller's inputs.  The launch is programmatically dependent on the prep launch
  // (it may start while that one drains): wait here, before the first read of what prep wrote.
  VSL_STAMP(2);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  VSL_STAMP(3);

  // This image's transforms.  Every lane reads the same words; the shuffle marks them warp-uniform, so K^-1 and
  // the V projection matrices sit in uniform registers and enter the FMAs as operands -- no shared memory, no
  // per-row reloads, no per-thread copies.
  float kinv[9], Pm[V][12];
  {
    const float* xf0 = reinterpret_cast<const float*>(P.xf + ((size_t)s * V) * P.B + b);
#pragma unroll
    for (int i = 0; i < 9; ++i) kinv[i] = __shfl_sync(0xffffffffu, __ldg(xf0 + i), 0);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float* xfv = reinterpret_cast<const float*>(P.xf + ((size_t)s * V + v) * P.B + b) + 9;
#pragma unroll
      for (int i = 0; i < 12; ++i) Pm[v][i] = __shfl_sync(0xffffffffu, __ldg(xfv + i), 0);
    }
  }
  // K^-1: the column that multiplies gx is folded per thread
  const float3 k0 = make_float3(kinv[0], kinv[1], kinv[2]), k1 = make_float3(kinv[3], kinv[4], kinv[5]),
               k2 = make_float3(kinv[6], kinv[7], kinv[8]);
  float kx0 = 0.f, kx1 = 0.f, kx2 = 0.f;
  if (EXACT) {
    kx0 = __fmul_rn(k0.x, gx); kx1 = __fmul_rn(k1.x, gx); kx2 = __fmul_rn(k2.x, gx);
  }
  // fast arithmetic: the folded projection of tap_issue -- per view Q[i][1] (uniform) and Q[i][0] gx + Q[i][2] (lane)
  float Qy[V][3], Al[V][3];
#pragma unroll
  for (int v = 0; v < V; ++v)
#pragma unroll
    for (int i = 0; i < 3; ++i) { Qy[v][i] = 0.f; Al[v][i] = 0.f; }
  if (!EXACT) {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float* xq = reinterpret_cast<const float*>(P.xq + ((size_t)s * V + v) * P.B + b);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float q0 = __shfl_sync(0xffffffffu, __ldg(xq + 3 * i), 0), q2 = __shfl_sync(0xffffffffu, __ldg(xq + 3 * i + 2), 0);
        Qy[v][i] = __shfl_sync(0xffffffffu, __ldg(xq + 3 * i + 1), 0);
        Al[v][i] = fmaf(q0, gx, q2);
      }
    }
  }

  float S2[V][3], S3[V][3], S4[V][3];  // sum du*d*gy, du*d, du   (sum du*d*gx = gx * S3: the column is fixed)
#pragma unroll
  for (int v = 0; v < V; ++v)
#pragma unroll
    for (int i = 0; i < 3; ++i) { S2[v][i] = 0.f; S3[v][i] = 0.f; S4[v][i] = 0.f; }

  // vertical neighbours of the smoothness gradient travel down in registers: the owners of the two rows above
  // the tile come first
  float b1, b2, c10, c11;
  {
    float a, bb, c, sm;
    owner_signs<EXACT>(qt + lane + kHalo, (unsigned)x, (unsigned)(y_base - 2), H, W, cxx, cxy, cyx, cyy, a, bb, c, sm);
    b2 = bb;
    owner_signs<EXACT>(qt + kQS + lane + kHalo, (unsigned)x, (unsigned)(y_base - 1), H, W, cxx, cxy, cyx, cyy, a, bb, c,
                       sm);
    b1 = bb;
    c10 = c;
    c11 = __shfl_up_sync(0xffffffffu, c, 1);
    if (lane == 0) c11 = shc[1 * 2 + 1];
  }

  // level base pointers (whole level, not this image: the image offset rides in the 32-bit pixel index)
  const float* __restrict__ tgt_img = P.tgt[s];
  const float* __restrict__ lg_img = use_lg ? P.logits[s] : nullptr;
  float* __restrict__ glg_img = use_lg ? P.g_logits[s] : nullptr;
  const float* __restrict__ mk_img = P.mask_mode == VSL_MASK_CONST ? P.mask[s] : nullptr;
  float* __restrict__ gx_img = P.g_x[s];
  const bool lg4 = (V % 2 == 0) && P.lg_vec4[s] != 0;
  const int smooth_inv = P.smooth_on_inverse, depth_inv = P.depth_is_inverse;

  struct Stream { float tt[3]; float lg[2 * V]; float mc; };               // streamed operands of one row
  struct Geo { float d, gy, dgy, r0, r1, r2, c0, c1, c2; };                 // per-pixel geometry of one row

  auto load_stream = [&](Stream& st, int pofs) {
#pragma unroll
    for (int c = 0; c < 3; ++c) st.tt[c] = __ldcs(tgt_img + pofs * 3 + c);
    if (use_lg) {
      if (lg4) {
#pragma unroll
        for (int k = 0; k < V / 2; ++k) {
          const float4 q = __ldcs(reinterpret_cast<const float4*>(lg_img + (size_t)pofs * (2 * V)) + k);
          st.lg[4 * k] = q.x; st.lg[4 * k + 1] = q.y; st.lg[4 * k + 2] = q.z; st.lg[4 * k + 3] = q.w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          const float2 q = __ldcs(reinterpret_cast<const float2*>(lg_img + (size_t)pofs * (2 * V)) + k);
          st.lg[2 * k] = q.x; st.lg[2 * k + 1] = q.y;
        }
      }
    }
    st.mc = 1.f;
    if (mk_img != nullptr) st.mc = __ldcs(mk_img + pofs);
  };
  auto make_geo = [&](Geo& g, int r) {
    const float qc = qt[(r + kHalo) * kQS + (xl - x_base) + kHalo];
    if (smooth_inv) g.d = depth_inv ? qc : (EXACT ? sxc[r * 32 + (xl - x_base)] : rcp_fast(qc));
    else g.d = !depth_inv ? qc : (EXACT ? __fdiv_rn(1.0f, qc) : rcp_fast(qc));
    if (EXACT) {  // pixel2cam's matmul (utils.py:114): sequential k, no contraction
      const float gy = grid_coord(y_base + r, H, hstep);
      g.r0 = __fadd_rn(__fadd_rn(kx0, __fmul_rn(k0.y, gy)), k0.z);
      g.r1 = __fadd_rn(__fadd_rn(kx1, __fmul_rn(k1.y, gy)), k1.z);
      g.r2 = __fadd_rn(__fadd_rn(kx2, __fmul_rn(k2.y, gy)), k2.z);
      g.c0 = __fmul_rn(g.r0, g.d); g.c1 = __fmul_rn(g.r1, g.d); g.c2 = __fmul_rn(g.r2, g.d);
      g.dgy = g.d * gy;
    } else {      // the ray is folded into the per-view projection (tap_issue); the row coordinate comes from the table
      g.gy = rtv[3 * r];
      g.dgy = g.d * g.gy;
    }
  };

  // One row = two phases.  Phase 1 holds every long-latency wait and issue: blend view v from its landed
  // gathers (reducing the 16 gathered floats to 5), re-issue that slot's gathers for the NEXT row at once, then
  // request the next row's streamed operands.  Phase 2 (softmax, gradients, accumulation, stores: over half
  // of the row's arithmetic) touches nothing in flight -- so whichever hardware scoreboards the compiler lets
  // these load groups share, no wait ever lands on a load that was only just issued.
  // what phase 2 needs of a view: sum|e| and dL/du up to the factor cpix * m (CONS: up to the factor m; Cz = |z - sampled
  // source depth|, gz = ccon * sign(z - sampled) * z, the <du, u> the closed form of d/d(depth) otherwise lacks)
  struct Keep { float E, u0, u1, u2, Cz, gz; };
  Tap tap[V];
  int pofs = pix0 + y_base * W + xl;            // pixel index inside the level

  auto row = [&](Stream& cur, Stream& nxt, Geo& gc, Geo& gn, int r) {
    const bool has_next = r + 1 < rows;
    if (has_next) make_geo(gn, r + 1);
    Keep keep[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      Tap& t = tap[v];
      // The padding channel of the four gathers carries zeros nobody needs.  Left to itself the register
      // allocator hands those registers to the very next instructions after the loads are issued, and the
      // write-after-write hazard then stalls the warp for the full memory latency.  OR-ing them into the
      // running |e| sum (a no-op on the value) keeps them reserved until the data is consumed: 2 LOP3 per view.
      const unsigned pad = CONS ? 0u : (__float_as_uint(t.A.w) | __float_as_uint(t.B.w) | __float_as_uint(t.C.w) |
                                        __float_as_uint(t.D.w));   // CONS: the fourth channel is data and is consumed
      const float cA[3] = {t.A.x, t.A.y, t.A.z}, cB[3] = {t.B.x, t.B.y, t.B.z},
                  cC[3] = {t.C.x, t.C.y, t.C.z}, cD[3] = {t.D.x, t.D.y, t.D.z};
      // E = sum_c |e_c|; J_k = sum_c sign(e_c) * corner_k[c]  (the channel sum commutes with d/dx, d/dy)
      // EXACT: the reference's four products.  Fast: one product, w11 = wx1 wy1, then w10 = wx1 - w11,
      // w01 = wy1 - w11, w00 = (1 - wx1) - w01
      constexpr bool kProd = EXACT;
      const float wx0 = EXACT ? t.wx0 : 1.0f - t.wx1, wy0 = EXACT ? t.wy0 : (kProd ? 1.0f - t.wy1 : 0.f);
      const float w11 = __fmul_rn(t.wx1, t.wy1);
      const float w10 = kProd ? __fmul_rn(t.wx1, wy0) : t.wx1 - w11, w01 = kProd ? __fmul_rn(wx0, t.wy1) : t.wy1 - w11;
      const float w00 = kProd ? __fmul_rn(wx0, wy0) : wx0 - w01;
      float E = __uint_as_float(pad), JA = 0.f, JB = 0.f, JC = 0.f, JD = 0.f;
      float sgc[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float wv = EXACT ? blend(w00, w01, w10, w11, cA[c], cC[c], cB[c], cD[c])
                               : fmaf(w11, cD[c], fmaf(w10, cB[c], fmaf(w01, cC[c], w00 * cA[c])));
        const float e = wv - cur.tt[c];
        E += fabsf(e);
        const float sg = EXACT ? signed_by(1.f, e) : sign_fast(e);
        JA = fmaf(sg, cA[c], JA); JB = fmaf(sg, cB[c], JB);
        JC = fmaf(sg, cC[c], JC); JD = fmaf(sg, cD[c], JD);
        if (DSRC) sgc[c] = sg;
      }
      if (DSRC) {
        // d/d(corner_k[c]) = cpix * m * sign(e_c) * w_k; a corner in the border lands in the border of the
        // gradient level, which nobody reads
        float m = cur.mc;
        if (use_lg) {
          const float l0 = cur.lg[2 * v], l1 = cur.lg[2 * v + 1];
          if (EXACT) {
            const float mx = fmaxf(l0, l1);
            const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
            m = e1 / (e0 + e1);
          } else {
            m = rcp_fast(1.f + ex2_fast(1.4426950408889634f * (l0 - l1)));   // softmax(l)[1]
          }
        }
        if (act) {
          float4* g = P.gsrc[v][s] + t.off;
          const float km = cpix * m;
          const float k00 = km * w00, k10 = km * w10, k01 = km * w01, k11 = km * w11;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g), "f"(k00 * sgc[0]), "f"(k00 * sgc[1]),
                       "f"(k00 * sgc[2]), "f"(0.f) : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g + 1), "f"(k10 * sgc[0]), "f"(k10 * sgc[1]),
                       "f"(k10 * sgc[2]), "f"(0.f) : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g + stride4), "f"(k01 * sgc[0]),
                       "f"(k01 * sgc[1]), "f"(k01 * sgc[2]), "f"(0.f) : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g + stride4 + 1), "f"(k11 * sgc[0]),
                       "f"(k11 * sgc[1]), "f"(k11 * sgc[2]), "f"(0.f) : "memory");
        }
      }
      // d/dx = wy0 (JB - JA) + wy1 (JD - JC), d/dy = wx0 (JC - JA) + wx1 (JD - JB); the fast path in lerp form
      float dx, dy;
      if (kProd) {
        dx = wy0 * (JB - JA) + t.wy1 * (JD - JC);
        dy = wx0 * (JC - JA) + t.wx1 * (JD - JB);
      } else {
        const float ax = JB - JA, ay = JC - JA;
        dx = fmaf(t.wy1, (JD - JC) - ax, ax);
        dy = fmaf(t.wx1, (JD - JB) - ay, ay);
      }
      if (CONS) {
        // consistent_depth_loss: sampled = bilinear(source depth) at the same corners, err = |z - sampled|
        const float sw = fmaf(w11, t.D.w, fmaf(w10, t.B.w, fmaf(w01, t.C.w, w00 * t.A.w)));
        const float ec = t.zu - sw;
        const float sc = sign_fast(ec) * ccon;                       // ccon * sign(z - sampled)
        const float axw = t.B.w - t.A.w, ayw = t.C.w - t.A.w;
        const float dxw = fmaf(t.wy1, (t.D.w - t.C.w) - axw, axw), dyw = fmaf(t.wx1, (t.D.w - t.B.w) - ayw, ayw);
        dx = fmaf(cpix, dx, -sc * dxw);                              // d/dx of (cpix E + ccon err) / m
        dy = fmaf(cpix, dy, -sc * dyw);
        keep[v].Cz = fabsf(ec);
        keep[v].gz = sc * t.zu;
        if (P.gsd[v][s] != nullptr && act) {
          // d/d(source depth corner k) = -ccon m sign w_k; the mask value is needed here already
          float m = cur.mc;
          if (use_lg) m = rcp_fast(1.f + ex2_fast(1.4426950408889634f * (cur.lg[2 * v] - cur.lg[2 * v + 1])));
          const float km = -sc * m;
          float* g = P.gsd[v][s] + t.off;
          asm volatile("red.global.add.f32 [%0], %1;" ::"l"(g), "f"(km * w00) : "memory");
          asm volatile("red.global.add.f32 [%0], %1;" ::"l"(g + 1), "f"(km * w10) : "memory");
          asm volatile("red.global.add.f32 [%0], %1;" ::"l"(g + stride4), "f"(km * w01) : "memory");
          asm volatile("red.global.add.f32 [%0], %1;" ::"l"(g + stride4 + 1), "f"(km * w11) : "memory");
        }
        keep[v].E = E; keep[v].u0 = dx * t.rz; keep[v].u1 = dy * t.rz;
        keep[v].u2 = sc - (t.qx * keep[v].u0 + t.qy * keep[v].u1);    // + d/dz of the consistency term
      } else {
        // the border zeros make the sampler's corner masks implicit: an outside corner contributes 0
        keep[v].E = E; keep[v].u0 = dx * t.rz; keep[v].u1 = dy * t.rz;
        keep[v].u2 = -(t.qx * keep[v].u0 + t.qy * keep[v].u1);
      }
      if (has_next)
        tap_issue<EXACT>(t, Pm[v], gn.c0, gn.c1, gn.c2, Al[v], Qy[v], gn.d, gn.gy, P.src[v][s], stride4, coff, Wf, Hf);
    }
#ifdef VSL_EXP_NOSTREAM   // timing experiment (wrong results): what do the streamed loads cost?
    nxt = cur;
#else
    if (has_next) load_stream(nxt, pofs + W);
#endif
#ifdef VSL_EXP_PREFETCH
    {
      const float* ta = tgt_img + (pofs + 3 * W) * 3;
      const float* la = use_lg ? lg_img + (size_t)(pofs + 3 * W) * (2 * V) : ta;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %2, %3;\n\t@p prefetch.global.L2 [%0];\n\t@p prefetch.global.L2 [%1];\n\t}" ::"l"(ta),
                   "l"(la), "r"(r + 3), "r"(rows));
    }
#endif

    // ---- phase 2
    // smoothness: the four differences this element owns, then the gradient = the weighted signs of the 10
    // stencils it is part of (own column from registers, the two columns to the left by shuffle)
    float g_q;
    {
      float a0, b0, c00;
      if (EXACT) {
        float sm;
        owner_signs<EXACT>(qt + (r + kHalo) * kQS + lane + kHalo, (unsigned)x, (unsigned)(y_base + r), H, W, cxx, cxy, cyx,
                           cyy, a0, b0, c00, sm);
        sm_sum += sm;
      } else {
        // the same differences with the border conditions as weights (lane: kxx_l, kxy_l; row: table entries y, z):
        // no compares, no selects; lanes past the image edge are dropped from the sum at the end
        const float* q = qt + (r + kHalo) * kQS + lane + kHalo;
        const float q00 = q[0], q01 = q[1], q02 = q[2], q10 = q[kQS], q11 = q[kQS + 1], q20 = q[2 * kQS];
        const float dx0 = q01 - q00, dy0 = q10 - q00;
        const float dxx = (q02 - q01) - dx0, dyy = (q20 - q10) - dy0, dxy = (q11 - q10) - dx0;
        const float kyy_r = rtv[3 * r + 1], kxy_r = kxy_l * rtv[3 * r + 2];
        a0 = kxx_l * sign_fast(dxx); b0 = kyy_r * sign_fast(dyy); c00 = kxy_r * sign_fast(dxy);
        sm_sum = fmaf(kxx_l, fabsf(dxx), fmaf(kyy_r, fabsf(dyy), fmaf(kxy_r, fabsf(dxy), sm_sum)));
      }
      float a1 = __shfl_up_sync(0xffffffffu, a0, 1), a2 = __shfl_up_sync(0xffffffffu, a0, 2);
      float c01 = __shfl_up_sync(0xffffffffu, c00, 1);
      {  // lanes 0 and 1 take what lies left of the tile from the halo columns: broadcast loads + selects, no branch
        const int o = (r + kHalo) * 2;
        const float h0 = sha[o], h1 = sha[o + 1], hc1 = shc[o + 1];
        a1 = lane == 0 ? h1 : a1;
        a2 = lane == 0 ? h0 : (lane == 1 ? h1 : a2);
        c01 = lane == 0 ? hc1 : c01;
      }
      g_q = (a0 - 2.f * a1 + a2) + (b0 - 2.f * b1 + b2) + (c00 - c01 - c10 + c11);
      b2 = b1; b1 = b0; c10 = c00; c11 = c01;
    }

    // mask values m (explainability softmax or constant), the regulariser and d/dlogits -- all views inside ONE
    // uniform branch, so their dependent MUFU chains interleave instead of running one after the other
    float mv[V];
    if (use_lg) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float l0 = cur.lg[2 * v], l1 = cur.lg[2 * v + 1];
        float p0, p1;
        if (EXACT) {
          const float mx = fmaxf(l0, l1);
          const float e0 = expf(l0 - mx), e1 = expf(l1 - mx), se = e0 + e1;
          p0 = e0 / se; p1 = e1 / se;
          exp_sum += (mx + logf(se)) - l1;
        } else {
          const float z = l0 - l1;
          const float e = ex2_fast(-1.4426950408889634f * fabsf(z)), se = 1.f + e, big = rcp_fast(se), small = e * big;
          p0 = z >= 0.f ? big : small;
          p1 = z >= 0.f ? small : big;
          // log(1 + e^-|z|) + max(z, 0); the logarithms of a tile are taken once, of the product of its 1 + e^-|z|
          // (at most 2^kRH per view: no overflow; 1e-7 relative on the sum)
          se_prod[v] *= se;
          exp_sum += fmaxf(z, 0.f);
        }
        mv[v] = p1;
        // d/dm of the mask-weighted terms: cpix E (+ ccon |z - sampled|)
        const float g0 = p0 * (cexp - (CONS ? fmaf(ccon, keep[v].Cz, cpix * keep[v].E) : cpix * keep[v].E) * p1);
        cur.lg[2 * v] = g0; cur.lg[2 * v + 1] = -g0;
      }
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) mv[v] = cur.mc;
    }
    float g_d = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float m = mv[v];
      pix_sum = fmaf(m, keep[v].E, pix_sum);
      if (CONS) con_sum = fmaf(m, keep[v].Cz, con_sum);
      const float k = CONS ? m : cpix * m;         // CONS: cpix / ccon already sit in u0..u2
      const float du0 = keep[v].u0 * k, du1 = keep[v].u1 * k, du2 = keep[v].u2 * k;
      const float* pp = Pm[v];
      if (EXACT) {
        const float gc0 = du0 * pp[0] + du1 * pp[4] + du2 * pp[8];
        const float gc1 = du0 * pp[1] + du1 * pp[5] + du2 * pp[9];
        const float gc2 = du0 * pp[2] + du1 * pp[6] + du2 * pp[10];
        g_d += gc0 * gc.r0 + gc1 * gc.r1 + gc2 * gc.r2;
      } else {
        g_d -= du0 * pp[3] + du1 * pp[7] + du2 * pp[11];  // <du, M ray> = <du, u - t> / d and <du, u> = 0 ...
        if (CONS) g_d = fmaf(m, keep[v].gz, g_d);         // ... but for d/dz of the consistency term: <du, u> = g_z z
      }
      S2[v][0] = fmaf(du0, gc.dgy, S2[v][0]); S2[v][1] = fmaf(du1, gc.dgy, S2[v][1]); S2[v][2] = fmaf(du2, gc.dgy, S2[v][2]);
      S3[v][0] = fmaf(du0, gc.d, S3[v][0]);   S3[v][1] = fmaf(du1, gc.d, S3[v][1]);   S3[v][2] = fmaf(du2, gc.d, S3[v][2]);
      S4[v][0] += du0;                        S4[v][1] += du1;                        S4[v][2] += du2;
    }
    if (!EXACT) {
      // d = 1 / x (and the tile holds x): 1 / d is the tile value itself, no reciprocal
      const float inv_d = (depth_inv && !smooth_inv) ? qt[(r + kHalo) * kQS + (xl - x_base) + kHalo] : rcp_fast(gc.d);
      g_d *= inv_d;
    }
    if (act) {
      // chain rules of depth = x or 1/x and of the smoothed quantity q = x or 1/x
      float dd_dx = 1.f, dq_dx = 1.f;
      if (smooth_inv) {
        const float qc = qt[(r + kHalo) * kQS + (xl - x_base) + kHalo];
        dq_dx = -qc * qc;
        if (depth_inv) dd_dx = dq_dx;
      } else if (depth_inv) {
        dd_dx = -gc.d * gc.d;
      }
      float g_out = g_d * dd_dx + g_q * dq_dx;
      if (P.x_is_logit != 0) {          // d(disp)/dx = scale * s (1 - s), s recovered from disp itself
        float disp;
        if (smooth_inv) disp = depth_inv ? rcp_fast(qt[(r + kHalo) * kQS + (xl - x_base) + kHalo]) : gc.d;
        else disp = qt[(r + kHalo) * kQS + (xl - x_base) + kHalo];
        const float u = disp - P.disp_min;
        g_out *= u * (P.disp_scale - u) * rcp_fast(P.disp_scale);
      }
      __stcs(gx_img + pofs, g_out);
      if (use_lg) {
        if (lg4) {
#pragma unroll
          for (int k = 0; k < V / 2; ++k)
            __stcs(reinterpret_cast<float4*>(glg_img + (size_t)pofs * (2 * V)) + k,
                   make_float4(cur.lg[4 * k], cur.lg[4 * k + 1], cur.lg[4 * k + 2], cur.lg[4 * k + 3]));
        } else {
#pragma unroll
          for (int k = 0; k < V; ++k)
            __stcs(reinterpret_cast<float2*>(glg_img + (size_t)pofs * (2 * V)) + k,
                   make_float2(cur.lg[2 * k], cur.lg[2 * k + 1]));
        }
      }
    }
    pofs += W;
  };

  {
    Stream st0, st1;                            // ping-pong: no register copies between rows
    Geo g0, g1;
    VSL_STAMP(4);
    load_stream(st0, pofs);
    make_geo(g0, 0);
#pragma unroll
    for (int v = 0; v < V; ++v)
      tap_issue<EXACT>(tap[v], Pm[v], g0.c0, g0.c1, g0.c2, Al[v], Qy[v], g0.d, g0.gy, P.src[v][s], stride4, coff, Wf, Hf);
    VSL_STAMP(5);
#ifdef VSL_EXP_ROWS   // timing experiment (wrong results): only the first VSL_EXP_ROWS rows of every tile
    const int rows_run = min(rows, VSL_EXP_ROWS);
#else
    const int rows_run = rows;
#endif
    for (int r = 0; r < rows_run; r += 2) {
      row(st0, st1, g0, g1, r);
      if (r + 1 < rows) row(st1, st0, g1, g0, r + 1);
    }
  }

  VSL_STAMP(6);
  // ---- 4. one warp reduction per tile: 3 loss sums + per view (gx sum du d, sum du d gy, sum du d, sum du)
  float vals[N];
  if (!EXACT) {
#pragma unroll
    for (int v = 0; v < V; ++v) exp_sum = fmaf(lg2_fast(se_prod[v]), 0.6931471805599453f, exp_sum);
  }
  vals[0] = pix_sum * cpix; vals[1] = sm_sum; vals[2] = exp_sum * cexp;
  vals[3] = CONS ? con_sum * ccon : 0.f;
  if (!act) { vals[0] = 0.f; vals[1] = 0.f; vals[2] = 0.f; vals[3] = 0.f; }   // lanes past the image edge recomputed the last column
#pragma unroll
  for (int v = 0; v < V; ++v)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      vals[kLossSlots + v * 12 + i] = gx * S3[v][i];
      vals[kLossSlots + v * 12 + 3 + i] = S2[v][i];
      vals[kLossSlots + v * 12 + 6 + i] = S3[v][i];
      vals[kLossSlots + v * 12 + 9 + i] = S4[v][i];
      if (!act) { vals[kLossSlots + v * 12 + i] = 0.f; vals[kLossSlots + v * 12 + 3 + i] = 0.f; vals[kLossSlots + v * 12 + 6 + i] = 0.f; vals[kLossSlots + v * 12 + 9 + i] = 0.f; }
    }
  using Z = BflySizes<N>;
  bfly_step<N, 16>(vals, lane);
  bfly_step<Z::h1, 8>(vals, lane);
  bfly_step<Z::h2, 4>(vals, lane);
  bfly_step<Z::h3, 2>(vals, lane);
  bfly_step<Z::h4, 1>(vals, lane);
  float* __restrict__ out = P.partials + (size_t)tile * N;
#pragma unroll
  for (int j = 0; j < Z::h5; ++j) {
    const int idx = bfly_index<N>(lane, j);
    if (idx >= 0) out[idx] = vals[j];
  }
#ifdef VSL_EXP_TIMELINE
  VSL_STAMP(7);
  __syncwarp();
  if (lane == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    for (int k = 0; k < 8; ++k) {
      out[2 * k] = __uint_as_float((unsigned)(stamps[k] & 0xffffffffull));
      out[2 * k + 1] = __uint_as_float((unsigned)(stamps[k] >> 32));
    }
    out[16] = __uint_as_float(smid);
    out[17] = __uint_as_float((unsigned)s);
  }
#endif
}

// grid = B + 1 blocks of T threads (256; 1024 -- 512 from three views up, for the shared-memory table -- when an image
// has more than kFinBigRows tile rows: every pass over the rows is a dependent round trip, and at 480 x 640 the
// 256-thread form makes five of them per image and ten in the loss block).  Block b < B: pose gradients of batch element b (all views).  The tile partials
// of (image, scale) are contiguous rows of N floats: warp w adds rows w, w + 8, ... with lane = column (one
// coalesced request per row, every request independent), the 8 warp sums are combined in warp order.
// Block B: the three loss scalars.  Every sum runs in a fixed order in double => deterministic.
// Launched programmatically dependent on the fused kernel: what does not depend on it (K_s^-1 from the caller's
// intrinsics) is done before griddepcontrol.wait, under the fused kernel's tail.
constexpr int kFinBigRows = 256;
template <int V, int kFinThreads>
__global__ void __launch_bounds__(kFinThreads)
loss_finalize_kernel(const LossParams P, const float* __restrict__ poses, const float* __restrict__ K_pyr,
                     int pose_format, float inv_loss_scale, float* __restrict__ losses, float* __restrict__ g_poses) {
  constexpr int N = NT<V>::value, NW = kFinThreads / 32;
  __shared__ double part[NW][VSL_MAX_SCALES][N];
  __shared__ double tsum[VSL_MAX_SCALES][N];
  __shared__ float sK[VSL_MAX_SCALES][9], sKinv[VSL_MAX_SCALES][9];
  __shared__ double sgT[VSL_MAX_VIEWS][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_items = P.item_begin[P.S];
  const int b = blockIdx.x;
  if (b < P.B && (int)threadIdx.x < P.S) {   // same LU inverse as the transform table of the prep launch
    const int sc = threadIdx.x;
    float K[9], Ki[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) K[i] = K_pyr[((size_t)b * P.S + sc) * 9 + i];
    inv3_lu(K, Ki);
#pragma unroll
    for (int i = 0; i < 9; ++i) { sK[sc][i] = K[i]; sKinv[sc][i] = Ki[i]; }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");   // from here on: the fused launch's partials

  if (b == P.B) {
    // losses: columns 0..3 of every tile row (pixel, smooth, exp, consist), thread-strided, then a fixed tree
    double acc[kLossSlots] = {0.0, 0.0, 0.0, 0.0};
    // rows in flight per thread: the loads of a batch are independent; a row's four loss slots are ONE 16-byte load
    // (a row is 16 (1 + 3 V) bytes), issued unconditionally from a clamped row -- a predicated load per row would cap
    // the batch at the seven predicate registers -- and dropped, if past the end, when it is added
    constexpr int U = 10;
    static_assert(kLossSlots == 4, "one float4 per tile row");
    for (int i0 = threadIdx.x; i0 < n_items; i0 += kFinThreads * U) {
      float4 q[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        q[u] = *reinterpret_cast<const float4*>(P.partials + (size_t)min(i0 + u * kFinThreads, n_items - 1) * N);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool in = i0 + u * kFinThreads < n_items;
        acc[0] += in ? (double)q[u].x : 0.0; acc[1] += in ? (double)q[u].y : 0.0;
        acc[2] += in ? (double)q[u].z : 0.0; acc[3] += in ? (double)q[u].w : 0.0;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < kLossSlots; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (lane == 0)
#pragma unroll
      for (int k = 0; k < kLossSlots; ++k) part[warp][0][k] = acc[k];
    __syncthreads();
    if (threadIdx.x < kLossSlots) {
      double t = 0.0;
      for (int w = 0; w < NW; ++w) t += part[w][0][threadIdx.x];
      losses[threadIdx.x] = (float)(t * (double)inv_loss_scale);   // [0..3] = pixel, smooth, exp, consist
      tsum[0][threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0)                                           // [4] = their sum
      losses[4] = (float)((tsum[0][0] + tsum[0][1] + tsum[0][2] + tsum[0][3]) * (double)inv_loss_scale);
    return;
  }

  // the tile rows of this image at every scale as ONE list (row k of scale sc sits at item_begin[sc] + b * per_b + k):
  // warp w takes rows w, w + NW, ... of the list, lane = column, and has all of its rows in flight at once -- one
  // round trip to L2 for the whole image instead of one per scale
  {
    int first[VSL_MAX_SCALES + 1];
    first[0] = 0;
#pragma unroll
    for (int sc = 0; sc < VSL_MAX_SCALES; ++sc) first[sc + 1] = first[sc] + (sc < P.S ? P.bands[sc] * P.strips[sc] : 0);
    const int total = first[VSL_MAX_SCALES];
    constexpr int U = 12;                       // rows in flight per lane
    for (int col = lane; col < N; col += 32) {  // N = 4 + 12 V columns: a second pass for three and four views
      double acc[VSL_MAX_SCALES];
#pragma unroll
      for (int sc = 0; sc < VSL_MAX_SCALES; ++sc) acc[sc] = 0.0;
      for (int k0 = warp; k0 < total; k0 += NW * U) {
        float q[U];
        int qs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int k = k0 + u * NW;
          int sc = 0, f0 = 0, pb = first[1];
#pragma unroll
          for (int t = 1; t < VSL_MAX_SCALES; ++t)
            if (t < P.S && k >= first[t]) { sc = t; f0 = first[t]; pb = first[t + 1] - first[t]; }
          qs[u] = k < total ? sc : -1;
          const int item = P.item_begin[sc] + b * pb + (k - f0);
          q[u] = k < total ? P.partials[(size_t)item * N + col] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int sc = 0; sc < VSL_MAX_SCALES; ++sc) acc[sc] += qs[u] == sc ? (double)q[u] : 0.0;
      }
#pragma unroll
      for (int sc = 0; sc < VSL_MAX_SCALES; ++sc)
        if (sc < P.S) part[warp][sc][col] = acc[sc];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < P.S * N; e += kFinThreads) {
    const int sc = e / N, c = e - sc * N;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) t += part[w][sc][c];
    tsum[sc][c] = t;
  }
  __syncthreads();
  // dT[v][k][j] = sum_s sum_i K_s[i][k] * dP_s[i][j]: one thread per matrix element
  if ((int)threadIdx.x < V * 16) {
    const int v = threadIdx.x >> 4, k = (threadIdx.x >> 2) & 3, j = threadIdx.x & 3;
    double acc = 0.0;
    if (k < 3) {
      for (int sc = 0; sc < P.S; ++sc) {
        const double* t = &tsum[sc][kLossSlots + v * 12];
        const float* ki = sKinv[sc];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double dPij = (j < 3) ? (double)ki[j * 3] * t[i] + (double)ki[j * 3 + 1] * t[3 + i] +
                                            (double)ki[j * 3 + 2] * t[6 + i]
                                      : t[9 + i];
          acc += (double)sK[sc][i * 3 + k] * dPij;
        }
      }
    }
    sgT[v][k * 4 + j] = acc;
  }
  __syncthreads();
  if ((int)threadIdx.x >= V) return;
  const int v = threadIdx.x;
  const int psz = (pose_format == VSL_POSE_MATRIX) ? 16 : 6;
  float gT[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) gT[i] = (float)sgT[v][i];
  float* out = g_poses + ((size_t)b * V + v) * psz;
  if (pose_format == VSL_POSE_MATRIX) {
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] = gT[i];
  } else {
    float g[6];
    pose_vec_grad<float>(poses + ((size_t)b * V + v) * 6, pose_format, gT, g);
#pragma unroll
    for (int i = 0; i < 6; ++i) out[i] = g[i];
  }
}

// =====================================================================================================
// Launch 1: pyramids, RGBA re-layout of the sources, transforms.
// A block stages a tile of kPrepPx level-0 pixels (RB rows x TW columns, RB = max(F, 8), F = 2^(S-1)) of one
// image in shared memory with 16-byte cp.async (rows are contiguous: every global access of the launch is a
// full-line stream), then every thread produces output ELEMENTS from the tile:
//   sources: level 0 as RGBA (float 4 of a pixel = 0) and every coarser level as RGBA,
//   target : coarser levels as RGB,
// one float per thread per step, consecutive threads -> consecutive floats, so the stores are full lines too.
// An element of level s is summed in the order of TF's ResizeArea (ComputePatchSum): the f = 2^s level-0
// values of a contributing row left to right, the f row sums top to bottom, times 1/f^2 -- always from
// LEVEL-0 values, hence bit-exact against the oracle, like pyramid_kernel (vsl_ops.cu).
// Blocks past the tiles zero the borders of the RGBA levels and fill the transform table.
// =====================================================================================================
constexpr int kPrepPx = 1024;       // level-0 pixels per tile
constexpr int kPrepThreads = 256;

struct PrepImgJob {
  const void* tgt;                                  // [B,H,W,3] float32, or uint8 (img_format != VSL_IMG_F32)
  const void* src[VSL_MAX_VIEWS];
  float* tgt_lvl[VSL_MAX_SCALES];                   // RGB levels; [0] = the float32 target, written only for uint8 images
  float img_div, img_sub;                           // uint8 images: value = (float)u8 / img_div - img_sub
  float4* src_lvl[VSL_MAX_VIEWS][VSL_MAX_SCALES];   // zero-bordered RGBA levels
  // consistency term: source view v's own network output at scale s ([B,Hs,Ws,1], nullable); its depth map (x, or 1/x
  // with src_x_inverse) becomes the fourth channel of src_lvl[v][s]
  const float* src_x[VSL_MAX_VIEWS][VSL_MAX_SCALES];
  int src_x_inverse;
  int V, B, H, W, S;
  int tiles_x, tiles_y, n_tiles;                    // per image; n_tiles = (V + 1) * B * tiles_y * tiles_x
  int n_blocks;                                     // persistent blocks of the launch
  int border_begin[VSL_MAX_SCALES + 1];             // prefix sums of border float4 per image over the scales
};

__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }

// fourth channel of a source level: zero padding, or the source view's depth map (consistency term)
struct WChan {
  const float* p;      // element (0, 0) of the tile inside the view's [B,Hs,Ws,1] level; nullptr = zeros
  int stride, inverse;
  VSL_DEV float at(int oy, int ox) const {
    if (p == nullptr) return 0.f;
    const float v = __ldg(p + oy * stride + ox);
    return inverse ? __fdiv_rn(1.0f, v) : v;
  }
};

// All elements of level SHIFT >= 3 that this tile covers: one float per (virtual) thread per step over a
// (pixel, 4) grid whose sides are powers of two (shifts only); RGB output leaves the 4th lane of a pixel idle.
// The caller hands this to the threads that are not busy with levels 1 and 2 (t = their index, nt = how many).
template <int SHIFT, int TW, int RB, bool C4>
VSL_DEV void prep_level(const float* tile, int rows, int cols, float* __restrict__ dst, int dst_row_stride, int t,
                        int nt, const WChan wc) {
  constexpr int f = 1 << SHIFT, C = C4 ? 4 : 3;
  constexpr int OW = TW >> SHIFT, OH = RB >> SHIFT;    // output pixels of a full tile
  constexpr int LOG2OW = ilog2(OW);
  const int ow = cols >> SHIFT, oh = rows >> SHIFT;
  const float scale = 1.0f / (float)(f * f);
  for (int o = t; o < OH * OW * 4; o += nt) {
    const int c = o & 3, px = o >> 2;
    const int oy = px >> LOG2OW, ox = px & (OW - 1);
    if (oy >= oh || ox >= ow || (!C4 && c == 3)) continue;
    float out = 0.f;
    if (c == 3) {
      out = wc.at(oy, ox);
    } else {
      const float* p = tile + (oy << SHIFT) * (TW * 3) + (ox << SHIFT) * 3 + c;
      float acc = 0.f;
#pragma unroll
      for (int ry = 0; ry < f; ++ry) {
        float rs = p[ry * (TW * 3)];
#pragma unroll
        for (int k = 1; k < f; ++k) rs = __fadd_rn(rs, p[ry * (TW * 3) + k * 3]);
        acc = ry == 0 ? rs : __fadd_rn(acc, rs);
      }
      out = __fmul_rn(acc, scale);
    }
    dst[oy * dst_row_stride + ox * C + c] = out;
  }
}

// Levels 1 and 2, one output PIXEL per thread with 8- / 16-byte shared-memory loads (the 2 x 3 or 4 x 3 floats of
// a contributing row are contiguous).  Level 1 has exactly kPrepThreads pixels per full tile, level 2 a quarter.
template <int SHIFT, int TW, int RB, bool C4>
VSL_DEV void prep_level12(const float* tile, int rows, int cols, float* __restrict__ dst, int dst_row_stride, int t,
                          const WChan wc) {
  static_assert(SHIFT == 1 || SHIFT == 2, "vector path covers levels 1 and 2");
  constexpr int f = 1 << SHIFT, OW = TW >> SHIFT, OH = RB >> SHIFT, LOG2OW = ilog2(OW);
  if (t >= OW * OH) return;
  const int oy = t >> LOG2OW, ox = t & (OW - 1);
  if (oy >= (rows >> SHIFT) || ox >= (cols >> SHIFT)) return;
  const float* p = tile + (oy << SHIFT) * (TW * 3) + (ox << SHIFT) * 3;
  float acc[3];
#pragma unroll
  for (int ry = 0; ry < f; ++ry) {
    float a[f * 3];
    if (SHIFT == 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float2 q = reinterpret_cast<const float2*>(p + ry * (TW * 3))[k];
        a[2 * k] = q.x; a[2 * k + 1] = q.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float4 q = reinterpret_cast<const float4*>(p + ry * (TW * 3))[k];
        a[4 * k] = q.x; a[4 * k + 1] = q.y; a[4 * k + 2] = q.z; a[4 * k + 3] = q.w;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float rs = a[c];
#pragma unroll
      for (int k = 1; k < f; ++k) rs = __fadd_rn(rs, a[k * 3 + c]);
      acc[c] = ry == 0 ? rs : __fadd_rn(acc[c], rs);
    }
  }
  const float scale = 1.0f / (float)(f * f);
  float* d = dst + oy * dst_row_stride + ox * (C4 ? 4 : 3);
  if (C4) {
    *reinterpret_cast<float4*>(d) = make_float4(__fmul_rn(acc[0], scale), __fmul_rn(acc[1], scale), __fmul_rn(acc[2], scale),
                                                wc.at(oy, ox));
  } else {
    d[0] = __fmul_rn(acc[0], scale); d[1] = __fmul_rn(acc[1], scale); d[2] = __fmul_rn(acc[2], scale);
  }
}

// Every level of one staged tile.  Level 1: all threads.  Level 2: the first quarter.  Levels >= 3: the rest.
template <int LOG2F, int TW, int RB, bool C4, typename DstOf, typename WOf>
VSL_DEV void prep_levels(const float* tile, int rows, int cols, DstOf dst_of, WOf w_of) {
  const int t = threadIdx.x;
  constexpr int n2 = (TW >> 2) * (RB >> 2);
  if constexpr (LOG2F >= 1) { int st; float* d = dst_of(1, st); prep_level12<1, TW, RB, C4>(tile, rows, cols, d, st, t, w_of(1)); }
  if constexpr (LOG2F >= 2) { int st; float* d = dst_of(2, st); prep_level12<2, TW, RB, C4>(tile, rows, cols, d, st, t, w_of(2)); }
  if (t >= n2) {
    if constexpr (LOG2F >= 3) { int st; float* d = dst_of(3, st); prep_level<3, TW, RB, C4>(tile, rows, cols, d, st, t - n2, kPrepThreads - n2, w_of(3)); }
    if constexpr (LOG2F >= 4) { int st; float* d = dst_of(4, st); prep_level<4, TW, RB, C4>(tile, rows, cols, d, st, t - n2, kPrepThreads - n2, w_of(4)); }
    if constexpr (LOG2F >= 5) { int st; float* d = dst_of(5, st); prep_level<5, TW, RB, C4>(tile, rows, cols, d, st, t - n2, kPrepThreads - n2, w_of(5)); }
  }
}

// Level 0 of a source as RGBA: one pixel per thread per step, a 16-byte store per pixel.
template <int TW, int RB>
VSL_DEV void prep_rgba0(const float* tile, int rows, int cols, float4* __restrict__ dst, int dst_row_stride4, const WChan wc) {
  constexpr int LOG2TW = TW == 32 ? 5 : TW == 64 ? 6 : 7;
  static_assert((1 << LOG2TW) == TW, "tile width must be 32, 64 or 128");
#pragma unroll
  for (int j = 0; j < (TW * RB) / kPrepThreads; ++j) {
    const int px = threadIdx.x + j * kPrepThreads;
    const int y = px >> LOG2TW, x = px & (TW - 1);
    if (y < rows && x < cols) {
      const float* p = tile + y * (TW * 3) + x * 3;
      dst[(size_t)y * dst_row_stride4 + x] = make_float4(p[0], p[1], p[2], wc.at(y, x));
    }
  }
}

// q = n / d for 0 <= n < 2^31, d >= 1 without the generic 20-instruction sequence: float estimate + one correction
VSL_DEV int fast_div(int n, int d, float inv_d) {
  int q = __float2int_rz(((float)n + 0.5f) * inv_d);
  const int r = n - q * d;
  q += (r >= d) - (r < 0);
  return q;
}

// Border zeros of the RGBA levels and the transform table, spread over all blocks of a prep launch and done first, so
// their serial chains (sin / cos, LU inverse) run under the image work instead of as a tail.
VSL_DEV void prep_borders_and_table(const PrepImgJob& job, const PrepJob& prep) {
  const int B = job.B, H = job.H, W = job.W;
  const int per_img = job.border_begin[job.S];
  const int n_border = job.V * B * per_img;
  for (int k0 = blockIdx.x * blockDim.x + threadIdx.x; k0 < n_border + prep.n; k0 += gridDim.x * blockDim.x) {
    int k = k0;
    if (k >= n_border) { prep_one(prep, k - n_border); continue; }
    const int vb = k / per_img;
    k -= vb * per_img;
    int s = 0;
    while (s + 1 < job.S && k >= job.border_begin[s + 1]) ++s;
    k -= job.border_begin[s];
    const int v = vb / B, b = vb - v * B;
    const int Hs = H >> s, Ws = W >> s, st = Ws + 2 * kPad;
    int row, col;
    if (k < 2 * kPad * st) {             // kPad full rows on top, kPad at the bottom
      row = k / st; col = k - row * st;
      if (row >= kPad) row += Hs;
    } else {                             // 2 * kPad columns beside each image row
      k -= 2 * kPad * st;
      row = kPad + k / (2 * kPad);
      const int c = k % (2 * kPad);
      col = c < kPad ? c : Ws + c;
    }
    job.src_lvl[v][s][((size_t)b * (Hs + 2 * kPad) + row) * st + col] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Persistent blocks, three tile buffers: the cp.async of tile i+1 is in flight while tile i is turned into its
// output levels, so a block never sits idle waiting for its load, and one barrier per tile is enough.
// U8: the images arrive as the loader's uint8 (imageselect_Dataloader.py:86-93: decode_jpeg -> to_float -> / 255.0).
// A tile's bytes are staged the same way (a quarter of the traffic), then turned into the float tile through a
// 256-entry table of (float)u8 / img_div - img_sub -- IEEE division and subtraction, i.e. the reference's own
// conversion, bit for bit -- and the float32 level 0 of the target is written out for the fused kernel.
template <int LOG2F, bool U8>
__global__ void __launch_bounds__(kPrepThreads, 4)
loss_prep_kernel(const PrepImgJob job, const PrepJob prep) {
  constexpr int F = 1 << LOG2F, RB = F > 8 ? F : 8, TW = kPrepPx / RB;
  constexpr int NBUF = U8 ? 2 : 3;                  // float tiles (uint8: the in-flight buffers are the byte tiles)
  __shared__ float4 tile4[NBUF][kPrepPx * 3 / 4];
  __shared__ uint4 raw4[U8 ? 3 : 1][U8 ? kPrepPx * 3 / 16 : 1];
  __shared__ float lut[U8 ? 256 : 1];
  if (U8) lut[threadIdx.x] = __fsub_rn(__fdiv_rn((float)threadIdx.x, job.img_div), job.img_sub);   // kPrepThreads == 256
  const int B = job.B, H = job.H, W = job.W;
  // launched programmatically dependent on whatever kernel precedes it in the stream (usually the previous step's
  // finalize, or the network that produced the inputs): only the launch latency overlaps, nothing is read before
  asm volatile("griddepcontrol.wait;" ::: "memory");

  prep_borders_and_table(job, prep);

  // the level-0 images are read exactly once: evict-first in L2, so that what this launch WRITES (the RGBA
  // levels the fused launch gathers from next) is what stays resident
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const int per_img = job.tiles_x * job.tiles_y;
  const float inv_per_img = 1.0f / (float)per_img, inv_tx = 1.0f / (float)job.tiles_x, inv_B = 1.0f / (float)B;

  struct Tile { int im, b, y0, x0, rows, cols; };
  auto decode = [&](int t) {
    Tile q;
    const int ib = fast_div(t, per_img, inv_per_img), rem = t - ib * per_img;   // ib = image * B + b
    const int tyi = fast_div(rem, job.tiles_x, inv_tx), txi = rem - tyi * job.tiles_x;
    q.im = fast_div(ib, B, inv_B); q.b = ib - q.im * B;
    q.y0 = tyi * RB; q.x0 = txi * TW;
    q.rows = min(RB, H - q.y0); q.cols = min(TW, W - q.x0);   // multiples of F
    return q;
  };
  // stage: row r of the tile = cols * 3 contiguous floats
  // uint8: row r of the tile = cols * 3 contiguous bytes at raw + r * TW * 3
  auto stage_u8 = [&](const Tile& q, unsigned char* raw) {
    const unsigned char* __restrict__ img = reinterpret_cast<const unsigned char*>(q.im == 0 ? job.tgt : job.src[q.im > 0 ? q.im - 1 : 0]);
    const unsigned char* __restrict__ g0 = img + (((size_t)q.b * H + q.y0) * W + q.x0) * 3;
    if ((W % 16 == 0) && (q.cols % 16 == 0) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0)) {
      const unsigned ts = (unsigned)__cvta_generic_to_shared(raw);
      const int q_row = q.cols * 3 / 16;                  // 16-byte pieces per row
      for (int e = threadIdx.x; e < q.rows * q_row; e += kPrepThreads) {
        const int r = e / q_row, c = e - r * q_row;
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(ts + (unsigned)(r * TW * 3 + c * 16)),
                     "l"(g0 + (size_t)r * W * 3 + c * 16), "l"(pol)
                     : "memory");
      }
    } else {
      const int n_row = q.cols * 3;
      for (int i = threadIdx.x; i < q.rows * n_row; i += kPrepThreads) {
        const int r = i / n_row, c = i - r * n_row;
        raw[r * TW * 3 + c] = __ldg(g0 + (size_t)r * W * 3 + c);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto stage = [&](const Tile& q, float* tile) {
    const float* __restrict__ img = reinterpret_cast<const float*>(q.im == 0 ? job.tgt : job.src[q.im > 0 ? q.im - 1 : 0]);
    const float* __restrict__ g0 = img + (((size_t)q.b * H + q.y0) * W + q.x0) * 3;
    if ((W % 4 == 0) && (q.cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0)) {
      const unsigned ts = (unsigned)__cvta_generic_to_shared(tile);
      if (q.rows == RB && q.cols == TW) {                 // full tile: compile-time trip counts
        constexpr int QR = TW * 3 / 4;
#pragma unroll
        for (int i = 0; i < (RB * QR) / kPrepThreads; ++i) {
          const int e = threadIdx.x + i * kPrepThreads;
          const int r = e / QR, c = e - r * QR;
          asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(ts + (unsigned)(r * TW * 3 + c * 4) * 4u),
                       "l"(g0 + (size_t)r * W * 3 + c * 4), "l"(pol)
                       : "memory");
        }
      } else {
        const int q_row = q.cols * 3 / 4;                 // float4 per row
        for (int r = threadIdx.x >> 5; r < q.rows; r += kPrepThreads / 32)
          for (int c = threadIdx.x & 31; c < q_row; c += 32)
            asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(ts + (unsigned)(r * TW * 3 + c * 4) * 4u),
                         "l"(g0 + (size_t)r * W * 3 + c * 4), "l"(pol)
                         : "memory");
      }
    } else {
      const int n_row = q.cols * 3;
      for (int i = threadIdx.x; i < q.rows * n_row; i += kPrepThreads) {
        const int r = i / n_row, c = i - r * n_row;
        tile[r * TW * 3 + c] = __ldg(g0 + (size_t)r * W * 3 + c);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int t = blockIdx.x;
  if (t >= job.n_tiles) return;
  Tile cur = decode(t);
  if (U8) stage_u8(cur, reinterpret_cast<unsigned char*>(raw4[0]));
  else stage(cur, reinterpret_cast<float*>(tile4[0]));
  for (int it = 0; t < job.n_tiles; ++it, t += gridDim.x) {
    const float* tile = reinterpret_cast<const float*>(tile4[it % NBUF]);
    const int tn = t + gridDim.x;
    Tile nxt = cur;
    if (tn < job.n_tiles) {
      nxt = decode(tn);
      if (U8) stage_u8(nxt, reinterpret_cast<unsigned char*>(raw4[(it + 1) % 3]));
      else stage(nxt, reinterpret_cast<float*>(tile4[(it + 1) % 3]));
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (U8) {
      // bytes -> floats through the table, four at a time (every row of the byte tile is a multiple of 4 bytes long:
      // cols is a multiple of F >= 1 ... the tail of a ragged row converts stale bytes nobody reads)
      const unsigned* rw = reinterpret_cast<const unsigned*>(raw4[it % 3]);
      float4* tw = tile4[it % NBUF];
#pragma unroll
      for (int j = 0; j < (kPrepPx * 3 / 4) / kPrepThreads; ++j) {
        const int e = threadIdx.x + j * kPrepThreads;
        const unsigned w = rw[e];
        tw[e] = make_float4(lut[w & 255u], lut[(w >> 8) & 255u], lut[(w >> 16) & 255u], lut[w >> 24]);
      }
      __syncthreads();
    }

    const int b = cur.b, y0 = cur.y0, x0 = cur.x0, rows = cur.rows, cols = cur.cols;
    if (cur.im == 0) {
      if (U8) {   // the float32 target at level 0 (the fused kernel streams it): rows of cols * 3 contiguous floats
        float* __restrict__ d0 = job.tgt_lvl[0] + (((size_t)b * H + y0) * W + x0) * 3;
        const int n_row = cols * 3;
        for (int i = threadIdx.x; i < rows * n_row; i += kPrepThreads) {
          const int r = i / n_row, c = i - r * n_row;
          d0[(size_t)r * W * 3 + c] = tile[r * TW * 3 + c];
        }
      }
      prep_levels<LOG2F, TW, RB, false>(tile, rows, cols, [&](int sh, int& stride) {
        const int Hs = H >> sh, Ws = W >> sh;
        stride = Ws * 3;
        return job.tgt_lvl[sh] + (size_t)(((b * Hs + (y0 >> sh)) * Ws + (x0 >> sh)) * 3);
      }, [](int) { return WChan{nullptr, 0, 0}; });
    } else {
      const int v = cur.im - 1;
      auto dst_of = [&](int sh, int& stride) {
        const int Hs = H >> sh, Ws = W >> sh, st = Ws + 2 * kPad;
        stride = st * 4;
        return reinterpret_cast<float*>(job.src_lvl[v][sh] +
                                        (size_t)((b * (Hs + 2 * kPad) + (y0 >> sh) + kPad) * st + (x0 >> sh) + kPad));
      };
      auto w_of = [&](int sh) {
        const float* sx = job.src_x[v][sh];
        const int Hs = H >> sh, Ws = W >> sh;
        return WChan{sx != nullptr ? sx + (size_t)((b * Hs + (y0 >> sh)) * Ws + (x0 >> sh)) : nullptr, Ws, job.src_x_inverse};
      };
      int st0;
      float* d0 = dst_of(0, st0);
      prep_rgba0<TW, RB>(tile, rows, cols, reinterpret_cast<float4*>(d0), st0 / 4, w_of(0));
      prep_levels<LOG2F, TW, RB, true>(tile, rows, cols, dst_of, w_of);
    }
    // three buffers: the one staged next was last read two tiles ago, before the barrier above -- no second barrier
    cur = nxt;
  }
}

// =====================================================================================================
// Launch 1, register form (float32 images, W % 4 == 0, at most 5 scales): no shared memory, no barrier.
// A WARP owns a patch of 32 columns x 16 rows of one image: lane = (ly, lx) = (lane / 8, lane % 8), a thread holds
// the 4 x 4 pixels at (4 ly, 4 lx) of the patch -- 12 aligned 16-byte loads, everything in flight at once -- and
// produces from its registers
//   level 0 as RGBA (sources): 16 stores of 16 bytes, 64 contiguous bytes per row and lane;
//   level 1: its 2 x 2 outputs, level 2: its one output -- thread-local;
//   level k >= 3: the ResizeArea order (a row's 2^k values left to right, the row sums top to bottom) is a CHAIN,
//   so the running row sum walks across the 2^(k-2) lanes of a row group by shuffles (lane j continues where lane
//   j - 1 stopped), then the running column sum down the lane rows: bit-identical to the serial order.
// 16 warp-instructions per 32 pixels instead of the staged kernel's 109: the launch goes from issue-bound to
// memory-bound.  Warps stride over the patches (persistent, no tail beyond one patch).
// =====================================================================================================
// Staging of a patch: TMA bulk copies (cp.async.bulk, one per patch row, byte-counted on an mbarrier per warp and stage)
// unless VSL_PREP_NO_BULK asks for the per-lane 16-byte cp.async form (A/B: 79.4 against 80.1 us per cfg2 step).
#if !defined(VSL_PREP_NO_BULK) && !defined(VSL_PREP_BULK)
#define VSL_PREP_BULK
#endif
constexpr int kPrepRegThreads = 128, kPrepRegBlocks = 4;
constexpr int kPatchF4 = 16 * 24;                 // float4 per staged patch: 16 rows x 32 pixels x 3 floats
constexpr size_t kPrepRegSmem = sizeof(float4) * 2 * kPatchF4 * (kPrepRegThreads / 32);

VSL_DEV float4 ldg_stream4(const float* p, unsigned long long pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}

// the chain over n lanes at lane distance `dist` (1: along a row group, 8: down the lane rows): lane j of a group
// continues the left-to-right sum where lane j - 1 stopped.  v[i]: this lane's own 4 addends of sum i, in order.
template <int n, int dist, int M>
VSL_DEV void chain_across(float (&run)[M], const float (&v)[M][4], int pos) {
#pragma unroll
  for (int j = 1; j < n; ++j) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const float t = __shfl_up_sync(0xffffffffu, run[i], dist);
      const float cand = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(t, v[i][0]), v[i][1]), v[i][2]), v[i][3]);
      run[i] = pos == j ? cand : run[i];
    }
  }
}

VSL_DEV void stg256(float* p, float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a0), "f"(a1), "f"(a2), "f"(a3),
               "f"(a4), "f"(a5), "f"(a6), "f"(a7) : "memory");
}

// Where the products of one patch go.  Source view v: straight to its zero-bordered RGBA levels, 32 bytes (two pixels)
// per store where the level allows -- a whole sector per lane; two 16-byte stores are two partial-sector requests on
// the SM's path to L2, which is what bounds this launch.  WCH: the fourth channel carries the view's depth map
// (consistency term) instead of zeros.  TGT: the target's RGB levels are 12 bytes per pixel -- written lane by lane
// they would be quarter-filled sectors, so they are assembled in the warp's spent stage buffer and leave as
// contiguous rows (prep_flush_tgt).
__host__ __device__ constexpr int prep_scr_off(int sh) { return sh <= 1 ? 0 : prep_scr_off(sh - 1) + (16 >> (sh - 1)) * (32 >> (sh - 1)) * 3; }

template <bool TGT, bool WCH>
struct PrepSink {
  const PrepImgJob& job;
  int v, b, H, W;
  float* scr;
  struct Lvl { float* d; const float* w; int st, wst; };
  // level sh: where this thread's first output pixel (y >> sh, x >> sh) goes, and the row strides (floats)
  VSL_DEV Lvl lvl(int sh, int y, int x) const {
    Lvl L;
    const int Hs = H >> sh, Ws = W >> sh;
    if (TGT) {
      L.st = (32 >> sh) * 3;
      L.d = scr + prep_scr_off(sh) + ((y & 15) >> sh) * L.st + ((x & 31) >> sh) * 3;
      L.w = nullptr; L.wst = 0;
    } else {
      const int st = Ws + 2 * kPad;
      L.st = st * 4;
      L.d = reinterpret_cast<float*>(job.src_lvl[v][sh] + ((size_t)b * (Hs + 2 * kPad) + (y >> sh) + kPad) * st + (x >> sh) + kPad);
      L.wst = Ws;
      L.w = WCH ? job.src_x[v][sh] + ((size_t)b * Hs + (y >> sh)) * Ws + (x >> sh) : nullptr;
    }
    return L;
  }
  VSL_DEV float w_at(const Lvl& L, int r, int j) const {
    if (!WCH) return 0.f;
    const float q = __ldg(L.w + r * L.wst + j);
    return job.src_x_inverse ? __fdiv_rn(1.0f, q) : q;
  }
  // pixel (row r, column j) relative to the thread's first one
  VSL_DEV void px(const Lvl& L, int r, int j, float c0, float c1, float c2) const {
    if (TGT) {
      float* d = L.d + r * L.st + j * 3;
      d[0] = c0; d[1] = c1; d[2] = c2;
    } else {
      *reinterpret_cast<float4*>(L.d + r * L.st + j * 4) = make_float4(c0, c1, c2, w_at(L, r, j));
    }
  }
  // two horizontally adjacent pixels starting at an even column
  VSL_DEV void px2(const Lvl& L, int r, int j, const float (&c)[6]) const {
    if (TGT) {
      float2* d = reinterpret_cast<float2*>(L.d + r * L.st + j * 3);
      d[0] = make_float2(c[0], c[1]); d[1] = make_float2(c[2], c[3]); d[2] = make_float2(c[4], c[5]);
    } else {
      stg256(L.d + r * L.st + j * 4, c[0], c[1], c[2], w_at(L, r, j), c[3], c[4], c[5], w_at(L, r, j + 1));
    }
  }
};

// The target levels of one patch, scratch -> global: consecutive lanes write consecutive words of a level row.
template <int LOG2F>
VSL_DEV void prep_flush_tgt(const PrepImgJob& job, const float* scr, int b, int x0, int y0, int lane) {
  const int H = job.H, W = job.W;
  const int rows = min(16, H - y0), cols = min(32, W - x0);
#pragma unroll
  for (int sh = 1; sh <= LOG2F; ++sh) {
    const int Ws = W >> sh, nr = rows >> sh, nf = (cols >> sh) * 3;             // valid rows, valid floats per row
    float* __restrict__ d = job.tgt_lvl[sh] + (((size_t)b * (H >> sh) + (y0 >> sh)) * Ws + (x0 >> sh)) * 3;
    const float* t = scr + prep_scr_off(sh);
    const int ts = (32 >> sh) * 3;
    if (((nf | Ws) & 1) == 0) {                   // 8-byte words where the level's rows (Ws * 12 bytes) keep them aligned
      const int per = nf >> 1;
      const float inv = 1.0f / (float)per;
      for (int i = lane; i < nr * per; i += 32) {
        const int r = fast_div(i, per, inv), c = i - r * per;
        *reinterpret_cast<float2*>(d + (size_t)r * Ws * 3 + 2 * c) = *reinterpret_cast<const float2*>(t + r * ts + 2 * c);
      }
    } else {
      const float inv = 1.0f / (float)nf;
      for (int i = lane; i < nr * nf; i += 32) {
        const int r = fast_div(i, nf, inv), c = i - r * nf;
        d[(size_t)r * Ws * 3 + c] = t[r * ts + c];
      }
    }
  }
}

// Every level of the 4 x 4 pixels a thread holds (a[row][pixel * 3 + channel]) at (y, x) of image b.
template <int LOG2F, bool TGT, bool WCH>
VSL_DEV void prep_patch(const PrepSink<TGT, WCH> out, const float (&a)[4][12], int x, int y, int lx, int ly, bool in_x) {
  const int H = out.H;
  using Lvl = typename PrepSink<TGT, WCH>::Lvl;
  // ---- level 0 as RGBA
  if (!TGT && in_x) {
    const Lvl L = out.lvl(0, y, x);
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (y + r < H) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float c[6] = {a[r][6 * j], a[r][6 * j + 1], a[r][6 * j + 2], a[r][6 * j + 3], a[r][6 * j + 4], a[r][6 * j + 5]};
          out.px2(L, r, 2 * j, c);
        }
      }
  }
  if constexpr (LOG2F >= 1) {
    // pair sums of every row: (a0 + a1), (a2 + a3) -- level 1's row sums, and the head of every longer chain
    float s2[4][2][3];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) s2[r][j][c] = __fadd_rn(a[r][6 * j + c], a[r][6 * j + 3 + c]);
    if (in_x) {
      const Lvl L = out.lvl(1, y, x);
#pragma unroll
      for (int ry = 0; ry < 2; ++ry)
        if (y + 2 * ry + 1 < H) {
          float c[6];
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) c[3 * j + ch] = __fmul_rn(__fadd_rn(s2[2 * ry][j][ch], s2[2 * ry + 1][j][ch]), 0.25f);
          out.px2(L, ry, 0, c);
        }
    }
    if constexpr (LOG2F >= 2) {
      // row sums of the thread's 4 pixels: ((a0 + a1) + a2) + a3
      float s4[12];                             // [row * 3 + channel]
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) s4[r * 3 + c] = __fadd_rn(__fadd_rn(s2[r][0][c], a[r][6 + c]), a[r][9 + c]);
      float t4[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) t4[c] = __fadd_rn(__fadd_rn(__fadd_rn(s4[c], s4[3 + c]), s4[6 + c]), s4[9 + c]);
      if (in_x && y + 3 < H)
        out.px(out.lvl(2, y, x), 0, 0, __fmul_rn(t4[0], 0.0625f), __fmul_rn(t4[1], 0.0625f), __fmul_rn(t4[2], 0.0625f));
      if constexpr (LOG2F >= 3) {
        // own addends of the row chains, [row * 3 + channel][pixel]
        float own[12][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) own[r * 3 + c][j] = a[r][3 * j + c];
        auto level = [&](auto kc) {
          constexpr int k = decltype(kc)::value, n = 1 << (k - 2);
          float run[12];
#pragma unroll
          for (int i = 0; i < 12; ++i) run[i] = s4[i];
          chain_across<n, 1, 12>(run, own, lx & (n - 1));           // lanes at the end of a row group: full row sums
          float col[3], rows4[3][4];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int r = 0; r < 4; ++r) rows4[c][r] = run[r * 3 + c];
            col[c] = __fadd_rn(__fadd_rn(__fadd_rn(rows4[c][0], rows4[c][1]), rows4[c][2]), rows4[c][3]);
          }
          chain_across<n, 8, 3>(col, rows4, ly & (n - 1));
          if ((lx & (n - 1)) == n - 1 && (ly & (n - 1)) == n - 1 && in_x && y + 3 < H) {
            constexpr float sc = 1.0f / (float)(1 << (2 * k));
            out.px(out.lvl(k, y, x), 0, 0, __fmul_rn(col[0], sc), __fmul_rn(col[1], sc), __fmul_rn(col[2], sc));
          }
        };
        level(std::integral_constant<int, 3>{});
        if constexpr (LOG2F >= 4) level(std::integral_constant<int, 4>{});
      }
    }
  }
}

// Warps stride over the patches.  A patch's 6 KB travel global -> shared memory as 16-byte cp.async, lane = chunk
// (fully coalesced, nothing held in registers on the way), into a two-stage buffer the warp owns: the next patch is in
// flight while this one is turned into its levels, so a warp's loads never wait for its arithmetic and stores.
// U8: the images arrive as the loader's uint8 (imageselect_Dataloader.py:86-93).  A patch is then 16 rows x 96 BYTES
// (six 16-byte chunks per row, the first quarter of a stage), every lane turns its 4 x 4 pixels into float32 through a
// 256-entry table of (float)u / img_div - img_sub (the reference's own IEEE division, computed once per block), and the
// target's level 0 leaves as float32 too (the fused kernel reads it as such); everything after that is the float32 path.
template <int LOG2F, bool U8 = false>
__global__ void __launch_bounds__(kPrepRegThreads, kPrepRegBlocks)
loss_prep_reg_kernel(const PrepImgJob job, const PrepJob prep) {
  static_assert(LOG2F <= 4, "levels above 4 chain across more lane rows than a warp has");
  extern __shared__ float4 smem4[];               // [warps][2 stages][16 rows][24 float4]
  __shared__ float lut[U8 ? 256 : 1];
  if (U8) {
    for (int i = threadIdx.x; i < 256; i += kPrepRegThreads) lut[i] = __fsub_rn(__fdiv_rn((float)i, job.img_div), job.img_sub);
    __syncthreads();
  }
#ifdef VSL_PREP_BULK
  // the staging as TMA bulk copies: one cp.async.bulk per patch ROW (384 bytes, 96 for uint8) instead of one 16-byte
  // cp.async per lane and chunk, completion counted in bytes on an mbarrier per warp and stage
  __shared__ __align__(8) unsigned long long mbar[kPrepRegThreads / 32][2];
  if ((threadIdx.x & 31) == 0) {
    const unsigned m0 = (unsigned)__cvta_generic_to_shared(&mbar[threadIdx.x >> 5][0]);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(m0) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(m0 + 8u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
#endif
  asm volatile("griddepcontrol.wait;" ::: "memory");
  prep_borders_and_table(job, prep);

  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const int B = job.B, H = job.H, W = job.W;
  const int lane = threadIdx.x & 31, lx = lane & 7, ly = lane >> 3;
  constexpr int wpb = kPrepRegThreads / 32;
  const int warp = threadIdx.x >> 5;
  const int warp0 = blockIdx.x * wpb + warp, n_warps = gridDim.x * wpb;
  const int px = (W + 31) >> 5, py = (H + 15) >> 4, per_img = px * py;
  const int n_patches = (job.V + 1) * B * per_img;
  const float inv_per_img = 1.0f / (float)per_img, inv_px = 1.0f / (float)px, inv_B = 1.0f / (float)B;
  float4* const stage0 = smem4 + warp * (2 * kPatchF4);
  const unsigned stage0_s = (unsigned)__cvta_generic_to_shared(stage0);

  struct Patch { int im, b, x0, y0; };
  auto decode = [&](int p) {
    Patch q;
    const int ib = fast_div(p, per_img, inv_per_img), rem = p - ib * per_img;     // ib = image * B + b
    const int pyi = fast_div(rem, px, inv_px), pxi = rem - pyi * px;
    q.im = fast_div(ib, B, inv_B); q.b = ib - q.im * B;
    q.x0 = pxi * 32; q.y0 = pyi * 16;
    return q;
  };
  auto issue = [&](const Patch& q, int stage) {
    const unsigned dst = stage0_s + (unsigned)(stage * kPatchF4 * 16);
#ifdef VSL_PREP_BULK
    {
      const char* __restrict__ img = reinterpret_cast<const char*>(q.im == 0 ? job.tgt : job.src[q.im > 0 ? q.im - 1 : 0]);
      constexpr int px_bytes = U8 ? 3 : 12;
      const char* __restrict__ g = img + (((size_t)q.b * H + q.y0) * W + q.x0) * px_bytes;
      const int rows = min(16, H - q.y0), row_bytes = min(32, W - q.x0) * px_bytes;
      const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar[warp][stage]);
      // the stage was read (and, for the target, written) through the generic proxy: order that before the async writes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (lane == 0) {
        unsigned long long st_;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 %0, [%1], %2;" : "=l"(st_) : "r"(mb), "r"(rows * row_bytes) : "memory");
      }
      __syncwarp();
      if (lane < rows)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(dst + (unsigned)lane * (U8 ? 96u : 384u)), "l"(g + (size_t)lane * W * px_bytes), "r"(row_bytes), "r"(mb), "l"(pol)
                     : "memory");
      return;
    }
#endif
    if (U8) {
      const unsigned char* __restrict__ img = reinterpret_cast<const unsigned char*>(q.im == 0 ? job.tgt : job.src[q.im > 0 ? q.im - 1 : 0]);
      const unsigned char* __restrict__ g = img + (((size_t)q.b * H + q.y0) * W + q.x0) * 3;
      const int rows = min(16, H - q.y0), row_c = min(32, W - q.x0) * 3 / 16;     // 16-byte chunks per row (W % 16 == 0)
#pragma unroll
      for (int k = 0; k < 3; ++k) {                 // 16 rows x 6 chunks = 96
        const int c = lane + 32 * k, r = c / 6, cc = c - r * 6;
        if (r < rows && cc < row_c)
          asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst + (unsigned)c * 16u),
                       "l"(g + (size_t)r * W * 3 + cc * 16), "l"(pol) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      return;
    }
    const float* __restrict__ img = reinterpret_cast<const float*>(q.im == 0 ? job.tgt : job.src[q.im > 0 ? q.im - 1 : 0]);
    const float* __restrict__ g = img + (((size_t)q.b * H + q.y0) * W + q.x0) * 3;
    const int rows = min(16, H - q.y0), row_f4 = min(32, W - q.x0) * 3 / 4;       // 16-byte chunks per row
#pragma unroll
    for (int k = 0; k < kPatchF4 / 32; ++k) {
      const int c = lane + 32 * k, r = c / 24, cc = c - r * 24;
      if (r < rows && cc < row_f4)
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst + (unsigned)c * 16u),
                     "l"(g + (size_t)r * W * 3 + cc * 4), "l"(pol) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int p = warp0;
  if (p >= n_patches) return;
  Patch cur = decode(p);
  issue(cur, 0);
  for (int it = 0; p < n_patches; ++it, p += n_warps) {
    const int pn = p + n_warps;
    Patch nxt = cur;
    if (pn < n_patches) {
      nxt = decode(pn);
      issue(nxt, (it + 1) & 1);
#ifndef VSL_PREP_BULK
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
    }
#ifdef VSL_PREP_BULK
    {   // stage (it & 1) is on its (it >> 1)-th use: wait for that phase of its barrier
      const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar[warp][it & 1]);
      const unsigned parity = (unsigned)(it >> 1) & 1u;
      unsigned done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mb), "r"(parity) : "memory");
    }
#endif
    __syncwarp();
    const int x = cur.x0 + lx * 4, y = cur.y0 + ly * 4;
    const bool in_x = x < W;                      // a thread's 4 columns are inside or outside together (W % 4 == 0)
    float a[4][12];                               // [row][pixel * 3 + channel]
    if (U8) {
      // a staged row is 96 bytes; this lane's 4 pixels are 12 of them = three 32-bit words
      const unsigned* t = reinterpret_cast<const unsigned*>(stage0 + (it & 1) * kPatchF4) + (ly * 4) * 24 + lx * 3;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const unsigned w = t[r * 24 + k];
          a[r][4 * k] = lut[w & 255u]; a[r][4 * k + 1] = lut[(w >> 8) & 255u];
          a[r][4 * k + 2] = lut[(w >> 16) & 255u]; a[r][4 * k + 3] = lut[w >> 24];
        }
      if (cur.im == 0 && in_x) {                  // the float32 target the fused kernel reads: 48 contiguous bytes per row
        float* d0 = job.tgt_lvl[0] + (((size_t)cur.b * H + y) * W + x) * 3;
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (y + r < H) {
            float4* d = reinterpret_cast<float4*>(d0 + (size_t)r * W * 3);
            d[0] = make_float4(a[r][0], a[r][1], a[r][2], a[r][3]);
            d[1] = make_float4(a[r][4], a[r][5], a[r][6], a[r][7]);
            d[2] = make_float4(a[r][8], a[r][9], a[r][10], a[r][11]);
          }
      }
    } else {
      const float4* t = stage0 + (it & 1) * kPatchF4 + (ly * 4) * 24 + lx * 3;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float4 q = t[r * 24 + k];
          a[r][4 * k] = q.x; a[r][4 * k + 1] = q.y; a[r][4 * k + 2] = q.z; a[r][4 * k + 3] = q.w;
        }
    }
    // the image kind is warp-uniform: one specialised body each (no per-store branches)
    const int im = cur.im, b = cur.b;
    if (im == 0) {
      float* scr = reinterpret_cast<float*>(stage0 + (it & 1) * kPatchF4);   // this stage is spent once every lane has read it
      __syncwarp();
      prep_patch<LOG2F, true, false>(PrepSink<true, false>{job, 0, b, H, W, scr}, a, x, y, lx, ly, in_x);
      __syncwarp();
      prep_flush_tgt<LOG2F>(job, scr, b, cur.x0, cur.y0, lane);
    } else if (job.src_x[im - 1][0] == nullptr) {
      prep_patch<LOG2F, false, false>(PrepSink<false, false>{job, im - 1, b, H, W, nullptr}, a, x, y, lx, ly, in_x);
    } else {
      prep_patch<LOG2F, false, true>(PrepSink<false, true>{job, im - 1, b, H, W, nullptr}, a, x, y, lx, ly, in_x);
    }
    __syncwarp();                                 // every lane has read this stage before the next issue refills it
    cur = nxt;
  }
}

// d/d(source image) of the fused step: the gradient levels (zero-bordered RGBA, scattered into by
// loss_fused_kernel<.., DSRC>) folded back onto the caller's [B,H,W,3] layout.  Level s is the 2^s x 2^s block
// mean of level 0 (resize_area), so every level-0 pixel receives 1/4^s of its level-s pixel's gradient.
// One thread per level-0 pixel.
struct FoldJob {
  const float4* glvl[VSL_MAX_VIEWS][VSL_MAX_SCALES];   // gradient levels per view
  float* g_src[VSL_MAX_VIEWS];                         // [B,H,W,3] per view
  int B, H, W, S;
};
// grid = (ceil(W / 256), H, V * B): one launch for every view
__global__ void __launch_bounds__(256)
loss_fold_src_grad_kernel(const FoldJob j) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const int v = blockIdx.z / j.B, b = blockIdx.z - v * j.B;
  if (x >= j.W) return;
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, w = 1.0f;
  for (int s = 0; s < j.S; ++s) {
    const int Hs = j.H >> s, Ws = j.W >> s, st = Ws + 2 * kPad;
    const float4 q = __ldg(j.glvl[v][s] + ((size_t)b * (Hs + 2 * kPad) + (y >> s) + kPad) * st + (x >> s) + kPad);
    g0 = fmaf(w, q.x, g0); g1 = fmaf(w, q.y, g1); g2 = fmaf(w, q.z, g2);
    w *= 0.25f;
  }
  float* o = j.g_src[v] + (((size_t)b * j.H + y) * j.W + x) * 3;
  __stcs(o, g0); __stcs(o + 1, g1); __stcs(o + 2, g2);
}

// d/d(source view's network output) of the consistency term: the interior of the zero-bordered planes the fused
// kernel scattered d/d(source depth) into, chained through depth = 1/x where the step says so.  One thread per
// level-0 position; it serves the same (y, x) of every level that has it.
struct CropJob {
  const float* gsd[VSL_MAX_VIEWS][VSL_MAX_SCALES];
  const float* src_x[VSL_MAX_VIEWS][VSL_MAX_SCALES];
  float* g_src_x[VSL_MAX_VIEWS][VSL_MAX_SCALES];
  int B, H, W, S, inverse;
};
// grid = (ceil(W / 256), H, V * B)
__global__ void __launch_bounds__(256)
loss_crop_src_depth_grad_kernel(const CropJob j) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const int v = blockIdx.z / j.B, b = blockIdx.z - v * j.B;
  for (int s = 0; s < j.S; ++s) {
    const int Hs = j.H >> s, Ws = j.W >> s;
    if (x >= Ws || y >= Hs) return;
    float g = __ldg(j.gsd[v][s] + ((size_t)b * (Hs + 2 * kPad) + y + kPad) * (Ws + 2 * kPad) + x + kPad);
    const size_t o = ((size_t)b * Hs + y) * Ws + x;
    if (j.inverse) {
      const float q = __fdiv_rn(1.0f, __ldg(j.src_x[v][s] + o));   // depth = 1/x: d(depth)/dx = -depth^2
      g *= -q * q;
    }
    __stcs(j.g_src_x[v][s] + o, g);
  }
}

}  // namespace vsl

using namespace vsl;

namespace {

struct WsLayout {
  size_t xf, xq, partials, tgt_pyr, src_pyr, gsrc_pyr, tgt0, gsd, total;  // byte offsets (gsrc_pyr only with want_src_grad, tgt0 only for uint8 images, gsd only with the consistency term)
  size_t tgt_off[VSL_MAX_SCALES];                    // floats, level s of the target pyramid (s >= 1)
  size_t src_off[VSL_MAX_SCALES];                    // float4, level s inside one view's RGBA block
  size_t src_view;                                   // float4 per view
  int n_items, item_begin[VSL_MAX_SCALES + 1], strips[VSL_MAX_SCALES], bands[VSL_MAX_SCALES];
};

int check_desc(const VslLossDesc* d) {
  VSL_REQUIRE(d, VSL_E_NULL);
  VSL_REQUIRE(d->S >= 1 && d->S <= VSL_MAX_SCALES && d->V >= 1 && d->V <= VSL_MAX_VIEWS, VSL_E_SHAPE);
  VSL_REQUIRE(d->B > 0 && d->B <= 65535 / (VSL_MAX_VIEWS + 1) && d->H > 0 && d->W > 0, VSL_E_SHAPE);
  const int F = 1 << (d->S - 1);
  VSL_REQUIRE(d->H % F == 0 && d->W % F == 0 && (d->H >> (d->S - 1)) >= 3 && (d->W >> (d->S - 1)) >= 3, VSL_E_SHAPE);
  // 32-bit pixel offsets inside one image and thread indices of the prep launch
  VSL_REQUIRE((long long)(d->H + 4) * (d->W + 4) < (1ll << 26), VSL_E_SHAPE);
  VSL_REQUIRE((long long)(d->V + 1) * d->B * (d->H + 4) * (d->W + 4) < (1ll << 31), VSL_E_SHAPE);
  VSL_REQUIRE((long long)d->B * (d->H + 4) * (d->W + 4) * 2 * (d->V > 2 ? d->V : 2) < (1ll << 31), VSL_E_SHAPE);  // float offsets
  VSL_REQUIRE(d->pose_format >= VSL_POSE_EULER && d->pose_format <= VSL_POSE_MATRIX, VSL_E_FORMAT);
  VSL_REQUIRE(d->mask_mode >= VSL_MASK_NONE && d->mask_mode <= VSL_MASK_CONST, VSL_E_FORMAT);
  VSL_REQUIRE(d->exact_coords >= 0 && d->exact_coords <= 2, VSL_E_FORMAT);
  VSL_REQUIRE(d->img_format >= VSL_IMG_F32 && d->img_format <= VSL_IMG_U8_RAW, VSL_E_FORMAT);
  VSL_REQUIRE(!(d->want_src_grad && d->img_format != VSL_IMG_F32), VSL_E_UNSUPPORTED);   // no gradient w.r.t. bytes
  VSL_REQUIRE(!(d->want_src_grad && d->exact_coords == 1), VSL_E_UNSUPPORTED);
  VSL_REQUIRE(!d->x_is_logit || d->disp_scale > 0.f, VSL_E_UNSUPPORTED);
  VSL_REQUIRE(d->consist_weight >= 0.f, VSL_E_UNSUPPORTED);
  // the consistency term rides on the fast arithmetic of the float32 entry
  VSL_REQUIRE(!(d->consist_weight > 0.f) || (d->exact_coords != 1 && !d->want_src_grad && !d->x_is_logit &&
                                             d->img_format == VSL_IMG_F32), VSL_E_UNSUPPORTED);
  VSL_REQUIRE(d->ssim_weight >= 0.f && d->ssim_weight <= 1.f, VSL_E_UNSUPPORTED);
  VSL_REQUIRE(!(d->ssim_weight > 0.f) || (d->exact_coords != 1 && !d->want_src_grad && !d->x_is_logit &&
                                          !(d->consist_weight > 0.f)), VSL_E_UNSUPPORTED);
  return VSL_OK;
}

void layout(const VslLossDesc* d, WsLayout* L) {
  const int nt = kLossSlots + 12 * d->V;
  int n = 0;
  size_t tl = 0, sl = 0;
  for (int s = 0; s < d->S; ++s) {
    const int H = d->H >> s, W = d->W >> s;
    L->strips[s] = (W + 31) / 32;
    L->bands[s] = (H + kRH - 1) / kRH;
    L->item_begin[s] = n;
    n += d->B * L->bands[s] * L->strips[s];
    L->tgt_off[s] = tl;
    if (s >= 1) tl += round_up((size_t)d->B * H * W * 3, 4);  // keep every level 16-byte aligned
    L->src_off[s] = sl;
    sl += (size_t)d->B * (H + 2 * kPad) * (W + 2 * kPad);
  }
  L->item_begin[d->S] = n;
  L->n_items = n;
  sl = round_up(sl, 2);        // every view's block starts 32-byte aligned (the prep launch stores pixel pairs)
  L->src_view = sl;
  L->xf = 0;
  L->xq = round_up(sizeof(Xform) * (size_t)d->S * d->V * d->B, 256);
  L->partials = L->xq + round_up(sizeof(XformQ) * (size_t)d->S * d->V * d->B, 256);
  L->tgt_pyr = L->partials + round_up(sizeof(float) * (size_t)n * nt, 256);
  L->src_pyr = L->tgt_pyr + round_up(sizeof(float) * tl, 256);
  L->gsrc_pyr = L->src_pyr + sizeof(float4) * sl * (size_t)d->V;
  L->tgt0 = L->gsrc_pyr + (d->want_src_grad ? sizeof(float4) * sl * (size_t)d->V : 0);
  L->gsd = L->tgt0 + (d->img_format != VSL_IMG_F32 ? round_up(sizeof(float) * (size_t)d->B * d->H * d->W * 3, 256) : 0);
  L->total = L->gsd + (d->consist_weight > 0.f ? round_up(sizeof(float) * sl * (size_t)d->V, 256) : 0);
}

template <typename Kern>
int launch_fused_as(Kern kern, size_t smem_bytes, const LossParams& P, cudaStream_t st) {
  // > 48 KB of dynamic shared memory needs the opt-in; idempotent and cheap, so set on every call (no state)
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return (int)e;
  const int n = P.item_begin[P.S];
  // programmatic dependent launch: blocks may be scheduled while the prep launch drains; the kernel itself
  // waits (griddepcontrol.wait) before touching anything prep produced
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((n + kWarps - 1) / kWarps);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, P);
  if (e != cudaSuccess) return (int)e;
  return VSL_OK;
}

template <int V, bool EXACT, bool DSRC, bool CONS = false>
int launch_fused(const LossParams& P, cudaStream_t st) {
  return launch_fused_as(loss_fused_kernel<V, EXACT, DSRC, CONS>, WarpSmem<V, EXACT>::block_bytes, P, st);
}
// fast arithmetic: an even number of views runs the view-paired kernel (packed fp32x2, vsl_loss_pair.cu: 23 % fewer
// instructions, same time at 128x416, 2 % faster at 480x640); exact_coords == 2 asks for the scalar kernel instead,
// so that the two can be compared on the same inputs
template <int V>
int launch_fused_fast(const LossParams& P, int scalar_only, cudaStream_t st) {
  if constexpr (V % 2 == 0) {
    if (!scalar_only) return launch_fused_pair(V, P, st);
  }
  return launch_fused<V, false, false>(P, st);
}

template <int V>
int run_loss(const VslLossDesc* d, LossParams& P, const float* poses, const float* K_pyr, float* losses,
             float* g_poses, cudaStream_t st) {
  if (d->ev_main_begin != nullptr) cudaEventRecord((cudaEvent_t)d->ev_main_begin, st);
  // d/d(source) rides on the fast arithmetic only (check_desc refuses exact_coords + want_src_grad)
#ifdef VSL_DEV_V2_ONLY
  const int rc = launch_fused_fast<V>(P, d->exact_coords == 2, st);
#else
  const int rc = d->consist_weight > 0.f ? launch_fused<V, false, false, true>(P, st)
                 : d->want_src_grad      ? launch_fused<V, false, true>(P, st)
                                         : (d->exact_coords == 1 ? launch_fused<V, true, false>(P, st)
                                                                 : launch_fused_fast<V>(P, d->exact_coords == 2, st));
#endif
  if (rc != VSL_OK) return rc;
  if (d->ev_main_end != nullptr) cudaEventRecord((cudaEvent_t)d->ev_main_end, st);
  if (d->ssim_weight > 0.f) {   // extension: the SSIM share of the photometric term, added to the same outputs
    const int rs = launch_ssim_term(d, P, st);
    if (rs != VSL_OK) return rs;
  }
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(d->B + 1);
    constexpr int kBig = V <= 2 ? 1024 : 512;
    // rows of ONE image over all scales: the image blocks want the wide form only when they would otherwise make
    // several passes themselves (cfg5: 405 rows); at 72 rows per image (128 x 416) the wide block's larger
    // reductions cost more than the loss block gains, whatever the batch (measured at B = 256: +5 us)
    int rows_per_image = 0;
    for (int sc = 0; sc < P.S; ++sc) rows_per_image += P.bands[sc] * P.strips[sc];
    const bool big = rows_per_image > kFinBigRows;
    cfg.blockDim = dim3(big ? kBig : 256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = big ? cudaLaunchKernelEx(&cfg, loss_finalize_kernel<V, kBig>, P, poses, K_pyr, d->pose_format,
                                                   1.0f / d->loss_scale, losses, g_poses)
                              : cudaLaunchKernelEx(&cfg, loss_finalize_kernel<V, 256>, P, poses, K_pyr, d->pose_format,
                                                   1.0f / d->loss_scale, losses, g_poses);
    if (e != cudaSuccess) return (int)e;
  }
  return launch_status();
}

int launch_prep(const PrepImgJob& job, const PrepJob& prep, bool u8, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(job.n_blocks);
  cfg.blockDim = dim3(kPrepThreads);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaSuccess;
  // images whose rows are 16-byte aligned take the register form (a warp per 32 x 16 patch): float32 needs W % 4 == 0,
  // the loader's uint8 W % 16 == 0 (a patch row is 96 bytes = six 16-byte chunks)
  bool reg_form = job.W % (u8 ? 16 : 4) == 0 && job.S <= 5 && aligned(job.tgt, 16);
  for (int v = 0; v < job.V; ++v) reg_form = reg_form && aligned(job.src[v], 16);
#ifdef VSL_PREP_STAGED
  reg_form = false;
#endif
  if (reg_form) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_patches = (job.V + 1) * job.B * ((job.W + 31) / 32) * ((job.H + 15) / 16);
    cfg.gridDim = dim3(std::min((n_patches + kPrepRegThreads / 32 - 1) / (kPrepRegThreads / 32), sms * kPrepRegBlocks));
    cfg.blockDim = dim3(kPrepRegThreads);
    cfg.dynamicSmemBytes = kPrepRegSmem;
    auto go = [&](auto kern) {
      cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPrepRegSmem);
      return e2 != cudaSuccess ? e2 : cudaLaunchKernelEx(&cfg, kern, job, prep);
    };
    if (u8) {
      switch (job.S) {
        case 1: e = go(loss_prep_reg_kernel<0, true>); break;
        case 2: e = go(loss_prep_reg_kernel<1, true>); break;
        case 3: e = go(loss_prep_reg_kernel<2, true>); break;
        case 4: e = go(loss_prep_reg_kernel<3, true>); break;
        default: e = go(loss_prep_reg_kernel<4, true>); break;
      }
    } else {
      switch (job.S) {
        case 1: e = go(loss_prep_reg_kernel<0>); break;
        case 2: e = go(loss_prep_reg_kernel<1>); break;
        case 3: e = go(loss_prep_reg_kernel<2>); break;
        case 4: e = go(loss_prep_reg_kernel<3>); break;
        default: e = go(loss_prep_reg_kernel<4>); break;
      }
    }
  } else if (u8) {
    switch (job.S) {
      case 1: e = cudaLaunchKernelEx(&cfg, loss_prep_kernel<0, true>, job, prep); break;
      case 2: e = cudaLaunchKernelEx(&cfg, loss_prep_kernel<1, true>, job, prep); break;
      case 3: e = cudaLaunchKernelEx(&cfg, loss_prep_kernel<2, true>, job, prep); break;
      case 4: e = cudaLaunchKernelEx(&cfg, loss_prep_kernel<3, true>, job, prep); break;
      case 5: e = cudaLaunchKernelEx(&cfg, loss_prep_kernel<4, true>, job, prep); break;
      default: e = cudaLaunchKernelEx(&cfg, loss_prep_kernel<5, true>, job, prep); break;
    }
  } else {
    switch (job.S) {
      case 1: e = cudaLaunchKernelEx(&cfg, loss_prep_kernel<0, false>, job, prep); break;
      case 2: e = cudaLaunchKernelEx(&cfg, loss_prep_kernel<1, false>, job, prep); break;
      case 3: e = cudaLaunchKernelEx(&cfg, loss_prep_kernel<2, false>, job, prep); break;
      case 4: e = cudaLaunchKernelEx(&cfg, loss_prep_kernel<3, false>, job, prep); break;
      case 5: e = cudaLaunchKernelEx(&cfg, loss_prep_kernel<4, false>, job, prep); break;
      default: e = cudaLaunchKernelEx(&cfg, loss_prep_kernel<5, false>, job, prep); break;
    }
  }
  if (e != cudaSuccess) return (int)e;
  return launch_status();
}

}  // namespace

extern "C" {

size_t vsl_loss_ws_bytes(const VslLossDesc* d) {
  if (check_desc(d) != VSL_OK) return 0;
  WsLayout L;
  layout(d, &L);
  return L.total;
}

int vsl_loss_ws_layout(const VslLossDesc* d, long long* out) {
  const int rc = check_desc(d);
  if (rc != VSL_OK) return rc;
  VSL_REQUIRE(out, VSL_E_NULL);
  WsLayout L;
  layout(d, &L);
  for (int s = 0; s < d->S; ++s) {
    out[s] = s == 0 ? (d->img_format != VSL_IMG_F32 ? (long long)L.tgt0 : -1ll)
                    : (long long)(L.tgt_pyr + sizeof(float) * L.tgt_off[s]);
    for (int v = 0; v < d->V; ++v)
      out[d->S + v * d->S + s] = (long long)(L.src_pyr + sizeof(float4) * (L.src_view * (size_t)v + L.src_off[s]));
  }
  return VSL_OK;
}

}  // extern "C"

namespace {
// tgt / srcs: float32 or uint8 images, as d->img_format says
int loss_fwd_bwd(const VslLossDesc* d, const void* tgt, const void* const* srcs, const float* const* x_pyr,
                 const float* poses, const float* K_pyr, const float* const* logits_pyr,
                 const float* const* mask_pyr, float* losses, float* const* g_x_pyr, float* g_poses,
                 float* const* g_logits_pyr, float* const* g_srcs, void* ws, vsl_stream_t stream,
                 const float* const* src_x_pyr = nullptr, float* const* g_src_x_pyr = nullptr) {
  int rc = VSL_OK;
  const bool cons = d->consist_weight > 0.f;
  VSL_REQUIRE(!cons || src_x_pyr, VSL_E_NULL);
  VSL_REQUIRE(tgt && srcs && x_pyr && poses && K_pyr && losses && g_x_pyr && g_poses && ws, VSL_E_NULL);
  VSL_REQUIRE(d->mask_mode != VSL_MASK_EXP || (logits_pyr && g_logits_pyr), VSL_E_NULL);
  VSL_REQUIRE(d->mask_mode != VSL_MASK_CONST || mask_pyr, VSL_E_NULL);
  VSL_REQUIRE(d->loss_scale != 0.f, VSL_E_UNSUPPORTED);
  VSL_REQUIRE(!d->want_src_grad || g_srcs, VSL_E_NULL);
  VSL_REQUIRE(aligned(ws, 256), VSL_E_ALIGN);
  cudaStream_t st = (cudaStream_t)stream;
  WsLayout L;
  layout(d, &L);
  char* base = reinterpret_cast<char*>(ws);
  Xform* xf = reinterpret_cast<Xform*>(base + L.xf);
  float* tgt_pyr = reinterpret_cast<float*>(base + L.tgt_pyr);
  float4* src_pyr = reinterpret_cast<float4*>(base + L.src_pyr);
  float4* gsrc_pyr = reinterpret_cast<float4*>(base + L.gsrc_pyr);
  const bool u8 = d->img_format != VSL_IMG_F32;

  LossParams P;
  P.B = d->B; P.H = d->H; P.W = d->W; P.S = d->S; P.V = d->V;
  P.mask_mode = d->mask_mode; P.depth_is_inverse = d->depth_is_inverse; P.smooth_on_inverse = d->smooth_on_inverse;
  P.x_is_logit = d->x_is_logit; P.disp_scale = d->disp_scale; P.disp_min = d->disp_min;
  P.xf = xf;
  P.xq = reinterpret_cast<const XformQ*>(base + L.xq);
  P.partials = reinterpret_cast<float*>(base + L.partials);
  for (int s = 0; s <= d->S; ++s) P.item_begin[s] = L.item_begin[s];
  for (int v = 0; v < d->V; ++v) VSL_REQUIRE(srcs[v], VSL_E_NULL);
  for (int s = 0; s < VSL_MAX_SCALES; ++s) {
    P.tgt[s] = nullptr; P.x[s] = nullptr; P.logits[s] = nullptr; P.mask[s] = nullptr;
    P.g_x[s] = nullptr; P.g_logits[s] = nullptr; P.lg_vec4[s] = 0; P.x_vec2[s] = 0; P.strips[s] = 0; P.bands[s] = 0;
    for (int v = 0; v < VSL_MAX_VIEWS; ++v) { P.src[v][s] = nullptr; P.gsrc[v][s] = nullptr; P.gsd[v][s] = nullptr; }
    P.ccon[s] = 0.f;
  }
  float* gsd = reinterpret_cast<float*>(base + L.gsd);
  for (int s = 0; s < d->S; ++s) {
    const int H = d->H >> s, W = d->W >> s;
    VSL_REQUIRE(x_pyr[s] && g_x_pyr[s], VSL_E_NULL);
    P.x[s] = x_pyr[s];
    P.g_x[s] = g_x_pyr[s];
    P.x_vec2[s] = (W % 2 == 0) && aligned(x_pyr[s], 8);
    if (d->mask_mode == VSL_MASK_EXP) {
      VSL_REQUIRE(logits_pyr[s] && g_logits_pyr[s], VSL_E_NULL);
      VSL_REQUIRE(aligned(logits_pyr[s], 8) && aligned(g_logits_pyr[s], 8), VSL_E_ALIGN);
      P.logits[s] = logits_pyr[s];
      P.g_logits[s] = g_logits_pyr[s];
      P.lg_vec4[s] = (d->V % 2 == 0) && aligned(logits_pyr[s], 16) && aligned(g_logits_pyr[s], 16);
    }
    if (d->mask_mode == VSL_MASK_CONST) {
      VSL_REQUIRE(mask_pyr[s], VSL_E_NULL);
      P.mask[s] = mask_pyr[s];
    }
    P.tgt[s] = (s == 0) ? (u8 ? reinterpret_cast<const float*>(base + L.tgt0) : reinterpret_cast<const float*>(tgt))
                        : tgt_pyr + L.tgt_off[s];
    for (int v = 0; v < d->V; ++v) {
      P.src[v][s] = src_pyr + L.src_view * (size_t)v + L.src_off[s];
      P.gsrc[v][s] = d->want_src_grad ? gsrc_pyr + L.src_view * (size_t)v + L.src_off[s] : nullptr;
      if (cons) {
        VSL_REQUIRE(src_x_pyr[v * d->S + s], VSL_E_NULL);
        VSL_REQUIRE(!g_src_x_pyr || g_src_x_pyr[v * d->S + s], VSL_E_NULL);
        P.gsd[v][s] = g_src_x_pyr ? gsd + L.src_view * (size_t)v + L.src_off[s] : nullptr;
      }
    }
    P.strips[s] = L.strips[s]; P.bands[s] = L.bands[s];
    P.Hs[s] = H; P.Ws[s] = W;
    P.stride4[s] = W + 2 * kPad;
    P.plane4[s] = (H + 2 * kPad) * (W + 2 * kPad);
    P.coff[s] = (int)((unsigned)(kPad * P.stride4[s] + kPad) - kMagicBits * (unsigned)(P.stride4[s] + 1));
    P.Wf[s] = (float)W; P.Hf[s] = (float)H;
    P.wstep[s] = 2.0f / (float)(W - 1);  // fp32 division, as grid_step() does on the device
    P.hstep[s] = 2.0f / (float)(H - 1);
    const double npx = (double)d->B * H * W;
    const double dw = d->pixel_scale_norm ? (double)d->data_weight / (double)(1 << s) : (double)d->data_weight;
    P.cpix[s] = (float)((double)d->loss_scale * dw * (1.0 - (double)d->ssim_weight) / (npx * 3.0));   // the L1 share
    P.cexp[s] = (float)((double)d->loss_scale * (double)d->explain_reg_weight / npx);
    P.ccon[s] = (float)((double)d->loss_scale * (double)d->consist_weight / npx);
    const double sw = (double)d->loss_scale * (double)d->smooth_weight / (double)(1 << s);
    P.cxx[s] = (float)(sw / ((double)d->B * H * (W - 2)));
    P.cxy[s] = (float)(sw / ((double)d->B * (H - 1) * (W - 1)));
    P.cyx[s] = P.cxy[s];
    P.cyy[s] = (float)(sw / ((double)d->B * (H - 2) * W));
  }

  // 0. (only with want_src_grad) the gradient levels start from zero
  if (d->want_src_grad) {
    for (int v = 0; v < d->V; ++v) VSL_REQUIRE(g_srcs[v], VSL_E_NULL);
    const cudaError_t e = cudaMemsetAsync(gsrc_pyr, 0, sizeof(float4) * L.src_view * (size_t)d->V, st);
    if (e != cudaSuccess) return (int)e;
  }
  if (cons && g_src_x_pyr) {
    const cudaError_t e = cudaMemsetAsync(gsd, 0, sizeof(float) * L.src_view * (size_t)d->V, st);
    if (e != cudaSuccess) return (int)e;
  }
  // 1. pyramids, RGBA source levels, transforms
  PrepJob prep = make_prep(poses, K_pyr, d->B, d->S, d->V, d->pose_format, xf, nullptr);
  prep.xq = reinterpret_cast<XformQ*>(base + L.xq);
  PrepImgJob job;
  job.tgt = tgt;
  // uint8 images: (float)u8 / 255 (imageselect_Dataloader.py:93), / 255 - 0.5 (imageselect_Dataloader_optflow_dim11.py:128)
  // or the raw value (imageselect_Dataloader_optflow.py:129, normalisation commented out)
  job.img_div = d->img_format == VSL_IMG_U8_RAW ? 1.0f : 255.0f;
  job.img_sub = d->img_format == VSL_IMG_U8_255_CENTRED ? 0.5f : 0.0f;
  job.V = d->V; job.B = d->B; job.H = d->H; job.W = d->W; job.S = d->S;
  job.src_x_inverse = d->depth_is_inverse;
  const int F = 1 << (d->S - 1);
  {
    const int RB = F > 8 ? F : 8, TW = kPrepPx / RB;
    job.tiles_x = (d->W + TW - 1) / TW;
    job.tiles_y = (d->H + RB - 1) / RB;
    job.n_tiles = (d->V + 1) * d->B * job.tiles_x * job.tiles_y;
  }
  job.border_begin[0] = 0;
  for (int s = 0; s < VSL_MAX_SCALES; ++s) {
    job.tgt_lvl[s] = (s >= 1 && s < d->S) ? tgt_pyr + L.tgt_off[s] : nullptr;
    if (s == 0 && u8) job.tgt_lvl[0] = reinterpret_cast<float*>(base + L.tgt0);
    for (int v = 0; v < VSL_MAX_VIEWS; ++v) {
      job.src[v] = v < d->V ? srcs[v] : nullptr;
      job.src_lvl[v][s] = (v < d->V && s < d->S) ? src_pyr + L.src_view * (size_t)v + L.src_off[s] : nullptr;
      job.src_x[v][s] = (cons && v < d->V && s < d->S) ? src_x_pyr[v * d->S + s] : nullptr;
    }
    if (s < d->S) {
      const int H = d->H >> s, W = d->W >> s;
      job.border_begin[s + 1] = job.border_begin[s] + (H + 2 * kPad) * (W + 2 * kPad) - H * W;
    } else {
      job.border_begin[s + 1] = job.border_begin[s];
    }
  }
  {
    // persistent blocks: as many as can be resident (2 x 12 KB of shared memory and <= 64 registers x 256 threads
    // each: 4 per SM), never more than there are tiles
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    job.n_blocks = std::min(job.n_tiles, sms * 4);
  }
  rc = launch_prep(job, prep, u8, st);
  if (rc != VSL_OK) return rc;
  // 2 + 3. fused loss and finalize
#ifdef VSL_DEV_V2_ONLY   // development builds only (never the shipped library): one instantiation, quick to compile
  rc = d->V == 2 ? run_loss<2>(d, P, poses, K_pyr, losses, g_poses, st) : VSL_E_UNSUPPORTED;
#else
  switch (d->V) {
    case 1: rc = run_loss<1>(d, P, poses, K_pyr, losses, g_poses, st); break;
    case 2: rc = run_loss<2>(d, P, poses, K_pyr, losses, g_poses, st); break;
    case 3: rc = run_loss<3>(d, P, poses, K_pyr, losses, g_poses, st); break;
    default: rc = run_loss<4>(d, P, poses, K_pyr, losses, g_poses, st); break;
  }
#endif
  if (rc == VSL_OK && cons && g_src_x_pyr) {
    // 4'. d/d(source views' network outputs) of the consistency term: interior of the scattered planes, chain rule
    CropJob cj;
    for (int v = 0; v < VSL_MAX_VIEWS; ++v)
      for (int s = 0; s < VSL_MAX_SCALES; ++s) {
        const bool on = v < d->V && s < d->S;
        cj.gsd[v][s] = on ? P.gsd[v][s] : nullptr;
        cj.src_x[v][s] = on ? src_x_pyr[v * d->S + s] : nullptr;
        cj.g_src_x[v][s] = on ? g_src_x_pyr[v * d->S + s] : nullptr;
      }
    cj.B = d->B; cj.H = d->H; cj.W = d->W; cj.S = d->S; cj.inverse = d->depth_is_inverse;
    loss_crop_src_depth_grad_kernel<<<dim3((d->W + 255) / 256, d->H, d->V * d->B), 256, 0, st>>>(cj);
    rc = launch_status();
  }
  if (rc != VSL_OK || !d->want_src_grad) return rc;
  // 4. d/d(source image): fold the gradient levels back to [B,H,W,3], all views in one launch
  {
    FoldJob fj;
    for (int v = 0; v < VSL_MAX_VIEWS; ++v) {
      fj.g_src[v] = v < d->V ? g_srcs[v] : nullptr;
      for (int s = 0; s < VSL_MAX_SCALES; ++s) fj.glvl[v][s] = (v < d->V && s < d->S) ? P.gsrc[v][s] : nullptr;
    }
    fj.B = d->B; fj.H = d->H; fj.W = d->W; fj.S = d->S;
    loss_fold_src_grad_kernel<<<dim3((d->W + 255) / 256, d->H, d->V * d->B), 256, 0, st>>>(fj);
  }
  return launch_status();
}
}  // namespace

extern "C" {

int vsl_loss_fwd_bwd(const VslLossDesc* d, const float* tgt, const float* const* srcs, const float* const* x_pyr,
                     const float* poses, const float* K_pyr, const float* const* logits_pyr,
                     const float* const* mask_pyr, float* losses, float* const* g_x_pyr, float* g_poses,
                     float* const* g_logits_pyr, float* const* g_srcs, void* ws, vsl_stream_t stream) {
  const int rc = check_desc(d);
  if (rc != VSL_OK) return rc;
  VSL_REQUIRE(d->img_format == VSL_IMG_F32, VSL_E_FORMAT);
  VSL_REQUIRE(!(d->consist_weight > 0.f), VSL_E_UNSUPPORTED);   // that step is vsl_loss_consist_fwd_bwd
  return loss_fwd_bwd(d, tgt, reinterpret_cast<const void* const*>(srcs), x_pyr, poses, K_pyr, logits_pyr, mask_pyr, losses,
                      g_x_pyr, g_poses, g_logits_pyr, g_srcs, ws, stream);
}

int vsl_loss_consist_fwd_bwd(const VslLossDesc* d, const float* tgt, const float* const* srcs,
                             const float* const* x_pyr, const float* const* src_x_pyr, const float* poses,
                             const float* K_pyr, const float* const* logits_pyr, const float* const* mask_pyr,
                             float* losses, float* const* g_x_pyr, float* const* g_src_x_pyr, float* g_poses,
                             float* const* g_logits_pyr, void* ws, vsl_stream_t stream) {
  const int rc = check_desc(d);
  if (rc != VSL_OK) return rc;
  VSL_REQUIRE(d->consist_weight > 0.f, VSL_E_UNSUPPORTED);
  return loss_fwd_bwd(d, tgt, reinterpret_cast<const void* const*>(srcs), x_pyr, poses, K_pyr, logits_pyr, mask_pyr, losses,
                      g_x_pyr, g_poses, g_logits_pyr, nullptr, ws, stream, src_x_pyr, g_src_x_pyr);
}

int vsl_loss_fwd_bwd_u8(const VslLossDesc* d, const unsigned char* tgt, const unsigned char* const* srcs,
                        const float* const* x_pyr, const float* poses, const float* K_pyr,
                        const float* const* logits_pyr, const float* const* mask_pyr, float* losses,
                        float* const* g_x_pyr, float* g_poses, float* const* g_logits_pyr, void* ws,
                        vsl_stream_t stream) {
  const int rc = check_desc(d);
  if (rc != VSL_OK) return rc;
  VSL_REQUIRE(d->img_format != VSL_IMG_F32, VSL_E_FORMAT);
  return loss_fwd_bwd(d, tgt, reinterpret_cast<const void* const*>(srcs), x_pyr, poses, K_pyr, logits_pyr, mask_pyr, losses,
                      g_x_pyr, g_poses, g_logits_pyr, nullptr, ws, stream);
}

}  // extern "C"
