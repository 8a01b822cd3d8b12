// libvsl: the fused multi-scale view-synthesis loss, forward AND backward in one pass.
//
// What the reference does per training step (train.py:107-135, train_depth_then_cam_lr.py:211-328) is a
// Python loop over S scales x V source views of ~150 TF ops each, then autodiff of all of it.  Here the
// whole thing is three launches:
//
//   1. pyramid_kernel       (vsl_ops.cu)  resize_area levels 1..S-1 of the target and source images, and
//                                         (riding along) K_s^-1 and P = K4_s . T_v per (scale, view, batch)
//   2. loss_fused_kernel    (this file)   every scale and view in ONE grid, one block per tile:
//        per target pixel: smoothness stencil (forward sum + gradient) on a shared-memory tile of x,
//        then per view: back-project, pose, project, bilinear gather of the source, L1 against the
//        target, explainability / validity mask, softmax cross-entropy regulariser -- and, because the
//        loss is a weighted sum of means whose upstream gradient is known (loss_scale), the gradients
//        d/dx, d/dlogits are written in the same pass and dP = sum du (x) [cam;1] is accumulated in
//        registers and reduced once per block.  No full-resolution intermediate is ever written.
//   3. loss_finalize_kernel (this file)   fixed-order reduction of the block partials (deterministic),
//        dT = K4^T dP summed over scales, pose chain rule, the three loss scalars.
//
// Work decomposition: a tile is 8 rows x (32*R) columns of one image at one scale; block = 8 warps, warp w
// owns row w, a thread visits R pixels 32 columns apart, so every global access of a warp is a run of 32
// consecutive pixels.  Tiles of all scales live in one 1-D grid, largest scale first, and the hardware block
// scheduler balances them (a persistent variant with static cost-balanced ranges measured 10 % slower).
//
// Streaming operands (target tile, logits tile) enter shared memory with 16-byte cp.async issued before the
// smoothness pass, so their latency is covered by it; d/dlogits is assembled in the logits tile in place and
// leaves with 16-byte stores.  Only the data-dependent bilinear gathers go through L1 as scalar loads.
#include <algorithm>

#include "vsl_common.cuh"
// Part of the single translation unit vsl_lib.cu (prep_xforms_kernel / pyramid_kernel come from vsl_ops.cu).

namespace vsl {

constexpr int kTH = 8;                          // tile rows = warps per block
constexpr int kThreads = 32 * kTH;
constexpr int kMaxR = 4;                        // pixels per thread
constexpr int kTW = 32 * kMaxR;                 // tile columns
constexpr int kHalo = 2;
constexpr int kTileH = kTH + 2 * kHalo;         // 12 rows of x
constexpr int kQStride = kTW + 2 * kHalo + 4;   // 136
constexpr int kOwnW = kTW + kHalo;              // owners: tile + 2 columns to the left
constexpr int kOwnH = kTH + kHalo;              //         tile + 2 rows above
constexpr int kOStride = kOwnW + 2;             // 132

struct LossParams {
  int B, H, W, S, V;
  int mask_mode, depth_is_inverse, smooth_on_inverse;
  const float* tgt[VSL_MAX_SCALES];
  const float* src[VSL_MAX_VIEWS][VSL_MAX_SCALES];
  const float* x[VSL_MAX_SCALES];
  const float* logits[VSL_MAX_SCALES];
  const float* mask[VSL_MAX_SCALES];
  float* g_x[VSL_MAX_SCALES];
  float* g_logits[VSL_MAX_SCALES];
  const Xform* xf;                 // [S][V][B]
  float* partials;                 // [n_items][NT]
  float cpix[VSL_MAX_SCALES];      // loss_scale * data_weight_s / (B Hs Ws 3)
  float cexp[VSL_MAX_SCALES];      // loss_scale * explain_reg_weight / (B Hs Ws)
  float csm[VSL_MAX_SCALES][4];    // loss_scale * smooth_weight / 2^s / count_k   (xx, xy, yx, yy)
  int item_begin[VSL_MAX_SCALES + 1];
  int tiles_x[VSL_MAX_SCALES], bands[VSL_MAX_SCALES], R[VSL_MAX_SCALES];
  int staged[VSL_MAX_SCALES];      // 1: rows of this scale are 16-byte aligned -> cp.async / vector path
  int img_begin[VSL_MAX_SCALES + 1];  // prefix sums of tiles per image over the scales (grid.x index -> scale)
  float inv_tx[VSL_MAX_SCALES];    // 1 / tiles_x
  float wstep[VSL_MAX_SCALES], hstep[VSL_MAX_SCALES];  // meshgrid linspace steps 2/(W-1), 2/(H-1) in fp32
};

template <int V> struct NT { static constexpr int value = 3 + 12 * V; };

// dynamic shared memory carve-up (floats)
template <int V> struct Smem {
  static constexpr int qt = 0;                                   // [kTileH][kQStride]
  static constexpr int sA = qt + kTileH * kQStride;              // [kOwnH][kOStride] x 3
  static constexpr int sB = sA + kOwnH * kOStride;
  static constexpr int sC = sB + kOwnH * kOStride;
  static constexpr int xf = sC + kOwnH * kOStride;               // V * 21 (padded to 24)
  static constexpr int scratch = xf + V * 24;                    // NT * kTH
  static constexpr int tgt = (scratch + NT<V>::value * kTH + 3) / 4 * 4;   // [kTH][kTW*3], 16-byte aligned
  static constexpr int lg = tgt + kTH * kTW * 3;                 // [kTH][kTW*2V]
  static constexpr int total = lg + kTH * kTW * 2 * V;
  static constexpr size_t bytes = sizeof(float) * total;
};

VSL_DEV float signed_by(float c, float v) {  // c * sign(v), sign(0) = 0
  return (v == 0.f) ? 0.f : copysignf(c, v);
}

VSL_DEV void cp_async16(float* smem_dst, const float* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
VSL_DEV void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// EXACT = true : coordinates, softmax and the warped value follow the reference's rounding sequence
//                (bit-identical sample positions to the oracle for matrix poses).
// EXACT = false: the same algebra with FMA contraction, MUFU reciprocal / exp / log and the closed form
//                d(depth) = -<du, t> / depth; differs from EXACT by a few ulp per quantity.
#ifndef VSL_FUSED_MIN_BLOCKS
#define VSL_FUSED_MIN_BLOCKS 3
#endif
template <int V, bool EXACT>
__global__ void __launch_bounds__(kThreads, VSL_FUSED_MIN_BLOCKS)
loss_fused_kernel(const LossParams P) {
  constexpr int N = NT<V>::value;
  using L = Smem<V>;
  extern __shared__ float4 smem4[];
  float* sm = reinterpret_cast<float*>(smem4);
  float (*qt)[kQStride] = reinterpret_cast<float (*)[kQStride]>(sm + L::qt);   // x or 1/x with a 2-pixel halo
  float (*sA)[kOStride] = reinterpret_cast<float (*)[kOStride]>(sm + L::sA);   // cxx * sign(dx2) per owner
  float (*sB)[kOStride] = reinterpret_cast<float (*)[kOStride]>(sm + L::sB);   // cyy * sign(dy2)
  float (*sC)[kOStride] = reinterpret_cast<float (*)[kOStride]>(sm + L::sC);   // cxy*sign(dxdy) + cyx*sign(dydx)
  float* sxf = sm + L::xf;                                                     // per view: kinv[9], p[12]
  float* scratch = sm + L::scratch;
  float* s_tgt = sm + L::tgt;
  float* s_lg = sm + L::lg;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- which tile: grid = (tiles per image over all scales, B); no integer division on the way
  const int t_img = blockIdx.x, b = blockIdx.y;
  int s = 0;
  while (s + 1 < P.S && t_img >= P.img_begin[s + 1]) ++s;
  const int rem = t_img - P.img_begin[s];
  const int band = __float2int_rz(((float)rem + 0.5f) * P.inv_tx[s]), tx = rem - band * P.tiles_x[s];
  const int tile = P.item_begin[s] + b * (P.img_begin[s + 1] - P.img_begin[s]) + rem;  // partial slot
  const int H = P.H >> s, W = P.W >> s, R = P.R[s];
  const int y_base = band * kTH, x_base = tx * 32 * R;
  const int cols = min(32 * R, W - x_base);            // valid columns of this tile
  const int tw = ((cols + 31) >> 5) << 5;               // rounded up to whole warp iterations
  const size_t img_off = (size_t)b * H * W;
  const int y = y_base + warp;                          // this warp's row
  const bool staged = P.staged[s] != 0;
  const bool use_lg = P.mask_mode == VSL_MASK_EXP;

  // ---- 1. streaming operands: target and logits rows of the tile -> shared memory, asynchronously
  if (staged && y < H) {
    const size_t row = img_off + (size_t)y * W + x_base;
    const float* __restrict__ gt = P.tgt[s] + row * 3;
    float* dt = s_tgt + warp * (kTW * 3);
    for (int k = lane; k < (cols * 3) >> 2; k += 32) cp_async16(dt + 4 * k, gt + 4 * k);
    if (use_lg) {
      const float* __restrict__ gl = P.logits[s] + row * (2 * V);
      float* dl = s_lg + warp * (kTW * 2 * V);
      for (int k = lane; k < (cols * 2 * V) >> 2; k += 32) cp_async16(dl + 4 * k, gl + 4 * k);
    }
  }

  // ---- 2. the x tile (+halo) and this image's V transforms
  const float* __restrict__ xs = P.x[s] + img_off;
  for (int ty = warp; ty < kTileH; ty += kTH) {
    const int gy = y_base + ty - kHalo;
    const bool rin = gy >= 0 && gy < H;
    const float* __restrict__ row = xs + (size_t)(rin ? gy : 0) * W;
    for (int tc = lane; tc < tw + 2 * kHalo; tc += 32) {
      const int gx = x_base + tc - kHalo;
      qt[ty][tc] = (rin && gx >= 0 && gx < W) ? row[gx] : 0.f;
    }
  }
  if (P.smooth_on_inverse) {  // each thread revisits exactly the elements it wrote
    for (int ty = warp; ty < kTileH; ty += kTH) {
      const int gy = y_base + ty - kHalo;
      for (int tc = lane; tc < tw + 2 * kHalo; tc += 32) {
        const int gx = x_base + tc - kHalo;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) qt[ty][tc] = __fdiv_rn(1.0f, qt[ty][tc]);
      }
    }
  }
  if (threadIdx.x < V * 21) {  // per view: K^-1 rows padded to float4 (12 floats), then the 3 rows of P (12 floats)
    const int v = threadIdx.x / 21, k = threadIdx.x - v * 21;
    const int dst = k < 9 ? (k / 3) * 4 + k % 3 : 12 + (k - 9);
    sxf[v * 24 + dst] = reinterpret_cast<const float*>(P.xf + ((size_t)s * V + v) * P.B + b)[k];
  }
  __syncthreads();

  const float cpix = P.cpix[s], cexp = P.cexp[s];
  float pix_sum = 0.f, exp_sum = 0.f, sm_sum = 0.f;

  // ---- 3. smoothness, pass 1: every element of (tile + 2 rows above + 2 columns left) evaluates the four
  // second differences it owns (it is their top-left corner) ONCE and publishes their weighted signs.
  {
    const float cxx = P.csm[s][0], cxy = P.csm[s][1], cyx = P.csm[s][2], cyy = P.csm[s][3];
    for (int oy = warp; oy < kOwnH; oy += kTH) {
      const int gy = y_base - kHalo + oy;
      const bool rin = gy >= 0 && gy < H;
      for (int ox = lane; ox < tw + kHalo; ox += 32) {
        const int gx = x_base - kHalo + ox;
        const bool in = rin && gx >= 0 && gx < W;
        const float q00 = qt[oy][ox], q01 = qt[oy][ox + 1], q02 = qt[oy][ox + 2];
        const float q10 = qt[oy + 1][ox], q11 = qt[oy + 1][ox + 1], q20 = qt[oy + 2][ox];
        const float dx0 = __fsub_rn(q01, q00), dy0 = __fsub_rn(q10, q00);
        float dxx = __fsub_rn(__fsub_rn(q02, q01), dx0);
        float dyy = __fsub_rn(__fsub_rn(q20, q10), dy0);
        float dxy = __fsub_rn(__fsub_rn(q11, q10), dx0);  // d/dy of dx
        float dyx = __fsub_rn(__fsub_rn(q11, q01), dy0);  // d/dx of dy
        if (!(in && gx + 2 < W)) dxx = 0.f;
        if (!(in && gy + 2 < H)) dyy = 0.f;
        if (!(in && gx + 1 < W && gy + 1 < H)) { dxy = 0.f; dyx = 0.f; }
        sA[oy][ox] = signed_by(cxx, dxx);
        sB[oy][ox] = signed_by(cyy, dyy);
        sC[oy][ox] = signed_by(cxy, dxy) + signed_by(cyx, dyx);
        if (oy >= kHalo && ox >= kHalo)
          sm_sum += cxx * fabsf(dxx) + cyy * fabsf(dyy) + cxy * fabsf(dxy) + cyx * fabsf(dyx);
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();

  // ---- 4. the pixels
  const float gy = grid_coord(y, H, P.hstep[s]);
  const float wstep = P.wstep[s];
  float S1[V][3], S3[V][3], S4[V][3];  // sum du*d*gx, du*d, du   (sum du*d*gy = gy * S3: the row is fixed)
#pragma unroll
  for (int v = 0; v < V; ++v)
#pragma unroll
    for (int i = 0; i < 3; ++i) { S1[v][i] = 0.f; S3[v][i] = 0.f; S4[v][i] = 0.f; }

  if (y < H) {
    for (int r = 0; r * 32 < tw; ++r) {
      const int xl = r * 32 + lane;                // column inside the tile
      const int x = x_base + xl;
      if (x >= W) continue;
      const int pofs = y * W + x;                  // pixel offset inside this image
      const size_t pix = img_off + pofs;
      const int oy = warp + kHalo, ox = xl + kHalo;

      // smoothness, pass 2: gradient = the published signs of the 10 stencils this element is part of
      const float g_q = (sA[oy][ox] - 2.f * sA[oy][ox - 1] + sA[oy][ox - 2]) +
                        (sB[oy][ox] - 2.f * sB[oy - 1][ox] + sB[oy - 2][ox]) +
                        (sC[oy][ox] - sC[oy][ox - 1] - sC[oy - 1][ox] + sC[oy - 1][ox - 1]);

      // depth of this pixel and d(depth)/dx, d(q)/dx
      const float qc = qt[oy][ox];
      float d, dd_dx, dq_dx;
      if (P.smooth_on_inverse) {
        dq_dx = -qc * qc;
        if (P.depth_is_inverse) { d = qc; dd_dx = dq_dx; }
        else { d = xs[pofs]; dd_dx = 1.f; }
      } else {
        dq_dx = 1.f;
        if (P.depth_is_inverse) { d = EXACT ? __fdiv_rn(1.0f, qc) : __fdividef(1.0f, qc); dd_dx = -d * d; }
        else { d = qc; dd_dx = 1.f; }
      }

      const float gx = grid_coord(x, W, wstep);
      float tt[3];
      if (staged) {
        const float* t = s_tgt + (warp * kTW + xl) * 3;
        tt[0] = t[0]; tt[1] = t[1]; tt[2] = t[2];
      } else {
        const float* __restrict__ t = P.tgt[s] + pix * 3;
        tt[0] = t[0]; tt[1] = t[1]; tt[2] = t[2];
      }
      const float dgx = d * gx;
      // K^-1 is the same for every view of a scale
      float r0, r1, r2;
      {
        const float4 k0 = *reinterpret_cast<const float4*>(sxf), k1 = *reinterpret_cast<const float4*>(sxf + 4),
                     k2 = *reinterpret_cast<const float4*>(sxf + 8);
        if (EXACT) {
          const float kk[9] = {k0.x, k0.y, k0.z, k1.x, k1.y, k1.z, k2.x, k2.y, k2.z};
          Ray ray = back_project(kk, gx, gy);
          r0 = ray.r0; r1 = ray.r1; r2 = ray.r2;
        } else {
          r0 = fmaf(k0.x, gx, fmaf(k0.y, gy, k0.z));
          r1 = fmaf(k1.x, gx, fmaf(k1.y, gy, k1.z));
          r2 = fmaf(k2.x, gx, fmaf(k2.y, gy, k2.z));
        }
      }
      const float c0 = __fmul_rn(r0, d), c1 = __fmul_rn(r1, d), c2 = __fmul_rn(r2, d);
      float g_d = 0.f;

#pragma unroll
      for (int v = 0; v < V; ++v) {
#ifdef VSL_XF_VEC
        const float4 P0 = *reinterpret_cast<const float4*>(sxf + v * 24 + 12),
                     P1 = *reinterpret_cast<const float4*>(sxf + v * 24 + 16),
                     P2 = *reinterpret_cast<const float4*>(sxf + v * 24 + 20);
        const float pp[12] = {P0.x, P0.y, P0.z, P0.w, P1.x, P1.y, P1.z, P1.w, P2.x, P2.y, P2.z, P2.w};
#else
        const float* __restrict__ pp = sxf + v * 24 + 12;
#endif
        float qx, qy, rz;
        if (EXACT) {
          Proj q = project(pp, c0, c1, c2);
          qx = q.x; qy = q.y; rz = 1.0f / q.zp;
        } else {
          const float u0 = fmaf(pp[0], c0, fmaf(pp[1], c1, fmaf(pp[2], c2, pp[3])));
          const float u1 = fmaf(pp[4], c0, fmaf(pp[5], c1, fmaf(pp[6], c2, pp[7])));
          const float u2 = fmaf(pp[8], c0, fmaf(pp[9], c1, fmaf(pp[10], c2, pp[11])));
          rz = __fdividef(1.0f, u2 + kEpsZ);
          qx = u0 * rz; qy = u1 * rz;
        }
        const Foot f = footprint(qx, qy, W, H);
        const float* __restrict__ p00 = P.src[v][s] + (img_off + (size_t)(f.y0 * W + f.x0)) * 3;
        const int dxo = (f.x1 - f.x0) * 3, dyo = (f.y1 - f.y0) * W * 3;
        const float* __restrict__ p10 = p00 + dxo;
        const float* __restrict__ p01 = p00 + dyo;
        const float* __restrict__ p11 = p01 + dxo;
        float i00[3], i01[3], i10[3], i11[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          i00[c] = __ldg(p00 + c); i10[c] = __ldg(p10 + c); i01[c] = __ldg(p01 + c); i11[c] = __ldg(p11 + c);
        }
        // mask value m (explainability softmax or constant) and the regulariser
        float m = 1.f, p0 = 0.f, p1 = 0.f;
        float* lgp = s_lg + (warp * kTW + xl) * (2 * V) + 2 * v;   // staged logits slot, reused for d/dlogits
        if (use_lg) {
          const float2 lg = staged ? *reinterpret_cast<const float2*>(lgp)
                                   : *reinterpret_cast<const float2*>(P.logits[s] + pix * (2 * V) + 2 * v);
          if (EXACT) {
            const float mx = fmaxf(lg.x, lg.y);
            const float e0 = expf(lg.x - mx), e1 = expf(lg.y - mx), se = e0 + e1;
            p0 = e0 / se; p1 = e1 / se;
            exp_sum += (mx + logf(se)) - lg.y;
          } else {
            const float z = lg.x - lg.y;
            const float t = __expf(-fabsf(z)), se = 1.f + t, big = __fdividef(1.f, se), small = t * big;
            p0 = z >= 0.f ? big : small;
            p1 = z >= 0.f ? small : big;
            exp_sum += __logf(se) + fmaxf(z, 0.f);
          }
          m = p1;
        } else if (P.mask_mode == VSL_MASK_CONST) {
          m = P.mask[s][pix];
        }
        const float w00 = __fmul_rn(f.wx0, f.wy0), w01 = __fmul_rn(f.wx0, f.wy1),
                    w10 = __fmul_rn(f.wx1, f.wy0), w11 = __fmul_rn(f.wx1, f.wy1);
        // E = sum_c |e_c|; J_k = sum_c sign(e_c) * corner_k[c]  (the channel sum commutes with d/dx, d/dy)
        float E = 0.f, J00 = 0.f, J01 = 0.f, J10 = 0.f, J11 = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float wv = EXACT ? blend(w00, w01, w10, w11, i00[c], i01[c], i10[c], i11[c])
                                 : fmaf(w11, i11[c], fmaf(w10, i10[c], fmaf(w01, i01[c], w00 * i00[c])));
          const float e = wv - tt[c];
          E += fabsf(e);
          const float sg = signed_by(1.f, e);
          J00 = fmaf(sg, i00[c], J00); J01 = fmaf(sg, i01[c], J01);
          J10 = fmaf(sg, i10[c], J10); J11 = fmaf(sg, i11[c], J11);
        }
        const float ex0 = f.mx1 * J10 - f.mx0 * J00, ex1 = f.mx1 * J11 - f.mx0 * J01;  // d/dx per row
        const float ey0 = f.my1 * J01 - f.my0 * J00, ey1 = f.my1 * J11 - f.my0 * J10;  // d/dy per column
        const float dx = f.wy0 * ex0 + f.wy1 * ex1;
        const float dy = f.wx0 * ey0 + f.wx1 * ey1;
        pix_sum = fmaf(m, E, pix_sum);
        if (use_lg) {
          const float g0 = p0 * (cexp - cpix * E * p1);
          if (staged) *reinterpret_cast<float2*>(lgp) = make_float2(g0, -g0);
          else *reinterpret_cast<float2*>(P.g_logits[s] + pix * (2 * V) + 2 * v) = make_float2(g0, -g0);
        }
        const float k = cpix * m * rz;
        const float du0 = dx * k, du1 = dy * k, du2 = -(qx * du0 + qy * du1);
        if (EXACT) {
          const float gc0 = du0 * pp[0] + du1 * pp[4] + du2 * pp[8];
          const float gc1 = du0 * pp[1] + du1 * pp[5] + du2 * pp[9];
          const float gc2 = du0 * pp[2] + du1 * pp[6] + du2 * pp[10];
          g_d += gc0 * r0 + gc1 * r1 + gc2 * r2;
        } else {
          g_d -= du0 * pp[3] + du1 * pp[7] + du2 * pp[11];  // <du, M ray> = <du, u - t> / d and <du, u> = 0
        }
        S1[v][0] = fmaf(du0, dgx, S1[v][0]); S1[v][1] = fmaf(du1, dgx, S1[v][1]); S1[v][2] = fmaf(du2, dgx, S1[v][2]);
        S3[v][0] = fmaf(du0, d, S3[v][0]);   S3[v][1] = fmaf(du1, d, S3[v][1]);   S3[v][2] = fmaf(du2, d, S3[v][2]);
        S4[v][0] += du0;                     S4[v][1] += du1;                     S4[v][2] += du2;
      }
      if (!EXACT) g_d = __fdividef(g_d, d);
      P.g_x[s][pix] = g_d * dd_dx + g_q * dq_dx;
    }
  }

  // ---- 5. d/dlogits leaves the tile with 16-byte stores (each warp wrote its own row: no block barrier)
  if (staged && use_lg && y < H) {
    __syncwarp();
    float* __restrict__ gl = P.g_logits[s] + (img_off + (size_t)y * W + x_base) * (2 * V);
    const float* sl = s_lg + warp * (kTW * 2 * V);
    for (int k = lane; k < (cols * 2 * V) >> 2; k += 32)
      *reinterpret_cast<float4*>(gl + 4 * k) = *reinterpret_cast<const float4*>(sl + 4 * k);
  }

  // ---- 6. one block reduction per tile: 3 loss sums + per view (sum du d gx, gy sum du d, sum du d, sum du)
  float vals[N];
  vals[0] = pix_sum * cpix; vals[1] = sm_sum; vals[2] = exp_sum * cexp;
#pragma unroll
  for (int v = 0; v < V; ++v)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      vals[3 + v * 12 + i] = S1[v][i];
      vals[3 + v * 12 + 3 + i] = gy * S3[v][i];
      vals[3 + v * 12 + 6 + i] = S3[v][i];
      vals[3 + v * 12 + 9 + i] = S4[v][i];
    }
  // transposed reduction through shared memory (the tile buffers are dead by now): thread j sums the 32 lanes of
  // one (value, warp) pair with a per-thread rotation that keeps every access bank-conflict free, then N
  // threads add the 8 warp totals in a fixed order.  ~3x fewer instructions than shuffles for N >= 15.
  static_assert(L::total >= N * kThreads + N * kTH, "reduction buffers must fit in the tile buffers");
  __syncthreads();
  float* red = sm;
  float* wsum = sm + N * kThreads;  // [N][kTH] warp totals
#pragma unroll
  for (int k = 0; k < N; ++k) red[k * kThreads + threadIdx.x] = vals[k];
  __syncthreads();
  for (int j = threadIdx.x; j < N * kTH; j += kThreads) {
    const float4* src = reinterpret_cast<const float4*>(red + j * 32);   // row (value j >> 3, warp j & 7)
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 q = src[(i + j) & 7];   // rotation: the 8 threads of a quarter-warp hit 8 different banks
      a += (q.x + q.y) + (q.z + q.w);
    }
    wsum[j] = a;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kTH; ++w) t += wsum[threadIdx.x * kTH + w];
    P.partials[(size_t)tile * N + threadIdx.x] = t;
  }
}

// grid = B + 1 blocks of 1024 threads.  Block b < B: pose gradients of batch element b (all views); one WARP per
// (scale, view, component) sums that image's partial slots (lane-strided, then a fixed shuffle tree).
// Block B: the three loss scalars.  Every sum runs in a fixed order in double => deterministic.
template <int V>
__global__ void __launch_bounds__(1024)
loss_finalize_kernel(const LossParams P, const float* __restrict__ poses, const float* __restrict__ K_pyr,
                     int pose_format, float inv_loss_scale, float* __restrict__ losses, float* __restrict__ g_poses) {
  constexpr int N = NT<V>::value;
  __shared__ double sh[32 * 3];
  __shared__ double tsum[VSL_MAX_SCALES * V * 12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  auto warp_dsum = [](double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  const int n_items = P.item_begin[P.S];
  if ((int)blockIdx.x == P.B) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
      const float* p = P.partials + (size_t)i * N;
      a0 += (double)p[0]; a1 += (double)p[1]; a2 += (double)p[2];
    }
    a0 = warp_dsum(a0); a1 = warp_dsum(a1); a2 = warp_dsum(a2);
    if (lane == 0) { sh[warp * 3] = a0; sh[warp * 3 + 1] = a1; sh[warp * 3 + 2] = a2; }
    __syncthreads();
    if (threadIdx.x < 3) {
      double t = 0.0;
      for (int w = 0; w < nwarp; ++w) t += sh[w * 3 + threadIdx.x];
      losses[threadIdx.x] = (float)(t * (double)inv_loss_scale);
    }
    return;
  }
  const int b = blockIdx.x;
  // loads of every round first (independent, all in flight), shuffle trees afterwards
  constexpr int kRounds = (VSL_MAX_SCALES * V * 12 + 31) / 32;
  double acc[kRounds];
#pragma unroll
  for (int r = 0; r < kRounds; ++r) {
    const int e = warp + r * nwarp;
    acc[r] = 0.0;
    if (e < P.S * V * 12) {
      const int k = e % 12, v = (e / 12) % V, s = e / (12 * V);
      const int per_b = P.bands[s] * P.tiles_x[s];
      const float* p = P.partials + (size_t)(P.item_begin[s] + b * per_b) * N + 3 + v * 12 + k;
      for (int i = lane; i < per_b; i += 32) acc[r] += (double)p[(size_t)i * N];
    }
  }
#pragma unroll
  for (int r = 0; r < kRounds; ++r) {
    const int e = warp + r * nwarp;
    const double a = warp_dsum(acc[r]);
    if (lane == 0 && e < P.S * V * 12) tsum[e] = a;
  }
  __syncthreads();
  // dT[v][k][j] = sum_s sum_i K_s[i][k] * dP_s[i][j]: one thread per matrix element, short chains
  __shared__ double sgT[VSL_MAX_VIEWS][16];
  if ((int)threadIdx.x < V * 16) {
    const int v = threadIdx.x >> 4, k = (threadIdx.x >> 2) & 3, j = threadIdx.x & 3;
    double acc = 0.0;
    if (k < 3) {
      for (int s = 0; s < P.S; ++s) {
        const double* t = tsum + (s * V + v) * 12;
        const Xform& xf = P.xf[((size_t)s * V + v) * P.B + b];
        const float* Ks = K_pyr + ((size_t)b * P.S + s) * 9;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double dPij = (j < 3) ? (double)xf.kinv[j * 3] * t[i] + (double)xf.kinv[j * 3 + 1] * t[3 + i] +
                                            (double)xf.kinv[j * 3 + 2] * t[6 + i]
                                      : t[9 + i];
          acc += (double)Ks[i * 3 + k] * dPij;
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

}  // namespace vsl

using namespace vsl;

namespace {

struct WsLayout {
  size_t xf, partials, pyr, total;
  size_t level_off[VSL_MAX_SCALES];  // offset (in floats) of level s inside one image's pyramid block
  size_t pyr_img;                    // floats per image pyramid (levels 1..S-1)
  int n_items, item_begin[VSL_MAX_SCALES + 1], tiles_x[VSL_MAX_SCALES], bands[VSL_MAX_SCALES], R[VSL_MAX_SCALES];
};

int check_desc(const VslLossDesc* d) {
  VSL_REQUIRE(d, VSL_E_NULL);
  VSL_REQUIRE(d->S >= 1 && d->S <= VSL_MAX_SCALES && d->V >= 1 && d->V <= VSL_MAX_VIEWS, VSL_E_SHAPE);
  VSL_REQUIRE(d->B > 0 && d->B <= 65535 / (VSL_MAX_VIEWS + 1) && d->H > 0 && d->W > 0, VSL_E_SHAPE);
  const int F = 1 << (d->S - 1);
  VSL_REQUIRE(d->H % F == 0 && d->W % F == 0 && (d->H >> (d->S - 1)) >= 3 && (d->W >> (d->S - 1)) >= 3, VSL_E_SHAPE);
  VSL_REQUIRE(d->pose_format >= VSL_POSE_EULER && d->pose_format <= VSL_POSE_MATRIX, VSL_E_FORMAT);
  VSL_REQUIRE(d->mask_mode >= VSL_MASK_NONE && d->mask_mode <= VSL_MASK_CONST, VSL_E_FORMAT);
  return VSL_OK;
}

void layout(const VslLossDesc* d, WsLayout* L) {
  const int nt = 3 + 12 * d->V;
  int n = 0;
  size_t lv = 0;
  for (int s = 0; s < d->S; ++s) {
    const int H = d->H >> s, W = d->W >> s;
    int R = (W + 31) / 32;
    R = R > kMaxR ? kMaxR : R;
    L->R[s] = R;
    L->tiles_x[s] = (W + 32 * R - 1) / (32 * R);
    L->bands[s] = (H + kTH - 1) / kTH;
    L->item_begin[s] = n;
    n += d->B * L->bands[s] * L->tiles_x[s];
    L->level_off[s] = lv;
    if (s >= 1) lv += round_up((size_t)d->B * H * W * 3, 4);  // keep every level 16-byte aligned
  }
  L->item_begin[d->S] = n;
  L->n_items = n;
  L->pyr_img = lv;
  L->xf = 0;
  L->partials = round_up(sizeof(Xform) * (size_t)d->S * d->V * d->B, 256);
  L->pyr = L->partials + round_up(sizeof(float) * (size_t)n * nt, 256);
  L->total = L->pyr + sizeof(float) * lv * (size_t)(d->V + 1);
}

template <int V, bool EXACT>
int launch_fused(const WsLayout& L, const LossParams& P, cudaStream_t st) {
  // > 48 KB of dynamic shared memory needs the opt-in; idempotent and cheap, so set on every call (no state)
  cudaError_t e = cudaFuncSetAttribute(loss_fused_kernel<V, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Smem<V>::bytes);
  if (e != cudaSuccess) return (int)e;
  loss_fused_kernel<V, EXACT><<<dim3(P.img_begin[P.S], P.B), kThreads, Smem<V>::bytes, st>>>(P);
  return VSL_OK;
}

template <int V>
int run_loss(const VslLossDesc* d, const WsLayout& L, LossParams& P, const float* poses, const float* K_pyr,
             float* losses, float* g_poses, cudaStream_t st) {
  if (d->ev_main_begin != nullptr) cudaEventRecord((cudaEvent_t)d->ev_main_begin, st);
  const int rc = d->exact_coords ? launch_fused<V, true>(L, P, st) : launch_fused<V, false>(L, P, st);
  if (rc != VSL_OK) return rc;
  if (d->ev_main_end != nullptr) cudaEventRecord((cudaEvent_t)d->ev_main_end, st);
  loss_finalize_kernel<V><<<d->B + 1, 1024, 0, st>>>(P, poses, K_pyr, d->pose_format, 1.0f / d->loss_scale,
                                                       losses, g_poses);
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

int vsl_loss_fwd_bwd(const VslLossDesc* d, const float* tgt, const float* const* srcs, const float* const* x_pyr,
                     const float* poses, const float* K_pyr, const float* const* logits_pyr,
                     const float* const* mask_pyr, float* losses, float* const* g_x_pyr, float* g_poses,
                     float* const* g_logits_pyr, void* ws, vsl_stream_t stream) {
  int rc = check_desc(d);
  if (rc != VSL_OK) return rc;
  VSL_REQUIRE(tgt && srcs && x_pyr && poses && K_pyr && losses && g_x_pyr && g_poses && ws, VSL_E_NULL);
  VSL_REQUIRE(d->mask_mode != VSL_MASK_EXP || (logits_pyr && g_logits_pyr), VSL_E_NULL);
  VSL_REQUIRE(d->mask_mode != VSL_MASK_CONST || mask_pyr, VSL_E_NULL);
  VSL_REQUIRE(d->loss_scale != 0.f, VSL_E_UNSUPPORTED);
  VSL_REQUIRE(aligned(ws, 256), VSL_E_ALIGN);
  cudaStream_t st = (cudaStream_t)stream;
  WsLayout L;
  layout(d, &L);
  char* base = reinterpret_cast<char*>(ws);
  Xform* xf = reinterpret_cast<Xform*>(base + L.xf);
  float* pyr = reinterpret_cast<float*>(base + L.pyr);

  LossParams P;
  P.B = d->B; P.H = d->H; P.W = d->W; P.S = d->S; P.V = d->V;
  P.mask_mode = d->mask_mode; P.depth_is_inverse = d->depth_is_inverse; P.smooth_on_inverse = d->smooth_on_inverse;
  P.xf = xf;
  P.partials = reinterpret_cast<float*>(base + L.partials);
  for (int s = 0; s <= d->S; ++s) P.item_begin[s] = L.item_begin[s];
  P.img_begin[0] = 0;
  for (int s = 0; s < d->S; ++s) P.img_begin[s + 1] = P.img_begin[s] + L.bands[s] * L.tiles_x[s];
  for (int v = 0; v < d->V; ++v) VSL_REQUIRE(srcs[v], VSL_E_NULL);
  for (int s = 0; s < VSL_MAX_SCALES; ++s) {
    P.tgt[s] = nullptr; P.x[s] = nullptr; P.logits[s] = nullptr; P.mask[s] = nullptr;
    P.g_x[s] = nullptr; P.g_logits[s] = nullptr; P.staged[s] = 0;
    for (int v = 0; v < VSL_MAX_VIEWS; ++v) P.src[v][s] = nullptr;
  }
  for (int s = 0; s < d->S; ++s) {
    const int H = d->H >> s, W = d->W >> s;
    VSL_REQUIRE(x_pyr[s] && g_x_pyr[s], VSL_E_NULL);
    P.x[s] = x_pyr[s];
    P.g_x[s] = g_x_pyr[s];
    if (d->mask_mode == VSL_MASK_EXP) {
      VSL_REQUIRE(logits_pyr[s] && g_logits_pyr[s], VSL_E_NULL);
      VSL_REQUIRE(aligned(logits_pyr[s], 8) && aligned(g_logits_pyr[s], 8), VSL_E_ALIGN);
      P.logits[s] = logits_pyr[s];
      P.g_logits[s] = g_logits_pyr[s];
    }
    if (d->mask_mode == VSL_MASK_CONST) {
      VSL_REQUIRE(mask_pyr[s], VSL_E_NULL);
      P.mask[s] = mask_pyr[s];
    }
    P.tgt[s] = (s == 0) ? tgt : pyr + L.level_off[s];
    for (int v = 0; v < d->V; ++v) P.src[v][s] = (s == 0) ? srcs[v] : pyr + L.pyr_img * (size_t)(v + 1) + L.level_off[s];
    P.tiles_x[s] = L.tiles_x[s]; P.bands[s] = L.bands[s]; P.R[s] = L.R[s];
    P.inv_tx[s] = 1.0f / (float)L.tiles_x[s];
    P.wstep[s] = 2.0f / (float)(W - 1);  // fp32 division, as grid_step() does on the device
    P.hstep[s] = 2.0f / (float)(H - 1);
    // 16-byte row alignment of the streamed operands: W % 4 == 0 makes every row start (and the tile's
    // x_base, a multiple of 32) a multiple of 4 pixels = 48 B of target / 8V*4 B of logits
    P.staged[s] = (W % 4 == 0) && aligned(P.tgt[s], 16) &&
                  (d->mask_mode != VSL_MASK_EXP || (aligned(P.logits[s], 16) && aligned(P.g_logits[s], 16)));
    const double npx = (double)d->B * H * W;
    const double dw = d->pixel_scale_norm ? (double)d->data_weight / (double)(1 << s) : (double)d->data_weight;
    P.cpix[s] = (float)((double)d->loss_scale * dw / (npx * 3.0));
    P.cexp[s] = (float)((double)d->loss_scale * (double)d->explain_reg_weight / npx);
    const double sw = (double)d->loss_scale * (double)d->smooth_weight / (double)(1 << s);
    P.csm[s][0] = (float)(sw / ((double)d->B * H * (W - 2)));
    P.csm[s][1] = (float)(sw / ((double)d->B * (H - 1) * (W - 1)));
    P.csm[s][2] = P.csm[s][1];
    P.csm[s][3] = (float)(sw / ((double)d->B * (H - 2) * W));
  }

  // 1. transforms and the image pyramids (target + V sources) in one launch
  const PrepJob prep = make_prep(poses, K_pyr, d->B, d->S, d->V, d->pose_format, xf, nullptr);
  if (d->S > 1) {
    PyrJob job;
    job.nimg = d->V + 1;
    for (int i = 0; i <= d->V; ++i) {
      job.img[i] = (i == 0) ? tgt : srcs[i - 1];
      for (int s = 0; s < VSL_MAX_SCALES; ++s)
        job.lvl[i][s] = (s >= 1 && s < d->S) ? pyr + L.pyr_img * (size_t)i + L.level_off[s] : nullptr;
    }
    rc = launch_pyramid(job, prep, d->B, d->H, d->W, 3, d->S, st);
    if (rc != VSL_OK) return rc;
  } else {
    prep_xforms_kernel<<<(prep.n + 63) / 64, 64, 0, st>>>(prep);
  }
  // 2 + 3. fused loss and finalize
  switch (d->V) {
    case 1: return run_loss<1>(d, L, P, poses, K_pyr, losses, g_poses, st);
    case 2: return run_loss<2>(d, L, P, poses, K_pyr, losses, g_poses, st);
    case 3: return run_loss<3>(d, L, P, poses, K_pyr, losses, g_poses, st);
    default: return run_loss<4>(d, L, P, poses, K_pyr, losses, g_poses, st);
  }
}

}  // extern "C"
