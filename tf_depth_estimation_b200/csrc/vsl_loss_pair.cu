// libvsl: the fused loss kernel for an EVEN number of source views in the fast arithmetic mode -- the kernel the
// BASELINE configuration (two source views, train.py:107-135) runs.  Same tiles, same pipeline and same results
// (to a few ulp) as loss_fused_kernel<V, false, false> in vsl_loss.cu; what differs is how the per-view arithmetic
// is issued:
//
//   * Views are processed in PAIRS with sm_100's packed fp32 instructions (fma/mul/add.f32x2 -> FFMA2 / FMUL2 /
//     FADD2): one issue slot does the same operation for view 2p and view 2p+1.  Everything a view computes that is
//     not a raw gathered texel -- projection, footprint, bilinear weights, error, sign, d/dx d/dy, du, the softmax
//     gradients and the 9 dP accumulators -- lives in 64-bit register pairs (lo = even view, hi = odd view).  Only
//     the 12 blend and 12 J multiply-adds per view read the LDG.128 destination registers and stay scalar (a pair
//     needs two ADJACENT registers, and a texel of view 0 never sits next to one of view 1); they write their
//     results straight into the halves of a pair, so no packing moves are needed.  Scalars shared by both views
//     (depth, grid row, target, weights) enter as broadcast operands (R.F32), per-view constants as uniform
//     register pairs.
//   * Projection folded per image: u = d * (Q [gx, gy, 1]) + t with Q = P[:, :3] K^-1 (XformQ table written by the
//     prep launch).  The column is fixed per lane, so Q [gx, ., 1] is a lane constant and a row costs
//     2 x 3 packed FMAs per pair instead of 3 + 9 per view.
//   * Smoothness: image-border conditions are data, not predicates: a lane carries its weights (0 where the
//     second difference it owns does not exist), a row its uniform ones; the grid row coordinate comes from a
//     32-entry table the warp fills once per tile.
//   * |e| = sign(e) * e, the bilinear weights from one product (w11, then differences), d/dx and d/dy in lerp form.
#include "vsl_loss_common.cuh"

namespace vsl {

// ---- packed fp32 pairs
#ifndef VSL_NO_F32X2
struct f2 { unsigned long long v; };
VSL_DEV f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
VSL_DEV float lo(f2 a) { float x; asm("{\n\t.reg .f32 t;\n\tmov.b64 {%0, t}, %1;\n\t}" : "=f"(x) : "l"(a.v)); return x; }
VSL_DEV float hi(f2 a) { float y; asm("{\n\t.reg .f32 t;\n\tmov.b64 {t, %0}, %1;\n\t}" : "=f"(y) : "l"(a.v)); return y; }
VSL_DEV f2 bc(float a) { return pk(a, a); }
VSL_DEV f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
VSL_DEV f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
VSL_DEV f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
VSL_DEV f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
VSL_DEV f2 add2_rm(f2 a, f2 b) { f2 r; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }

#else   // the same algorithm with scalar instructions (timing experiment: what do the packed instructions buy?)
struct f2 { float x, y; };
VSL_DEV f2 pk(float lo, float hi) { f2 r; r.x = lo; r.y = hi; return r; }
VSL_DEV float lo(f2 a) { return a.x; }
VSL_DEV float hi(f2 a) { return a.y; }
VSL_DEV f2 bc(float a) { return pk(a, a); }
VSL_DEV f2 fma2(f2 a, f2 b, f2 c) { return pk(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
VSL_DEV f2 mul2(f2 a, f2 b) { return pk(a.x * b.x, a.y * b.y); }
VSL_DEV f2 add2(f2 a, f2 b) { return pk(a.x + b.x, a.y + b.y); }
VSL_DEV f2 sub2(f2 a, f2 b) { return pk(a.x - b.x, a.y - b.y); }
VSL_DEV f2 add2_rm(f2 a, f2 b) { return pk(__fadd_rd(a.x, b.x), __fadd_rd(a.y, b.y)); }
#endif

// (a, b) if z >= 0 else (b, a), as one compare and two selects (the C++ ternaries on pair halves compile to
// predicated moves, two per select)
VSL_DEV void swap_unless_ge0(float z, float a, float b, float& first, float& second) {
  asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %2, 0f00000000;\n\tselp.f32 %0, %3, %4, p;\n\tselp.f32 %1, %4, %3, p;\n\t}"
      : "=f"(first), "=f"(second) : "f"(z), "f"(a), "f"(b));
}

// What one view PAIR has in flight between issuing its gathers and consuming them.
struct TapPair {
  float4 A[2], B[2], C[2], D[2];   // corners (x0,y0) (x1,y0) (x0,y1) (x1,y1) of the even / odd view
  f2 wx1, wy1;                     // (x - x0), (y - y0)
  f2 qx, qy, rz;                   // projected coordinates (unclamped) and 1 / (z + eps)
};

constexpr int kPF = 3;   // rows ahead of their load that the streamed operands are prefetched into L2

template <int V> struct PairSmem {   // the general kernel's layout: the row table is WarpSmem<V, false>::rt
  static constexpr int gy = WarpSmem<V, false>::rt;
  static constexpr int total = WarpSmem<V, false>::total;
  static constexpr size_t block_bytes = WarpSmem<V, false>::block_bytes;
};

template <int V>
#ifdef VSL_FUSED_MAXNREG   // timing experiments: an explicit register cap instead of the one launch bounds imply
__global__ void __maxnreg__(VSL_FUSED_MAXNREG)
#else
__global__ void __launch_bounds__(kThreads, (V <= 2 ? VSL_FUSED_MIN_BLOCKS : VSL_FUSED_MIN_BLOCKS / 2))
#endif
loss_fused_pair_kernel(const LossParams P) {
  static_assert(V % 2 == 0, "views are processed in pairs");
  constexpr int NP = V / 2, N = NT<V>::value;
  using L = WarpSmem<V, false>;
  extern __shared__ float4 smem4[];
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int tile = blockIdx.x * kWarps + warp;
  if (tile >= P.item_begin[P.S]) return;   // warps are independent: no block barrier anywhere below
  float* wsm = reinterpret_cast<float*>(smem4) + warp * PairSmem<V>::total;
  float* qt = wsm + L::qt;
  float* sha = wsm + L::ha;
  float* shc = wsm + L::hc;
  float* gyt = wsm + PairSmem<V>::gy;
  // the same table through a per-thread address: a load the compiler believes to be warp-uniform is followed by a
  // register -> uniform-register move that waits for it; as an ordinary per-lane value the row constants just feed
  // the broadcast operands
  const float* gyv = reinterpret_cast<const float*>(smem4) + (threadIdx.x >> 5) * PairSmem<V>::total + PairSmem<V>::gy;

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
  const int pix0 = b * H * W;
  const bool use_lg = P.mask_mode == VSL_MASK_EXP;

  // ---- 1. the x tile (+halo), zero outside the image (identical to loss_fused_kernel)
  const float* __restrict__ xs = P.x[s] + pix0;
  {
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
        tp += 32 - kPairs;                       // 32 pairs on: one row further (18 pairs) and 14 more
        ++ty;
        if (tp >= kPairs) { tp -= kPairs; ++ty; }
      }
    } else {
      int ty = 0, tc = lane;
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
    // the grid row coordinates of the tile (fp32 linspace formula of utils.py:153-159), one per lane
    gyt[3 * lane] = grid_coord(y_base + lane, H, P.hstep[s]);
    gyt[3 * lane + 1] = y_base + lane < H - 2 ? P.cyy[s] : 0.f;
    gyt[3 * lane + 2] = y_base + lane < H - 1 ? 1.f : 0.f;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (P.x_is_logit != 0) {
      int ty = 0, tc = lane;
      for (int i = lane; i < kQH * kQS; i += 32) {
        const int gy = y_base - kHalo + ty, gx = x_base - kHalo + tc;
        if ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) {
          const float v = qt[i];
          const float sg = rcp_fast(1.0f + ex2_fast(-1.4426950408889634f * v));
          qt[i] = fmaf(P.disp_scale, sg, P.disp_min);
        }
        tc += 32;
        if (tc >= kQS) { tc -= kQS; ++ty; }
      }
      __syncwarp();
    }
    if (P.smooth_on_inverse != 0) {
      for (int i = lane; i < kQH * kQS; i += 32) {
        const float v = qt[i];
        qt[i] = v != 0.f ? __fdiv_rn(1.0f, v) : 0.f;
      }
    }
  }
  __syncwarp();

  // ---- 2. smoothness of the two columns LEFT of the tile (weighted signs the gradient of columns 0 and 1 needs)
  const float cxx = P.cxx[s], cxy = P.cxy[s], cyx = P.cyx[s], cyy = P.cyy[s];
  for (int i = lane; i < kOH * 2; i += 32) {
    const int oy = i >> 1, ox = i & 1;
    float a, bb, c, sm;
    owner_signs<false>(qt + oy * kQS + ox, (unsigned)(x_base - kHalo + ox), (unsigned)(y_base - kHalo + oy), H, W, cxx,
                       cxy, cyx, cyy, a, bb, c, sm);
    sha[i] = a;
    shc[i] = c;
  }
  __syncwarp();

  // ---- 3. the pixels
  const int xl = min(x, W - 1);
  const float gx = grid_coord(xl, W, P.wstep[s]);
  const float Wf = P.Wf[s], Hf = P.Hf[s];
  const int stride4 = P.stride4[s];
  const int coff = P.coff[s] + b * P.plane4[s];
  const float cpix = P.cpix[s], cexp = P.cexp[s];
  // a lane's smoothness weights: zero where the second difference owned by its column does not exist
  const float kxx_l = x < W - 2 ? cxx : 0.f;
  const float kxy_l = x < W - 1 ? cxy + cyx : 0.f;

  asm volatile("griddepcontrol.wait;" ::: "memory");   // from here on: what the prep launch wrote

  // This image's folded transforms: per pair and output row i, Q[i][1] and t[i] as uniform pairs (even, odd view);
  // Q[i][0] * gx + Q[i][2] as a lane constant.
  float Qy[V][3], Tt[V][3];
  f2 Al[NP][3];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const float* xq = reinterpret_cast<const float*>(P.xq + ((size_t)s * V + v) * P.B + b);
    float q[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) q[i] = __shfl_sync(0xffffffffu, __ldg(xq + i), 0);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      Qy[v][i] = q[i * 3 + 1];
      Tt[v][i] = q[9 + i];
      const float al = fmaf(q[i * 3], gx, q[i * 3 + 2]);
      if (v & 1) Al[v >> 1][i] = pk(lo(Al[v >> 1][i]), al);
      else Al[v >> 1][i] = pk(al, 0.f);
    }
  }

  f2 S2[NP][3], S3[NP][3], S4[NP][3];   // sum du*d*gy, du*d, du per pair (row 2 holds -du2: the sign is restored below)
#pragma unroll
  for (int p = 0; p < NP; ++p)
#pragma unroll
    for (int i = 0; i < 3; ++i) { S2[p][i] = bc(0.f); S3[p][i] = bc(0.f); S4[p][i] = bc(0.f); }
  f2 pix2 = bc(0.f), exp2 = bc(0.f);
  float sm_sum = 0.f;

  float b1, b2, c10, c11;
  {
    float a, bb, c, sm;
    owner_signs<false>(qt + lane + kHalo, (unsigned)x, (unsigned)(y_base - 2), H, W, cxx, cxy, cyx, cyy, a, bb, c, sm);
    b2 = bb;
    owner_signs<false>(qt + kQS + lane + kHalo, (unsigned)x, (unsigned)(y_base - 1), H, W, cxx, cxy, cyx, cyy, a, bb, c,
                       sm);
    b1 = bb;
    c10 = c;
    c11 = __shfl_up_sync(0xffffffffu, c, 1);
    if (lane == 0) c11 = shc[1 * 2 + 1];
  }

  const float* __restrict__ tgt_img = P.tgt[s];
  const float* __restrict__ lg_img = use_lg ? P.logits[s] : nullptr;
  float* __restrict__ glg_img = use_lg ? P.g_logits[s] : nullptr;
  const float* __restrict__ mk_img = P.mask_mode == VSL_MASK_CONST ? P.mask[s] : nullptr;
  float* __restrict__ gx_img = P.g_x[s];
  const bool lg4 = P.lg_vec4[s] != 0;
  const int smooth_inv = P.smooth_on_inverse, depth_inv = P.depth_is_inverse;

  struct Stream { float tt[3]; float lg[2 * V]; float mc; };
  struct Geo { float d, gy, dgy; };

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
  // One row of the streamed operands into L2, `ahead` rows before it is loaded: these are read exactly once, so
  // every load of them would otherwise pay the full DRAM latency at the top of the row that consumes it.
  auto prefetch_stream = [&](int pf, int row, int n_rows) {
    const float* ta = tgt_img + pf * 3;
    const float* la = use_lg ? lg_img + (size_t)pf * (2 * V) : (mk_img != nullptr ? mk_img + pf : ta);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %2, %3;\n\t@p prefetch.global.L2 [%0];\n\t@p prefetch.global.L2 [%1];\n\t}" ::"l"(ta),
                 "l"(la), "r"(row), "r"(n_rows));
  };
  auto make_geo = [&](Geo& g, int r) {
    const float qc = qt[(r + kHalo) * kQS + (xl - x_base) + kHalo];
    if (smooth_inv) g.d = depth_inv ? qc : rcp_fast(qc);   // the tile holds 1/x, the warp wants x
    else g.d = !depth_inv ? qc : rcp_fast(qc);
    g.gy = gyv[3 * r];
    g.dgy = g.d * g.gy;
  };

  // Projection, footprint and the 2 x 4 gathers of one view pair for the row described by g.
  auto tap_issue = [&](TapPair& t, int p, const Geo& g) {
    const f2 gyb = bc(g.gy), db = bc(g.d);
    const f2 u0 = fma2(db, fma2(gyb, pk(Qy[2 * p][0], Qy[2 * p + 1][0]), Al[p][0]), pk(Tt[2 * p][0], Tt[2 * p + 1][0]));
    const f2 u1 = fma2(db, fma2(gyb, pk(Qy[2 * p][1], Qy[2 * p + 1][1]), Al[p][1]), pk(Tt[2 * p][1], Tt[2 * p + 1][1]));
    const f2 u2 = fma2(db, fma2(gyb, pk(Qy[2 * p][2], Qy[2 * p + 1][2]), Al[p][2]), pk(Tt[2 * p][2], Tt[2 * p + 1][2]));
    const f2 z = add2(u2, bc(kEpsZ));
    t.rz = pk(rcp_fast(lo(z)), rcp_fast(hi(z)));
    t.qx = mul2(u0, t.rz);
    t.qy = mul2(u1, t.rz);
    // beyond [-2, size] all four corners are border zeros; clamping there changes neither value nor gradient
    const f2 xc = pk(fminf(fmaxf(lo(t.qx), -2.0f), Wf), fminf(fmaxf(hi(t.qx), -2.0f), Wf));
    const f2 yc = pk(fminf(fmaxf(lo(t.qy), -2.0f), Hf), fminf(fmaxf(hi(t.qy), -2.0f), Hf));
    const f2 tx = add2_rm(xc, bc(kMagic)), ty = add2_rm(yc, bc(kMagic));   // floor in the low mantissa bits
    t.wx1 = sub2(xc, sub2(tx, bc(kMagic)));
    t.wy1 = sub2(yc, sub2(ty, bc(kMagic)));
    {
      const int off = (int)(__float_as_uint(lo(ty)) * (unsigned)stride4 + __float_as_uint(lo(tx)) + (unsigned)coff);
#ifdef VSL_EXP_L1GATHER   // timing experiment (wrong results): every gather hits the same line
      const float4* __restrict__ gp = P.src[2 * p][s] + (off & 1);
#else
      const float4* __restrict__ gp = P.src[2 * p][s] + off;
#endif
      t.A[0] = __ldg(gp); t.B[0] = __ldg(gp + 1);
      t.C[0] = __ldg(gp + stride4); t.D[0] = __ldg(gp + stride4 + 1);
    }
    {
      const int off = (int)(__float_as_uint(hi(ty)) * (unsigned)stride4 + __float_as_uint(hi(tx)) + (unsigned)coff);
#ifdef VSL_EXP_L1GATHER
      const float4* __restrict__ gp = P.src[2 * p + 1][s] + (off & 1);
#else
      const float4* __restrict__ gp = P.src[2 * p + 1][s] + off;
#endif
      t.A[1] = __ldg(gp); t.B[1] = __ldg(gp + 1);
      t.C[1] = __ldg(gp + stride4); t.D[1] = __ldg(gp + stride4 + 1);
    }
  };

  struct Keep { f2 E, u0, u1, w2; };   // per pair: sum|e| and dL/du up to the factor cpix * m (w2 = -u2)
  TapPair tap[NP];
  int pofs = pix0 + y_base * W + xl;

  auto row = [&](Stream& cur, Stream& nxt, Geo& gc, Geo& gn, int r) {
    const bool has_next = r + 1 < rows;
    make_geo(gn, min(r + 1, rows - 1));   // unconditional (no branch): its LDS / MUFU latencies hide under the blend
    Keep keep[NP];
    // ---- phase 1: consume the landed gathers of each pair, re-issue them for the next row
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      TapPair& t = tap[p];
      // bilinear weights from one product: w11 = wx1 wy1, w10 = wx1 - w11, w01 = wy1 - w11, w00 = 1 - wx1 - w01
      const f2 w11 = mul2(t.wx1, t.wy1);
      const f2 w10 = sub2(t.wx1, w11), w01 = sub2(t.wy1, w11);
      const f2 w00 = sub2(sub2(bc(1.0f), t.wx1), w01);
      float wv[2][3], pad[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float a00 = h ? hi(w00) : lo(w00), a01 = h ? hi(w01) : lo(w01), a10 = h ? hi(w10) : lo(w10),
                    a11 = h ? hi(w11) : lo(w11);
        const float cA[3] = {t.A[h].x, t.A[h].y, t.A[h].z}, cB[3] = {t.B[h].x, t.B[h].y, t.B[h].z},
                    cC[3] = {t.C[h].x, t.C[h].y, t.C[h].z}, cD[3] = {t.D[h].x, t.D[h].y, t.D[h].z};
#pragma unroll
        for (int c = 0; c < 3; ++c) wv[h][c] = fmaf(a11, cD[c], fmaf(a10, cB[c], fmaf(a01, cC[c], a00 * cA[c])));
        // keep the padding words of the four gathers reserved until the data is consumed (see loss_fused_kernel)
        pad[h] = __uint_as_float(__float_as_uint(t.A[h].w) | __float_as_uint(t.B[h].w) | __float_as_uint(t.C[h].w) |
                                 __float_as_uint(t.D[h].w));
      }
      f2 sg[3];
      f2 E = pk(pad[0], pad[1]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const f2 e = sub2(pk(wv[0][c], wv[1][c]), bc(cur.tt[c]));
        // sign(e) = 2 sat(e 2^100 + 1/2) - 1, and |e| = sign(e) e
        const f2 st = pk(__saturatef(fmaf(lo(e), 1.2676506e30f, 0.5f)), __saturatef(fmaf(hi(e), 1.2676506e30f, 0.5f)));
        sg[c] = fma2(st, bc(2.0f), bc(-1.0f));
        E = fma2(sg[c], e, E);
      }
      float J[4][2];   // J_k = sum_c sign(e_c) corner_k[c]
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float s0 = h ? hi(sg[0]) : lo(sg[0]), s1 = h ? hi(sg[1]) : lo(sg[1]), s2 = h ? hi(sg[2]) : lo(sg[2]);
        J[0][h] = fmaf(s2, t.A[h].z, fmaf(s1, t.A[h].y, s0 * t.A[h].x));
        J[1][h] = fmaf(s2, t.B[h].z, fmaf(s1, t.B[h].y, s0 * t.B[h].x));
        J[2][h] = fmaf(s2, t.C[h].z, fmaf(s1, t.C[h].y, s0 * t.C[h].x));
        J[3][h] = fmaf(s2, t.D[h].z, fmaf(s1, t.D[h].y, s0 * t.D[h].x));
      }
      const f2 JA = pk(J[0][0], J[0][1]), JB = pk(J[1][0], J[1][1]), JC = pk(J[2][0], J[2][1]), JD = pk(J[3][0], J[3][1]);
      // d/dx = wy0 (JB - JA) + wy1 (JD - JC), d/dy = wx0 (JC - JA) + wx1 (JD - JB), in lerp form
      const f2 ax = sub2(JB, JA), bx = sub2(JD, JC), ay = sub2(JC, JA), by = sub2(JD, JB);
      const f2 dx = fma2(t.wy1, sub2(bx, ax), ax);
      const f2 dy = fma2(t.wx1, sub2(by, ay), ay);
      keep[p].E = E;
      keep[p].u0 = mul2(dx, t.rz);
      keep[p].u1 = mul2(dy, t.rz);
      keep[p].w2 = fma2(t.qy, keep[p].u1, mul2(t.qx, keep[p].u0));
      if (has_next) tap_issue(t, p, gn);
    }
#ifdef VSL_EXP_NOSTREAM   // timing experiment (wrong results): what do the streamed loads cost?
    nxt = cur;
#else
    if (has_next) load_stream(nxt, pofs + W);
#endif
    prefetch_stream(pofs + kPF * W, r + kPF, rows);

    // ---- phase 2: nothing below touches a load in flight
    // smoothness, first half: the second differences this element owns, their weighted signs, and the requests
    // (shuffles, halo loads) for those of the two columns to the left -- consumed after the per-view block below
    float gq0, a1, a2, c01;    // gq0: the part of the gradient that needs no neighbour
    {
      const float* q = qt + (r + kHalo) * kQS + lane + kHalo;
      const float q00 = q[0], q01 = q[1], q02 = q[2], q10 = q[kQS], q11 = q[kQS + 1], q20 = q[2 * kQS];
      const float dx0 = q01 - q00, dy0 = q10 - q00;
      const float dxx = (q02 - q01) - dx0, dyy = (q20 - q10) - dy0, dxy = (q11 - q10) - dx0;
      // row conditions come as weights from the row table: yy needs y < H-2, the mixed difference y < H-1
      const float kyy_r = gyv[3 * r + 1], kxy_r = kxy_l * gyv[3 * r + 2];
      const float a0 = kxx_l * sign_fast(dxx), b0 = kyy_r * sign_fast(dyy), c00 = kxy_r * sign_fast(dxy);
      sm_sum = fmaf(kxx_l, fabsf(dxx), fmaf(kyy_r, fabsf(dyy), fmaf(kxy_r, fabsf(dxy), sm_sum)));
      a1 = __shfl_up_sync(0xffffffffu, a0, 1); a2 = __shfl_up_sync(0xffffffffu, a0, 2);
      c01 = __shfl_up_sync(0xffffffffu, c00, 1);
      gq0 = a0 + (b0 - 2.f * b1 + b2) + (c00 - c10 + c11);
      b2 = b1; b1 = b0; c10 = c00;
    }

    f2 gacc = bc(0.f);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      f2 m;
      if (use_lg) {
        const float l00 = cur.lg[4 * p], l01 = cur.lg[4 * p + 1], l10 = cur.lg[4 * p + 2], l11 = cur.lg[4 * p + 3];
        const float z0 = l00 - l01, z1 = l10 - l11;
        const f2 e = pk(ex2_fast(-1.4426950408889634f * fabsf(z0)), ex2_fast(-1.4426950408889634f * fabsf(z1)));
        const f2 se = add2(e, bc(1.0f));
        const f2 big = pk(rcp_fast(lo(se)), rcp_fast(hi(se)));
        const f2 small = mul2(e, big);
        // softmax cross-entropy against [0, 1] = log(1 + e^-|z|) + max(z, 0)
        exp2 = fma2(pk(lg2_fast(lo(se)), lg2_fast(hi(se))), bc(0.6931471805599453f), exp2);
        exp2 = add2(exp2, pk(fmaxf(z0, 0.f), fmaxf(z1, 0.f)));
        float p1a, p0a, p1b, p0b;                            // softmax(l)[1], softmax(l)[0] of the even / odd view
        swap_unless_ge0(z0, lo(small), lo(big), p1a, p0a);
        swap_unless_ge0(z1, hi(small), hi(big), p1b, p0b);
        m = pk(p1a, p1b);
        const f2 k = mul2(m, bc(cpix));
        const f2 tq = sub2(bc(cexp), mul2(keep[p].E, k));     // cexp - cpix E p1
        const float g0a = p0a * lo(tq), g0b = p0b * hi(tq);
        cur.lg[4 * p] = g0a; cur.lg[4 * p + 1] = -g0a; cur.lg[4 * p + 2] = g0b; cur.lg[4 * p + 3] = -g0b;
      } else {
        m = bc(cur.mc);
      }
      pix2 = fma2(m, keep[p].E, pix2);
      const f2 k = mul2(m, bc(cpix));
      const f2 du0 = mul2(keep[p].u0, k), du1 = mul2(keep[p].u1, k), dw2 = mul2(keep[p].w2, k);   // dw2 = -du2
      // <du, t>, the numerator of d/d(depth) (closed form: <du, u> = 0)
      gacc = fma2(du0, pk(Tt[2 * p][0], Tt[2 * p + 1][0]), gacc);
      gacc = fma2(du1, pk(Tt[2 * p][1], Tt[2 * p + 1][1]), gacc);
      gacc = fma2(dw2, pk(-Tt[2 * p][2], -Tt[2 * p + 1][2]), gacc);
      const f2 dgy = bc(gc.dgy), dd = bc(gc.d);
      S2[p][0] = fma2(du0, dgy, S2[p][0]); S2[p][1] = fma2(du1, dgy, S2[p][1]); S2[p][2] = fma2(dw2, dgy, S2[p][2]);
      S3[p][0] = fma2(du0, dd, S3[p][0]);  S3[p][1] = fma2(du1, dd, S3[p][1]);  S3[p][2] = fma2(dw2, dd, S3[p][2]);
      S4[p][0] = add2(S4[p][0], du0);      S4[p][1] = add2(S4[p][1], du1);      S4[p][2] = add2(S4[p][2], dw2);
    }
    const float g_d = -(lo(gacc) + hi(gacc)) * rcp_fast(gc.d);
    // smoothness, second half: gradient = the weighted signs of the 10 stencils the element is part of (own column
    // carried down in registers, lanes 0 and 1 take what lies left of the tile from the halo columns)
    float g_q;
    {
      const int o = (r + kHalo) * 2;
      const float h0 = sha[o], h1 = sha[o + 1], hc1 = shc[o + 1];
      a1 = lane == 0 ? h1 : a1;
      a2 = lane == 0 ? h0 : (lane == 1 ? h1 : a2);
      c01 = lane == 0 ? hc1 : c01;
      g_q = gq0 + (a2 - 2.f * a1) - c01;
      c11 = c01;
    }
    if (act) {
      float dd_dx = 1.f, dq_dx = 1.f;
      if (smooth_inv) {
        const float qc = qt[(r + kHalo) * kQS + (xl - x_base) + kHalo];
        dq_dx = -qc * qc;
        if (depth_inv) dd_dx = dq_dx;
      } else if (depth_inv) {
        dd_dx = -gc.d * gc.d;
      }
      float g_out = g_d * dd_dx + g_q * dq_dx;
      if (P.x_is_logit != 0) {
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
    Stream st0, st1;
    Geo g0, g1;
#pragma unroll
    for (int a = 1; a < kPF; ++a) prefetch_stream(pofs + a * W, a, rows);
    load_stream(st0, pofs);
    make_geo(g0, 0);
#pragma unroll
    for (int p = 0; p < NP; ++p) tap_issue(tap[p], p, g0);
    for (int r = 0; r < rows; r += 2) {
      row(st0, st1, g0, g1, r);
      if (r + 1 < rows) row(st1, st0, g1, g0, r + 1);
    }
  }

  // ---- 4. one warp reduction per tile (same slots as loss_fused_kernel)
  // recomputed: not worth a register in the loop.  (The compiler keeps min(x, W - 1) on the stack across the row loop
  // for this -- 4 bytes, reloaded here once per tile, ~3 % of the kernel in the r6 capture.  Rebuilding the column from
  // the running pixel offset, or from %laneid, frees that slot and ptxas then spills a value INSIDE the loop: measured 2 us
  // and 1 us slower.)
  const float gx_end = grid_coord(min(x_base + lane, W - 1), W, P.wstep[s]);
  float vals[N];
  float pix_sum = 0.f, exp_sum = 0.f;
  pix_sum = lo(pix2) + hi(pix2);
  exp_sum = lo(exp2) + hi(exp2);
  vals[0] = pix_sum * cpix; vals[1] = sm_sum; vals[2] = exp_sum * cexp;
  vals[3] = 0.f;                                                // consistency term: the scalar kernel only
  if (!act) { vals[0] = 0.f; vals[1] = 0.f; vals[2] = 0.f; }   // lanes past the image edge recomputed the last column
#pragma unroll
  for (int v = 0; v < V; ++v)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int p = v >> 1;
      const float sgn_i = i == 2 ? -1.f : 1.f;               // row 2 accumulated -du2
      const float s2 = sgn_i * ((v & 1) ? hi(S2[p][i]) : lo(S2[p][i]));
      const float s3 = sgn_i * ((v & 1) ? hi(S3[p][i]) : lo(S3[p][i]));
      const float s4 = sgn_i * ((v & 1) ? hi(S4[p][i]) : lo(S4[p][i]));
      vals[kLossSlots + v * 12 + i] = act ? gx_end * s3 : 0.f;
      vals[kLossSlots + v * 12 + 3 + i] = act ? s2 : 0.f;
      vals[kLossSlots + v * 12 + 6 + i] = act ? s3 : 0.f;
      vals[kLossSlots + v * 12 + 9 + i] = act ? s4 : 0.f;
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
}


namespace {
template <int V>
int launch_pair(const LossParams& P, cudaStream_t st) {
  // > 48 KB of dynamic shared memory needs the opt-in; idempotent and cheap, so set on every call (no state)
  cudaError_t e = cudaFuncSetAttribute(loss_fused_pair_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)PairSmem<V>::block_bytes);
  if (e != cudaSuccess) return (int)e;
  const int n = P.item_begin[P.S];
  // programmatic dependent launch, as for loss_fused_kernel: the kernel waits (griddepcontrol.wait) before it
  // touches anything the prep launch produced
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((n + kWarps - 1) / kWarps);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = PairSmem<V>::block_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, loss_fused_pair_kernel<V>, P);
  return e == cudaSuccess ? VSL_OK : (int)e;
}
}  // namespace

int launch_fused_pair(int V, const LossParams& P, cudaStream_t st) {
  switch (V) {
    case 2: return launch_pair<2>(P, st);
#ifndef VSL_DEV_V2_ONLY
    case 4: return launch_pair<4>(P, st);
#endif
    default: return VSL_E_UNSUPPORTED;
  }
}

}  // namespace vsl
