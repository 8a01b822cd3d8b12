// libvsl: the SSIM part of the photometric term INSIDE the fused step (BASELINE.json north_star: "photometric L1+SSIM
// reconstruction term ... stages the warped and target tiles plus the 3x3 halo in shared memory for SSIM").
// EXTENSION: the reference has no SSIM; semantics are this library's own, oracle =
// oracle/vsl_oracle.py view_synthesis_loss(flags.ssim_weight), parity unpinned.
//
// With VslLossDesc.ssim_weight = a in (0, 1] the photometric term of scale s and view v becomes
//     data_weight_s * [ (1 - a) * mean_{B,H,W,3}(|warp - tgt| * m)  +  a * mean_{B,H-2,W-2,3}(D * m_centre) ],
//     D = clip((1 - SSIM_3x3(warp, tgt)) / 2, 0, 1)   (VALID 3 x 3 windows, C1 = 0.01^2, C2 = 0.03^2),
// m the explainability / validity mask at the window's centre.  The L1 part stays in the fused loss kernel (its weight
// scaled by 1 - a); this kernel adds the SSIM part -- value AND gradient, no full-resolution intermediate:
//
//   one block = a 32 x 16 tile of one image at one scale (every scale in one grid; 256 threads, two rows each in P3),
//   looping over the source views
//   P1  the warped and target colours of the tile and a 2-pixel ring (36 x 20) -> shared memory.  Same sample
//       positions as the fused kernel's fast arithmetic: folded projection u = d (Q [gx, gy, 1]) + t, clamp to
//       [-2, size], four 16-byte gathers from the zero-bordered RGBA level the prep launch wrote.
//   P2  the windows centred on the tile and a 1-pixel ring (34 x 18): SSIM per channel, and the three coefficients of
//       its derivative -- a window's dS/d(warp_k) is AFFINE in that pixel's (warp, tgt) values, alpha + beta warp_k +
//       gamma tgt_k -- times the upstream weight (-a data_weight_s m / 2 count, 0 where the clip is active), to shared
//       memory; the tile's share of the loss.
//   P3  every pixel gathers the (up to) nine windows it belongs to -> d/d(warp colour), and chains it through the
//       sampler and the projection exactly as the fused kernel does: d/dx, d/dy, du, d/d(depth) summed over the views
//       and added to the g_x the fused kernel wrote (this launch runs after it; one owner per pixel), the 12 sums of
//       du (x) [d gx, d gy, d, 1] and the loss added to the owning tile's partial slot (13 float reductions per block:
//       the one place of this term whose summation order is not fixed), the mask's share added to g_logits.
#include "vsl_loss_common.cuh"
#include "vsl_ssim.cuh"

namespace vsl {

#ifndef VSL_SSIM_TH
#define VSL_SSIM_TH 16
#endif
constexpr int kStW = 32, kStH = VSL_SSIM_TH;      // (8-row tiles: 1.69 ring samples and 1.33 windows per pixel; 16: 1.41 / 1.20)
constexpr int kSvW = kStW + 4, kSvH = kStH + 4;   // colour values: tile + 2-pixel ring
constexpr int kScW = kStW + 2, kScH = kStH + 2;   // window centres: tile + 1-pixel ring
constexpr int kStThreads = 256;
constexpr int kStRows = kStW * kStH / kStThreads; // tile rows a thread owns in P3 (rows ly, ly + 8, ...)
static_assert(kStRows >= 1 && kStRows * kStThreads == kStW * kStH, "whole rows per thread");

struct SsimSmem {                                  // 49.6 KB at 16-row tiles: dynamic shared memory (opt-in above 48 KB)
  float va[3][kSvH * kSvW], vt[3][kSvH * kSvW];
  float coef[3][3][kScH * kScW];
  float dmap[3][kScH * kScW];   // D per channel and window
  float msk[kScH * kScW];       // mask at the window's centre
  float scratch[13 * (kStThreads / 32)];
  float tot[13];
  XformQ sxq;
};

// One target pixel's sample in source view v: everything the forward value and the backward chain need.
struct SsimTap {
  float4 A, B, C, D;        // corners (x0,y0) (x1,y0) (x0,y1) (x1,y1)
  float wx1, wy1;           // (x - x0), (y - y0)
  float qx, qy, rz, d;      // projected coordinates (unclamped), 1 / (z + eps), depth
  float gx, gy;             // grid coordinates of the target pixel
};

VSL_DEV SsimTap ssim_tap(const LossParams& P, const XformQ& xq, int s, int v, int b, int y, int x) {
  SsimTap t;
  const int H = P.Hs[s], W = P.Ws[s];
  const float xr = __ldg(P.x[s] + ((size_t)b * H + y) * W + x);
  t.d = P.depth_is_inverse ? __fdividef(1.0f, xr) : xr;
  t.gx = grid_coord(x, W, P.wstep[s]);
  t.gy = grid_coord(y, H, P.hstep[s]);
  const float u0 = fmaf(t.d, fmaf(xq.q[0], t.gx, fmaf(xq.q[1], t.gy, xq.q[2])), xq.t[0]);
  const float u1 = fmaf(t.d, fmaf(xq.q[3], t.gx, fmaf(xq.q[4], t.gy, xq.q[5])), xq.t[1]);
  const float u2 = fmaf(t.d, fmaf(xq.q[6], t.gx, fmaf(xq.q[7], t.gy, xq.q[8])), xq.t[2]);
  t.rz = __fdividef(1.0f, u2 + kEpsZ);
  t.qx = u0 * t.rz; t.qy = u1 * t.rz;
  // beyond [-2, size] all four corners are border zeros; clamping there changes neither value nor gradient
  const float xc = fminf(fmaxf(t.qx, -2.0f), (float)W), yc = fminf(fmaxf(t.qy, -2.0f), (float)H);
  const float fx = floorf(xc), fy = floorf(yc);
  t.wx1 = xc - fx; t.wy1 = yc - fy;
  const int stride4 = P.stride4[s];
  const float4* g = P.src[v][s] + (size_t)b * P.plane4[s] + ((int)fy + kPad) * stride4 + ((int)fx + kPad);
  t.A = __ldg(g); t.B = __ldg(g + 1); t.C = __ldg(g + stride4); t.D = __ldg(g + stride4 + 1);
  return t;
}

struct SsimGrid {
  int item_begin[VSL_MAX_SCALES + 1];   // blocks of scale s: [item_begin[s], item_begin[s+1])
  int tiles_x[VSL_MAX_SCALES], tiles_y[VSL_MAX_SCALES];
  float css[VSL_MAX_SCALES];            // loss_scale * a * data_weight_s / (B (Hs-2) (Ws-2) 3)
  int V;
};

#ifndef VSL_SSIM_MINB
#define VSL_SSIM_MINB 4
#endif
__global__ void __launch_bounds__(kStThreads, VSL_SSIM_MINB)
loss_ssim_kernel(const LossParams P, const SsimGrid G) {
  extern __shared__ float4 ssim_smem_raw[];
  SsimSmem& sm = *reinterpret_cast<SsimSmem*>(ssim_smem_raw);
  float (&va)[3][kSvH * kSvW] = sm.va;
  float (&vt)[3][kSvH * kSvW] = sm.vt;
  float (&coef)[3][3][kScH * kScW] = sm.coef;
  float (&dmap)[3][kScH * kScW] = sm.dmap;
  float (&msk)[kScH * kScW] = sm.msk;
  float* scratch = sm.scratch;
  float* tot = sm.tot;
  XformQ& sxq = sm.sxq;
  const int item = blockIdx.x;
  int s = 0;
  while (s + 1 < P.S && item >= G.item_begin[s + 1]) ++s;
  const int H = P.Hs[s], W = P.Ws[s];
  const int per_img = G.tiles_x[s] * G.tiles_y[s];
  const int rem = item - G.item_begin[s];
  const int b = rem / per_img, r2 = rem - b * per_img;
  const int ty = r2 / G.tiles_x[s], tx = r2 - ty * G.tiles_x[s];
  const int y_base = ty * kStH, x_base = tx * kStW;
  const float css = G.css[s];
  const int nlg = 2 * G.V;

  float gx_add[kStRows];  // a pixel's d/dx summed over the views in view order: one owner, one store, deterministic
#pragma unroll
  for (int rr = 0; rr < kStRows; ++rr) gx_add[rr] = 0.f;
  for (int v = 0; v < G.V; ++v) {
  __syncthreads();        // the previous view's P3 has finished reading the tiles
  if (threadIdx.x < 12)
    reinterpret_cast<float*>(&sxq)[threadIdx.x] =
        reinterpret_cast<const float*>(P.xq + ((size_t)s * G.V + v) * P.B + b)[threadIdx.x];
  __syncthreads();

  // ---- P1: colours of the tile + 2-pixel ring
  for (int p = threadIdx.x; p < kSvH * kSvW; p += kStThreads) {
    const int ly = p / kSvW, lx = p - ly * kSvW;
    const int y = y_base - 2 + ly, x = x_base - 2 + lx;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, t0 = 0.f, t1 = 0.f, t2 = 0.f;
    if ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) {
      const SsimTap t = ssim_tap(P, sxq, s, v, b, y, x);
      const float w11 = t.wx1 * t.wy1, w10 = t.wx1 - w11, w01 = t.wy1 - w11, w00 = 1.0f - t.wx1 - w01;
      a0 = fmaf(w11, t.D.x, fmaf(w10, t.B.x, fmaf(w01, t.C.x, w00 * t.A.x)));
      a1 = fmaf(w11, t.D.y, fmaf(w10, t.B.y, fmaf(w01, t.C.y, w00 * t.A.y)));
      a2 = fmaf(w11, t.D.z, fmaf(w10, t.B.z, fmaf(w01, t.C.z, w00 * t.A.z)));
      const float* tp = P.tgt[s] + (((size_t)b * H + y) * W + x) * 3;
      t0 = __ldg(tp); t1 = __ldg(tp + 1); t2 = __ldg(tp + 2);
    }
    va[0][p] = a0; va[1][p] = a1; va[2][p] = a2;
    vt[0][p] = t0; vt[1][p] = t1; vt[2][p] = t2;
  }
  __syncthreads();

  // the mask at a pixel: softmax(logits)[1] of this view, a constant weight, or 1
  auto mask_at = [&](int y, int x) -> float {
    const size_t pix = ((size_t)b * H + y) * W + x;
    if (P.mask_mode == VSL_MASK_EXP) {
      const float2 l = __ldg(reinterpret_cast<const float2*>(P.logits[s] + pix * nlg) + v);
      return __fdividef(1.0f, 1.0f + __expf(l.x - l.y));
    }
    if (P.mask_mode == VSL_MASK_CONST) return __ldg(P.mask[s] + pix);
    return 1.0f;
  };

  // ---- P2: windows centred on the tile + 1-pixel ring.  First the mask at every centre (0 where no window exists),
  // then one (centre, channel) pair per thread and trip: 3 x 612 pairs over the block's 256 threads
  for (int c = threadIdx.x; c < kScH * kScW; c += kStThreads) {
    const int cy = c / kScW, cx = c - cy * kScW;
    const int y = y_base - 1 + cy, x = x_base - 1 + cx;
    const bool valid = y >= 1 && y <= H - 2 && x >= 1 && x <= W - 2;
    msk[c] = valid ? mask_at(y, x) : -1.0f;          // -1: no window centred here
  }
  __syncthreads();
  float acc[13];
#pragma unroll
  for (int k = 0; k < 13; ++k) acc[k] = 0.f;
  for (int it = threadIdx.x; it < 3 * kScH * kScW; it += kStThreads) {
    const int ch = it / (kScH * kScW), c = it - ch * (kScH * kScW);
    const int cy = c / kScW, cx = c - cy * kScW;
    const float m = msk[c];
    float k0 = 0.f, k1 = 0.f, k2 = 0.f, dch = 0.f;
    if (m >= 0.0f) {
      float xv[9], yv[9];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          xv[i * 3 + j] = va[ch][(cy + i) * kSvW + cx + j];
          yv[i * 3 + j] = vt[ch][(cy + i) * kSvW + cx + j];
        }
      const SsimWin w = ssim_stats<true, true>(xv, yv);
      const float raw = 0.5f * (1.0f - w.S);
      dch = fminf(fmaxf(raw, 0.0f), 1.0f);
      const float up = (raw >= 0.0f && raw <= 1.0f) ? -0.5f * css * m : 0.0f;   // d(clip)/d(raw) * d(raw)/dS * weight
      k0 = up * w.ax; k1 = up * w.beta; k2 = up * w.gamma;
      if (cy >= 1 && cy <= kStH && cx >= 1 && cx <= kStW) acc[0] += css * m * dch;   // a window this tile owns
    }
    coef[ch][0][c] = k0; coef[ch][1][c] = k1; coef[ch][2][c] = k2;
    dmap[ch][c] = dch;
  }
  __syncthreads();

  // ---- P3: the tile's own pixels
#pragma unroll
  for (int rr = 0; rr < kStRows; ++rr) {
    const int lx = threadIdx.x & (kStW - 1), ly = threadIdx.x / kStW + rr * (kStThreads / kStW);
    const int y = y_base + ly, x = x_base + lx;
    if (y < H && x < W) {
      const int pv = (ly + 2) * kSvW + lx + 2;          // this pixel in the colour tile
      const int pc = (ly + 1) * kScW + lx + 1;          // the window centred on it
      float ga[3];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const int c = pc + dy * kScW + dx;
            s0 += coef[ch][0][c]; s1 += coef[ch][1][c]; s2 += coef[ch][2][c];
          }
        ga[ch] = fmaf(s1, va[ch][pv], fmaf(s2, vt[ch][pv], s0));
      }
      const size_t pix = ((size_t)b * H + y) * W + x;
      if (ga[0] != 0.f || ga[1] != 0.f || ga[2] != 0.f) {
        const SsimTap t = ssim_tap(P, sxq, s, v, b, y, x);
        const float JA = fmaf(ga[2], t.A.z, fmaf(ga[1], t.A.y, ga[0] * t.A.x));
        const float JB = fmaf(ga[2], t.B.z, fmaf(ga[1], t.B.y, ga[0] * t.B.x));
        const float JC = fmaf(ga[2], t.C.z, fmaf(ga[1], t.C.y, ga[0] * t.C.x));
        const float JD = fmaf(ga[2], t.D.z, fmaf(ga[1], t.D.y, ga[0] * t.D.x));
        // d/dx = wy0 (JB - JA) + wy1 (JD - JC), d/dy = wx0 (JC - JA) + wx1 (JD - JB)
        const float ax = JB - JA, bx = JD - JC, ay = JC - JA, by = JD - JB;
        const float dxs = fmaf(t.wy1, bx - ax, ax), dys = fmaf(t.wx1, by - ay, ay);
        const float du0 = dxs * t.rz, du1 = dys * t.rz, du2 = -(t.qx * du0 + t.qy * du1);
        // d/d(depth) = -<du, t> / depth (closed form: <du, u> = 0), then the chain to the network output x
        const float g_d = -(du0 * sxq.t[0] + du1 * sxq.t[1] + du2 * sxq.t[2]) * __fdividef(1.0f, t.d);
        const float chain = P.depth_is_inverse ? -t.d * t.d : 1.0f;
        gx_add[rr] += g_d * chain;
        const float dgx = t.d * t.gx, dgy = t.d * t.gy;
        acc[1] += du0 * dgx; acc[2] += du1 * dgx; acc[3] += du2 * dgx;
        acc[4] += du0 * dgy; acc[5] += du1 * dgy; acc[6] += du2 * dgy;
        acc[7] += du0 * t.d; acc[8] += du1 * t.d; acc[9] += du2 * t.d;
        acc[10] += du0; acc[11] += du1; acc[12] += du2;
      }
      if (P.mask_mode == VSL_MASK_EXP && y >= 1 && y <= H - 2 && x >= 1 && x <= W - 2) {
        const float m = msk[pc];                          // P2 left the mask of this very centre there
        const float g1 = css * (dmap[0][pc] + dmap[1][pc] + dmap[2][pc]) * m * (1.0f - m);      // d(m)/d(l1) = m (1 - m), d(m)/d(l0) = -m (1 - m)
        float2* gl = reinterpret_cast<float2*>(P.g_logits[s] + pix * nlg) + v;
        float2 cur = *gl;
        cur.x -= g1; cur.y += g1;
        *gl = cur;
      }
    }
  }
  block_sum_bfly<13>(acc, scratch, tot);
  __syncthreads();
  if (threadIdx.x < 13) {
    // the partial slot of the fused kernel's tile (32 columns x kRH rows) this block lies in
    const int tile = P.item_begin[s] + (b * P.bands[s] + y_base / kRH) * P.strips[s] + tx;
    const int N = kLossSlots + 12 * G.V;
    const int col = threadIdx.x == 0 ? 0 : kLossSlots + v * 12 + (threadIdx.x - 1);
    const float val = tot[threadIdx.x];
    if (val != 0.f) atomicAdd(P.partials + (size_t)tile * N + col, val);
  }
  }   // views
#pragma unroll
  for (int rr = 0; rr < kStRows; ++rr) {
    const int lx = threadIdx.x & (kStW - 1), ly = threadIdx.x / kStW + rr * (kStThreads / kStW);
    const int y = y_base + ly, x = x_base + lx;
    if (y < H && x < W && gx_add[rr] != 0.f) P.g_x[s][((size_t)b * H + y) * W + x] += gx_add[rr];
  }
}

int launch_ssim_term(const VslLossDesc* d, const LossParams& P, cudaStream_t st) {
  SsimGrid G;
  G.V = d->V;
  int n = 0;
  for (int s = 0; s < VSL_MAX_SCALES; ++s) {
    G.item_begin[s] = n;
    if (s < d->S) {
      const int H = d->H >> s, W = d->W >> s;
      G.tiles_x[s] = (W + kStW - 1) / kStW;
      G.tiles_y[s] = (H + kStH - 1) / kStH;
      n += d->B * G.tiles_x[s] * G.tiles_y[s];
      const double dw = d->pixel_scale_norm ? (double)d->data_weight / (double)(1 << s) : (double)d->data_weight;
      G.css[s] = (float)((double)d->loss_scale * (double)d->ssim_weight * dw /
                         ((double)d->B * (H - 2) * (W - 2) * 3.0));
    } else {
      G.tiles_x[s] = G.tiles_y[s] = 0;
      G.css[s] = 0.f;
    }
    G.item_begin[s + 1] = n;
  }
  static_assert(kRH % kStH == 0, "an SSIM tile lies inside one tile of the fused kernel");
  // > 48 KB of dynamic shared memory needs the opt-in; idempotent and cheap, so set on every call (no state)
  const cudaError_t ea = cudaFuncSetAttribute(loss_ssim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)sizeof(SsimSmem));
  if (ea != cudaSuccess) return (int)ea;
  loss_ssim_kernel<<<n, kStThreads, sizeof(SsimSmem), st>>>(P, G);
  return launch_status();
}

}  // namespace vsl
