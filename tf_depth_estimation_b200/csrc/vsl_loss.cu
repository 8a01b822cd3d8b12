// libvsl: the fused multi-scale view-synthesis loss, forward AND backward in one pass.
//
// What the reference does per training step (train.py:107-135, train_depth_then_cam_lr.py:211-328) is a
// Python loop over S scales x V source views of ~150 TF ops each, then autodiff of all of it.  Here the
// whole thing is three launches:
//
//   1. prep_xforms_kernel   (vsl_ops.cu)  K_s^-1 and P = K4_s . T_v per (scale, view, batch element)
//   2. pyramid_kernel       (vsl_ops.cu)  resize_area levels 1..S-1 of the target and source images
//   3. loss_fused_kernel    (this file)   every scale and view in ONE grid:
//        per target pixel: smoothness stencil (forward sum + gradient) on a shared-memory tile of x,
//        then per view: back-project, pose, project, bilinear gather of the source, L1 against the
//        target, explainability / validity mask, softmax cross-entropy regulariser -- and, because the
//        loss is a weighted sum of means whose upstream gradient is known (loss_scale), the gradients
//        d/dx, d/dlogits are written in the same pass and dP = sum du (x) [cam;1] is accumulated in
//        registers and reduced once per block.  No full-resolution intermediate is ever written.
//   4. loss_finalize_kernel (this file)   fixed-order reduction of the block partials (deterministic),
//        dT = K4^T dP summed over scales, pose chain rule, the three loss scalars.
//
// Work decomposition: an "item" is an 8-row x (32*R)-column tile of one image at one scale; block = 8
// warps, warp w owns row w, a thread visits R pixels 32 columns apart, so every global access of a warp
// is a run of 32 consecutive pixels.  Items of all scales live in one 1-D grid, largest scale first.
#include "vsl_common.cuh"
// Part of the single translation unit vsl_lib.cu (prep_xforms_kernel / pyramid_kernel come from vsl_ops.cu).

namespace vsl {

constexpr int kTH = 8;      // tile rows = warps per block
constexpr int kMaxR = 4;    // pixels per thread
constexpr int kHalo = 2;
constexpr int kTileW = 32 * kMaxR + 2 * kHalo;
constexpr int kTileH = kTH + 2 * kHalo;

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
};

template <int V> struct NT { static constexpr int value = 3 + 12 * V; };

template <int V>
__global__ void __launch_bounds__(256, 2)
loss_fused_kernel(const LossParams P) {
  __shared__ float qt[kTileH][kTileW];          // smoothness operand (x or 1/x) with a 2-pixel halo
  __shared__ Xform sxf[V];
  __shared__ float scratch[NT<V>::value * kTH];

  // ---- which item
  int s = 0;
  while (s + 1 < P.S && (int)blockIdx.x >= P.item_begin[s + 1]) ++s;
  const int local = blockIdx.x - P.item_begin[s];
  const int per_b = P.bands[s] * P.tiles_x[s];
  const int b = local / per_b, rem = local - b * per_b;
  const int band = rem / P.tiles_x[s], tx = rem - band * P.tiles_x[s];
  const int H = P.H >> s, W = P.W >> s, R = P.R[s];
  const int y_base = band * kTH, x_base = tx * 32 * R;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // ---- stage the x tile (+halo) and this image's V transforms
  const float* __restrict__ xs = P.x[s] + (size_t)b * H * W;
  const int tw = 32 * R + 2 * kHalo;
  for (int e = threadIdx.x; e < kTileH * tw; e += blockDim.x) {
    const int ty = e / tw, txx = e - ty * tw;
    const int gy = y_base + ty - kHalo, gx = x_base + txx - kHalo;
    float v = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      v = xs[(size_t)gy * W + gx];
      if (P.smooth_on_inverse) v = __fdiv_rn(1.0f, v);
    }
    qt[ty][txx] = v;
  }
  for (int e = threadIdx.x; e < V * 21; e += blockDim.x)
    reinterpret_cast<float*>(sxf)[e] =
        reinterpret_cast<const float*>(P.xf + ((size_t)s * V + e / 21) * P.B + b)[e % 21];
  __syncthreads();

  const int y = y_base + warp;
  const bool row_ok = y < H;
  const float gy = grid_coord(y, H, grid_step(H));
  const float wstep = grid_step(W);
  const float cpix = P.cpix[s], cexp = P.cexp[s];
  const float cxx = P.csm[s][0], cxy = P.csm[s][1], cyx = P.csm[s][2], cyy = P.csm[s][3];

  float pix_sum = 0.f, exp_sum = 0.f, sm_sum = 0.f;
  float S1[V][3], S2[V][3], S3[V][3];  // sum du*d*gx, sum du*d, sum du
#pragma unroll
  for (int v = 0; v < V; ++v)
#pragma unroll
    for (int i = 0; i < 3; ++i) { S1[v][i] = 0.f; S2[v][i] = 0.f; S3[v][i] = 0.f; }

  for (int r = 0; r < R; ++r) {
    const int x = x_base + r * 32 + lane;
    if (!(row_ok && x < W)) continue;
    const size_t pix = ((size_t)b * H + y) * W + x;
    const int ty = warp + kHalo, txx = r * 32 + lane + kHalo;

    // ---- smoothness: forward terms owned by (y,x), gradient gathered from the 14 stencils touching it
    float g_q;
    {
      auto Q = [&](int dy, int dx) { return qt[ty + dy][txx + dx]; };
      auto d2 = [&](float a, float bb, float c) { return __fsub_rn(__fsub_rn(c, bb), __fsub_rn(bb, a)); };
      auto mixed = [&](int dy, int dx, float& vxy, float& vyx) {
        const float q00 = Q(dy, dx), q01 = Q(dy, dx + 1), q10 = Q(dy + 1, dx), q11 = Q(dy + 1, dx + 1);
        vxy = __fsub_rn(__fsub_rn(q11, q10), __fsub_rn(q01, q00));
        vyx = __fsub_rn(__fsub_rn(q11, q01), __fsub_rn(q10, q00));
      };
      const bool xm2 = x >= 2 && x < W, xm1 = x >= 1 && x + 1 < W, x0 = x + 2 < W;
      const bool ym2 = y >= 2, ym1 = y >= 1 && y + 1 < H, y0 = y + 2 < H;
      const float qc = Q(0, 0);
      const float vx0 = x0 ? d2(qc, Q(0, 1), Q(0, 2)) : 0.f;
      const float vx1 = xm1 ? d2(Q(0, -1), qc, Q(0, 1)) : 0.f;
      const float vx2 = xm2 ? d2(Q(0, -2), Q(0, -1), qc) : 0.f;
      const float vy0 = y0 ? d2(qc, Q(1, 0), Q(2, 0)) : 0.f;
      const float vy1 = ym1 ? d2(Q(-1, 0), qc, Q(1, 0)) : 0.f;
      const float vy2 = ym2 ? d2(Q(-2, 0), Q(-1, 0), qc) : 0.f;
      float a00 = 0.f, b00 = 0.f, a01 = 0.f, b01 = 0.f, a10 = 0.f, b10 = 0.f, a11 = 0.f, b11 = 0.f;
      const bool cx0 = x + 1 < W, cx1 = x >= 1, cy0 = y + 1 < H, cy1 = y >= 1;
      if (cy0 && cx0) mixed(0, 0, a00, b00);
      if (cy0 && cx1) mixed(0, -1, a01, b01);
      if (cy1 && cx0) mixed(-1, 0, a10, b10);
      if (cy1 && cx1) mixed(-1, -1, a11, b11);
      sm_sum += cxx * fabsf(vx0) + cyy * fabsf(vy0) + cxy * fabsf(a00) + cyx * fabsf(b00);
      g_q = cxx * (sgn(vx0) - 2.f * sgn(vx1) + sgn(vx2)) + cyy * (sgn(vy0) - 2.f * sgn(vy1) + sgn(vy2)) +
            cxy * (sgn(a00) - sgn(a01) - sgn(a10) + sgn(a11)) + cyx * (sgn(b00) - sgn(b01) - sgn(b10) + sgn(b11));
    }

    // ---- depth of this pixel and d(depth)/dx, d(q)/dx
    const float qc = qt[ty][txx];
    float d, dd_dx, dq_dx;
    if (P.smooth_on_inverse) {
      dq_dx = -qc * qc;
      if (P.depth_is_inverse) { d = qc; dd_dx = dq_dx; }
      else { d = xs[(size_t)y * W + x]; dd_dx = 1.f; }
    } else {
      dq_dx = 1.f;
      if (P.depth_is_inverse) { d = __fdiv_rn(1.0f, qc); dd_dx = -d * d; }
      else { d = qc; dd_dx = 1.f; }
    }

    const float gx = grid_coord(x, W, wstep);
    const float* __restrict__ tp = P.tgt[s] + pix * 3;
    const float t0 = tp[0], t1 = tp[1], t2 = tp[2];
    const float dgx = d * gx;
    float g_d = 0.f;

#pragma unroll
    for (int v = 0; v < V; ++v) {
      const Xform& xf = sxf[v];
      Ray ray = back_project(xf.kinv, gx, gy);
      Proj q = project(xf.p, __fmul_rn(ray.r0, d), __fmul_rn(ray.r1, d), __fmul_rn(ray.r2, d));
      Foot f = footprint(q.x, q.y, W, H);
      const float* __restrict__ sb = P.src[v][s] + (size_t)b * H * W * 3;
      const float* __restrict__ p00 = sb + ((size_t)f.y0 * W + f.x0) * 3;
      const float* __restrict__ p01 = sb + ((size_t)f.y1 * W + f.x0) * 3;
      const float* __restrict__ p10 = sb + ((size_t)f.y0 * W + f.x1) * 3;
      const float* __restrict__ p11 = sb + ((size_t)f.y1 * W + f.x1) * 3;
      float i00[3], i01[3], i10[3], i11[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        i00[c] = __ldg(p00 + c); i01[c] = __ldg(p01 + c); i10[c] = __ldg(p10 + c); i11[c] = __ldg(p11 + c);
      }
      // mask value m and its logits
      float m = 1.f, p0 = 0.f, p1 = 0.f;
      float2 lg = make_float2(0.f, 0.f);
      if (P.mask_mode == VSL_MASK_EXP) {
        lg = *reinterpret_cast<const float2*>(P.logits[s] + pix * (2 * V) + 2 * v);
        const float mx = fmaxf(lg.x, lg.y);
        const float e0 = expf(lg.x - mx), e1 = expf(lg.y - mx), se = e0 + e1;
        p0 = e0 / se; p1 = e1 / se;
        m = p1;
        exp_sum += (mx + logf(se)) - lg.y;
      } else if (P.mask_mode == VSL_MASK_CONST) {
        m = P.mask[s][pix];
      }
      const float w00 = __fmul_rn(f.wx0, f.wy0), w01 = __fmul_rn(f.wx0, f.wy1), w10 = __fmul_rn(f.wx1, f.wy0),
                  w11 = __fmul_rn(f.wx1, f.wy1);
      const float tt[3] = {t0, t1, t2};
      float E = 0.f, dx = 0.f, dy = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float e = blend(w00, w01, w10, w11, i00[c], i01[c], i10[c], i11[c]) - tt[c];
        E += fabsf(e);
        const float g = sgn(e);
        dx += g * (f.wy0 * (f.mx1 * i10[c] - f.mx0 * i00[c]) + f.wy1 * (f.mx1 * i11[c] - f.mx0 * i01[c]));
        dy += g * (f.wx0 * (f.my1 * i01[c] - f.my0 * i00[c]) + f.wx1 * (f.my1 * i11[c] - f.my0 * i10[c]));
      }
      pix_sum += m * E;
      if (P.mask_mode == VSL_MASK_EXP) {
        const float g0 = p0 * (cexp - cpix * E * p1);
        *reinterpret_cast<float2*>(P.g_logits[s] + pix * (2 * V) + 2 * v) = make_float2(g0, -g0);
      }
      const float k = cpix * m / q.zp;
      const float du0 = dx * k, du1 = dy * k, du2 = -(q.x * du0 + q.y * du1);
      const float gc0 = du0 * xf.p[0] + du1 * xf.p[4] + du2 * xf.p[8];
      const float gc1 = du0 * xf.p[1] + du1 * xf.p[5] + du2 * xf.p[9];
      const float gc2 = du0 * xf.p[2] + du1 * xf.p[6] + du2 * xf.p[10];
      g_d += gc0 * ray.r0 + gc1 * ray.r1 + gc2 * ray.r2;
      S1[v][0] += du0 * dgx; S1[v][1] += du1 * dgx; S1[v][2] += du2 * dgx;
      S2[v][0] += du0 * d;   S2[v][1] += du1 * d;   S2[v][2] += du2 * d;
      S3[v][0] += du0;       S3[v][1] += du1;       S3[v][2] += du2;
    }
    P.g_x[s][pix] = g_d * dd_dx + g_q * dq_dx;
  }

  // ---- one block reduction per item: 3 loss sums + per view (sum du d gx, gy sum du d, sum du d, sum du)
  float vals[NT<V>::value];
  vals[0] = pix_sum * cpix; vals[1] = sm_sum; vals[2] = exp_sum * cexp;
#pragma unroll
  for (int v = 0; v < V; ++v)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      vals[3 + v * 12 + i] = S1[v][i];
      vals[3 + v * 12 + 3 + i] = gy * S2[v][i];
      vals[3 + v * 12 + 6 + i] = S2[v][i];
      vals[3 + v * 12 + 9 + i] = S3[v][i];
    }
  block_sum<NT<V>::value>(vals, scratch, P.partials + (size_t)blockIdx.x * NT<V>::value);
}

// grid = B + 1 blocks of 128 threads.  Block b < B: pose gradients of batch element b (all views).
// Block B: the three loss scalars.  Everything is summed in a fixed order in double.
template <int V>
__global__ void __launch_bounds__(128)
loss_finalize_kernel(const LossParams P, const float* __restrict__ poses, const float* __restrict__ K_pyr,
                     int pose_format, float inv_loss_scale, float* __restrict__ losses, float* __restrict__ g_poses) {
  constexpr int N = NT<V>::value;
  __shared__ double sh[128];
  auto block_dsum = [&](double v) -> double {
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    double r = sh[0];
    __syncthreads();
    return r;
  };
  const int n_items = P.item_begin[P.S];
  if ((int)blockIdx.x == P.B) {
    for (int t = 0; t < 3; ++t) {
      double s = 0.0;
      for (int i = threadIdx.x; i < n_items; i += blockDim.x) s += (double)P.partials[(size_t)i * N + t];
      s = block_dsum(s);
      if (threadIdx.x == 0) losses[t] = (float)(s * (double)inv_loss_scale);
    }
    return;
  }
  const int b = blockIdx.x;
  const int psz = (pose_format == VSL_POSE_MATRIX) ? 16 : 6;
  for (int v = 0; v < V; ++v) {
    double gT[16];
    for (int i = 0; i < 16; ++i) gT[i] = 0.0;
    for (int s = 0; s < P.S; ++s) {
      const int per_b = P.bands[s] * P.tiles_x[s];
      const int first = P.item_begin[s] + b * per_b;
      double t[12];
      for (int k = 0; k < 12; ++k) {
        double a = 0.0;
        for (int i = threadIdx.x; i < per_b; i += blockDim.x)
          a += (double)P.partials[(size_t)(first + i) * N + 3 + v * 12 + k];
        t[k] = block_dsum(a);
      }
      if (threadIdx.x == 0) {
        const Xform& xf = P.xf[((size_t)s * V + v) * P.B + b];
        const float* Ks = K_pyr + ((size_t)b * P.S + s) * 9;
        double dP[12];
        for (int i = 0; i < 3; ++i) {
          for (int j = 0; j < 3; ++j)
            dP[i * 4 + j] = (double)xf.kinv[j * 3] * t[i] + (double)xf.kinv[j * 3 + 1] * t[3 + i] +
                            (double)xf.kinv[j * 3 + 2] * t[6 + i];
          dP[i * 4 + 3] = t[9 + i];
        }
        for (int k = 0; k < 3; ++k)
          for (int j = 0; j < 4; ++j)
            gT[k * 4 + j] += (double)Ks[k] * dP[j] + (double)Ks[3 + k] * dP[4 + j] + (double)Ks[6 + k] * dP[8 + j];
      }
    }
    if (threadIdx.x == 0) {
      float* out = g_poses + ((size_t)b * V + v) * psz;
      if (pose_format == VSL_POSE_MATRIX) {
        for (int i = 0; i < 16; ++i) out[i] = (float)gT[i];
      } else {
        float g[6];
        pose_vec_grad(poses + ((size_t)b * V + v) * 6, pose_format, gT, g);
        for (int i = 0; i < 6; ++i) out[i] = g[i];
      }
    }
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
  VSL_REQUIRE(d->B > 0 && d->B <= 65535 && d->H > 0 && d->W > 0, VSL_E_SHAPE);
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
    if (s >= 1) lv += (size_t)d->B * H * W * 3;
  }
  L->item_begin[d->S] = n;
  L->n_items = n;
  L->pyr_img = lv;
  L->xf = 0;
  L->partials = round_up(sizeof(Xform) * (size_t)d->S * d->V * d->B, 256);
  L->pyr = L->partials + round_up(sizeof(float) * (size_t)n * nt, 256);
  L->total = L->pyr + sizeof(float) * lv * (size_t)(d->V + 1);
}

template <int V>
int run_loss(const VslLossDesc* d, const WsLayout& L, LossParams& P, const float* poses, const float* K_pyr,
             float* losses, float* g_poses, cudaStream_t st) {
  if (d->ev_main_begin != nullptr) cudaEventRecord((cudaEvent_t)d->ev_main_begin, st);
  loss_fused_kernel<V><<<L.n_items, 256, 0, st>>>(P);
  if (d->ev_main_end != nullptr) cudaEventRecord((cudaEvent_t)d->ev_main_end, st);
  loss_finalize_kernel<V><<<d->B + 1, 128, 0, st>>>(P, poses, K_pyr, d->pose_format, 1.0f / d->loss_scale,
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
  for (int v = 0; v < d->V; ++v) VSL_REQUIRE(srcs[v], VSL_E_NULL);
  for (int s = 0; s < VSL_MAX_SCALES; ++s) {
    P.tgt[s] = nullptr; P.x[s] = nullptr; P.logits[s] = nullptr; P.mask[s] = nullptr;
    P.g_x[s] = nullptr; P.g_logits[s] = nullptr;
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

  // 1. transforms
  const int nx = d->S * d->V * d->B;
  prep_xforms_kernel<<<(nx + 63) / 64, 64, 0, st>>>(poses, K_pyr, d->B, d->S, d->V, d->pose_format, xf, nullptr);
  // 2. image pyramids (target + V sources)
  if (d->S > 1) {
    const int F = 1 << (d->S - 1);
    const int TC = 1024 / F > 128 ? 128 : 1024 / F;
    dim3 grid((d->W + TC - 1) / TC, d->H / F, d->B);
    const size_t smem = sizeof(float) * (size_t)F * TC * 3;
    for (int i = 0; i <= d->V; ++i) {
      PyrLevels lv;
      for (int s = 0; s < VSL_MAX_SCALES; ++s) lv.p[s] = nullptr;
      for (int s = 1; s < d->S; ++s) lv.p[s] = pyr + L.pyr_img * (size_t)i + L.level_off[s];
      pyramid_kernel<<<grid, 256, smem, st>>>(i == 0 ? tgt : srcs[i - 1], d->H, d->W, 3, d->S, F, TC, lv);
    }
  }
  // 3 + 4. fused loss and finalize
  switch (d->V) {
    case 1: return run_loss<1>(d, L, P, poses, K_pyr, losses, g_poses, st);
    case 2: return run_loss<2>(d, L, P, poses, K_pyr, losses, g_poses, st);
    case 3: return run_loss<3>(d, L, P, poses, K_pyr, losses, g_poses, st);
    default: return run_loss<4>(d, L, P, poses, K_pyr, losses, g_poses, st);
  }
}

}  // extern "C"
