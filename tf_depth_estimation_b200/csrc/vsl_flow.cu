// libvsl: the flow-and-depth loss of the DeMoN-pair family (train_optflow_combine.py:138-240; BASELINE configs[3];
// SURVEY 8f.1 "same sampler core, different coordinate source"), forward AND backward in one pass.
//
// Per scale the reference runs three smoothness terms (compute_smooth_loss of the predicted inverse depth and of
// both flow channels, :142-150), a supervised |label - pred_depth| (:163-164), THREE samplers -- the right image
// warped by the ground-truth depth (only its validity mask and coordinates are used, :169-176), by the predicted
// depth (:178-187) and by the predicted flow (optflow_warp, :190-197) -- and |pred_flow - depth_optflow(ground-truth
// coordinates)| (:204-210): about 400 TF ops and 60 full-resolution intermediates per scale, forward alone.  Here:
//
//   launch 1-2  resize_area pyramids of left / right (RGB) and of the label (1 channel), bit-exact (vsl_pyramid.cuh);
//               the transform table (K_s^-1, K4_s . T) rides on the first.
//   launch 3    flow_loss_kernel: one block per 32 x 32 tile of one image at one scale (a warp walks four consecutive
//               rows), tiles of every scale in one grid.  The three predicted maps enter shared memory once with a 2-pixel halo: the smoothness value
//               (each second difference owned by its top-left element) and its gradient (gather form: the signs of
//               the ten stencils an element is part of) come from there.  Each thread then does its pixels: the
//               ground-truth projection (mask + flow target, no gather), the two gathers with the reference's
//               zero-padding masks, the four absolute errors and every gradient; nothing full-resolution is written
//               but the three gradient maps.  Coordinates follow the reference's rounding sequence
//               (vsl_geom.cuh) -- the sampled positions and the mask are bit-identical to the stand-alone warp op.
//   launch 4    fixed-order double-precision sum of the tile partials -> losses[5].
//
// The pose is the loader's 4x4 matrix (tgt2src_projs[:,0], :173) and the label is data: no gradient flows to
// either; the mask of the ground-truth warp carries none (it depends on data only).
#include "vsl_common.cuh"
#include "vsl_prep.cuh"
#include "vsl_pyramid.cuh"

namespace vsl {

constexpr int kFlowTW = 32, kFlowTH = 32, kFlowHalo = 2;
constexpr int kFlowWarps = 8, kFlowRows = kFlowTH / kFlowWarps;    // a warp owns kFlowRows consecutive rows of the tile
constexpr int kFlowThreads = 32 * kFlowWarps;
constexpr int kFlowSW = kFlowTW + 2 * kFlowHalo, kFlowSH = kFlowTH + 2 * kFlowHalo;   // 36 x 36 per map
constexpr int kFlowTerms = 4;   // depth, smooth, optflow, pixel (the order the script prints them, :240)

struct FlowParams {
  int B, S;
  int Hs[VSL_MAX_SCALES], Ws[VSL_MAX_SCALES];
  int tiles_x[VSL_MAX_SCALES], tiles_y[VSL_MAX_SCALES];
  int item_begin[VSL_MAX_SCALES + 1];
  const float* left[VSL_MAX_SCALES];    // RGB levels
  const float* right[VSL_MAX_SCALES];
  const float* label[VSL_MAX_SCALES];   // 1 channel
  const float* pd[VSL_MAX_SCALES];      // predicted inverse depth, flow x, flow y: [B,Hs,Ws,1]
  const float* fx[VSL_MAX_SCALES];
  const float* fy[VSL_MAX_SCALES];
  float* g_pd[VSL_MAX_SCALES];
  float* g_fx[VSL_MAX_SCALES];
  float* g_fy[VSL_MAX_SCALES];
  const Xform* xf;                      // [S][B]
  float* partials;                      // [n_items][kFlowTerms]
  float c_depth[VSL_MAX_SCALES];        // depth_weight / 2^s / (B Hs Ws)
  float c_pixel[VSL_MAX_SCALES];        // data_weight / 2^s / (B Hs Ws 3)
  float c_flow[VSL_MAX_SCALES];         // optflow_weight / 2^s / (B Hs Ws)
  float c_xx[VSL_MAX_SCALES], c_xy[VSL_MAX_SCALES], c_yy[VSL_MAX_SCALES];   // smooth_weight / 2^s / count_k (xy = yx)
  float wstep[VSL_MAX_SCALES], hstep[VSL_MAX_SCALES];   // meshgrid linspace steps 2/(W-1), 2/(H-1) in fp32
  float loss_scale;
};

// sign(e) in {-1, 0, 1} with two FMA-pipe instructions: sat(e * 2^100 + 0.5) is 0, 0.5 or 1 (exact for every float32
// that is not denormal)
VSL_DEV float sign2(float e) { return fmaf(__saturatef(fmaf(e, 1.2676506e30f, 0.5f)), 2.0f, -1.0f); }

// Owners of second differences (my_losses.py:27-36): element (oy, ox) owns dxx = q[ox+2] - 2 q[ox+1] + q[ox] of its
// row, dyy of its column and the two mixed differences of the 2 x 2 cell to its lower right.  A tile needs the owners
// two rows above and two columns left of it as well: (kFlowTH + 2) x (kFlowTW + 2) owners per map.
constexpr int kFlowGW = kFlowTW + kFlowHalo, kFlowGH = kFlowTH + kFlowHalo;
struct FlowSmem {
  float tile[3][kFlowSH * kFlowSW];        // the three predicted maps with their halo, zero outside the image
  float sg[3][3][kFlowGH * kFlowGW];       // per map: sign(dxx), sign(dyy), sign(dxy) + sign(dyx) of the owner, 0 if absent
  Xform sx;
  float scratch[kFlowTerms * kFlowWarps];
};

// bilinear_sampler (utils.py:219-308) of a packed RGB level at (x, y) against `tgt`, weighted by `w`:
// returns sum_c |sample_c - tgt_c| and d(that sum)/dx, d/dy (the sampler's backward, SURVEY 8a "Backward semantics").
VSL_DEV float sample_error(const float* __restrict__ img, int H, int W, float x, float y, const float (&tgt)[3],
                           float& dx, float& dy) {
  const Foot f = footprint(x, y, W, H);
  const float w00 = __fmul_rn(f.wx0, f.wy0), w01 = __fmul_rn(f.wx0, f.wy1), w10 = __fmul_rn(f.wx1, f.wy0),
              w11 = __fmul_rn(f.wx1, f.wy1);
  const float* p00 = img + ((size_t)f.y0 * W + f.x0) * 3;
  const float* p01 = img + ((size_t)f.y1 * W + f.x0) * 3;
  const float* p10 = img + ((size_t)f.y0 * W + f.x1) * 3;
  const float* p11 = img + ((size_t)f.y1 * W + f.x1) * 3;
  float err = 0.f;
  dx = 0.f; dy = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float i00 = __ldg(p00 + c), i01 = __ldg(p01 + c), i10 = __ldg(p10 + c), i11 = __ldg(p11 + c);
    const float e = __fsub_rn(blend(w00, w01, w10, w11, i00, i01, i10, i11), tgt[c]);
    const float g = sign2(e);
    err += fabsf(e);
    dx += g * (f.wy0 * (f.mx1 * i10 - f.mx0 * i00) + f.wy1 * (f.mx1 * i11 - f.mx0 * i01));
    dy += g * (f.wx0 * (f.my1 * i01 - f.my0 * i00) + f.wx1 * (f.my1 * i11 - f.my0 * i10));
  }
  return err;
}

#ifndef VSL_FLOW_MINB
#define VSL_FLOW_MINB 3
#endif
__global__ void __launch_bounds__(kFlowThreads, VSL_FLOW_MINB)
flow_loss_kernel(const FlowParams P) {
  extern __shared__ float4 flow_smem_raw[];
  FlowSmem& sm = *reinterpret_cast<FlowSmem*>(flow_smem_raw);
  float (&tile)[3][kFlowSH * kFlowSW] = sm.tile;
  Xform& sx = sm.sx;
  float* scratch = sm.scratch;
  const int item = blockIdx.x;
  int s = 0;
  while (s + 1 < P.S && item >= P.item_begin[s + 1]) ++s;
  const int H = P.Hs[s], W = P.Ws[s];
  const int per_b = P.tiles_x[s] * P.tiles_y[s];
  const int rem = item - P.item_begin[s];
  const int b = rem / per_b, r2 = rem - b * per_b;
  const int ty = r2 / P.tiles_x[s], tx = r2 - ty * P.tiles_x[s];
  const int y_base = ty * kFlowTH, x_base = tx * kFlowTW;
  const size_t img0 = (size_t)b * H * W;

  if (threadIdx.x < 21)
    reinterpret_cast<float*>(&sx)[threadIdx.x] = reinterpret_cast<const float*>(P.xf + (size_t)s * P.B + b)[threadIdx.x];
  {
    // 36 x 36 elements per map, every slot shared by the three maps; rows of the tile below the image are not loaded
    const float* __restrict__ m0 = P.pd[s] + img0;
    const float* __restrict__ m1 = P.fx[s] + img0;
    const float* __restrict__ m2 = P.fy[s] + img0;
    const int rows = min(kFlowSH, H - y_base + 2 * kFlowHalo);
    for (int r = threadIdx.x; r < rows * kFlowSW; r += kFlowThreads) {
      const int ly = r / kFlowSW, lx = r - ly * kFlowSW;
      const int gy = y_base - kFlowHalo + ly, gx = x_base - kFlowHalo + lx;
      const bool in = (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
      const int o = in ? gy * W + gx : 0;
      const float a = __ldg(m0 + o), bq = __ldg(m1 + o), c = __ldg(m2 + o);
      tile[0][r] = in ? a : 0.f;
      tile[1][r] = in ? bq : 0.f;
      tile[2][r] = in ? c : 0.f;
    }
  }
  __syncthreads();

  float acc[kFlowTerms] = {0.f, 0.f, 0.f, 0.f};
  const float c_xx = P.c_xx[s], c_xy = P.c_xy[s], c_yy = P.c_yy[s];
  {
    // every second difference once, at its owner: value (owners inside the tile proper) and sign (all owners).  Rows
    // of the tile below the image were not loaded: a difference that does not exist is SELECTED away, never multiplied
    constexpr int PT = kFlowSW;
    float val = 0.f;
    for (int e = threadIdx.x; e < kFlowGH * kFlowGW; e += kFlowThreads) {
      const int ol = e / kFlowGW, oc = e - ol * kFlowGW;
      const int oy = y_base - kFlowHalo + ol, ox = x_base - kFlowHalo + oc;
      const bool rin = (unsigned)oy < (unsigned)H, cin = (unsigned)ox < (unsigned)W;
      const bool exx = rin && ox >= 0 && ox + 2 < W;
      const bool eyy = cin && oy >= 0 && oy + 2 < H;
      const bool exy = oy >= 0 && ox >= 0 && oy + 1 < H && ox + 1 < W;
      const bool own = ol >= kFlowHalo && oc >= kFlowHalo;
      const int t0 = ol * PT + oc;
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const float* q = &tile[m][t0];
        const float r0 = q[0], r1 = q[1], r2 = q[2], c1 = q[PT], c2 = q[2 * PT], d = q[PT + 1];
        const float dx0 = __fsub_rn(r1, r0), dy0 = __fsub_rn(c1, r0);
        const float xx = __fsub_rn(__fsub_rn(r2, r1), dx0);
        const float yy = __fsub_rn(__fsub_rn(c2, c1), dy0);
        const float a = __fsub_rn(__fsub_rn(d, c1), dx0);      // d/dy of dx
        const float b = __fsub_rn(__fsub_rn(d, r1), dy0);      // d/dx of dy
        sm.sg[m][0][e] = exx ? sign2(xx) : 0.f;
        sm.sg[m][1][e] = eyy ? sign2(yy) : 0.f;
        sm.sg[m][2][e] = exy ? sign2(a) + sign2(b) : 0.f;
        if (own) {
          float v = exx ? c_xx * fabsf(xx) : 0.f;
          v = eyy ? fmaf(c_yy, fabsf(yy), v) : v;
          v = exy ? fmaf(c_xy, fabsf(a) + fabsf(b), v) : v;
          val += v;
        }
      }
    }
    acc[1] = val;
  }
  __syncthreads();

  const int lx = threadIdx.x & 31, wy = (threadIdx.x >> 5) * kFlowRows;
  const int j = x_base + lx;
  if (j < W) {
    const float gx = grid_coord(j, W, P.wstep[s]);
    const float* right = P.right[s] + img0 * 3;
    const float c_depth = P.c_depth[s], c_pixel = P.c_pixel[s], c_flow = P.c_flow[s];
    const float hstep = P.hstep[s], loss_scale = P.loss_scale;
    const float* __restrict__ labels = P.label[s];
    const float* __restrict__ lefts = P.left[s];
    float* __restrict__ o_pd = P.g_pd[s];
    float* __restrict__ o_fx = P.g_fx[s];
    float* __restrict__ o_fy = P.g_fy[s];
#pragma unroll 1
    for (int rr = 0; rr < kFlowRows; ++rr) {
      const int ly = wy + rr, i = y_base + ly;
      if (i >= H) break;
      const size_t pix = img0 + (size_t)i * W + j;
      const int c0 = (ly + kFlowHalo) * kFlowSW + lx + kFlowHalo;
      const float lab = __ldg(labels + pix);
      float tgt[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) tgt[c] = __ldg(lefts + pix * 3 + c);
      const float pd = tile[0][c0], fx = tile[1][c0], fy = tile[2][c0];
      float g_pd, g_fx, g_fy;
      {   // gradients of the three smoothness terms: the signs of the ten differences the element is part of, times
          // its place in the stencil (1, -2, 1 / 1, -1, -1, 1); small integers, so the bracketed sums are exact
        const int g0 = (ly + kFlowHalo) * kFlowGW + lx + kFlowHalo;
        auto grad = [&](int m) {
          const float* xx = &sm.sg[m][0][g0];
          const float* yy = &sm.sg[m][1][g0];
          const float* xy = &sm.sg[m][2][g0];
          const float gxx = fmaf(-2.f, xx[-1], xx[0] + xx[-2]);
          const float gyy = fmaf(-2.f, yy[-kFlowGW], yy[0] + yy[-2 * kFlowGW]);
          const float gxy = (xy[0] - xy[-1]) - (xy[-kFlowGW] - xy[-kFlowGW - 1]);
          return fmaf(c_xx, gxx, fmaf(c_yy, gyy, c_xy * gxy));
        };
        g_pd = grad(0);
        g_fx = grad(1);
        g_fy = grad(2);
      }
      const float gy = grid_coord(i, H, hstep);
      const Ray ray = back_project(sx.kinv, gx, gy);
      {   // supervised inverse-depth error, :163-164
        const float e = __fsub_rn(lab, pd);
        acc[0] = fmaf(c_depth, fabsf(e), acc[0]);
        g_pd -= c_depth * sgn(e);
      }
      float wmask, tfx, tfy;
      {   // the ground-truth warp: validity mask and flow target only, :169-176, :204
        const float d = __fdiv_rn(1.0f, lab);
        const Proj q = project(sx.p, __fmul_rn(ray.r0, d), __fmul_rn(ray.r1, d), __fmul_rn(ray.r2, d));
        const Foot f = footprint(q.x, q.y, W, H);
        wmask = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(f.wx0, f.wy0), __fmul_rn(f.wx0, f.wy1)), __fmul_rn(f.wx1, f.wy0)),
                          __fmul_rn(f.wx1, f.wy1));
        tfx = __fsub_rn(q.x, gx);     // depth_optflow, utils.py:321-338
        tfy = __fsub_rn(q.y, gy);
      }
      const float kp = c_pixel * wmask;
      {   // the right image warped by the predicted depth, :178-187
        const float d = __fdiv_rn(1.0f, pd);
        const Proj q = project(sx.p, __fmul_rn(ray.r0, d), __fmul_rn(ray.r1, d), __fmul_rn(ray.r2, d));
        float dx, dy;
        const float err = sample_error(right, H, W, q.x, q.y, tgt, dx, dy);
        acc[3] = fmaf(kp, err, acc[3]);
        const float rz = __frcp_rn(q.zp);
        const float du0 = dx * rz, du1 = dy * rz;
        const float du2 = -(q.x * du0 + q.y * du1);
        const float gc0 = du0 * sx.p[0] + du1 * sx.p[4] + du2 * sx.p[8];
        const float gc1 = du0 * sx.p[1] + du1 * sx.p[5] + du2 * sx.p[9];
        const float gc2 = du0 * sx.p[2] + du1 * sx.p[6] + du2 * sx.p[10];
        const float g_d = gc0 * ray.r0 + gc1 * ray.r1 + gc2 * ray.r2;
        g_pd -= kp * g_d * d * d;                           // depth = 1 / pred_depth
      }
      {   // ... and by the predicted flow (optflow_warp, utils.py:201-217), :190-197
        float dx, dy;
        const float err = sample_error(right, H, W, __fadd_rn(gx, fx), __fadd_rn(gy, fy), tgt, dx, dy);
        acc[3] = fmaf(kp, err, acc[3]);
        g_fx += kp * dx;
        g_fy += kp * dy;
      }
      {   // flow against the flow the ground-truth depth implies, :204-210
        const float ex = __fsub_rn(fx, tfx), ey = __fsub_rn(fy, tfy);
        acc[2] = fmaf(c_flow, fabsf(ex) + fabsf(ey), acc[2]);
        g_fx += c_flow * sgn(ex);
        g_fy += c_flow * sgn(ey);
      }
      o_pd[pix] = g_pd * loss_scale;
      o_fx[pix] = g_fx * loss_scale;
      o_fy[pix] = g_fy * loss_scale;
    }
  }
  block_sum<kFlowTerms>(acc, scratch, P.partials + (size_t)item * kFlowTerms);
}

// =====================================================================================================
// The loader's image path on the device (imageselect_Dataloader_optflow.py:120-133, :218-236): the decoded JPEG strip
// (uint8, two frames side by side) -> tf.image.resize_images(strip, [H, 2 W]) (bilinear, TF1 semantics) ->
// tf.to_float -> unpack_image_sequence (target = columns [0, W), source = columns [W, 2 W)).  One thread per output
// pixel and frame, all channels.  TF's ResizeBilinear (un-vendored, unpinned TF 1.x: kernels/resize_bilinear_op.cc,
// align_corners = false, no half-pixel centres): in = out_index * (in_size / (float) out_size); lower = floor(in),
// upper = min(ceil(in), in_size - 1), lerp = in - lower; top = tl + (tr - tl) * x_lerp, bottom likewise,
// out = top + (bottom - top) * y_lerp, all in float32.
// =====================================================================================================
__global__ void __launch_bounds__(256)
unpack_strip_kernel(const unsigned char* __restrict__ strip, int B, int h, int w, int H, int W, float hscale,
                    float wscale, float* __restrict__ tgt, float* __restrict__ src) {
  const size_t n = (size_t)B * H * 2 * W;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int X = (int)(e % (size_t)(2 * W)), Y = (int)((e / (size_t)(2 * W)) % (size_t)H), b = (int)(e / ((size_t)2 * W * H));
  const float iy = __fmul_rn((float)Y, hscale), ix = __fmul_rn((float)X, wscale);
  const float fy0 = floorf(iy), fx0 = floorf(ix);
  const int y0 = (int)fy0, x0 = (int)fx0;
  const int y1 = min((int)ceilf(iy), h - 1), x1 = min((int)ceilf(ix), w - 1);
  const float ly = __fsub_rn(iy, fy0), lx = __fsub_rn(ix, fx0);
  const unsigned char* img = strip + (size_t)b * h * w * 3;
  const unsigned char* tl = img + ((size_t)y0 * w + x0) * 3;
  const unsigned char* tr = img + ((size_t)y0 * w + x1) * 3;
  const unsigned char* bl = img + ((size_t)y1 * w + x0) * 3;
  const unsigned char* br = img + ((size_t)y1 * w + x1) * 3;
  float* out = (X < W ? tgt : src) + (((size_t)b * H + Y) * W + (X < W ? X : X - W)) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float a = (float)tl[c], bq = (float)tr[c], cq = (float)bl[c], d = (float)br[c];
    const float top = __fadd_rn(a, __fmul_rn(__fsub_rn(bq, a), lx));
    const float bot = __fadd_rn(cq, __fmul_rn(__fsub_rn(d, cq), lx));
    out[c] = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
  }
}

__global__ void flow_xforms_kernel(const PrepJob j) {   // single-scale call: no pyramid launch to ride on
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < j.n) prep_one(j, idx);
}

// losses[0..3] = depth, smooth, optflow, pixel; [4] = their sum (total_loss, :240).  One block, fixed order.
constexpr int kFlowFinThreads = 1024;
__global__ void __launch_bounds__(kFlowFinThreads)
flow_finalize_kernel(const float* __restrict__ partials, int n_items, float* __restrict__ losses) {
  __shared__ double sh[kFlowFinThreads / 32][kFlowTerms];
  double a[kFlowTerms] = {0.0, 0.0, 0.0, 0.0};
  constexpr int U = 4;                                   // rows in flight per thread
  for (int i0 = threadIdx.x; i0 < n_items; i0 += kFlowFinThreads * U) {
    float4 q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * kFlowFinThreads;
      q[u] = i < n_items ? reinterpret_cast<const float4*>(partials)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) { a[0] += (double)q[u].x; a[1] += (double)q[u].y; a[2] += (double)q[u].z; a[3] += (double)q[u].w; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < kFlowTerms; ++k) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int k = 0; k < kFlowTerms; ++k) sh[threadIdx.x >> 5][k] = a[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[kFlowTerms] = {0.0, 0.0, 0.0, 0.0};
    for (int w = 0; w < kFlowFinThreads / 32; ++w)
#pragma unroll
      for (int k = 0; k < kFlowTerms; ++k) t[k] += sh[w][k];
#pragma unroll
    for (int k = 0; k < kFlowTerms; ++k) losses[k] = (float)t[k];
    losses[4] = (float)(((t[0] + t[1]) + t[2]) + t[3]);
  }
}

namespace {

struct FlowLayout {
  size_t xf, partials, left[VSL_MAX_SCALES], right[VSL_MAX_SCALES], label[VSL_MAX_SCALES], total;
  int n_items;
};

int check_flow(const VslFlowLossDesc* d) {
  VSL_REQUIRE(d != nullptr, VSL_E_NULL);
  VSL_REQUIRE(d->S >= 1 && d->S <= VSL_MAX_SCALES, VSL_E_SHAPE);
  const int F = 1 << (d->S - 1);
  VSL_REQUIRE(d->B > 0 && d->B <= 65535 / (VSL_MAX_VIEWS + 1) && d->H > 0 && d->W > 0 && d->H % F == 0 && d->W % F == 0,
              VSL_E_SHAPE);
  VSL_REQUIRE((d->H >> (d->S - 1)) >= 3 && (d->W >> (d->S - 1)) >= 3, VSL_E_SHAPE);   // the four means need elements
  VSL_REQUIRE((long long)d->B * d->H * d->W * 3 < (1ll << 31), VSL_E_SHAPE);
  return VSL_OK;
}

FlowLayout flow_layout(const VslFlowLossDesc* d) {
  FlowLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += round_up(bytes, 256); return o; };
  L.xf = take(sizeof(Xform) * (size_t)d->S * d->B);
  L.n_items = 0;
  for (int s = 0; s < d->S; ++s) {
    const int Hs = d->H >> s, Ws = d->W >> s;
    L.n_items += d->B * ((Hs + kFlowTH - 1) / kFlowTH) * ((Ws + kFlowTW - 1) / kFlowTW);
  }
  L.partials = take(sizeof(float) * kFlowTerms * (size_t)L.n_items);
  for (int s = 1; s < d->S; ++s) {
    const size_t n = (size_t)d->B * (d->H >> s) * (d->W >> s);
    L.left[s] = take(n * 3 * sizeof(float));
    L.right[s] = take(n * 3 * sizeof(float));
    L.label[s] = take(n * sizeof(float));
  }
  L.total = off;
  return L;
}

}  // namespace
}  // namespace vsl

using namespace vsl;

extern "C" {

int vsl_unpack_strip(const unsigned char* strip, int B, int h, int w, int H, int W, float* tgt, float* src,
                     vsl_stream_t stream) {
  VSL_REQUIRE(strip && tgt && src, VSL_E_NULL);
  VSL_REQUIRE(B > 0 && h > 0 && w > 1 && H > 0 && W > 0 && (long long)B * H * W * 6 < (1ll << 40), VSL_E_SHAPE);
  const size_t n = (size_t)B * H * 2 * W;
  // CalculateResizeScale: in_size / static_cast<float>(out_size)
  const float hscale = (float)h / (float)H, wscale = (float)w / (float)(2 * W);
  unpack_strip_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(strip, B, h, w, H, W, hscale, wscale,
                                                                                   tgt, src);
  return launch_status();
}

size_t vsl_flow_loss_ws_bytes(const VslFlowLossDesc* d) {
  if (check_flow(d) != VSL_OK) return 0;
  return flow_layout(d).total;
}

int vsl_flow_loss_fwd_bwd(const VslFlowLossDesc* d, const float* left, const float* right, const float* label,
                          const float* const* depth_pyr, const float* const* flowx_pyr, const float* const* flowy_pyr,
                          const float* proj, const float* K_pyr, float* losses, float* const* g_depth_pyr,
                          float* const* g_flowx_pyr, float* const* g_flowy_pyr, void* ws, vsl_stream_t stream) {
  const int rc = check_flow(d);
  if (rc != VSL_OK) return rc;
  VSL_REQUIRE(left && right && label && depth_pyr && flowx_pyr && flowy_pyr && proj && K_pyr && losses && g_depth_pyr &&
                  g_flowx_pyr && g_flowy_pyr && ws, VSL_E_NULL);
  VSL_REQUIRE(aligned(ws, 256), VSL_E_ALIGN);
  for (int s = 0; s < d->S; ++s)
    VSL_REQUIRE(depth_pyr[s] && flowx_pyr[s] && flowy_pyr[s] && g_depth_pyr[s] && g_flowx_pyr[s] && g_flowy_pyr[s],
                VSL_E_NULL);
  cudaStream_t st = (cudaStream_t)stream;
  const FlowLayout L = flow_layout(d);
  char* base = static_cast<char*>(ws);

  FlowParams P;
  P.B = d->B; P.S = d->S;
  P.xf = reinterpret_cast<const Xform*>(base + L.xf);
  P.partials = reinterpret_cast<float*>(base + L.partials);
  P.loss_scale = d->loss_scale;
  int items = 0;
  for (int s = 0; s < VSL_MAX_SCALES; ++s) {
    if (s >= d->S) {
      P.Hs[s] = P.Ws[s] = P.tiles_x[s] = P.tiles_y[s] = 0;
      P.left[s] = P.right[s] = P.label[s] = P.pd[s] = P.fx[s] = P.fy[s] = nullptr;
      P.g_pd[s] = P.g_fx[s] = P.g_fy[s] = nullptr;
      P.c_depth[s] = P.c_pixel[s] = P.c_flow[s] = P.c_xx[s] = P.c_xy[s] = P.c_yy[s] = 0.f;
      P.wstep[s] = P.hstep[s] = 0.f;
      P.item_begin[s + 1] = items;
      continue;
    }
    const int Hs = d->H >> s, Ws = d->W >> s;
    P.Hs[s] = Hs; P.Ws[s] = Ws;
    P.tiles_x[s] = (Ws + kFlowTW - 1) / kFlowTW;
    P.tiles_y[s] = (Hs + kFlowTH - 1) / kFlowTH;
    P.item_begin[s] = items;
    items += d->B * P.tiles_x[s] * P.tiles_y[s];
    P.item_begin[s + 1] = items;
    P.left[s] = s == 0 ? left : reinterpret_cast<const float*>(base + L.left[s]);
    P.right[s] = s == 0 ? right : reinterpret_cast<const float*>(base + L.right[s]);
    P.label[s] = s == 0 ? label : reinterpret_cast<const float*>(base + L.label[s]);
    P.pd[s] = depth_pyr[s]; P.fx[s] = flowx_pyr[s]; P.fy[s] = flowy_pyr[s];
    P.g_pd[s] = g_depth_pyr[s]; P.g_fx[s] = g_flowx_pyr[s]; P.g_fy[s] = g_flowy_pyr[s];
    const double k = 1.0 / (double)(1 << s), n = (double)d->B * Hs * Ws;
    P.c_depth[s] = (float)(d->depth_weight * k / n);
    P.c_pixel[s] = (float)(d->data_weight * k / (n * 3.0));
    P.c_flow[s] = (float)(d->optflow_weight * k / n);
    P.c_xx[s] = (float)(d->smooth_weight * k / ((double)d->B * Hs * (Ws - 2)));
    P.c_yy[s] = (float)(d->smooth_weight * k / ((double)d->B * (Hs - 2) * Ws));
    P.c_xy[s] = (float)(d->smooth_weight * k / ((double)d->B * (Hs - 1) * (Ws - 1)));
    P.wstep[s] = 2.0f / (float)(Ws - 1);
    P.hstep[s] = 2.0f / (float)(Hs - 1);
  }

  // launches 1-2: pyramids (+ the transform table: one matrix pose per image, every scale)
  PrepJob prep = make_prep(proj, K_pyr, d->B, d->S, 1, VSL_POSE_MATRIX, reinterpret_cast<Xform*>(base + L.xf), nullptr);
  if (d->S > 1) {
    PyrJob rgb;
    rgb.nimg = 2;
    rgb.img[0] = left; rgb.img[1] = right;
    for (int s = 0; s < VSL_MAX_SCALES; ++s) {
      rgb.lvl[0][s] = (s >= 1 && s < d->S) ? reinterpret_cast<float*>(base + L.left[s]) : nullptr;
      rgb.lvl[1][s] = (s >= 1 && s < d->S) ? reinterpret_cast<float*>(base + L.right[s]) : nullptr;
    }
    int e = launch_pyramid(rgb, prep, d->B, d->H, d->W, 3, d->S, st);
    if (e != VSL_OK) return e;
    PyrJob lab;
    lab.nimg = 1;
    lab.img[0] = label;
    for (int s = 0; s < VSL_MAX_SCALES; ++s)
      lab.lvl[0][s] = (s >= 1 && s < d->S) ? reinterpret_cast<float*>(base + L.label[s]) : nullptr;
    PrepJob none = prep;
    none.n = 0;
    e = launch_pyramid(lab, none, d->B, d->H, d->W, 1, d->S, st);
    if (e != VSL_OK) return e;
  } else {
    flow_xforms_kernel<<<(prep.n + 127) / 128, 128, 0, st>>>(prep);
    const int e = launch_status();
    if (e != VSL_OK) return e;
  }
  {
    // > 48 KB of dynamic shared memory needs the opt-in; idempotent and cheap, so set on every call (no state)
    const cudaError_t ea = cudaFuncSetAttribute(flow_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)sizeof(FlowSmem));
    if (ea != cudaSuccess) return (int)ea;
  }
  flow_loss_kernel<<<items, kFlowThreads, sizeof(FlowSmem), st>>>(P);
  int e = launch_status();
  if (e != VSL_OK) return e;
  flow_finalize_kernel<<<1, kFlowFinThreads, 0, st>>>(P.partials, items, losses);
  return launch_status();
}

}  // extern "C"
