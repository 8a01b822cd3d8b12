// Flagged-off EXTENSIONS of the path: terms BASELINE.json's north_star names (SSIM in the photometric term,
// edge-aware disparity smoothness) but the reference does NOT implement (SURVEY.md D1/D2, row "ext" of §8a).
// Nothing in the reference-parity path calls these; their oracle is oracle/vsl_oracle.py ssim_dissimilarity /
// edge_aware_smooth_loss ("parity unpinned -- no reference implementation").
#pragma once
#include "vsl_common.cuh"
#include "vsl_ssim.cuh"

namespace vsl {

// =====================================================================================================
// SSIM dissimilarity: 3x3 VALID average pools, C1 = 0.01^2, C2 = 0.03^2, clip((1 - SSIM) / 2, 0, 1).
// An image is treated as B x H rows of W*C floats ("flat columns"): the horizontal neighbour of flat column f
// is f + C, so one code path serves every channel count with coalesced rows.
// Forward : block = 8 rows x 128 flat columns of WINDOWS; input tile (+2 rows, +2C columns) in shared memory.
// Backward: block = 8 rows x 128 flat columns of PIXELS.  A window's derivative with respect to its 9 pixels is
//           affine in the pixel values, dS/dx_k = ax + beta * x_k + gamma * y_k (dS/dy_k = ay + beta * y_k +
//           gamma * x_k), so the block first turns the (8+2) x (128+2C) windows around its pixels into four
//           coefficients each (times the upstream gradient), then every pixel GATHERS the 9 windows it belongs
//           to: deterministic, no atomics.
// Variances are accumulated centred (sum (x - mean)^2), not as E[x^2] - mean^2: no cancellation in fp32.
// =====================================================================================================
constexpr int kSsTY = 8, kSsTX = 128, kSsMaxC = 4, kSsThreads = 256;
constexpr int kSsInW = kSsTX + 4 * kSsMaxC, kSsInH = kSsTY + 4;
constexpr int kSsCfW = kSsTX + 2 * kSsMaxC, kSsCfH = kSsTY + 2;

struct SsimDims {
  int B, H, W, C;
  int WC, OWC;      // W*C, (W-2)*C
  float inv_n;      // 1 / (B (H-2) (W-2) C)
};

// rows [row0, row0+nrows) x flat columns [col0, col0+ncols) of image b into dst (row stride `ld`), zero outside
template <int LD, int NROWS>   // LD: row pitch of the tile (a constant, so the index split is a multiply-shift)
VSL_DEV void ssim_load_tile(float* dst, const float* __restrict__ img, int row0, int col0, int ncols, int H, int WC) {
  for (int i = threadIdx.x; i < NROWS * LD; i += kSsThreads) {
    const int r = i / LD, c = i - r * LD;
    if (c >= ncols) continue;
    const int gr = row0 + r, gc = col0 + c;
    float v = 0.f;
    if ((unsigned)gr < (unsigned)H && (unsigned)gc < (unsigned)WC) v = __ldg(img + (size_t)gr * WC + gc);
    dst[i] = v;
  }
}

// Window whose top-left element is at x[0] / y[0]; row stride ld, column stride C.
template <bool GRAD>
VSL_DEV SsimWin ssim_window(const float* x, const float* y, int ld, int C) {
  float xv[9], yv[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { xv[i * 3 + j] = x[i * ld + j * C]; yv[i * 3 + j] = y[i * ld + j * C]; }
  return ssim_stats<GRAD>(xv, yv);
}

__global__ void __launch_bounds__(kSsThreads)
ssim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, SsimDims d, float* __restrict__ map,
                float* __restrict__ partial) {
  __shared__ float tx[(kSsTY + 2) * kSsCfW], ty[(kSsTY + 2) * kSsCfW];
  __shared__ float scratch[kSsThreads / 32];
  const int b = blockIdx.z, i0 = blockIdx.y * kSsTY, f0 = blockIdx.x * kSsTX;
  const size_t img = (size_t)b * d.H * d.WC;
  const int ncols = kSsTX + 2 * d.C;
  ssim_load_tile<kSsCfW, kSsTY + 2>(tx, x + img, i0, f0, ncols, d.H, d.WC);
  ssim_load_tile<kSsCfW, kSsTY + 2>(ty, y + img, i0, f0, ncols, d.H, d.WC);
  __syncthreads();
  float acc[1] = {0.f};
  {
    // a thread owns one flat column and walks down half of the tile's rows with the window in registers: each
    // step loads only the new bottom row (6 shared-memory reads per window instead of 18)
    constexpr int kRowsPer = kSsTY / (kSsThreads / kSsTX);
    const int c = threadIdx.x % kSsTX, r0 = (threadIdx.x / kSsTX) * kRowsPer;
    const int wf = f0 + c, C = d.C;
    float xv[9], yv[9];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      xv[3 + j] = tx[r0 * kSsCfW + c + j * C];       yv[3 + j] = ty[r0 * kSsCfW + c + j * C];
      xv[6 + j] = tx[(r0 + 1) * kSsCfW + c + j * C]; yv[6 + j] = ty[(r0 + 1) * kSsCfW + c + j * C];
    }
#pragma unroll
    for (int k = 0; k < kRowsPer; ++k) {
      const int r = r0 + k, wi = i0 + r;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        xv[j] = xv[3 + j]; xv[3 + j] = xv[6 + j]; xv[6 + j] = tx[(r + 2) * kSsCfW + c + j * C];
        yv[j] = yv[3 + j]; yv[3 + j] = yv[6 + j]; yv[6 + j] = ty[(r + 2) * kSsCfW + c + j * C];
      }
      if (wi < d.H - 2 && wf < d.OWC) {
        const SsimWin w = ssim_stats<false>(xv, yv);
        const float v = fminf(fmaxf(0.5f * (1.0f - w.S), 0.f), 1.f);
        if (map != nullptr) map[((size_t)b * (d.H - 2) + wi) * d.OWC + wf] = v;
        acc[0] += v;
      }
    }
  }
  acc[0] *= d.inv_n;
  if (partial != nullptr)
    block_sum<1>(acc, scratch, partial + ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
}

__global__ void __launch_bounds__(kSsThreads)
ssim_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, SsimDims d,
                const float* __restrict__ g_map, const float* __restrict__ g_loss, int mean_path,
                float* __restrict__ g_x, float* __restrict__ g_y) {
  __shared__ float tx[kSsInH * kSsInW], ty[kSsInH * kSsInW];
  __shared__ float cax[kSsCfH * kSsCfW], cay[kSsCfH * kSsCfW], cb[kSsCfH * kSsCfW], cg[kSsCfH * kSsCfW];
  const int b = blockIdx.z, i0 = blockIdx.y * kSsTY, f0 = blockIdx.x * kSsTX;
  const int C = d.C;
  const size_t img = (size_t)b * d.H * d.WC;
  // pixels of rows [i0, i0+8), flat columns [f0, f0+128): windows of rows [i0-2, i0+8), columns [f0-2C, f0+128),
  // inputs of rows [i0-2, i0+10), columns [f0-2C, f0+128+2C)
  ssim_load_tile<kSsInW, kSsInH>(tx, x + img, i0 - 2, f0 - 2 * C, kSsTX + 4 * C, d.H, d.WC);
  ssim_load_tile<kSsInW, kSsInH>(ty, y + img, i0 - 2, f0 - 2 * C, kSsTX + 4 * C, d.H, d.WC);
  __syncthreads();
  const float g_mean = mean_path ? (g_loss != nullptr ? g_loss[0] : 1.0f) * d.inv_n : 0.f;
  const int wcols = kSsTX + 2 * C;
  for (int e = threadIdx.x; e < kSsCfH * kSsCfW; e += kSsThreads) {
    const int r = e / kSsCfW, c = e - r * kSsCfW;
    if (c >= wcols) continue;
    const int wi = i0 - 2 + r, wf = f0 - 2 * C + c;
    float ax = 0.f, ay = 0.f, be = 0.f, ga = 0.f;
    if ((unsigned)wi < (unsigned)(d.H - 2) && (unsigned)wf < (unsigned)d.OWC) {
      const SsimWin w = ssim_window<true>(tx + r * kSsInW + c, ty + r * kSsInW + c, kSsInW, C);
      const float v = 0.5f * (1.0f - w.S);
      float G = g_mean;
      if (g_map != nullptr) G += g_map[((size_t)b * (d.H - 2) + wi) * d.OWC + wf];
      // d clip(v, 0, 1) / dS = -1/2 inside the closed interval, 0 outside
      const float k = (v >= 0.f && v <= 1.f) ? -0.5f * G : 0.f;
      ax = k * w.ax; ay = k * w.ay; be = k * w.beta; ga = k * w.gamma;
    }
    const int o = r * kSsCfW + c;
    cax[o] = ax; cay[o] = ay; cb[o] = be; cg[o] = ga;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kSsTY * kSsTX; e += kSsThreads) {
    const int r = e / kSsTX, c = e - r * kSsTX;
    const int pi = i0 + r, pf = f0 + c;
    if (pi >= d.H || pf >= d.WC) continue;
    // pixel (r, c) sits at window-tile position (r + 2, c + 2C); its windows: rows r..r+2, columns c, c+C, c+2C
    float sax = 0.f, say = 0.f, sb = 0.f, sg = 0.f;
#pragma unroll
    for (int di = 0; di < 3; ++di)
#pragma unroll
      for (int dj = 0; dj < 3; ++dj) {
        const int o = (r + di) * kSsCfW + c + dj * C;
        sax += cax[o]; say += cay[o]; sb += cb[o]; sg += cg[o];
      }
    const float xv = tx[(r + 2) * kSsInW + c + 2 * C], yv = ty[(r + 2) * kSsInW + c + 2 * C];
    const size_t o = img + (size_t)pi * d.WC + pf;
    if (g_x != nullptr) g_x[o] = sax + sb * xv + sg * yv;
    if (g_y != nullptr) g_y[o] = say + sb * yv + sg * xv;
  }
}

// =====================================================================================================
// Edge-aware first-order smoothness: mean(|d_x disp| * exp(-mean_c |d_x img|)) + the same along y.
// One thread per pixel; the gradients are gather stencils (deterministic).
// =====================================================================================================
struct EdgeDims {
  int B, H, W, C;
  float cx, cy, inv_c;   // 1 / (B H (W-1)), 1 / (B (H-1) W), 1 / C
};

// edge weight between pixel p and its neighbour at +step (in pixels): exp(-mean_c |I(p+step) - I(p)|)
VSL_DEV float edge_weight(const float* __restrict__ img, size_t p, size_t step, int C, float inv_c) {
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += fabsf(img[(p + step) * C + c] - img[p * C + c]);
  return expf(-s * inv_c);
}

__global__ void __launch_bounds__(256)
edge_smooth_fwd_kernel(const float* __restrict__ disp, const float* __restrict__ img, EdgeDims d,
                       float* __restrict__ partial) {
  __shared__ float scratch[8];
  const size_t n = (size_t)d.B * d.H * d.W;
  float acc[1] = {0.f};
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(p % d.W), i = (int)((p / d.W) % d.H);
    if (j + 1 < d.W) acc[0] += d.cx * fabsf(disp[p + 1] - disp[p]) * edge_weight(img, p, 1, d.C, d.inv_c);
    if (i + 1 < d.H) acc[0] += d.cy * fabsf(disp[p + d.W] - disp[p]) * edge_weight(img, p, d.W, d.C, d.inv_c);
  }
  block_sum<1>(acc, scratch, partial + blockIdx.x);
}

__global__ void __launch_bounds__(256)
edge_smooth_bwd_kernel(const float* __restrict__ disp, const float* __restrict__ img, EdgeDims d,
                       const float* __restrict__ g_loss, float* __restrict__ g_disp, float* __restrict__ g_img) {
  const size_t n = (size_t)d.B * d.H * d.W;
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int j = (int)(p % d.W), i = (int)((p / d.W) % d.H);
  const float g = g_loss != nullptr ? g_loss[0] : 1.0f;
  // the (up to) four differences this pixel takes part in: (p, p+1), (p-1, p), (p, p+W), (p-W, p)
  const bool has[4] = {j + 1 < d.W, j > 0, i + 1 < d.H, i > 0};
  const size_t lo[4] = {p, p - 1, p, p - (size_t)d.W};          // lower element of the pair (only used when has[k])
  const size_t st[4] = {1, 1, (size_t)d.W, (size_t)d.W};
  const float sg_self[4] = {-1.f, 1.f, -1.f, 1.f};              // d(diff)/d(this pixel)
  const float cw[4] = {d.cx, d.cx, d.cy, d.cy};
  float gd = 0.f;
  float gi[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < 4; ++k) {
    if (!has[k]) continue;
    const float dd = disp[lo[k] + st[k]] - disp[lo[k]];
    const float w = edge_weight(img, lo[k], st[k], d.C, d.inv_c);
    gd += cw[k] * sg_self[k] * sgn(dd) * w;
    if (g_img != nullptr) {
      const float t = -cw[k] * fabsf(dd) * w * d.inv_c;         // d/d(mean_c |dI|) spread over the channels
      for (int c = 0; c < d.C && c < 4; ++c)
        gi[c] += t * sg_self[k] * sgn(img[(lo[k] + st[k]) * d.C + c] - img[lo[k] * d.C + c]);
    }
  }
  g_disp[p] = gd * g;
  if (g_img != nullptr)
    for (int c = 0; c < d.C && c < 4; ++c) g_img[p * d.C + c] = gi[c] * g;
}



}  // namespace vsl

using namespace vsl;

extern "C" {

static SsimDims ssim_dims(int B, int H, int W, int C) {
  SsimDims d;
  d.B = B; d.H = H; d.W = W; d.C = C;
  d.WC = W * C; d.OWC = (W - 2) * C;
  d.inv_n = (float)(1.0 / ((double)B * (H - 2) * (W - 2) * C));
  return d;
}
static dim3 ssim_grid(int B, int rows, int flat_cols) {
  return dim3((unsigned)((flat_cols + kSsTX - 1) / kSsTX), (unsigned)((rows + kSsTY - 1) / kSsTY), (unsigned)B);
}
static int ssim_check(int B, int H, int W, int C) {
  VSL_REQUIRE(B > 0 && B <= 65535 && H >= 3 && W >= 3 && C >= 1 && C <= kSsMaxC, VSL_E_SHAPE);
  VSL_REQUIRE((long long)W * C < (1LL << 30) && (H + kSsTY - 1) / kSsTY <= 65535, VSL_E_SHAPE);
  return VSL_OK;
}

size_t vsl_ssim_ws_bytes(int B, int H, int W, int C) {
  if (ssim_check(B, H, W, C) != VSL_OK) return 0;
  const dim3 g = ssim_grid(B, H - 2, (W - 2) * C);
  return sizeof(float) * (size_t)g.x * g.y * g.z;
}

int vsl_ssim_fwd(const float* x, const float* y, int B, int H, int W, int C, float* map, float* loss, void* ws,
                 vsl_stream_t stream) {
  VSL_REQUIRE(x && y && (map || loss), VSL_E_NULL);
  VSL_REQUIRE(!loss || ws, VSL_E_NULL);
  const int rc = ssim_check(B, H, W, C);
  if (rc != VSL_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 g = ssim_grid(B, H - 2, (W - 2) * C);
  const long long nblocks = (long long)g.x * g.y * g.z;
  VSL_REQUIRE(nblocks < (1LL << 31), VSL_E_SHAPE);
  ssim_fwd_kernel<<<g, kSsThreads, 0, st>>>(x, y, ssim_dims(B, H, W, C), map, loss ? (float*)ws : nullptr);
  if (loss) { const int es = launch_sum_partials((const float*)ws, (int)nblocks, loss, st); if (es != 0) return es; }
  return launch_status();
}

int vsl_ssim_bwd(const float* x, const float* y, int B, int H, int W, int C, const float* g_map,
                 const float* g_loss, int mean_path, float* g_x, float* g_y, vsl_stream_t stream) {
  VSL_REQUIRE(x && y && (g_x || g_y), VSL_E_NULL);
  VSL_REQUIRE(g_map || mean_path, VSL_E_NULL);
  const int rc = ssim_check(B, H, W, C);
  if (rc != VSL_OK) return rc;
  ssim_bwd_kernel<<<ssim_grid(B, H, W * C), kSsThreads, 0, (cudaStream_t)stream>>>(
      x, y, ssim_dims(B, H, W, C), g_map, g_loss, mean_path, g_x, g_y);
  return launch_status();
}

static EdgeDims edge_dims(int B, int H, int W, int C) {
  EdgeDims d;
  d.B = B; d.H = H; d.W = W; d.C = C;
  d.cx = (float)(1.0 / ((double)B * H * (W - 1)));
  d.cy = (float)(1.0 / ((double)B * (H - 1) * W));
  d.inv_c = 1.0f / (float)C;
  return d;
}
static const int kEdgeBlocks = 148 * 8;

size_t vsl_edge_smooth_ws_bytes(int, int, int) { return sizeof(float) * kEdgeBlocks; }

int vsl_edge_smooth_fwd(const float* disp, const float* img, int B, int H, int W, int C, float* loss, void* ws,
                        vsl_stream_t stream) {
  VSL_REQUIRE(disp && img && loss && ws, VSL_E_NULL);
  VSL_REQUIRE(B > 0 && H >= 2 && W >= 2 && C >= 1 && C <= 4, VSL_E_SHAPE);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)B * H * W;
  const int blocks = (int)((n + 255) / 256 < (size_t)kEdgeBlocks ? (n + 255) / 256 : kEdgeBlocks);
  edge_smooth_fwd_kernel<<<blocks, 256, 0, st>>>(disp, img, edge_dims(B, H, W, C), (float*)ws);
  { const int es = launch_sum_partials((const float*)ws, blocks, loss, st); if (es != 0) return es; }
  return launch_status();
}

int vsl_edge_smooth_bwd(const float* disp, const float* img, int B, int H, int W, int C, const float* g_loss,
                        float* g_disp, float* g_img, vsl_stream_t stream) {
  VSL_REQUIRE(disp && img && g_disp, VSL_E_NULL);
  VSL_REQUIRE(B > 0 && H >= 2 && W >= 2 && C >= 1 && C <= 4, VSL_E_SHAPE);
  const size_t n = (size_t)B * H * W;
  edge_smooth_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      disp, img, edge_dims(B, H, W, C), g_loss, g_disp, g_img);
  return launch_status();
}

}  // extern "C"
