// The 3 x 3 SSIM window statistics shared by the stand-alone SSIM op (vsl_ext.cu) and the SSIM term of the fused step
// (vsl_loss_ssim.cu).  Extension: SSIM is absent from the reference; oracle = oracle/vsl_oracle.py
// ssim_dissimilarity, parity unpinned.
#pragma once
#include "vsl_common.cuh"

namespace vsl {

struct SsimWin { float S, ax, ay, beta, gamma; };

// The 9 + 9 values of one window -> SSIM (and, with GRAD, the coefficients of its derivative).  FAST: one approximate
// reciprocal (MUFU.RCP, 1 ulp) of B1 B2 instead of two IEEE divisions -- the in-step term, whose bars are 1e-5.
template <bool GRAD, bool FAST = false>
VSL_DEV SsimWin ssim_stats(const float (&xv)[9], const float (&yv)[9]) {
  const float C1 = 1e-4f, C2 = 9e-4f, ninth = 1.0f / 9.0f;
  float sx = 0.f, sy = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { sx += xv[k]; sy += yv[k]; }
  const float mx = sx * ninth, my = sy * ninth;
  float vxx = 0.f, vyy = 0.f, vxy = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float dx = xv[k] - mx, dy = yv[k] - my;
    vxx = fmaf(dx, dx, vxx); vyy = fmaf(dy, dy, vyy); vxy = fmaf(dx, dy, vxy);
  }
  vxx *= ninth; vyy *= ninth; vxy *= ninth;
  const float A1 = 2.f * mx * my + C1, A2 = 2.f * vxy + C2;
  const float B1 = mx * mx + my * my + C1, B2 = vxx + vyy + C2;
  float i1, i2, inv;
  if (FAST) {
    inv = __fdividef(1.0f, B1 * B2);
    i1 = inv * B2; i2 = inv * B1;
  } else {
    i1 = 1.0f / B1; i2 = 1.0f / B2; inv = i1 * i2;
  }
  SsimWin w;
  w.S = A1 * A2 * inv;
  if (GRAD) {
    const float k = 2.0f * ninth;
    const float t = (A2 - A1) * inv, u = w.S * (i1 - i2);
    w.ax = k * (my * t - mx * u);
    w.ay = k * (mx * t - my * u);
    w.beta = -k * w.S * i2;
    w.gamma = k * A1 * inv;
  }
  return w;
}

}  // namespace vsl
