// tf.image.resize_area pyramids in one launch: shared by the stand-alone op (vsl_ops.cu) and the flow-and-depth
// loss step (vsl_flow.cu).
#pragma once
#include <algorithm>

#include "vsl_prep.cuh"

namespace vsl {

// =====================================================================================================
// resize_area pyramid: ONE launch reads level 0 of up to VSL_MAX_VIEWS+1 images once and writes every
// coarser level.  One THREAD owns one F x F block of level-0 pixels (F = 2^(S-1)), all channels: it streams
// the block row by row with 16-byte loads (the F*C floats of a block row are contiguous), keeps the running
// sums of every level in registers and stores each coarser element the moment its last row has been added.
// No shared memory, no barriers, all offsets compile-time.
// Summation order = TF's ResizeArea (ComputePatchSum) with unit overlap weights: for every output element
// the f level-0 values of each contributing row are summed left to right, the f row sums are then
// accumulated top to bottom, and the total is multiplied by 1/f^2 -- always from LEVEL-0 values, never from
// a coarser level, hence bit-exact against the oracle.
// The first few blocks also fill the per-(scale, view, batch) transform table when a PrepJob rides along.
// =====================================================================================================
struct PyrJob {
  const float* img[VSL_MAX_VIEWS + 1];
  float* lvl[VSL_MAX_VIEWS + 1][VSL_MAX_SCALES];  // [image][scale], scale 0 unused
  int nimg;
};

// level-s sums of one block row: out[j*C + c] = a[(j*f)*C + c] + ... + a[(j*f + f-1)*C + c], left to right
template <int C, int F, int SHIFT>
VSL_DEV void row_sums(const float* a, float* out) {
  constexpr int f = 1 << SHIFT;
#pragma unroll
  for (int j = 0; j < F / f; ++j)
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float acc = a[(j * f) * C + c];
#pragma unroll
      for (int k = 1; k < f; ++k) acc = __fadd_rn(acc, a[(j * f + k) * C + c]);
      out[j * C + c] = acc;
    }
}

template <int C, int LOG2F, int SHIFT>
struct PyrLevel {
  static constexpr int F = 1 << LOG2F, f = 1 << SHIFT, n = (F / f) * C;
  float acc[n];
  // add block row `r` (0-based) of level-0 values `a`; emit the finished output row when r closes a band
  VSL_DEV void add_row(const float* a, int r, float* __restrict__ dst_row0, int Ws) {
    float rs[n];
    row_sums<C, F, SHIFT>(a, rs);
    if ((r & (f - 1)) == 0) {
#pragma unroll
      for (int i = 0; i < n; ++i) acc[i] = rs[i];
    } else {
#pragma unroll
      for (int i = 0; i < n; ++i) acc[i] = __fadd_rn(acc[i], rs[i]);
    }
    if ((r & (f - 1)) == f - 1) {
      const float scale = 1.0f / (float)(f * f);
      float* __restrict__ d = dst_row0 + (size_t)(r >> SHIFT) * Ws * C;
      if ((n % 4 == 0) && ((reinterpret_cast<uintptr_t>(d) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < n; i += 4)
          *reinterpret_cast<float4*>(d + i) = make_float4(__fmul_rn(acc[i], scale), __fmul_rn(acc[i + 1], scale),
                                                          __fmul_rn(acc[i + 2], scale), __fmul_rn(acc[i + 3], scale));
      } else {
#pragma unroll
        for (int i = 0; i < n; ++i) d[i] = __fmul_rn(acc[i], scale);
      }
    }
  }
};

template <int C, int LOG2F>
__global__ void __launch_bounds__(128)
pyramid_kernel(const PyrJob job, const PrepJob prep, int B, int H, int W) {
  constexpr int F = 1 << LOG2F, NF = F * C;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (prep.n > 0 && gid < prep.n) prep_one(prep, gid);
  const int nbx = W >> LOG2F, nby = H >> LOG2F;
  const int per_img = nbx * nby;
  const int total = job.nimg * B * per_img;
  if (gid >= total) return;
  const int ib = gid / per_img, rem = gid - ib * per_img;   // ib = image * B + b
  const int by = rem / nbx, bx = rem - by * nbx;
  const int im = ib / B, b = ib - im * B;
  const float* __restrict__ src = job.img[im] + (((size_t)b * H + (size_t)by * F) * W + (size_t)bx * F) * C;
  const bool vec = (NF % 4 == 0) && ((W * C) % 4 == 0) && ((reinterpret_cast<uintptr_t>(job.img[im]) & 15) == 0);

  // destination of this block's first output row at every level
  PyrLevel<C, LOG2F, 1> l1;
  PyrLevel<C, LOG2F, (LOG2F >= 2 ? 2 : 1)> l2;
  PyrLevel<C, LOG2F, (LOG2F >= 3 ? 3 : 1)> l3;
  PyrLevel<C, LOG2F, (LOG2F >= 4 ? 4 : 1)> l4;
  PyrLevel<C, LOG2F, (LOG2F >= 5 ? 5 : 1)> l5;
  auto dst0 = [&](int s) -> float* {
    const int Hs = H >> s, Ws = W >> s;
    return job.lvl[im][s] + (((size_t)b * Hs + (size_t)by * (F >> s)) * Ws + (size_t)bx * (F >> s)) * C;
  };
  float* d1 = dst0(1);
  float* d2 = LOG2F >= 2 ? dst0(2) : nullptr;
  float* d3 = LOG2F >= 3 ? dst0(3) : nullptr;
  float* d4 = LOG2F >= 4 ? dst0(4) : nullptr;
  float* d5 = LOG2F >= 5 ? dst0(5) : nullptr;

#pragma unroll
  for (int r = 0; r < F; ++r) {
    float a[NF];
    const float* __restrict__ row = src + (size_t)r * W * C;
    if (vec) {
#pragma unroll
      for (int k = 0; k < NF / 4; ++k) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row) + k);
        a[4 * k] = v.x; a[4 * k + 1] = v.y; a[4 * k + 2] = v.z; a[4 * k + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < NF; ++k) a[k] = __ldg(row + k);
    }
    l1.add_row(a, r, d1, W >> 1);
    if (LOG2F >= 2) l2.add_row(a, r, d2, W >> 2);
    if (LOG2F >= 3) l3.add_row(a, r, d3, W >> 3);
    if (LOG2F >= 4) l4.add_row(a, r, d4, W >> 4);
    if (LOG2F >= 5) l5.add_row(a, r, d5, W >> 5);
  }
}

template <int C>
inline int launch_pyramid_c(const PyrJob& job, const PrepJob& prep, int B, int H, int W, int S, cudaStream_t st) {
  const int F = 1 << (S - 1);
  const long long threads = (long long)job.nimg * B * (H / F) * (W / F);
  const long long need = std::max<long long>(threads, prep.n);
  const unsigned grid = (unsigned)((need + 127) / 128);
  switch (S) {
    case 2: pyramid_kernel<C, 1><<<grid, 128, 0, st>>>(job, prep, B, H, W); break;
    case 3: pyramid_kernel<C, 2><<<grid, 128, 0, st>>>(job, prep, B, H, W); break;
    case 4: pyramid_kernel<C, 3><<<grid, 128, 0, st>>>(job, prep, B, H, W); break;
    case 5: pyramid_kernel<C, 4><<<grid, 128, 0, st>>>(job, prep, B, H, W); break;
    default: pyramid_kernel<C, 5><<<grid, 128, 0, st>>>(job, prep, B, H, W); break;
  }
  return launch_status();
}

inline int launch_pyramid(const PyrJob& job, const PrepJob& prep, int B, int H, int W, int C, int S, cudaStream_t st) {
  switch (C) {
    case 1: return launch_pyramid_c<1>(job, prep, B, H, W, S, st);
    case 2: return launch_pyramid_c<2>(job, prep, B, H, W, S, st);
    case 3: return launch_pyramid_c<3>(job, prep, B, H, W, S, st);
    default: return launch_pyramid_c<4>(job, prep, B, H, W, S, st);
  }
}

}  // namespace vsl
