// Host/device odds and ends shared by the translation units of libvsl.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vsl.h"
#include "vsl_geom.cuh"

namespace vsl {

#define VSL_REQUIRE(cond, code) \
  do {                          \
    if (!(cond)) return (code); \
  } while (0)

// Launch errors come back as positive cudaError_t values; nothing is thrown across the C ABI.
inline int launch_status() {
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? VSL_OK : (int)e;
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
inline size_t round_up(size_t n, size_t a) { return (n + a - 1) / a * a; }

VSL_DEV float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum N per-thread values over the block; the totals land in out[0..N) of thread 0 only.
// `scratch` holds N * (blockDim.x / 32) floats.  Fixed order => deterministic.
template <int N>
VSL_DEV void block_sum(float (&v)[N], float* scratch, float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float s = warp_sum(v[i]);
    if (lane == 0) scratch[warp * N + i] = s;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float s = 0.f;
    for (int w = 0; w < nwarp; ++w) s += scratch[w * N + threadIdx.x];
    out[threadIdx.x] = s;
  }
}

// Butterfly ("transpose") reduction of N per-thread values: every step halves the number of values a lane
// carries by trading one half for the partner's other half, so the warp needs ~N shuffles instead of 5 N.
// After the five steps lane l holds BflySizes<N>::h5 slots; bfly_index<N>(l, j) names the value in slot j.
template <int SZ, int OFF>
VSL_DEV void bfly_step(float* v, int lane) {
  constexpr int H = (SZ + 1) / 2;
  const bool up = (lane & OFF) != 0;  // upper lanes keep [H, SZ), lower lanes keep [0, H)
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const float lo = v[i], hi = (i + H < SZ) ? v[i + H] : 0.f;
    const float keep = up ? hi : lo, send = up ? lo : hi;
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}
template <int N> struct BflySizes {
  static constexpr int h1 = (N + 1) / 2, h2 = (h1 + 1) / 2, h3 = (h2 + 1) / 2, h4 = (h3 + 1) / 2, h5 = (h4 + 1) / 2;
};
// Value index of slot j of this lane after the five steps, or -1 if the slot is padding (a level of odd
// size gives its upper half one slot less).
template <int N>
VSL_DEV int bfly_index(int lane, int j) {
  using Z = BflySizes<N>;
  int li = j + ((lane & 1) ? Z::h5 : 0);
  bool ok = li < Z::h4;
  li += (lane & 2) ? Z::h4 : 0; ok = ok && li < Z::h3;
  li += (lane & 4) ? Z::h3 : 0; ok = ok && li < Z::h2;
  li += (lane & 8) ? Z::h2 : 0; ok = ok && li < Z::h1;
  li += (lane & 16) ? Z::h1 : 0; ok = ok && li < N;
  return ok ? li : -1;
}
// Block sum with the butterfly inside each warp; totals land in out[0..N) (written by threads < N).
// `scratch` holds N * (blockDim.x / 32) floats.  Fixed order => deterministic.
template <int N>
VSL_DEV void block_sum_bfly(float (&v)[N], float* scratch, float* out) {
  using Z = BflySizes<N>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  bfly_step<N, 16>(v, lane);
  bfly_step<Z::h1, 8>(v, lane);
  bfly_step<Z::h2, 4>(v, lane);
  bfly_step<Z::h3, 2>(v, lane);
  bfly_step<Z::h4, 1>(v, lane);
#pragma unroll
  for (int j = 0; j < Z::h5; ++j) {
    const int idx = bfly_index<N>(lane, j);
    if (idx >= 0) scratch[warp * N + idx] = v[j];
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float s = 0.f;
    for (int w = 0; w < nwarp; ++w) s += scratch[w * N + threadIdx.x];
    out[threadIdx.x] = s;
  }
}

// Sums `n` partials in double in a fixed order; out[0] = result.  One block.
static __global__ void sum_partials_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  __shared__ double sh[256];
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched programmatically dependent on the producer of `partial`
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)sh[0];
}

// The partial-sum launch behind a producer kernel on the same stream: programmatic dependent launch, so its block is
// resident (and past its launch latency) when the producer's last block retires; it waits inside (griddepcontrol.wait).
static inline int launch_sum_partials(const float* partial, int n, float* out, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, sum_partials_kernel, partial, n, out);
}

VSL_DEV float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }  // tf.abs' gradient

}  // namespace vsl
