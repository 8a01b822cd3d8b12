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

VSL_DEV float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }  // tf.abs' gradient

}  // namespace vsl
