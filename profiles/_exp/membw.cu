// Memory-system floors for the prep launch's traffic shape (cfg2): read R MB (streamed, evict-first), write Wr MB.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membw membw.cu ; run on the B200.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void k_write(float4* dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = make_float4(1.f, 2.f, 3.f, 0.f);
}
__global__ void k_write256(float* dst, size_t n8) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(dst + 8 * i), "f"(1.f) : "memory");
}
__global__ void k_read(const float4* src, size_t n, float* out) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(src + i);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 123.456f) *out = acc;
}
// 3 float4 in, 4 float4 out per thread-iteration (RGB -> RGBA shape), both coalesced
__global__ void k_rw(const float4* src, float4* dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = __ldcs(src + i), b = __ldcs(src + n + i), c = __ldcs(src + 2 * n + i);
    dst[i] = make_float4(a.x, a.y, a.z, 0.f);
    dst[n + i] = make_float4(a.w, b.x, b.y, 0.f);
    dst[2 * n + i] = make_float4(b.z, b.w, c.x, 0.f);
    dst[3 * n + i] = make_float4(c.y, c.z, c.w, 0.f);
  }
}
__global__ void k_flush(float4* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_float4(0, 0, 0, 0);
}

int main() {
  const size_t MB = 1 << 20;
  const size_t rd = 61 * MB, wr = 82 * MB, fl = 512 * MB;
  float4 *src, *dst, *flush; float* out;
  CK(cudaMalloc(&src, rd)); CK(cudaMalloc(&dst, wr)); CK(cudaMalloc(&flush, fl)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(src, 0, rd));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int blocks_per_sm = 2; blocks_per_sm <= 8; blocks_per_sm *= 2) {
    const int grid = 148 * blocks_per_sm, thr = 256;
    for (int which = 0; which < 4; ++which) {
      float best = 1e9f;
      for (int rep = 0; rep < 6; ++rep) {
        k_flush<<<1184, 256>>>(flush, fl / 16);
        cudaEventRecord(e0);
        if (which == 0) k_write<<<grid, thr>>>(dst, wr / 16);
        if (which == 1) k_write256<<<grid, thr>>>((float*)dst, wr / 32);
        if (which == 2) k_read<<<grid, thr>>>(src, rd / 16, out);
        if (which == 3) k_rw<<<grid, thr>>>(src, dst, rd / 48);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      const char* names[4] = {"write 82 MB (16B stores)", "write 82 MB (32B stores)", "read 61 MB", "read 61 + write 81 MB"};
      const double bytes = which < 2 ? (double)wr : which == 2 ? (double)rd : (double)rd * 7.0 / 3.0;
      printf("blocks/SM %d  %-28s %7.2f us  %6.2f TB/s\n", blocks_per_sm, names[which], best * 1e3, bytes / best / 1e9);
    }
  }
  return 0;
}
