// The step AFTER the path: tf.train.AdamOptimizer(lr, beta1).minimize(...) as the training
// scripts use it (train_depth_then_cam_lr.py:413, train.py:148) -- one launch over a FLAT arena holding every
// parameter, so a data-parallel step is "all-reduce a bucket of the gradient arena, then this kernel on the same
// range" (tf_depth_estimation_b200/dist.py).  The arithmetic is TensorFlow's ApplyAdam (un-vendored, unpinned TF 1.x:
// training/adam.py, kernels/training_ops.cc):
//     lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
//     m   += (g - m) * (1 - beta1);  v += (g*g - v) * (1 - beta2);  p -= lr_t * m / (sqrt(v) + eps)
// (epsilon outside the bias correction, unlike torch.optim.Adam).  HBM-bound: 16 B read + 12 B written per
// parameter; 16-byte accesses, gradient read evict-first (single use), grid = a multiple of the SM count.
#pragma once
#include "vsl_common.cuh"

namespace vsl {

struct AdamConsts { float lr_t, omb1, omb2, eps, gscale; };

VSL_DEV void adam_one(float& p, float g, float& m, float& v, const AdamConsts& c) {
  g *= c.gscale;
  m = fmaf(g - m, c.omb1, m);
  v = fmaf(fmaf(g, g, -v), c.omb2, v);
  p -= c.lr_t * m / (sqrtf(v) + c.eps);
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, long long head, AdamConsts c) {
  // [0, head): scalars up to the first 16-byte boundary; then float4s; then the tail
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  const long long n4 = (n - head) / 4;
  float4* p4 = reinterpret_cast<float4*>(p + head);
  const float4* g4 = reinterpret_cast<const float4*>(g + head);
  float4* m4 = reinterpret_cast<float4*>(m + head);
  float4* v4 = reinterpret_cast<float4*>(v + head);
  for (long long i = tid; i < n4; i += nth) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = __ldcs(g4 + i);
    adam_one(pp.x, gg.x, mm.x, vv.x, c); adam_one(pp.y, gg.y, mm.y, vv.y, c);
    adam_one(pp.z, gg.z, mm.z, vv.z, c); adam_one(pp.w, gg.w, mm.w, vv.w, c);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  const long long rest = head + (n - head - 4 * n4);   // scalar elements: [0, head) and [head + 4 n4, n)
  for (long long k = tid; k < rest; k += nth) {
    const long long i = k < head ? k : head + 4 * n4 + (k - head);
    adam_one(p[i], g[i], m[i], v[i], c);
  }
}

}  // namespace vsl

using namespace vsl;

extern "C" {

int vsl_adam_step(float* param, const float* grad, float* m, float* v, long long n, float lr, float beta1,
                  float beta2, float eps, int step, float grad_scale, vsl_stream_t stream) {
  VSL_REQUIRE(param && grad && m && v, VSL_E_NULL);
  VSL_REQUIRE(n > 0 && step >= 1, VSL_E_SHAPE);
  VSL_REQUIRE(aligned(param, 4) && aligned(grad, 4) && aligned(m, 4) && aligned(v, 4), VSL_E_ALIGN);
  // the four ranges must share their phase inside a 16-byte line for the vector body (true for ranges cut at the
  // same offset out of 16-byte aligned arenas)
  const uintptr_t ph = reinterpret_cast<uintptr_t>(param) % 16;
  VSL_REQUIRE(reinterpret_cast<uintptr_t>(grad) % 16 == ph && reinterpret_cast<uintptr_t>(m) % 16 == ph &&
                  reinterpret_cast<uintptr_t>(v) % 16 == ph, VSL_E_ALIGN);
  long long head = ph == 0 ? 0 : (long long)((16 - ph) / 4);
  if (head > n) head = n;
  AdamConsts c;
  c.lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step)));
  c.omb1 = 1.0f - beta1; c.omb2 = 1.0f - beta2; c.eps = eps; c.gscale = grad_scale;
  const long long want = ((n - head) / 4 + 255) / 256 + 1;
  const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, m, v, n, head, c);
  return launch_status();
}

}  // extern "C"
