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
#include <stdlib.h>
#include <string.h>

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


// =====================================================================================================
// Data-parallel step as ONE kernel over NVLink peer memory (instead of NCCL all-reduce -> Adam):
// rank r owns the contiguous shard [lo, hi) of the flat arena.  For its shard it
//   1. loads the gradient of EVERY rank (its own from HBM, the others' by peer-to-peer 16-byte loads over
//      NVLink / NVSwitch) and sums them in rank order            -- the reduce-scatter;
//   2. applies Adam with moments that exist for the shard only   -- 8 B/param of optimiser state / world;
//   3. stores the new parameters into every rank's arena (peer-to-peer 16-byte stores) -- the all-gather.
// Only the owner ever computes an element, so all replicas hold bit-identical parameters, and the sum order is
// fixed => deterministic.  Per rank and step the links carry (world-1)/world of the arena in and out -- what a
// ring all-reduce moves, in one pass and with no intermediate buffer -- and HBM sees 1/world of Adam's traffic.
// Two flag barriers (peer_barrier_kernel) order it against the ranks' gradient producers and parameter readers.
// =====================================================================================================
constexpr int kMaxPeers = 16;
struct PeerPtrs {
  const float* grad[kMaxPeers];
  float* param[kMaxPeers];
};

// Hyper-parameters as the host passes them; lr_t is formed on the device when the step count lives there.
struct AdamHyper { float lr, beta1, beta2, eps, gscale; int step; };

VSL_DEV AdamConsts adam_consts(const AdamHyper& h, int t) {
  AdamConsts c;
  c.lr_t = (float)((double)h.lr * sqrt(1.0 - pow((double)h.beta2, (double)t)) / (1.0 - pow((double)h.beta1, (double)t)));
  c.omb1 = 1.0f - h.beta1; c.omb2 = 1.0f - h.beta2; c.eps = h.eps; c.gscale = h.gscale;
  return c;
}

// state (nullable): device int[4] = {barrier epoch, Adam step count t, timed-out mirror, -}, advanced by peer_barrier_kernel, so
// that the whole barrier -> step -> barrier sequence takes no per-step host argument and replays from a CUDA graph.
// timed_out (nullable; may be pinned host memory): set by a barrier that gave up -- then this kernel does NOTHING
// (no update, no peer store): a late peer must never be summed half-written (the step fails instead, see dist.py).
#ifdef VSL_DP_TIMING_EXPERIMENTS
#define VSL_DP_DBG(x) (x)
#else
#define VSL_DP_DBG(x) 0
#endif
template <int WORLD>   // 0 = run-time world
__global__ void __launch_bounds__(256)
dp_adam_kernel(PeerPtrs pp, int world_rt, int rank, float* __restrict__ m, float* __restrict__ v, long long lo,
               long long n4, AdamHyper hy, const int* __restrict__ state, const volatile int* timed_out, int dbg) {
  constexpr int U = 4;    // float4s per thread and trip: U * world peer loads in flight before the first use
  const int world = WORLD ? WORLD : world_rt;
  __shared__ AdamConsts sc;
  __shared__ int s_dead;
  if (threadIdx.x == 0) {
    // with a state block the failure flag is mirrored in device memory (state[2]): no PCIe read on this kernel's path
    s_dead = state != nullptr ? (state[2] != 0) : ((timed_out != nullptr && *timed_out != 0) ? 1 : 0);
    sc = adam_consts(hy, state != nullptr ? state[1] + 1 : hy.step);
  }
  __syncthreads();
  if (s_dead) return;
  const AdamConsts c = sc;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const float4* own_p = reinterpret_cast<const float4*>(pp.param[rank]);
  const long long lo4 = lo / 4;
  for (long long i0 = tid; i0 < n4; i0 += nth * U) {
    float4 g[U], p[U], mm[U], vv[U];
    // the far loads first (peers in rank order, this rank's own gradient among them) ...
#pragma unroll
    for (int u = 0; u < U; ++u) g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < (WORLD ? WORLD : kMaxPeers); ++r) {
      if (r < world) {
        const float4* src = reinterpret_cast<const float4*>(pp.grad[VSL_DP_DBG(dbg & 2) ? rank : r]) + lo4;
        float4 t[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long i = i0 + u * nth;
          t[u] = i < n4 ? __ldcs(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { g[u].x += t[u].x; g[u].y += t[u].y; g[u].z += t[u].z; g[u].w += t[u].w; }
      }
    }
    // ... then this rank's parameters and moments
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * nth;
      if (i < n4) { p[u] = own_p[lo4 + i]; mm[u] = m4[i]; vv[u] = v4[i]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * nth;
      if (i < n4) {
        adam_one(p[u].x, g[u].x, mm[u].x, vv[u].x, c); adam_one(p[u].y, g[u].y, mm[u].y, vv[u].y, c);
        adam_one(p[u].z, g[u].z, mm[u].z, vv[u].z, c); adam_one(p[u].w, g[u].w, mm[u].w, vv[u].w, c);
        m4[i] = mm[u]; v4[i] = vv[u];
#pragma unroll
        for (int r = 0; r < (WORLD ? WORLD : kMaxPeers); ++r)
          if (r < world && (!VSL_DP_DBG(dbg & 1) || r == rank)) reinterpret_cast<float4*>(pp.param[r])[lo4 + i] = p[u];
      }
    }
  }
}

// The same step through the NVSwitch's multicast engine (NVLS): the gradient and parameter arenas of all ranks are
// bound to one multicast object, and rank r touches its shard through the multicast address only --
//   1. multimem.ld_reduce.add.v4.f32: ONE load returns the sum over every rank's copy, added inside the switch
//      -- the reduce-scatter without world - 1 separate pulls;
//   2. Adam on the shard;
//   3. multimem.st.v4.f32: ONE store lands in every rank's parameter arena, replicated by the switch -- the all-gather.
// Per rank the links carry about one arena in and one out instead of 2 (world - 1) / world each way.  The in-switch
// sum has its own (fixed) association order: deterministic, replicas bit-identical (one owner per element), but
// not bit-identical to the rank-ordered sum of dp_adam_kernel.
// Blocks per SM of the multicast form: the reduce phase loads the links outbound (every GPU feeds the switch its copy
// of every shard), the broadcast phase inbound -- with the whole shard in flight as ONE wave the phases run one after
// the other on every GPU at once; fewer resident blocks looping over the shard let the stores of one trip overlap the
// loads of the next.
constexpr int kMcBlocksPerSm = 8;
__global__ void __launch_bounds__(256)
dp_adam_mc_kernel(const float* __restrict__ mc_grad, float* __restrict__ mc_param, const float* __restrict__ own_param,
                  float* __restrict__ m, float* __restrict__ v, long long lo, long long n4, AdamHyper hy,
                  const int* __restrict__ state, const volatile int* timed_out) {
  constexpr int U = 4;
  __shared__ AdamConsts sc;
  __shared__ int s_dead;
  if (threadIdx.x == 0) {
    s_dead = state != nullptr ? (state[2] != 0) : ((timed_out != nullptr && *timed_out != 0) ? 1 : 0);
    sc = adam_consts(hy, state != nullptr ? state[1] + 1 : hy.step);
  }
  __syncthreads();
  if (s_dead) return;
  const AdamConsts c = sc;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const long long lo4 = lo / 4;
  const float4* own_p = reinterpret_cast<const float4*>(own_param) + lo4;
  const float4* g_mc = reinterpret_cast<const float4*>(mc_grad) + lo4;
  float4* p_mc = reinterpret_cast<float4*>(mc_param) + lo4;
  for (long long i0 = tid; i0 < n4; i0 += nth * U) {
    float4 g[U], p[U], mm[U], vv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * nth;
      g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < n4)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(g[u].x), "=f"(g[u].y), "=f"(g[u].z), "=f"(g[u].w) : "l"(g_mc + i) : "memory");
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * nth;
      if (i < n4) { p[u] = own_p[i]; mm[u] = m4[i]; vv[u] = v4[i]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * nth;
      if (i < n4) {
        adam_one(p[u].x, g[u].x, mm[u].x, vv[u].x, c); adam_one(p[u].y, g[u].y, mm[u].y, vv[u].y, c);
        adam_one(p[u].z, g[u].z, mm[u].z, vv[u].z, c); adam_one(p[u].w, g[u].w, mm[u].w, vv[u].w, c);
        m4[i] = mm[u]; v4[i] = vv[u];
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p_mc + i), "f"(p[u].x),
                     "f"(p[u].y), "f"(p[u].z), "f"(p[u].w) : "memory");
      }
    }
  }
}

// Barrier across the GPUs of a node through peer-mapped flag words.  flags of rank q: unsigned[kMaxPeers], word r is
// written by rank r only.  One block of `world` threads: thread r publishes `epoch` into rank r's word [rank], then
// waits until this GPU's word [r] has reached `epoch`.  Everything this GPU wrote before (peer stores of the
// previous kernel included) is fenced system-wide first.  The spin is bounded in wall-clock time (%globaltimer):
// if a peer does not arrive within timeout_ns the kernel sets *timed_out -- which makes every later dp_adam_kernel
// a no-op and the host-side step raise -- instead of hanging the GPU or letting the step run on half-written data.
// With `state` the epoch is state[0] + 1 (and is stored back; bump_step also advances the Adam step count
// state[1]), so the launch takes no per-step host argument.
struct PeerFlags { unsigned* f[kMaxPeers]; };

VSL_DEV unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void peer_barrier_kernel(PeerFlags pf, int rank, int world, unsigned epoch_arg, int* state, int bump_step,
                                    volatile int* timed_out, unsigned long long timeout_ns) {
  const int r = threadIdx.x;
  const unsigned epoch = state != nullptr ? (unsigned)state[0] + 1u : epoch_arg;
  __syncthreads();
  if (r < world) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.f[r] + rank), "r"(epoch) : "memory");
    const unsigned* mine = pf.f[rank] + r;
    const unsigned long long t0 = global_ns();
    unsigned seen;
    unsigned spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
      if ((int)(seen - epoch) >= 0) break;
      if ((++spins & 1023u) == 0u) {
        if (state != nullptr ? (*(volatile int*)(state + 2) != 0) : (timed_out != nullptr && *timed_out != 0))
          break;                                                     // a barrier of this rank already failed
        if (global_ns() - t0 > timeout_ns) {
          if (timed_out != nullptr) *timed_out = 1;
          if (state != nullptr) atomicExch(state + 2, 1);
          break;
        }
      }
    } while (true);
    __threadfence_system();
  }
  __syncthreads();
  if (r == 0 && state != nullptr) {
    state[0] = (int)epoch;
    if (bump_step) state[1] += 1;
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

// ---- peer arenas: the one place this library allocates, because memory that other processes map has to come
// straight from cudaMalloc (an IPC handle names a whole allocation) and must outlive every mapping of it.
int vsl_peer_alloc(size_t bytes, void** ptr) {
  VSL_REQUIRE(ptr, VSL_E_NULL);
  VSL_REQUIRE(bytes > 0, VSL_E_SHAPE);
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(*ptr, 0, bytes);
  return e == cudaSuccess ? VSL_OK : (int)e;
}
int vsl_peer_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  return e == cudaSuccess ? VSL_OK : (int)e;
}
int vsl_ipc_get_handle(void* ptr, unsigned char* handle64) {
  VSL_REQUIRE(ptr && handle64, VSL_E_NULL);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) return (int)e;
  memcpy(handle64, &h, 64);
  return VSL_OK;
}
// Opened with the CURRENT device as the accessor: the peer's allocation is mapped for kernels of this GPU
// (cudaIpcMemLazyEnablePeerAccess), which a mapping made under the owner's device would not be.
int vsl_ipc_open(const unsigned char* handle64, void** ptr) {
  VSL_REQUIRE(handle64 && ptr, VSL_E_NULL);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  return e == cudaSuccess ? VSL_OK : (int)e;
}
int vsl_ipc_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  return e == cudaSuccess ? VSL_OK : (int)e;
}

namespace {

constexpr long long kDefaultBarrierTimeoutMs = 120000;   // 2 minutes: rank skew (checkpoint, loader stall) is not a fault

int launch_barrier(unsigned* const* peer_flags, int rank, int world, unsigned epoch, int* state, int bump_step,
                   int* timed_out, long long timeout_ms, cudaStream_t st) {
  VSL_REQUIRE(peer_flags, VSL_E_NULL);
  VSL_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world && timeout_ms >= 0, VSL_E_SHAPE);
  PeerFlags pf;
  for (int r = 0; r < kMaxPeers; ++r) pf.f[r] = r < world ? peer_flags[r] : nullptr;
  for (int r = 0; r < world; ++r) VSL_REQUIRE(pf.f[r] != nullptr, VSL_E_NULL);
  const unsigned long long ns = (unsigned long long)(timeout_ms > 0 ? timeout_ms : kDefaultBarrierTimeoutMs) * 1000000ull;
  peer_barrier_kernel<<<1, 32, 0, st>>>(pf, rank, world, epoch, state, bump_step, timed_out, ns);
  return launch_status();
}

int launch_dp_adam(const float* const* peer_grads, float* const* peer_params, int rank, int world, float* m_shard,
                   float* v_shard, long long lo, long long hi, const AdamHyper& hy, const int* state,
                   const int* timed_out, cudaStream_t st) {
  VSL_REQUIRE(peer_grads && peer_params && m_shard && v_shard, VSL_E_NULL);
  VSL_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world && (state != nullptr || hy.step >= 1),
              VSL_E_SHAPE);
  VSL_REQUIRE(lo >= 0 && hi >= lo && lo % 4 == 0 && hi % 4 == 0, VSL_E_SHAPE);
  PeerPtrs pp;
  for (int r = 0; r < kMaxPeers; ++r) {
    pp.grad[r] = r < world ? peer_grads[r] : nullptr;
    pp.param[r] = r < world ? peer_params[r] : nullptr;
    if (r < world) {
      VSL_REQUIRE(pp.grad[r] && pp.param[r], VSL_E_NULL);
      VSL_REQUIRE(aligned(pp.grad[r], 16) && aligned(pp.param[r], 16), VSL_E_ALIGN);
    }
  }
  VSL_REQUIRE(aligned(m_shard, 16) && aligned(v_shard, 16), VSL_E_ALIGN);
  if (hi == lo) return VSL_OK;
  const long long n4 = (hi - lo) / 4;
  const long long want = (n4 + 4 * 256 - 1) / (4 * 256);
  const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
  int dbg = 0;
#ifdef VSL_DP_TIMING_EXPERIMENTS   // never defined for the shipped library: 1 = no remote stores, 2 = no remote loads
  const char* de = getenv("VSL_DP_DEBUG");
  dbg = de ? atoi(de) : 0;
#endif
  switch (world) {
    case 1: dp_adam_kernel<1><<<blocks, 256, 0, st>>>(pp, world, rank, m_shard, v_shard, lo, n4, hy, state, timed_out, dbg); break;
    case 2: dp_adam_kernel<2><<<blocks, 256, 0, st>>>(pp, world, rank, m_shard, v_shard, lo, n4, hy, state, timed_out, dbg); break;
    case 4: dp_adam_kernel<4><<<blocks, 256, 0, st>>>(pp, world, rank, m_shard, v_shard, lo, n4, hy, state, timed_out, dbg); break;
    case 8: dp_adam_kernel<8><<<blocks, 256, 0, st>>>(pp, world, rank, m_shard, v_shard, lo, n4, hy, state, timed_out, dbg); break;
    default: dp_adam_kernel<0><<<blocks, 256, 0, st>>>(pp, world, rank, m_shard, v_shard, lo, n4, hy, state, timed_out, dbg); break;
  }
  return launch_status();
}

}  // namespace

int vsl_peer_barrier(unsigned* const* peer_flags, int rank, int world, unsigned epoch, int* timed_out,
                     long long timeout_ms, vsl_stream_t stream) {
  return launch_barrier(peer_flags, rank, world, epoch, nullptr, 0, timed_out, timeout_ms, (cudaStream_t)stream);
}

int vsl_dp_adam_step(const float* const* peer_grads, float* const* peer_params, int rank, int world, float* m_shard,
                     float* v_shard, long long lo, long long hi, float lr, float beta1, float beta2, float eps,
                     int step, float grad_scale, const int* timed_out, vsl_stream_t stream) {
  const AdamHyper hy = {lr, beta1, beta2, eps, grad_scale, step};
  return launch_dp_adam(peer_grads, peer_params, rank, world, m_shard, v_shard, lo, hi, hy, nullptr, timed_out,
                        (cudaStream_t)stream);
}

int vsl_dp_step(unsigned* const* peer_flags, const float* const* peer_grads, float* const* peer_params, int rank,
                int world, float* m_shard, float* v_shard, long long lo, long long hi, float lr, float beta1, float beta2,
                float eps, float grad_scale, int* state, int* timed_out, long long timeout_ms, vsl_stream_t stream) {
  VSL_REQUIRE(state, VSL_E_NULL);
  cudaStream_t st = (cudaStream_t)stream;
  const AdamHyper hy = {lr, beta1, beta2, eps, grad_scale, 0};
  int rc = VSL_OK;
  // CUDA loads kernels lazily, and loading one may wait for running work.  The update kernel must therefore be
  // resident BEFORE this rank's first barrier starts to spin: otherwise a process that drives several ranks (one
  // host thread, one stream per rank) would sit in the load behind its own barrier, never launch the next rank,
  // and the barrier would time out.
  {
    cudaFuncAttributes fa;
    const void* fn = world == 1 ? (const void*)dp_adam_kernel<1> : world == 2 ? (const void*)dp_adam_kernel<2>
                   : world == 4 ? (const void*)dp_adam_kernel<4> : world == 8 ? (const void*)dp_adam_kernel<8>
                                                                              : (const void*)dp_adam_kernel<0>;
    const cudaError_t e = cudaFuncGetAttributes(&fa, fn);
    if (e != cudaSuccess) return (int)e;
  }
  if (world > 1) {
    rc = launch_barrier(peer_flags, rank, world, 0u, state, 0, timed_out, timeout_ms, st);
    if (rc != VSL_OK) return rc;
  }
  rc = launch_dp_adam(peer_grads, peer_params, rank, world, m_shard, v_shard, lo, hi, hy, state, timed_out, st);
  if (rc != VSL_OK) return rc;
  // the trailing barrier also advances the Adam step count (after every block of the update has read it)
  return launch_barrier(peer_flags, rank, world, 0u, state, 1, timed_out, timeout_ms, st);
}

int vsl_dp_step_mc(unsigned* const* peer_flags, const float* mc_grads, float* mc_params, const float* own_params, int rank,
                   int world, float* m_shard, float* v_shard, long long lo, long long hi, float lr, float beta1, float beta2,
                   float eps, float grad_scale, int* state, int* timed_out, long long timeout_ms, vsl_stream_t stream) {
  VSL_REQUIRE(state && mc_grads && mc_params && own_params && m_shard && v_shard, VSL_E_NULL);
  VSL_REQUIRE(world >= 2 && world <= kMaxPeers && rank >= 0 && rank < world, VSL_E_SHAPE);
  VSL_REQUIRE(lo >= 0 && hi >= lo && lo % 4 == 0 && hi % 4 == 0, VSL_E_SHAPE);
  VSL_REQUIRE(aligned(mc_grads, 16) && aligned(mc_params, 16) && aligned(own_params, 16) && aligned(m_shard, 16) &&
                  aligned(v_shard, 16), VSL_E_ALIGN);
  cudaStream_t st = (cudaStream_t)stream;
  const AdamHyper hy = {lr, beta1, beta2, eps, grad_scale, 0};
  {   // resident before the first barrier spins (see vsl_dp_step)
    cudaFuncAttributes fa;
    const cudaError_t e = cudaFuncGetAttributes(&fa, (const void*)dp_adam_mc_kernel);
    if (e != cudaSuccess) return (int)e;
  }
  int rc = launch_barrier(peer_flags, rank, world, 0u, state, 0, timed_out, timeout_ms, st);
  if (rc != VSL_OK) return rc;
  if (hi > lo) {
    const long long n4 = (hi - lo) / 4;
    const long long want = (n4 + 4 * 256 - 1) / (4 * 256);
    int per_sm = kMcBlocksPerSm;
#ifdef VSL_DP_TIMING_EXPERIMENTS   // never defined for the shipped library
    if (getenv("VSL_MC_BLOCKS_PER_SM")) per_sm = atoi(getenv("VSL_MC_BLOCKS_PER_SM"));
#endif
    const int blocks = (int)(want < 148 * per_sm ? want : 148 * per_sm);
    dp_adam_mc_kernel<<<blocks, 256, 0, st>>>(mc_grads, mc_params, own_params, m_shard, v_shard, lo, n4, hy, state, timed_out);
    rc = launch_status();
    if (rc != VSL_OK) return rc;
  }
  return launch_barrier(peer_flags, rank, world, 0u, state, 1, timed_out, timeout_ms, st);
}

}  // extern "C"
