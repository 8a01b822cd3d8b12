// Device-side geometry shared by every kernel of libvsl: the fp32 pixel grid, back-projection,
// projection, the zero-padded bilinear footprint, small-matrix algebra and the pose parameterisations.
//
// Arithmetic contract (DESIGN.md "numerics"): everything that decides WHERE a sample lands -- grid,
// K^-1 p, depth scaling, P [cam;1], the perspective divide, floor/clip/equal -- is evaluated with
// explicitly rounded, never-contracted fp32 operations in the reference's association order, so that the
// sampled coordinates are bit-identical to the oracle's when the pose is given as a matrix.  Gradient
// arithmetic downstream of the coordinates is free to use FMA.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace vsl {

#define VSL_DEV __device__ __forceinline__

constexpr float kEpsZ = 1e-10f;  // utils.py:136-137

// ---- meshgrid (utils.py:142-166): (linspace(-1,1,n)[i] + 1) * 0.5 * (n-1), TF1 linspace = start+step*i
VSL_DEV float grid_step(int n) { return __fdiv_rn(2.0f, (float)(n - 1)); }
VSL_DEV float grid_coord(int i, int n, float step) {
  float lin = __fadd_rn(-1.0f, __fmul_rn(step, (float)i));
  return __fmul_rn(__fmul_rn(__fadd_rn(lin, 1.0f), 0.5f), (float)(n - 1));
}

// Per (batch element, view) transform: K^-1 (row-major 3x3) and the first three rows of P = K4 . T.
struct Xform {
  float kinv[9];
  float p[12];
};

// The same transform folded for the view-paired fused kernel: u = depth * (Q [gx, gy, 1]) + t with
// Q = P[:, :3] . K^-1 (formed in double from the float32 P and K^-1, rounded once) and t = P[:, 3].
struct XformQ {
  float q[9];
  float t[3];
};

struct Ray { float r0, r1, r2; };     // K^-1 [gx, gy, 1]
struct Proj { float x, y, z, zp; };   // sampled coords, u2, u2 + eps

// pixel2cam's matmul (utils.py:114): sequential k, no contraction.
VSL_DEV Ray back_project(const float* __restrict__ kinv, float gx, float gy) {
  Ray r;
  r.r0 = __fadd_rn(__fadd_rn(__fmul_rn(kinv[0], gx), __fmul_rn(kinv[1], gy)), kinv[2]);
  r.r1 = __fadd_rn(__fadd_rn(__fmul_rn(kinv[3], gx), __fmul_rn(kinv[4], gy)), kinv[5]);
  r.r2 = __fadd_rn(__fadd_rn(__fmul_rn(kinv[6], gx), __fmul_rn(kinv[7], gy)), kinv[8]);
  return r;
}

VSL_DEV float row4(const float* __restrict__ p, float c0, float c1, float c2) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p[0], c0), __fmul_rn(p[1], c1)), __fmul_rn(p[2], c2)), p[3]);
}

// cam = ray * depth (utils.py:114); u = P [cam;1] (utils.py:132); x = u0/(u2+eps), y = u1/(u2+eps).
VSL_DEV Proj project(const float* __restrict__ p, float c0, float c1, float c2) {
  Proj q;
  float u0 = row4(p, c0, c1, c2), u1 = row4(p + 4, c0, c1, c2);
  q.z = row4(p + 8, c0, c1, c2);
  q.zp = __fadd_rn(q.z, kEpsZ);
  q.x = __fdiv_rn(u0, q.zp);
  q.y = __fdiv_rn(u1, q.zp);
  return q;
}

// ---- bilinear footprint (utils.py:252-270, 297-300): corner indices (clipped) and per-axis weights with
// the "corner == clipped corner" zero-padding masks folded in.
struct Foot {
  int x0, x1, y0, y1;        // clipped, always valid indices
  float wx0, wx1, wy0, wy1;  // (x1-x)*m, (x-x0)*m, ...
  float mx0, mx1, my0, my1;  // the 0/1 masks (needed by d/dx, d/dy)
};

VSL_DEV Foot footprint(float x, float y, int Ws, int Hs) {
  Foot f;
  const float xmax = (float)(Ws - 1), ymax = (float)(Hs - 1);
  float x0 = floorf(x), y0 = floorf(y);
  float x1 = __fadd_rn(x0, 1.0f), y1 = __fadd_rn(y0, 1.0f);
  float x0s = fminf(fmaxf(x0, 0.0f), xmax), x1s = fminf(fmaxf(x1, 0.0f), xmax);
  float y0s = fminf(fmaxf(y0, 0.0f), ymax), y1s = fminf(fmaxf(y1, 0.0f), ymax);
  f.mx0 = (x0 == x0s) ? 1.0f : 0.0f;
  f.mx1 = (x1 == x1s) ? 1.0f : 0.0f;
  f.my0 = (y0 == y0s) ? 1.0f : 0.0f;
  f.my1 = (y1 == y1s) ? 1.0f : 0.0f;
  f.wx0 = __fmul_rn(__fsub_rn(x1, x), f.mx0);
  f.wx1 = __fmul_rn(__fsub_rn(x, x0), f.mx1);
  f.wy0 = __fmul_rn(__fsub_rn(y1, y), f.my0);
  f.wy1 = __fmul_rn(__fsub_rn(y, y0), f.my1);
  f.x0 = (int)x0s; f.x1 = (int)x1s; f.y0 = (int)y0s; f.y1 = (int)y1s;
  return f;
}

// output = add_n([w00*im00, w01*im01, w10*im10, w11*im11]) left to right (utils.py:302-305)
VSL_DEV float blend(float w00, float w01, float w10, float w11, float i00, float i01, float i10, float i11) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w00, i00), __fmul_rn(w01, i01)), __fmul_rn(w10, i10)),
                   __fmul_rn(w11, i11));
}

// ---- small matrices, row-major, sequential-k products without contraction
VSL_DEV void mm3(const float* a, const float* b, float* c) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c[i * 3 + j] = __fadd_rn(__fadd_rn(__fmul_rn(a[i * 3], b[j]), __fmul_rn(a[i * 3 + 1], b[3 + j])),
                               __fmul_rn(a[i * 3 + 2], b[6 + j]));
}

// tf.matrix_inverse (utils.py:114) = Eigen PartialPivLU: row-pivoted LU, then per-column substitution.
VSL_DEV void inv3_lu(const float* m, float* inv) {
  float a[9];
  int perm[3] = {0, 1, 2};
#pragma unroll
  for (int i = 0; i < 9; ++i) a[i] = m[i];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int p = k;
    float best = fabsf(a[k * 3 + k]);
    for (int i = k + 1; i < 3; ++i) {
      float v = fabsf(a[i * 3 + k]);
      if (v > best) { best = v; p = i; }
    }
    if (p != k) {
      for (int j = 0; j < 3; ++j) { float t = a[k * 3 + j]; a[k * 3 + j] = a[p * 3 + j]; a[p * 3 + j] = t; }
      int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
    }
    for (int i = k + 1; i < 3; ++i) {
      a[i * 3 + k] = __fdiv_rn(a[i * 3 + k], a[k * 3 + k]);
      for (int j = k + 1; j < 3; ++j)
        a[i * 3 + j] = __fsub_rn(a[i * 3 + j], __fmul_rn(a[i * 3 + k], a[k * 3 + j]));
    }
  }
  for (int c = 0; c < 3; ++c) {
    float y[3];
    for (int i = 0; i < 3; ++i) {
      float s = (perm[i] == c) ? 1.0f : 0.0f;
      for (int j = 0; j < i; ++j) s = __fsub_rn(s, __fmul_rn(a[i * 3 + j], y[j]));
      y[i] = s;
    }
    for (int i = 2; i >= 0; --i) {
      float s = y[i];
      for (int j = i + 1; j < 3; ++j) s = __fsub_rn(s, __fmul_rn(a[i * 3 + j], inv[j * 3 + c]));
      inv[i * 3 + c] = __fdiv_rn(s, a[i * 3 + i]);
    }
  }
}

// P = K4 . T (utils.py:190-196), rows 0..2.  K4 = [[K,0],[0,0,0,1]], so the k = 3 term is 0 * T[3][j].
VSL_DEV void proj_rows(const float* K, const float* T, float* p) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s = __fadd_rn(__fadd_rn(__fmul_rn(K[i * 3], T[j]), __fmul_rn(K[i * 3 + 1], T[4 + j])),
                          __fmul_rn(K[i * 3 + 2], T[8 + j]));
      p[i * 4 + j] = __fadd_rn(s, __fmul_rn(0.0f, T[12 + j]));
    }
}

// ---- pose parameterisations: forward in reference-faithful fp32
constexpr float kPi = 3.14159265358979323846f;

VSL_DEV void euler_rot(float rx, float ry, float rz, float* R) {  // utils.py:26-75: R = Rx . Ry . Rz
  float x = fminf(fmaxf(rx, -kPi), kPi), y = fminf(fmaxf(ry, -kPi), kPi), z = fminf(fmaxf(rz, -kPi), kPi);
  float cz = cosf(z), sz = sinf(z), cy = cosf(y), sy = sinf(y), cx = cosf(x), sx = sinf(x);
  float Z[9] = {cz, -sz, 0.f, sz, cz, 0.f, 0.f, 0.f, 1.f};
  float Y[9] = {cy, 0.f, sy, 0.f, 1.f, 0.f, -sy, 0.f, cy};
  float X[9] = {1.f, 0.f, 0.f, 0.f, cx, -sx, 0.f, sx, cx};
  float XY[9];
  mm3(X, Y, XY);
  mm3(XY, Z, R);
}

VSL_DEV void angleaxis_rot(float r0, float r1, float r2, float* R) {  // utils_lr.py:77-103,128-135
  float th = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r0, r0), __fmul_rn(r1, r1)), __fmul_rn(r2, r2)));
  float a0 = __fdiv_rn(r0, th), a1 = __fdiv_rn(r1, th), a2 = __fdiv_rn(r2, th);  // NaN at th == 0
  float A[9] = {0.f, -a2, a1, a2, 0.f, -a0, -a1, a0, 0.f};
  float AA[9];
  mm3(A, A, AA);
  float s = sinf(th), omc = __fsub_rn(1.0f, cosf(th));
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    float e = (i % 4 == 0) ? 1.0f : 0.0f;
    R[i] = __fadd_rn(__fadd_rn(e, __fmul_rn(s, A[i])), __fmul_rn(omc, AA[i]));
  }
}

// T (row-major 4x4) from a pose in any format
VSL_DEV void pose_to_mat(const float* pose, int format, float* T) {
  if (format == 2) {
#pragma unroll
    for (int i = 0; i < 16; ++i) T[i] = pose[i];
    return;
  }
  float R[9];
  if (format == 0) euler_rot(pose[3], pose[4], pose[5], R);
  else angleaxis_rot(pose[3], pose[4], pose[5], R);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    T[i * 4] = R[i * 3]; T[i * 4 + 1] = R[i * 3 + 1]; T[i * 4 + 2] = R[i * 3 + 2]; T[i * 4 + 3] = pose[i];
  }
  T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
}

// ---- pose backward.  gT: upstream gradient of the 4x4 matrix (row-major); g: gradient of the 6-vector.
// T = float keeps the serial chain short (FP64 issue latency dominates otherwise); the sums that feed gT are
// accumulated in double by the callers.
template <typename T>
__device__ inline void pose_vec_grad(const float* pose, int format, const T* gT, float* g) {
  g[0] = (float)gT[3]; g[1] = (float)gT[7]; g[2] = (float)gT[11];
  T G[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) G[i * 3 + j] = gT[i * 4 + j];
  auto mm = [](const T* p, const T* q, T* o) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        T s = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) s += p[i * 3 + k] * q[k * 3 + j];
        o[i * 3 + j] = s;
      }
  };
  auto dot = [](const T* p, const T* q) {
    T s = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) s += p[i] * q[i];
    return s;
  };
  if (format == 0) {
    const float pi = kPi;  // the clip bounds are fp32(pi) (utils.py:40-42)
    float a[3] = {pose[3], pose[4], pose[5]};  // x, y, z
    bool in[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { in[i] = (a[i] >= -pi && a[i] <= pi); a[i] = fminf(fmaxf(a[i], -pi), pi); }
    const T cx = cosf(a[0]), sx = sinf(a[0]), cy = cosf(a[1]), sy = sinf(a[1]), cz = cosf(a[2]), sz = sinf(a[2]);
    const T X[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx}, Y[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy},
            Z[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
    const T dX[9] = {0, 0, 0, 0, -sx, -cx, 0, cx, -sx}, dY[9] = {-sy, 0, cy, 0, 0, 0, -cy, 0, -sy},
            dZ[9] = {-sz, -cz, 0, cz, -sz, 0, 0, 0, 0};
    T t1[9], t2[9];
    mm(dX, Y, t1); mm(t1, Z, t2); g[3] = in[0] ? (float)dot(G, t2) : 0.f;
    mm(X, dY, t1); mm(t1, Z, t2); g[4] = in[1] ? (float)dot(G, t2) : 0.f;
    mm(X, Y, t1); mm(t1, dZ, t2); g[5] = in[2] ? (float)dot(G, t2) : 0.f;
  } else {
    const T r[3] = {pose[3], pose[4], pose[5]};
    const T th = sqrtf((float)(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]));
    const T a[3] = {r[0] / th, r[1] / th, r[2] / th};
    const T A[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
    T AA[9];
    mm(A, A, AA);
    const T s = sinf((float)th), c = cosf((float)th);
    T gth = dot(G, A) * c + dot(G, AA) * s;
    // dA = s G + (1-c) (G A^T + A^T G)
    T dA[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        T t = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) t += G[i * 3 + k] * A[j * 3 + k] + A[k * 3 + i] * G[k * 3 + j];
        dA[i * 3 + j] = s * G[i * 3 + j] + (1 - c) * t;
      }
    // A = M - M^T with M01 = -a2, M02 = a1, M12 = -a0
    const T da[3] = {-(dA[5] - dA[7]), (dA[2] - dA[6]), -(dA[1] - dA[3])};
    const T dotar = da[0] * r[0] + da[1] * r[1] + da[2] * r[2];
    gth -= dotar / (th * th);
#pragma unroll
    for (int i = 0; i < 3; ++i) g[3 + i] = (float)(da[i] / th + gth * r[i] / th);
  }
}

}  // namespace vsl
