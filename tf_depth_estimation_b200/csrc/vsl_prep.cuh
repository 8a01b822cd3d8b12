// The per-(scale, view, batch) transform tables and the job that fills them: shared by the stand-alone warp ops
// (vsl_ops.cu) and the prep launch of the fused step (vsl_loss.cu).
#pragma once
#include "vsl_common.cuh"

namespace vsl {

// Per (scale, view, batch) transform table: K_s^-1 and rows 0..2 of K4_s . T_v.
// xf[(s*V + v)*B + b];  K_pyr is [B,S,3,3];  poses is [B,V,6] or [B,V,4,4].
struct PrepJob {
  const float* poses;
  const float* K_pyr;
  Xform* xf;
  float* pose_mat;  // [B,V,4,4], nullable
  int B, S, V, format, n;
  XformQ* xq;       // folded table of the view-paired fused kernel, same indexing as xf; nullable
};

VSL_DEV void prep_one(const PrepJob& j, int idx) {
  const int b = idx % j.B, v = (idx / j.B) % j.V, s = idx / (j.B * j.V);
  const int psz = (j.format == VSL_POSE_MATRIX) ? 16 : 6;
  float T[16], K[9];
  pose_to_mat(j.poses + (size_t)(b * j.V + v) * psz, j.format, T);
#pragma unroll
  for (int i = 0; i < 9; ++i) K[i] = j.K_pyr[(size_t)(b * j.S + s) * 9 + i];
  Xform o;
  inv3_lu(K, o.kinv);
  proj_rows(K, T, o.p);
  j.xf[idx] = o;
  if (j.xq != nullptr) {
    XformQ q;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        q.q[i * 3 + c] = (float)((double)o.p[i * 4] * (double)o.kinv[c] + (double)o.p[i * 4 + 1] * (double)o.kinv[3 + c] +
                                 (double)o.p[i * 4 + 2] * (double)o.kinv[6 + c]);
      q.t[i] = o.p[i * 4 + 3];
    }
    j.xq[idx] = q;
  }
  if (j.pose_mat != nullptr && s == 0)
    for (int i = 0; i < 16; ++i) j.pose_mat[(size_t)(b * j.V + v) * 16 + i] = T[i];
}

inline PrepJob make_prep(const float* poses, const float* K_pyr, int B, int S, int V, int format, Xform* xf,
                                float* pose_mat) {
  PrepJob j;
  j.poses = poses; j.K_pyr = K_pyr; j.xf = xf; j.pose_mat = pose_mat;
  j.B = B; j.S = S; j.V = V; j.format = format; j.n = B * S * V;
  j.xq = nullptr;
  return j;
}

}  // namespace vsl
