// TensorFlow 1.x custom ops over libvsl's C ABI (include/vsl.h): what north_star calls "registered as a TF custom
// op and gradient".  NOT BUILT AND NOT TESTED HERE -- TensorFlow cannot be installed in this image (Python 3.12, no
// network), so this file has never met TF's headers.  It is the shell a maintainer with a TF 1.x toolchain compiles:
//
//   TF_INC=$(python -c 'import tensorflow as tf; print(tf.sysconfig.get_include())')
//   TF_LIB=$(python -c 'import tensorflow as tf; print(tf.sysconfig.get_lib())')
//   g++ -std=c++11 -shared -fPIC -O2 vsl_tf_ops.cc -o libvsl_tf.so -I$TF_INC -I../../include \
//       -L$TF_LIB -ltensorflow_framework -L.. -lvsl -DGOOGLE_CUDA=1
//
// and the reference-side edit is vsl_tf.py beside this file (utils_lr.py:222-256 -> one load_op_library line).
// Every Compute() does three things only: shape checks, allocate_output / allocate_temp, one vsl_* call on the
// op's own CUDA stream.  No arithmetic lives here.
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/util/stream_executor_util.h"

#include "vsl.h"

namespace tf = tensorflow;

namespace {

// the cudaStream_t TF runs this op's kernels on
inline vsl_stream_t StreamOf(tf::OpKernelContext* ctx) {
  auto* stream = ctx->op_device_context()->stream();
  return reinterpret_cast<vsl_stream_t>(stream->implementation()->GpuStreamMemberHack());
}
inline const float* In(tf::OpKernelContext* ctx, int i) { return ctx->input(i).flat<float>().data(); }

tf::Status AllocWs(tf::OpKernelContext* ctx, size_t bytes, tf::Tensor* ws) {
  return ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({static_cast<tf::int64>(bytes + 256)}), ws);
}
inline void* Align256(tf::Tensor* ws) {
  auto p = reinterpret_cast<uintptr_t>(ws->flat<tf::uint8>().data());
  return reinterpret_cast<void*>((p + 255) / 256 * 256);
}

}  // namespace

// ---------------------------------------------------------------- projective_inverse_warp  utils_lr.py:222-256
REGISTER_OP("VslProjectiveInverseWarp")
    .Input("img: float")          // [B,H,W,C]
    .Input("depth: float")        // [B,H,W]
    .Input("pose: float")         // [B,6] or [B,4,4]
    .Input("intrinsics: float")   // [B,3,3]
    .Attr("format: int = 0")      // VSL_POSE_EULER / ANGLEAXIS / MATRIX
    .Output("out_img: float")
    .Output("coords: float")
    .Output("wmask: float")
    .Output("src_depth: float")
    .Output("pose_mat: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      tf::shape_inference::ShapeHandle img = c->input(0);
      c->set_output(0, img);
      auto B = c->Dim(img, 0), H = c->Dim(img, 1), W = c->Dim(img, 2);
      c->set_output(1, c->MakeShape({B, H, W, 2}));
      c->set_output(2, c->MakeShape({B, H, W, 1}));
      c->set_output(3, c->MakeShape({B, H, W, 1}));
      c->set_output(4, c->MakeShape({B, 4, 4}));
      return tf::Status::OK();
    });

class VslProjectiveInverseWarpOp : public tf::OpKernel {
 public:
  explicit VslProjectiveInverseWarpOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("format", &format_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& img = ctx->input(0);
    OP_REQUIRES(ctx, img.dims() == 4, tf::errors::InvalidArgument("img must be [B,H,W,C]"));
    const int B = img.dim_size(0), H = img.dim_size(1), W = img.dim_size(2), C = img.dim_size(3);
    tf::Tensor *out, *coords, *wmask, *z, *pm, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, img.shape(), &out));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, H, W, 2}), &coords));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, tf::TensorShape({B, H, W, 1}), &wmask));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, tf::TensorShape({B, H, W, 1}), &z));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(4, tf::TensorShape({B, 4, 4}), &pm));
    OP_REQUIRES_OK(ctx, AllocWs(ctx, vsl_warp_ws_bytes(B, H, W), &ws));
    const int rc = vsl_warp_fwd(In(ctx, 0), In(ctx, 1), In(ctx, 2), In(ctx, 3), B, H, W, C, format_,
                                out->flat<float>().data(), coords->flat<float>().data(), wmask->flat<float>().data(),
                                z->flat<float>().data(), pm->flat<float>().data(), Align256(&ws), StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, tf::errors::Internal("vsl_warp_fwd: ", vsl_strerror(rc)));
  }

 private:
  int format_;
};
REGISTER_KERNEL_BUILDER(Name("VslProjectiveInverseWarp").Device(tf::DEVICE_GPU), VslProjectiveInverseWarpOp);

REGISTER_OP("VslProjectiveInverseWarpGrad")
    .Input("img: float").Input("depth: float").Input("pose: float").Input("intrinsics: float")
    .Input("g_out_img: float").Input("g_coords: float").Input("g_wmask: float").Input("g_src_depth: float")
    .Input("g_pose_mat: float")
    .Attr("format: int = 0")
    .Output("g_img: float").Output("g_depth: float").Output("g_pose: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(0)); c->set_output(1, c->input(1)); c->set_output(2, c->input(2));
      return tf::Status::OK();
    });

class VslProjectiveInverseWarpGradOp : public tf::OpKernel {
 public:
  explicit VslProjectiveInverseWarpGradOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("format", &format_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& img = ctx->input(0);
    const int B = img.dim_size(0), H = img.dim_size(1), W = img.dim_size(2), C = img.dim_size(3);
    tf::Tensor *g_img, *g_depth, *g_pose, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, img.shape(), &g_img));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, ctx->input(1).shape(), &g_depth));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, ctx->input(2).shape(), &g_pose));
    OP_REQUIRES_OK(ctx, AllocWs(ctx, vsl_warp_ws_bytes(B, H, W), &ws));
    const int rc = vsl_warp_bwd(In(ctx, 0), In(ctx, 1), In(ctx, 2), In(ctx, 3), B, H, W, C, format_, In(ctx, 4),
                                In(ctx, 5), In(ctx, 6), In(ctx, 7), In(ctx, 8), g_img->flat<float>().data(),
                                g_depth->flat<float>().data(), g_pose->flat<float>().data(), Align256(&ws),
                                StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, tf::errors::Internal("vsl_warp_bwd: ", vsl_strerror(rc)));
  }

 private:
  int format_;
};
REGISTER_KERNEL_BUILDER(Name("VslProjectiveInverseWarpGrad").Device(tf::DEVICE_GPU), VslProjectiveInverseWarpGradOp);

// ---------------------------------------------------------------- the fused multi-scale loss (vsl_loss_fwd_bwd)
// One op returns the three losses AND every gradient (they are produced in the same pass); vsl_tf.py registers a
// gradient that multiplies them by the upstream scalar.  S and V are the lengths of the list inputs.
REGISTER_OP("VslViewSynthesisLoss")
    .Input("tgt: float")                  // [B,H,W,3]
    .Input("srcs: V * float")             // V x [B,H,W,3]
    .Input("x_pyr: S * float")            // S x [B,Hs,Ws,1]
    .Input("poses: float")                // [B,V,6] or [B,V,4,4]
    .Input("k_pyr: float")                // [B,S,3,3]
    .Input("logits_pyr: S * float")       // S x [B,Hs,Ws,2V]
    .Attr("S: int >= 1").Attr("V: int >= 1")
    .Attr("pose_format: int = 0").Attr("pixel_scale_norm: int = 1").Attr("depth_is_inverse: int = 1")
    .Attr("smooth_on_inverse: int = 0").Attr("data_weight: float = 1.0").Attr("smooth_weight: float = 0.5")
    .Attr("explain_reg_weight: float = 0.2")
    .Output("losses: float")              // [5] = pixel, smooth, exp, consist, total
    .Output("g_x_pyr: S * float")
    .Output("g_poses: float")
    .Output("g_logits_pyr: S * float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      int S, V;
      TF_RETURN_IF_ERROR(c->GetAttr("S", &S));
      TF_RETURN_IF_ERROR(c->GetAttr("V", &V));
      c->set_output(0, c->Vector(3));
      for (int s = 0; s < S; ++s) c->set_output(1 + s, c->input(1 + V + s));
      c->set_output(1 + S, c->input(1 + V + S));
      for (int s = 0; s < S; ++s) c->set_output(2 + S + s, c->input(3 + V + S + s));
      return tf::Status::OK();
    });

class VslViewSynthesisLossOp : public tf::OpKernel {
 public:
  explicit VslViewSynthesisLossOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    memset(&d_, 0, sizeof(d_));
    OP_REQUIRES_OK(c, c->GetAttr("S", &d_.S));
    OP_REQUIRES_OK(c, c->GetAttr("V", &d_.V));
    OP_REQUIRES_OK(c, c->GetAttr("pose_format", &d_.pose_format));
    OP_REQUIRES_OK(c, c->GetAttr("pixel_scale_norm", &d_.pixel_scale_norm));
    OP_REQUIRES_OK(c, c->GetAttr("depth_is_inverse", &d_.depth_is_inverse));
    OP_REQUIRES_OK(c, c->GetAttr("smooth_on_inverse", &d_.smooth_on_inverse));
    OP_REQUIRES_OK(c, c->GetAttr("data_weight", &d_.data_weight));
    OP_REQUIRES_OK(c, c->GetAttr("smooth_weight", &d_.smooth_weight));
    OP_REQUIRES_OK(c, c->GetAttr("explain_reg_weight", &d_.explain_reg_weight));
    d_.mask_mode = VSL_MASK_EXP;
    d_.loss_scale = 1.0f;
  }
  void Compute(tf::OpKernelContext* ctx) override {
    VslLossDesc d = d_;
    const tf::Tensor& tgt = ctx->input(0);
    d.B = tgt.dim_size(0); d.H = tgt.dim_size(1); d.W = tgt.dim_size(2);
    const int S = d.S, V = d.V;
    const size_t ws_bytes = vsl_loss_ws_bytes(&d);
    OP_REQUIRES(ctx, ws_bytes > 0, tf::errors::InvalidArgument("unsupported shape for vsl_loss_fwd_bwd"));
    const float *srcs[VSL_MAX_VIEWS], *xs[VSL_MAX_SCALES], *lgs[VSL_MAX_SCALES];
    float *g_x[VSL_MAX_SCALES], *g_lg[VSL_MAX_SCALES];
    for (int v = 0; v < V; ++v) srcs[v] = In(ctx, 1 + v);
    tf::Tensor *losses, *g_poses, *t, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({5}), &losses));
    for (int s = 0; s < S; ++s) {
      xs[s] = In(ctx, 1 + V + s);
      lgs[s] = In(ctx, 3 + V + S + s);
      OP_REQUIRES_OK(ctx, ctx->allocate_output(1 + s, ctx->input(1 + V + s).shape(), &t));
      g_x[s] = t->flat<float>().data();
      OP_REQUIRES_OK(ctx, ctx->allocate_output(2 + S + s, ctx->input(3 + V + S + s).shape(), &t));
      g_lg[s] = t->flat<float>().data();
    }
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1 + S, ctx->input(1 + V + S).shape(), &g_poses));
    OP_REQUIRES_OK(ctx, AllocWs(ctx, ws_bytes, &ws));
    const int rc = vsl_loss_fwd_bwd(&d, In(ctx, 0), srcs, xs, In(ctx, 1 + V + S), In(ctx, 2 + V + S), lgs, nullptr,
                                    losses->flat<float>().data(), g_x, g_poses->flat<float>().data(), g_lg, nullptr,
                                    Align256(&ws), StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, tf::errors::Internal("vsl_loss_fwd_bwd: ", vsl_strerror(rc)));
  }

 private:
  VslLossDesc d_;
};
REGISTER_KERNEL_BUILDER(Name("VslViewSynthesisLoss").Device(tf::DEVICE_GPU), VslViewSynthesisLossOp);
