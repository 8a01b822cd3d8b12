"""Reference-side binding for a TensorFlow 1.x toolchain (NOT importable in this image: no TensorFlow).

Drop-in for the reference's `utils_lr.py`: same function names and return tuples (utils_lr.py:222-256,
utils.py:168-199), the arithmetic in libvsl through the custom ops of vsl_tf_ops.cc.  Untested here; see the header of
that file for the build line.
"""
import os

import tensorflow as tf
from tensorflow.python.framework import ops as _ops

_mod = tf.load_op_library(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libvsl_tf.so'))
_FORMAT = {'eular': 0, 'angleaxis': 1, 'matrix': 2}


def projective_inverse_warp(img, depth, pose, intrinsics, format='eular'):
    """utils_lr.py:222-256 -> (output_img, src_pixel_coords, wmask, src_depth, pose_mat)."""
    return _mod.vsl_projective_inverse_warp(img, depth, pose, intrinsics, format=_FORMAT[format])


@_ops.RegisterGradient('VslProjectiveInverseWarp')
def _warp_grad(op, g_out, g_coords, g_wmask, g_z, g_pose_mat):
    g_img, g_depth, g_pose = _mod.vsl_projective_inverse_warp_grad(
        op.inputs[0], op.inputs[1], op.inputs[2], op.inputs[3], g_out, g_coords, g_wmask, g_z, g_pose_mat,
        format=op.get_attr('format'))
    return g_img, g_depth, g_pose, None      # intrinsics are data


def view_synthesis_loss(tgt, srcs, x_pyr, poses, k_pyr, logits_pyr, FLAGS, pose_format='eular'):
    """The per-scale loop of train.py:107-135 / train_depth_then_cam_lr.py:297-328 as one op.
    -> (total, pixel_loss, smooth_loss, exp_loss).  Only `total` (= pixel + smooth + exp, the sum every train script
    forms, term weights already inside FLAGS) is differentiable, wrt x_pyr, poses, logits_pyr: the op produces ONE set
    of gradients, those of the sum.  The three terms are returned under stop_gradient for summaries; re-weighting
    them after the fact would silently use the wrong gradients, so it is made impossible instead."""
    out = _mod.vsl_view_synthesis_loss(tgt, srcs, x_pyr, poses, k_pyr, logits_pyr, pose_format=_FORMAT[pose_format],
                                       data_weight=FLAGS.data_weight, smooth_weight=FLAGS.smooth_weight,
                                       explain_reg_weight=FLAGS.explain_reg_weight)
    terms = tf.stop_gradient(out.losses[:3])
    return out.losses[4], terms[0], terms[1], terms[2]


@_ops.RegisterGradient('VslViewSynthesisLoss')
def _loss_grad(op, g_losses, *_unused):
    # losses = [pixel, smooth, exp, total]; the op's precomputed gradients are those of `total`, so only the upstream
    # entry of `total` may be non-zero (the wrapper above hands the other three out under stop_gradient)
    S, V = op.get_attr('S'), op.get_attr('V')
    g = g_losses[4]
    g_x = [g * t for t in op.outputs[1:1 + S]]
    g_poses = g * op.outputs[1 + S]
    g_lg = [g * t for t in op.outputs[2 + S:2 + 2 * S]]
    return [None] + [None] * V + g_x + [g_poses, None] + g_lg
