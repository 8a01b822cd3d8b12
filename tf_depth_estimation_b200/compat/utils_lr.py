"""Drop-in for the reference's utils_lr.py (API v2): same function names, argument order and
return tuples, backed by libvsl's CUDA kernels.  Tensors are CUDA float32 torch tensors in the reference's
layouts; every function is differentiable where TF autodiff would be (intrinsics are data)."""
from tf_depth_estimation_b200.ops import (axis_angle_to_rotation_matrix, bilinear_sampler, cam2pixel,  # noqa: F401
                                          consistent_depth_loss, depth_optflow, euler2mat, meshgrid, optflow_warp,
                                          pixel2cam)
from tf_depth_estimation_b200 import ops as _ops

__all__ = ['euler2mat', 'axis_angle_to_rotation_matrix', 'pose_vec2mat', 'pixel2cam', 'cam2pixel', 'meshgrid',
           'projective_inverse_warp', 'optflow_warp', 'bilinear_sampler', 'consistent_depth_loss', 'depth_optflow']


def pose_vec2mat(vec, format):
    """utils_lr.py:106-149.  format: 'eular' | 'angleaxis' (the reference's debug-only 'test' is not provided)."""
    return _ops.pose_vec2mat(vec, format)


def projective_inverse_warp(img, depth, pose, intrinsics, format='eular'):
    """utils_lr.py:222-256 -> (output_img, src_pixel_coords, wmask, src_depth, pose).
    format 'matrix' takes pose as a ready [B,4,4] transform (my_losses.py:224-239)."""
    return _ops.projective_inverse_warp(img, depth, pose, intrinsics, format)
