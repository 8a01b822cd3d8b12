"""Drop-in for the reference's utils.py (API v1): Euler poses only, 3-tuple warp result."""
from tf_depth_estimation_b200.ops import (bilinear_sampler, cam2pixel as _cam2pixel, depth_optflow, euler2mat,  # noqa: F401
                                          meshgrid, optflow_warp, pixel2cam)
from tf_depth_estimation_b200 import ops as _ops

__all__ = ['euler2mat', 'pose_vec2mat', 'pixel2cam', 'cam2pixel', 'meshgrid', 'projective_inverse_warp',
           'optflow_warp', 'bilinear_sampler', 'depth_optflow']


def pose_vec2mat(vec):
    """utils.py:79-98."""
    return _ops.pose_vec2mat(vec, 'eular')


def cam2pixel(cam_coords, proj):
    """utils.py:121-140 -> pixel coords only."""
    return _cam2pixel(cam_coords, proj)[0]


def projective_inverse_warp(img, depth, pose, intrinsics):
    """utils.py:168-199 -> (output_img, src_pixel_coords, wmask).  A [B,4,4] pose (the v0 callers, e.g.
    train.py:127-134, pass loader matrices) is accepted as well."""
    fmt = 'matrix' if pose.dim() == 3 else 'eular'
    return _ops.projective_inverse_warp(img, depth, pose, intrinsics, fmt)[:3]
