"""Drop-in for the reference's my_losses.py: the smoothness, explainability-regulariser and reference-mask helpers
(my_losses.py:14-43), the two composite losses compute_loss_single_depth / compute_loss_pairwise_depth
(my_losses.py:46, :101) and the fused multi-scale entry the per-script loss loops collapse into.

In the composites everything that is on the hot path -- the resize_area pyramids, pose_vec2mat, the matrix-format
projective_inverse_warp -- runs in libvsl's CUDA kernels; the terms the reference itself delegates to third-party
TF ops of the DeMoN project (scale_invariant_gradient, pointwise_l2_loss, sops.replace_nonfinite, ease_out_quad:
not vendored, not pinned) and the tiny camera loss are delegated to framework (torch) ops here too."""
import torch

from tf_depth_estimation_b200 import ops as _ops
from tf_depth_estimation_b200.ops import LossFlags, view_synthesis_loss  # noqa: F401

__all__ = ['get_reference_explain_mask', 'compute_smooth_loss', 'compute_exp_reg_loss', 'compute_loss_single_depth',
           'compute_loss_pairwise_depth', 'view_synthesis_loss', 'LossFlags']


def get_reference_explain_mask(downscaling, FLAGS, device='cuda'):
    """my_losses.py:14-23: constant [0,1] labels, [batch_size, H/2^s, W/2^s, 2]."""
    m = torch.zeros(FLAGS.batch_size, int(FLAGS.resizedheight / (2 ** downscaling)),
                    int(FLAGS.resizedwidth / (2 ** downscaling)), 2, device=device)
    m[..., 1] = 1.0
    return m


def compute_smooth_loss(pred_disp):
    """my_losses.py:27-36."""
    return _ops.compute_smooth_loss(pred_disp)


def compute_exp_reg_loss(pred, ref):
    """my_losses.py:39-43.  `ref` must be the constant mask of get_reference_explain_mask (the only labels the
    reference ever passes); anything else raises."""
    if ref is not None:
        r = ref.reshape(-1, 2)
        if not (bool((r[:, 0] == 0).all()) and bool((r[:, 1] == 1).all())):
            raise ValueError('compute_exp_reg_loss supports the reference explainability mask [0,1] only')
    return _ops.compute_exp_reg_loss(pred)


# ---- the DeMoN project's ops as the reference calls them (framework ops; see the module docstring)
def _replace_nonfinite(x):
    return torch.where(torch.isfinite(x), x, torch.zeros_like(x))


def _scale_invariant_gradient(u, deltas, weights, epsilon=0.001):
    """u [N,C,H,W] -> [N, 2 C len(deltas), H, W]: w (u(p+d) - u(p)) / (|u(p+d)| + |u(p)| + eps) along x then y,
    zero where p + d leaves the image."""
    planes = []
    for d, w in zip(deltas, weights):
        for dim in (3, 2):
            out = torch.zeros_like(u)
            n = u.shape[dim]
            if d < n:
                far, near = u.narrow(dim, d, n - d), u.narrow(dim, 0, n - d)
                out.narrow(dim, 0, n - d).copy_(w * (far - near) / (far.abs() + near.abs() + epsilon))
            planes.append(out)
    return torch.cat(planes, dim=1)


def _pointwise_l2_loss(inp, gt, epsilon):
    diff = _replace_nonfinite(inp - gt.detach())
    return torch.sqrt((diff * diff).sum(dim=1) + epsilon).mean()


def _ease_out_quad(current_time, start_value, change_value, duration):
    u = min(max(float(current_time) / duration, 0.0), 1.0)
    return -change_value * u * (u - 2.0) + start_value


def _area_levels(x, FLAGS, scales):
    """tf.image.resize_area(x, [H/2^s, W/2^s]) for the requested scales (CUDA pyramid: exact block means)."""
    need = max(scales) + 1
    levels = _ops.image_pyramid(x, need)
    for s in scales:
        want = (int(FLAGS.resizedheight / (2 ** s)), int(FLAGS.resizedwidth / (2 ** s)))
        if tuple(levels[s].shape[1:3]) != want:
            raise ValueError('label / image size does not match FLAGS.resizedheight x resizedwidth')
    return levels


def compute_loss_single_depth(pred_depth, label, global_step, FLAGS):
    """my_losses.py:46-97 -> (depth_loss, smooth_loss, loss_depth_sig)."""
    S = FLAGS.num_scales
    w_sig = _ease_out_quad(global_step, 0, FLAGS.depth_sig_weight, float(FLAGS.max_steps // 3))
    labels = _area_levels(label, FLAGS, range(S))
    depth_loss, smooth_loss, loss_depth_sig = 0, 0, 0
    for s in range(S):
        pre = _scale_invariant_gradient(pred_depth[s].permute(0, 3, 1, 2), [2], [1])
        gt = _scale_invariant_gradient(labels[s].permute(0, 3, 1, 2), [2], [1])
        loss_depth_sig = loss_depth_sig + w_sig * _pointwise_l2_loss(pre, gt, 0.000001)
        diff = _replace_nonfinite(labels[s] - pred_depth[s])
        depth_loss = depth_loss + diff.abs().mean() * FLAGS.depth_weight / (2 ** s)
    return depth_loss, smooth_loss, loss_depth_sig


def compute_loss_pairwise_depth(image_left, image_right, pred_depth_left, pred_poses_right, pred_exp_logits_left,
                                pred_depth_right, pred_poses_left, pred_exp_logits_right, gt_right_cam, intrinsics,
                                label, FLAGS, global_step):
    """my_losses.py:101-330 -> the reference's 11-tuple (depth_loss, cam_loss, pixel_loss, consist_loss,
    loss_depth_sig, exp_loss, left_image_all, right_image_all, proj_image_left_all, proj_image_right_all,
    proj_error_stack_all).  As in the reference, pixel / consistency / explainability terms are switched off
    (commented out there) and scales 2..num_scales-1 use pred_depth_*[s-2]."""
    S = FLAGS.num_scales
    depth_loss, cam_loss, pixel_loss, consist_loss, loss_depth_sig, exp_loss = 0, 0, 0, 0, 0, 0
    left_all, right_all, proj_left_all, proj_right_all, proj_err_all = [], [], [], [], []
    GT_l2r = _ops.pose_vec2mat(gt_right_cam, 'angleaxis')
    GT_r2l = torch.linalg.inv(GT_l2r)
    w_sig = _ease_out_quad(global_step, 0, FLAGS.depth_sig_weight, float(FLAGS.max_steps // 3))
    proj_l2r = _ops.pose_vec2mat(pred_poses_right[:, 0, :], 'angleaxis')
    proj_r2l = _ops.pose_vec2mat(pred_poses_left[:, 0, :], 'angleaxis')
    cam_loss = cam_loss + ((GT_l2r[:, 0:3, 0:3] - proj_l2r[:, 0:3, 0:3]) ** 2).mean() * FLAGS.cam_weight_rot
    cam_loss = cam_loss + ((GT_r2l[:, 0:3, 0:3] - proj_r2l[:, 0:3, 0:3]) ** 2).mean() * FLAGS.cam_weight_rot
    cam_loss = cam_loss + ((GT_l2r[:, 0:3, 3] - proj_l2r[:, 0:3, 3]) ** 2).mean() * FLAGS.cam_weight_tran
    cam_loss = cam_loss + ((GT_r2l[:, 0:3, 3] - proj_r2l[:, 0:3, 3]) ** 2).mean() * FLAGS.cam_weight_tran
    scales = range(2, S)
    if len(scales) == 0:
        return (depth_loss, cam_loss, pixel_loss, consist_loss, loss_depth_sig, exp_loss, left_all, right_all,
                proj_left_all, proj_right_all, proj_err_all)
    labels = _area_levels(label, FLAGS, scales)
    lefts = _area_levels(image_left, FLAGS, scales)
    rights = _area_levels(image_right, FLAGS, scales)
    for s in scales:
        pdl, pdr = pred_depth_left[s - 2], pred_depth_right[s - 2]
        pre = _scale_invariant_gradient(pdl.permute(0, 3, 1, 2), [2], [1])
        gt = _scale_invariant_gradient(labels[s].permute(0, 3, 1, 2), [2], [1])
        loss_depth_sig = loss_depth_sig + w_sig * _pointwise_l2_loss(pre, gt, 0.000001)
        diff = _replace_nonfinite(labels[s] - pdl)
        depth_loss = depth_loss + diff.abs().mean() * FLAGS.depth_weight / (2 ** s)
        K = intrinsics[:, s, :, :].contiguous()
        proj_left = _ops.projective_inverse_warp(rights[s], (1.0 / labels[s]).squeeze(3), GT_l2r, K, 'matrix')[0]
        proj_right = _ops.projective_inverse_warp(lefts[s], (1.0 / pdr).squeeze(3), GT_r2l, K, 'matrix')[0]
        left_all.append(lefts[s])
        right_all.append(rights[s])
        proj_left_all.append(proj_left)
        proj_right_all.append(proj_right)
        proj_err_all.append((proj_right - rights[s]).abs())
    return (depth_loss, cam_loss, pixel_loss, consist_loss, loss_depth_sig, exp_loss, left_all, right_all,
            proj_left_all, proj_right_all, proj_err_all)
