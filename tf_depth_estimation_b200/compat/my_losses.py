"""Drop-in for the in-scope part of the reference's my_losses.py: the smoothness, explainability-regulariser
and reference-mask helpers (my_losses.py:14-43) plus the fused multi-scale entry the per-script loss loops
collapse into.  compute_loss_single_depth / compute_loss_pairwise_depth mix these with DeMoN's third-party
scale-invariant-gradient ops (depthmotionnet, lmbspecialops: not vendored by the reference) and
are therefore not provided."""
import torch

from tf_depth_estimation_b200 import ops as _ops
from tf_depth_estimation_b200.ops import LossFlags, view_synthesis_loss  # noqa: F401

__all__ = ['get_reference_explain_mask', 'compute_smooth_loss', 'compute_exp_reg_loss', 'view_synthesis_loss',
           'LossFlags']


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
