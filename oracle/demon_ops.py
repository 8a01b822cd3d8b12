"""TEST INFRASTRUCTURE ONLY -- restatement of the THIRD-PARTY ops my_losses.py:46-330 calls.

compute_loss_single_depth / compute_loss_pairwise_depth (my_losses.py:46, :101) use four functions of the DeMoN
project that the reference neither vendors nor pins: depthmotionnet.v2.losses
scale_invariant_gradient / pointwise_l2_loss, lmbspecialops replace_nonfinite (`sops`), tfutils ease_out_quad.
They are restated here from their published definitions so that the reference's own function bodies can be
executed (oracle/ref_loader.load_functions(..., extra_globals=DEMON_GLOBALS)).  PARITY UNPINNED for these four:
no source, fixture or test of them exists under /root/reference.

  scale_invariant_gradient(u [N,C,H,W], deltas, weights, epsilon): per delta d, weight w, two planes per channel
      gx = w (u(x+d, y) - u(x, y)) / (|u(x+d, y)| + |u(x, y)| + epsilon),   0 where x+d leaves the image
      gy = w (u(x, y+d) - u(x, y)) / (|u(x, y+d)| + |u(x, y)| + epsilon),   0 where y+d leaves the image
      concatenated over deltas along the channel axis (x plane first).
  pointwise_l2_loss(inp, gt, epsilon): mean over pixels of sqrt(sum_c replace_nonfinite(inp - stop_grad(gt))^2 + epsilon).
  replace_nonfinite(x): x where finite, 0 elsewhere (gradient 0 there).
  ease_out_quad(t, start, change, duration): u = clip(t / duration, 0, 1);  -change u (u - 2) + start.
"""
import torch


def replace_nonfinite(x):
    return torch.where(torch.isfinite(x), x, torch.zeros_like(x))


def _sig_one(u, d, w, eps):
    gx = torch.zeros_like(u)
    gy = torch.zeros_like(u)
    if d < u.shape[3]:
        a, b = u[:, :, :, d:], u[:, :, :, :-d]
        gx = torch.cat([w * (a - b) / (a.abs() + b.abs() + eps), gx[:, :, :, :d]], dim=3)
    if d < u.shape[2]:
        a, b = u[:, :, d:, :], u[:, :, :-d, :]
        gy = torch.cat([w * (a - b) / (a.abs() + b.abs() + eps), gy[:, :, :d, :]], dim=2)
    return torch.cat([gx, gy], dim=1)


def scale_invariant_gradient(inp, deltas, weights, epsilon=0.001):
    assert len(deltas) == len(weights)
    return torch.cat([_sig_one(inp, int(d), float(w), float(epsilon)) for d, w in zip(deltas, weights)], dim=1)


def pointwise_l2_loss(inp, gt, epsilon, data_format='NCHW'):
    diff = replace_nonfinite(inp - gt.detach())
    axis = 1 if data_format == 'NCHW' else 3
    return torch.sqrt((diff * diff).sum(dim=axis) + epsilon).mean()


def ease_out_quad(current_time, start_value, change_value, duration):
    u = torch.clamp(torch.as_tensor(current_time) / duration, 0, 1)
    return -change_value * u * (u - 2) + start_value


class _Sops(object):
    replace_nonfinite = staticmethod(replace_nonfinite)


DEMON_GLOBALS = {'scale_invariant_gradient': scale_invariant_gradient, 'pointwise_l2_loss': pointwise_l2_loss,
                 'sops': _Sops, 'ease_out_quad': ease_out_quad}
