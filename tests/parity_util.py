"""Shared helpers for the GPU parity tests."""
import torch

from oracle import vsl_oracle as O


def smooth_pixels(tgt, srcs, x_pyr, poses, K_pyr, flags, err_eps=2e-4, coord_eps=2e-3, src_x_pyr=None):
    """Per scale, which target pixels lie AWAY from the loss's gradient discontinuities (SURVEY.md section 7,
    "Discontinuities"): |warp - tgt| has a kink at 0 and the bilinear footprint switches cell at integer
    source coordinates, so an implementation whose coordinates differ from the oracle's by a float32 ulp may
    legitimately land on the other side there.  Gradient comparisons are made on the returned masks
    ([B,Hs,Ws] bool per view); the share of excluded pixels is asserted to be small by the callers.
    With src_x_pyr (the consistency term) the kink of |z - sampled source depth| at 0 is excluded as well.
    Computed with the float64 oracle."""
    B, H, W, _ = tgt.shape
    out = []
    for s in range(flags.num_scales):
        hs, ws = H >> s, W >> s
        x = x_pyr[s].double()
        depth = (1.0 / x if flags.depth_is_inverse else x).squeeze(3)
        tgt_s = O.resize_area(tgt.double(), hs, ws)
        per_view = []
        for v, src in enumerate(srcs):
            src_s = O.resize_area(src.double(), hs, ws)
            warped, coords, _, z_u, _ = O.projective_inverse_warp(src_s, depth, poses[:, v].double(), K_pyr[:, s].double(),
                                                                  flags.pose_format)
            frac = coords - torch.floor(coords)
            ok = ((warped - tgt_s).abs() > err_eps).all(3) & ((frac > coord_eps) & (frac < 1 - coord_eps)).all(3)
            if src_x_pyr is not None:
                sx = src_x_pyr[v][s].double()
                cerr = O.consistent_depth_loss(1.0 / sx if flags.depth_is_inverse else sx, z_u, coords)
                ok = ok & (cerr > err_eps).all(3)
            per_view.append(ok)
        out.append(per_view)
    return out


def masked_rel_err(a, b, mask):
    """max |a-b| over mask / max |b| over everything."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    m = mask.expand_as(a) if mask.dim() == a.dim() else mask
    return float(((a - b).abs() * m).max() / b.abs().max().clamp_min(1e-30))


def lr_flags(make, fl):
    """The left-right trainer's configuration of the fused step (train_depth_then_cam_lr.py:211-340): angle-axis
    poses, smoothness on 1/x, warp depth 1/x, no 1/2^s on the pixel term, consistency weighted by FLAGS.depth_weight."""
    return make(num_scales=fl['num_scales'], smooth_weight=fl['smooth_weight'], data_weight=fl['data_weight'],
                explain_reg_weight=fl['explain_reg_weight'], consist_weight=fl['depth_weight'],
                pose_format='angleaxis', pixel_scale_norm=False, smooth_on_inverse=True, depth_is_inverse=True)


def lr_two_directions(loss_fn, c, conv, flags):
    """The left-right loss as two calls of a view-synthesis step with the consistency term: left target / right
    source and the reverse; each direction's source-depth pyramid is the other direction's prediction.
    loss_fn(tgt, srcs, x_pyr, poses, K_pyr, logits_pyr, src_x_pyr, flags) -> 4 terms.  conv(tensor, grad) makes the
    leaf.  -> (terms[4] summed over the directions, leaves dict)"""
    S = flags.num_scales
    pl = [conv(c['pred_left%d' % s], True) for s in range(S)]
    pr = [conv(c['pred_right%d' % s], True) for s in range(S)]
    po_r, po_l = conv(c.pose_right, True), conv(c.pose_left, True)
    ll = [conv(c['lg_left%d' % s], True) for s in range(S)]
    lr = [conv(c['lg_right%d' % s], True) for s in range(S)]
    left, right, K = conv(c.image_left, False), conv(c.image_right, False), conv(c.K_pyr, False)
    a = loss_fn(left, [right], pl, po_r.unsqueeze(1), K, ll, [pr], flags)
    b = loss_fn(right, [left], pr, po_l.unsqueeze(1), K, lr, [pl], flags)
    return [x + y for x, y in zip(a, b)], dict(pl=pl, pr=pr, po_r=po_r, po_l=po_l, ll=ll, lr=lr)
