"""Shared helpers for the GPU parity tests."""
import torch

from oracle import vsl_oracle as O


def smooth_pixels(tgt, srcs, x_pyr, poses, K_pyr, flags, err_eps=2e-4, coord_eps=2e-3):
    """Per scale, which target pixels lie AWAY from the loss's gradient discontinuities (SURVEY.md section 7,
    "Discontinuities"): |warp - tgt| has a kink at 0 and the bilinear footprint switches cell at integer
    source coordinates, so an implementation whose coordinates differ from the oracle's by a float32 ulp may
    legitimately land on the other side there.  Gradient comparisons are made on the returned masks
    ([B,Hs,Ws] bool per view); the share of excluded pixels is asserted to be small by the callers.
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
            warped, coords, _, _, _ = O.projective_inverse_warp(src_s, depth, poses[:, v].double(), K_pyr[:, s].double(),
                                                                flags.pose_format)
            frac = coords - torch.floor(coords)
            ok = ((warped - tgt_s).abs() > err_eps).all(3) & ((frac > coord_eps) & (frac < 1 - coord_eps)).all(3)
            per_view.append(ok)
        out.append(per_view)
    return out


def masked_rel_err(a, b, mask):
    """max |a-b| over mask / max |b| over everything."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    m = mask.expand_as(a) if mask.dim() == a.dim() else mask
    return float(((a - b).abs() * m).max() / b.abs().max().clamp_min(1e-30))
