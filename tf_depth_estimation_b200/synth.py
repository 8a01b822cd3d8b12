"""Synthetic frame snippets of the shapes BASELINE.json names (no dataset is reachable offline).

Value domains follow what the reference's loaders and nets produce (tail):
images in [0, 1] (imageselect_Dataloader.py:93), disparity = 4*sigmoid(.) (nets_optflow_depth.py:143-144)
clipped away from 0, small 6-DoF poses (nets.py:47 scales by 0.01), KITTI-like intrinsics halved per
scale (Demon_Data_loader.py:25-39), N(0,1) explainability logits.

Generated on the CPU with a seeded torch.Generator so that every rank / test sees identical data.
"""
import math

import torch

CONFIGS = {
    # name: (B, H, W, S, V)   -- BASELINE.json "configs", per GPU
    'cfg1': (4, 128, 416, 4, 2),
    'cfg2': (32, 128, 416, 4, 2),     # headline microbench
    'cfg3_per_gpu32': (32, 128, 416, 4, 2),
    'cfg4': (64, 192, 256, 4, 1),
    'cfg5': (64, 480, 640, 4, 2),
}


def _texture(g, B, H, W, C, n_waves=8, noise=0.02):
    ys = torch.arange(H, dtype=torch.float32).reshape(1, H, 1, 1)
    xs = torch.arange(W, dtype=torch.float32).reshape(1, 1, W, 1)
    img = torch.zeros(B, H, W, C)
    for _ in range(n_waves):
        fx = (torch.rand(B, 1, 1, C, generator=g) - 0.5) * 0.12
        fy = (torch.rand(B, 1, 1, C, generator=g) - 0.5) * 0.12
        ph = torch.rand(B, 1, 1, C, generator=g) * (2 * math.pi)
        img += torch.sin(xs * fx + ys * fy + ph)
    img += noise * n_waves * (torch.rand(B, H, W, C, generator=g) - 0.5)
    lo = img.amin(dim=(1, 2, 3), keepdim=True)
    hi = img.amax(dim=(1, 2, 3), keepdim=True)
    return ((img - lo) / (hi - lo)).contiguous()


def _lowpass(g, B, H, W):
    ch, cw = max(H // 16, 2), max(W // 16, 2)
    coarse = torch.randn(B, 1, ch, cw, generator=g)
    return torch.nn.functional.interpolate(coarse, size=(H, W), mode='bilinear', align_corners=True)


def intrinsics(B, H, W):
    K = torch.zeros(B, 3, 3)
    K[:, 0, 0] = 0.58 * W
    K[:, 1, 1] = 0.58 * W
    K[:, 0, 2] = W / 2.0
    K[:, 1, 2] = H / 2.0
    K[:, 2, 2] = 1.0
    return K


def intrinsics_pyramid(K, S):
    """[B,3,3] -> [B,S,3,3] (Demon_Data_loader.py:25-39)."""
    out = []
    for s in range(S):
        Ks = K.clone()
        Ks[:, 0, 0] = K[:, 0, 0] / (2 ** s)
        Ks[:, 1, 1] = K[:, 1, 1] / (2 ** s)
        Ks[:, 0, 2] = K[:, 0, 2] / (2 ** s)
        Ks[:, 1, 2] = K[:, 1, 2] / (2 ** s)
        out.append(Ks)
    return torch.stack(out, dim=1)


def make_snippets(B, H, W, S=4, V=2, seed=1234, motion=1.0, hard=False):
    """One batch of synthetic training inputs for the view-synthesis loss.

    -> dict(tgt [B,H,W,3], srcs list V x [B,H,W,3], disp_pyr list S x [B,Hs,Ws,1],
            poses [B,V,6], K [B,3,3], K_pyr [B,S,3,3], logits_pyr list S x [B,Hs,Ws,2V])
    `motion` scales the pose magnitudes (>= 5 puts ~30 % of the pixels out of view);
    `hard` swaps the smooth textures for i.i.d. U[0,1] (throughput-only runs).
    """
    g = torch.Generator().manual_seed(seed)
    if hard:
        tgt = torch.rand(B, H, W, 3, generator=g)
        srcs = [torch.rand(B, H, W, 3, generator=g) for _ in range(V)]
    else:
        tgt = _texture(g, B, H, W, 3)
        srcs = [_texture(g, B, H, W, 3) for _ in range(V)]
    disp0 = torch.clamp(4.0 * torch.sigmoid(_lowpass(g, B, H, W)), 0.05, 4.0)  # [B,1,H,W]
    disp_pyr = []
    for s in range(S):
        hs, ws = H // 2 ** s, W // 2 ** s
        d = torch.nn.functional.adaptive_avg_pool2d(disp0, (hs, ws))
        d = d + 0.01 * torch.randn(B, 1, hs, ws, generator=g)
        disp_pyr.append(torch.clamp(d, 0.05, 4.0).permute(0, 2, 3, 1).contiguous())
    mean_depth = (1.0 / disp0).mean().item()
    t = (torch.rand(B, V, 3, generator=g) * 2 - 1) * 0.1 * mean_depth * motion
    r = (torch.rand(B, V, 3, generator=g) * 2 - 1) * 0.02 * motion
    poses = torch.cat([t, r], dim=2).contiguous()
    K = intrinsics(B, H, W)
    logits_pyr = [torch.randn(B, H // 2 ** s, W // 2 ** s, 2 * V, generator=g) for s in range(S)]
    return dict(tgt=tgt, srcs=srcs, disp_pyr=disp_pyr, poses=poses, K=K,
                K_pyr=intrinsics_pyramid(K, S), logits_pyr=logits_pyr)
