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


def _rigid(t, r):
    """[B,3] translation + [B,3] small rotation angles -> [B,4,4] (Rz Ry Rx, what a loader's matrix holds)."""
    B = t.shape[0]
    cx, sx = torch.cos(r[:, 0]), torch.sin(r[:, 0])
    cy, sy = torch.cos(r[:, 1]), torch.sin(r[:, 1])
    cz, sz = torch.cos(r[:, 2]), torch.sin(r[:, 2])
    one, zero = torch.ones(B), torch.zeros(B)
    Rx = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], 1).reshape(B, 3, 3)
    Ry = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], 1).reshape(B, 3, 3)
    Rz = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], 1).reshape(B, 3, 3)
    T = torch.zeros(B, 4, 4)
    T[:, :3, :3] = Rz @ Ry @ Rx
    T[:, :3, 3] = t
    T[:, 3, 3] = 1.0
    return T.contiguous()


def make_flow_pairs(B, H, W, S=4, seed=4321, motion=1.0):
    """One batch of the DeMoN-pair family (train_optflow_combine.py:85-110; BASELINE configs[3]).

    -> dict(left, right [B,H,W,3]; label [B,H,W,1] ground-truth inverse depth; depth_pyr / flowx_pyr / flowy_pyr
            lists of S network outputs [B,Hs,Ws,1] (inverse depth, flow in pixels of that scale); proj [B,4,4] the
            loader's target-to-source transform; K [B,3,3]; K_pyr [B,S,3,3])
    """
    g = torch.Generator().manual_seed(seed)
    left, right = _texture(g, B, H, W, 3), _texture(g, B, H, W, 3)
    label = torch.clamp(4.0 * torch.sigmoid(_lowpass(g, B, H, W)), 0.05, 4.0).permute(0, 2, 3, 1).contiguous()
    disp0 = torch.clamp(label.permute(0, 3, 1, 2) * (1.0 + 0.2 * _lowpass(g, B, H, W)), 0.05, 4.0)
    fx0, fy0 = 4.0 * motion * _lowpass(g, B, H, W), 3.0 * motion * _lowpass(g, B, H, W)
    depth_pyr, fx_pyr, fy_pyr = [], [], []
    for s in range(S):
        hs, ws = H // 2 ** s, W // 2 ** s
        pool = lambda x: torch.nn.functional.adaptive_avg_pool2d(x, (hs, ws))
        noise = lambda a: a * torch.randn(B, 1, hs, ws, generator=g)
        nhwc = lambda x: x.permute(0, 2, 3, 1).contiguous()
        depth_pyr.append(nhwc(torch.clamp(pool(disp0) + noise(0.01), 0.05, 4.0)))
        fx_pyr.append(nhwc(pool(fx0) / 2 ** s + noise(0.05)))
        fy_pyr.append(nhwc(pool(fy0) / 2 ** s + noise(0.05)))
    mean_depth = (1.0 / label).mean().item()
    t = (torch.rand(B, 3, generator=g) * 2 - 1) * 0.1 * mean_depth * motion
    r = (torch.rand(B, 3, generator=g) * 2 - 1) * 0.02 * motion
    K = intrinsics(B, H, W)
    return dict(left=left, right=right, label=label, depth_pyr=depth_pyr, flowx_pyr=fx_pyr, flowy_pyr=fy_pyr,
                proj=_rigid(t, r), K=K, K_pyr=intrinsics_pyramid(K, S))
