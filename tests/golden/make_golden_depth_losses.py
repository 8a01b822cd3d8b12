"""Generates tests/golden/depth_losses_golden.npz by EXECUTING THE REFERENCE'S OWN function bodies
compute_loss_single_depth / compute_loss_pairwise_depth (my_losses.py:46, :101) over the TF1 shim, with
/root/reference/utils_lr.py imported unmodified for pose_vec2mat / projective_inverse_warp and the four
un-vendored DeMoN ops supplied by oracle/demon_ops.py (restated, parity unpinned).

Run in the build container only:  python tests/golden/make_golden_depth_losses.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import demon_ops, ref_loader  # noqa: E402
from tf_depth_estimation_b200 import synth  # noqa: E402

assert ref_loader.available(), 'needs /root/reference'
ref_lr = ref_loader.load_module('utils_lr.py', 'ref_utils_lr')
import tensorflow as tf  # noqa: E402  (the shim)

g = dict(demon_ops.DEMON_GLOBALS)
g.update(pose_vec2mat=ref_lr.pose_vec2mat, projective_inverse_warp=ref_lr.projective_inverse_warp)
fns = ref_loader.load_functions('my_losses.py', ['compute_loss_single_depth', 'compute_loss_pairwise_depth'], g)


class Flags(object):
    num_scales, resizedheight, resizedwidth, batch_size = 4, 32, 64, 2
    depth_weight, depth_sig_weight, max_steps = 10.0, 2.0, 300
    cam_weight_rot, cam_weight_tran = 3.0, 5.0


OUT = {}


def put(case, **arrays):
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().numpy()
        OUT['%s/%s' % (case, k)] = np.ascontiguousarray(v)


def T(x, dt, grad=False):
    t = x.detach().clone().to(dt).as_subclass(tf.Tensor)
    t.requires_grad_(grad)
    return t


plain = lambda t: t.detach().as_subclass(torch.Tensor)
F = Flags
B, H, W, S = F.batch_size, F.resizedheight, F.resizedwidth, F.num_scales
d = synth.make_snippets(B, H, W, S=S, V=1, seed=4242)
gen = torch.Generator().manual_seed(99)
label = (1.0 / d['disp_pyr'][0]) * (1.0 + 0.05 * torch.randn(B, H, W, 1, generator=gen))       # ground-truth map
label[0, 3, 5, 0] = float('inf')                                                                 # a hole: replace_nonfinite
pred = [(1.0 / x) for x in d['disp_pyr']]
step = 40

# ---- compute_loss_single_depth
put('single', label=label, step=np.array(step), **{'pred%d' % s: pred[s] for s in range(S)})
for dt, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
    tf.set_float(dt)
    ps = [T(p, dt, True) for p in pred]
    depth_loss, smooth_loss, sig = fns['compute_loss_single_depth'](ps, T(label, dt), torch.tensor(step), F)
    grads = torch.autograd.grad(depth_loss + sig, ps)
    if tag == 'f32':
        put('single', depth_loss=plain(depth_loss), sig_loss=plain(sig), smooth_loss=np.array(float(smooth_loss)))
    put('single', **{'g_pred%d_%s' % (s, tag): plain(gr) for s, gr in enumerate(grads)})

# ---- compute_loss_pairwise_depth: scales 2..S-1 use pred_depth_*[s-2]
img_l, img_r = d['tgt'], d['srcs'][0]
K_pyr = d['K_pyr']
pdl = [pred[s] for s in range(2, S)]
pdr = [pred[s] * 1.03 for s in range(2, S)]
gt_cam = d['poses'][:, 0].clone()
poses_r = (gt_cam + 0.01 * torch.randn(B, 6, generator=gen)).reshape(B, 1, 6)
poses_l = (-gt_cam + 0.01 * torch.randn(B, 6, generator=gen)).reshape(B, 1, 6)
label_pair = torch.where(torch.isfinite(label), label, torch.ones_like(label))   # 1/label is warped with: keep finite
put('pair', image_left=img_l, image_right=img_r, gt_right_cam=gt_cam, pred_poses_right=poses_r, pred_poses_left=poses_l,
    intrinsics=K_pyr, label=label_pair, step=np.array(step),
    **{'pred_depth_left%d' % i: t for i, t in enumerate(pdl)}, **{'pred_depth_right%d' % i: t for i, t in enumerate(pdr)})
for dt, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
    tf.set_float(dt)
    a_l = [T(t, dt, True) for t in pdl]
    a_r = [T(t, dt, True) for t in pdr]
    p_r, p_l = T(poses_r, dt, True), T(poses_l, dt, True)
    out = fns['compute_loss_pairwise_depth'](T(img_l, dt), T(img_r, dt), a_l, p_r, None, a_r, p_l, None, T(gt_cam, dt),
                                             T(K_pyr, dt), T(label_pair, dt), F, torch.tensor(step))
    depth_loss, cam_loss, pixel_loss, consist_loss, sig, exp_loss = out[:6]
    Rr = torch.Generator().manual_seed(7)
    # a linear functional of the right-from-left warps keeps pred_depth_right in the graph (the reference returns
    # them for visualisation only)
    extra = sum((w * torch.randn(w.shape, generator=Rr).to(dt)).sum() for w in out[9]) * 1e-3
    grads = torch.autograd.grad(depth_loss + cam_loss + sig + extra, a_l + a_r + [p_r, p_l], allow_unused=True)
    if tag == 'f32':
        put('pair', depth_loss=plain(depth_loss), cam_loss=plain(cam_loss), sig_loss=plain(sig),
            zeros=np.array([float(pixel_loss), float(consist_loss), float(exp_loss)]))
        for name, lst in zip(('left_image', 'right_image', 'proj_image_left', 'proj_image_right', 'proj_error'), out[6:]):
            put('pair', **{'%s%d' % (name, i): plain(t) for i, t in enumerate(lst)})
    names = ['g_pdl%d' % i for i in range(len(a_l))] + ['g_pdr%d' % i for i in range(len(a_r))] + ['g_poses_r', 'g_poses_l']
    put('pair', **{'%s_%s' % (n, tag): plain(gr) if gr is not None else np.zeros(1) for n, gr in zip(names, grads)})
tf.set_float(torch.float32)
OUT['flags'] = np.array(repr({k: getattr(F, k) for k in dir(F) if not k.startswith('_')}))
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'depth_losses_golden.npz'), **OUT)
print('wrote', len(OUT), 'arrays')
for k in ('single/depth_loss', 'single/sig_loss', 'pair/depth_loss', 'pair/cam_loss', 'pair/sig_loss', 'pair/zeros'):
    print(k, OUT[k])
