"""Generates tests/golden/consist_golden.npz by EXECUTING THE REFERENCE'S OWN SOURCE.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_consist.py

The left-right loss of train_depth_then_cam_lr.py:211-340 -- smoothness on 1/x of both views, photometric error of
both warp directions weighted by the explainability masks, the mask regulariser, and the left-right DEPTH
CONSISTENCY term (consistent_depth_loss on the warp's projected depth and coordinates, :336-340) -- re-assembled from
the reference's functions (utils_lr.py imported unmodified over the torch-backed TF1 shim, the loss helpers compiled
out of my_losses.py) in the order the script applies them.  The terms of that loop that are outside the path
(supervised depth_loss, cam_loss, the `single` depth branches) are left out.

Stored: inputs, the four loss terms and the gradients of their sum w.r.t. both depth pyramids, both poses and both
logit pyramids, in fp32 and fp64 (shim float switch).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tf_depth_estimation_b200 import synth  # noqa: E402

assert ref_loader.available(), 'needs /root/reference'
ref_lr = ref_loader.load_module('utils_lr.py', 'ref_utils_lr')
import tensorflow as tf  # noqa: E402  (the shim)

losses = ref_loader.load_functions(
    'my_losses.py', ['compute_smooth_loss', 'compute_exp_reg_loss', 'get_reference_explain_mask'])

OUT = {}
FLAGS = dict(num_scales=4, smooth_weight=2.0, data_weight=10.0, explain_reg_weight=0.5, depth_weight=3.0)


def put(**arrays):
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().numpy()
        OUT['lr_consist/%s' % k] = np.ascontiguousarray(v)


def T(x, dtype, grad=False):
    t = x.detach().clone().to(dtype).as_subclass(tf.Tensor)
    t.requires_grad_(grad)
    return t


def plain(t):
    return t.detach().as_subclass(torch.Tensor)


def lr_loss(image_left, image_right, pred_left, pred_right, pose_right, pose_left, lg_left, lg_right, intrinsics):
    """train_depth_then_cam_lr.py:211-340 (the path's terms), names as in the script."""
    B, H, W, _ = image_left.shape

    class F(object):
        batch_size, resizedheight, resizedwidth = B, H, W
    pixel_loss = smooth_loss = exp_loss = consist_loss = 0
    for s in range(FLAGS['num_scales']):
        smooth_loss += FLAGS['smooth_weight'] / (2 ** s) * losses['compute_smooth_loss'](1.0 / pred_left[s])
        smooth_loss += FLAGS['smooth_weight'] / (2 ** s) * losses['compute_smooth_loss'](1.0 / pred_right[s])
        hs, ws = int(H / (2 ** s)), int(W / (2 ** s))
        curr_image_left = tf.image.resize_area(image_left, [hs, ws])
        curr_image_right = tf.image.resize_area(image_right, [hs, ws])
        proj_left, coords_right, _, warp_depth_right, _ = ref_lr.projective_inverse_warp(
            curr_image_right, tf.squeeze(1.0 / pred_left[s], axis=3), pose_right, intrinsics[:, s, :, :],
            format='angleaxis')
        err_left = tf.abs(proj_left - curr_image_left)
        proj_right, coords_left, _, warp_depth_left, _ = ref_lr.projective_inverse_warp(
            curr_image_left, tf.squeeze(1.0 / pred_right[s], axis=3), pose_left, intrinsics[:, s, :, :],
            format='angleaxis')
        err_right = tf.abs(proj_right - curr_image_right)
        ref_exp_mask = losses['get_reference_explain_mask'](s, F)
        logits = tf.slice(lg_left[s], [0, 0, 0, 0], [-1, -1, -1, 2])
        exp_loss += FLAGS['explain_reg_weight'] * losses['compute_exp_reg_loss'](logits, ref_exp_mask)
        exp_left = tf.nn.softmax(logits)
        pixel_loss += tf.reduce_mean(err_left * tf.expand_dims(exp_left[:, :, :, 1], -1)) * FLAGS['data_weight']
        logits = tf.slice(lg_right[s], [0, 0, 0, 0], [-1, -1, -1, 2])
        exp_loss += FLAGS['explain_reg_weight'] * losses['compute_exp_reg_loss'](logits, ref_exp_mask)
        exp_right = tf.nn.softmax(logits)
        pixel_loss += tf.reduce_mean(err_right * tf.expand_dims(exp_right[:, :, :, 1], -1)) * FLAGS['data_weight']
        right_err = ref_lr.consistent_depth_loss(1.0 / pred_right[s], warp_depth_right, coords_right)
        left_err = ref_lr.consistent_depth_loss(1.0 / pred_left[s], warp_depth_left, coords_left)
        consist_loss += tf.reduce_mean(right_err * tf.expand_dims(exp_left[:, :, :, 1], -1)) * FLAGS['depth_weight']
        consist_loss += tf.reduce_mean(left_err * tf.expand_dims(exp_right[:, :, :, 1], -1)) * FLAGS['depth_weight']
    return pixel_loss, smooth_loss, exp_loss, consist_loss


def main():
    S = FLAGS['num_scales']
    d = synth.make_snippets(2, 32, 48, S=S, V=1, seed=91, motion=2.0)
    e = synth.make_snippets(2, 32, 48, S=S, V=1, seed=92, motion=2.0)      # the right view's own network outputs
    left, right = d['tgt'], d['srcs'][0]
    pose_right, pose_left = d['poses'][:, 0], e['poses'][:, 0]
    put(image_left=left, image_right=right, pose_right=pose_right, pose_left=pose_left, K_pyr=d['K_pyr'],
        **{'pred_left%d' % s: x for s, x in enumerate(d['disp_pyr'])},
        **{'pred_right%d' % s: x for s, x in enumerate(e['disp_pyr'])},
        **{'lg_left%d' % s: l for s, l in enumerate(d['logits_pyr'])},
        **{'lg_right%d' % s: l for s, l in enumerate(e['logits_pyr'])})
    OUT['lr_consist/flags'] = np.array(repr(FLAGS))
    for dt, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
        tf.set_float(dt)
        pl = [T(x, dt, True) for x in d['disp_pyr']]
        pr = [T(x, dt, True) for x in e['disp_pyr']]
        po_r, po_l = T(pose_right, dt, True), T(pose_left, dt, True)
        ll = [T(l, dt, True) for l in d['logits_pyr']]
        lr = [T(l, dt, True) for l in e['logits_pyr']]
        terms = lr_loss(T(left, dt), T(right, dt), pl, pr, po_r, po_l, ll, lr, T(d['K_pyr'], dt))
        wrt = pl + pr + [po_r, po_l] + ll + lr
        grads = torch.autograd.grad(sum(terms), wrt)
        put(**{'%s_%s' % (k, tag): plain(v) for k, v in zip(('pixel', 'smooth', 'exp', 'consist'), terms)})
        put(**{'g_pred_left%d_%s' % (s, tag): plain(grads[s]) for s in range(S)})
        put(**{'g_pred_right%d_%s' % (s, tag): plain(grads[S + s]) for s in range(S)})
        put(**{'g_pose_right_' + tag: plain(grads[2 * S]), 'g_pose_left_' + tag: plain(grads[2 * S + 1])})
        put(**{'g_lg_left%d_%s' % (s, tag): plain(grads[2 * S + 2 + s]) for s in range(S)})
        put(**{'g_lg_right%d_%s' % (s, tag): plain(grads[3 * S + 2 + s]) for s in range(S)})
    tf.set_float(torch.float32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'consist_golden.npz')
    np.savez_compressed(path, **OUT)
    print('wrote %s: %d arrays, %.1f KiB' % (path, len(OUT), os.path.getsize(path) / 1024.0))


if __name__ == '__main__':
    torch.set_num_threads(1)
    main()
