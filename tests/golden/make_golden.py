"""Generates tests/golden/vsl_golden.npz by EXECUTING THE REFERENCE'S OWN SOURCE.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

/root/reference/utils_lr.py and utils.py are imported unmodified over the torch-backed TF1 shim
(oracle/tf1_shim); compute_smooth_loss / compute_exp_reg_loss / get_reference_explain_mask are compiled
out of my_losses.py and get_multi_scale_intrinsics / make_intrinsics_matrix out of Demon_Data_loader.py.
The multi-scale loss loop is not a function in the reference (it is inlined in every train script), so it
is re-assembled here from those reference functions following train.py:107-135 (photometric, smoothness,
1/2^s weights) and train_depth_then_cam_lr.py:297-328 (explainability mask + regulariser).

Every case stores its inputs, its fp32 outputs, and gradients of a fixed random linear functional of the
outputs obtained by torch autograd through the reference graph, in fp32 and in fp64 (shim float switch).
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
ref_v1 = ref_loader.load_module('utils.py', 'ref_utils_v1')
import tensorflow as tf  # noqa: E402  (the shim)

losses = ref_loader.load_functions(
    'my_losses.py', ['compute_smooth_loss', 'compute_exp_reg_loss', 'get_reference_explain_mask'])
intr = ref_loader.load_functions(
    'Demon_Data_loader.py', ['make_intrinsics_matrix', 'get_multi_scale_intrinsics'])

OUT = {}


def put(case, **arrays):
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().numpy()
        OUT['%s/%s' % (case, k)] = np.ascontiguousarray(v)


def T(x, dtype, grad=False):
    t = x.detach().clone().to(dtype).as_subclass(tf.Tensor)
    t.requires_grad_(grad)
    return t


def plain(t):
    return t.detach().as_subclass(torch.Tensor)


def run_warp(case, img, depth, pose, K, fmt, api='lr', seed=0):
    """Forward in fp32; gradients of sum_i <R_i, out_i> wrt img/depth/pose in fp32 and fp64."""
    g = torch.Generator().manual_seed(1000 + seed)
    B, H, W, C = img.shape
    R = [torch.randn(B, H, W, 3, generator=g), 0.01 * torch.randn(B, H, W, 2, generator=g),
         torch.randn(B, H, W, 1, generator=g), 0.1 * torch.randn(B, H, W, 1, generator=g)]
    put(case, img=img, depth=depth, pose=pose, K=K, R_img=R[0], R_coords=R[1], R_wmask=R[2], R_z=R[3])
    OUT['%s/format' % case] = np.array(fmt)
    for dt, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
        tf.set_float(dt)
        a = [T(img, dt, True), T(depth, dt, True), T(pose, dt, True), T(K, dt)]
        if api == 'lr':
            outs = ref_lr.projective_inverse_warp(a[0], a[1], a[2], a[3], fmt)
        else:
            outs = ref_v1.projective_inverse_warp(a[0], a[1], a[2], a[3])
        L = sum((o * r.to(dt)).sum() for o, r in zip(outs[:4 if api == 'lr' else 3], R))
        gi, gd, gp = torch.autograd.grad(L, a[:3])
        if tag == 'f32':
            names = ['out', 'coords', 'wmask', 'z', 'pose_mat'][:len(outs)]
            put(case, **{n: plain(o) for n, o in zip(names, outs)})
        put(case, **{'g_img_' + tag: plain(gi), 'g_depth_' + tag: plain(gd), 'g_pose_' + tag: plain(gp)})
    tf.set_float(torch.float32)


def case_warps():
    d = synth.make_snippets(2, 16, 52, S=1, V=2, seed=11)          # cfg2 scale-3 shape
    depth = (1.0 / d['disp_pyr'][0]).squeeze(3)
    run_warp('warp_eular', d['srcs'][0], depth, d['poses'][:, 0], d['K'], 'eular', seed=1)
    run_warp('warp_angleaxis', d['srcs'][1], depth, d['poses'][:, 1], d['K'], 'angleaxis', seed=2)
    run_warp('warp_v1_eular', d['srcs'][0], depth, d['poses'][:, 1], d['K'], 'eular', api='v1', seed=3)
    # large motion: a good share of the samples leave the source image (zero padding, wmask < 1)
    d = synth.make_snippets(3, 24, 32, S=1, V=1, seed=12, motion=6.0)  # cfg4 scale-3 shape
    depth = (1.0 / d['disp_pyr'][0]).squeeze(3)
    mat = plain(ref_lr.pose_vec2mat(T(d['poses'][:, 0], torch.float32), 'eular'))
    run_warp('warp_matrix_far', d['srcs'][0], depth, mat, d['K'], 'matrix', seed=4)
    # identity pose: coords == fp32 grid, warp == src, wmask == 1 (analytic KAT, also stored)
    d = synth.make_snippets(1, 30, 40, S=1, V=1, seed=13)           # cfg5 scale-4-ish shape
    depth = (1.0 / d['disp_pyr'][0]).squeeze(3)
    run_warp('warp_identity', d['srcs'][0], depth, torch.eye(4).reshape(1, 4, 4), d['K'], 'matrix', seed=5)
    # behind-camera / z ~ 0 points are NOT special-cased by the reference (utils.py:136): rotate by ~pi/2
    d = synth.make_snippets(2, 28, 28, S=1, V=1, seed=14)           # 224x224 scale-3 shape
    depth = (1.0 / d['disp_pyr'][0]).squeeze(3)
    pose = torch.tensor([[0.3, -0.2, -0.6, 0.1, 1.2, -0.3], [0.0, 0.1, -1.5, 3.5, -0.2, 0.4]])  # rx clipped to pi
    run_warp('warp_eular_wild', d['srcs'][0], depth, pose, d['K'], 'eular', seed=6)


def case_pose():
    g = torch.Generator().manual_seed(21)
    vec = torch.cat([torch.randn(6, 3, generator=g), 0.7 * torch.randn(6, 3, generator=g)], 1)
    vec[0, 3:] = torch.tensor([3.3, -3.4, 0.5])  # beyond +-pi: clipped, zero gradient
    R = torch.randn(6, 4, 4, generator=g)
    put('pose', vec=vec, R=R)
    for fmt in ('eular', 'angleaxis'):
        for dt, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
            tf.set_float(dt)
            v = T(vec, dt, True)
            m = ref_lr.pose_vec2mat(v, fmt)
            gv, = torch.autograd.grad((m * R.to(dt)).sum(), v)
            if tag == 'f32':
                put('pose', **{'mat_' + fmt: plain(m)})
            put('pose', **{'g_%s_%s' % (fmt, tag): plain(gv)})
    tf.set_float(torch.float32)
    z = T(torch.zeros(2, 6), torch.float32)
    put('pose', mat_angleaxis_zero=plain(ref_lr.pose_vec2mat(z, 'angleaxis')))  # NaN rotation block


def case_sampler():
    g = torch.Generator().manual_seed(31)
    for C, name in ((3, 'sampler_c3'), (1, 'sampler_c1')):
        B, Hs, Ws, Ht, Wt = 2, 9, 13, 7, 11
        imgs = torch.rand(B, Hs, Ws, C, generator=g)
        coords = torch.stack([torch.rand(B, Ht, Wt, generator=g) * (Ws + 3) - 2,
                              torch.rand(B, Ht, Wt, generator=g) * (Hs + 3) - 2], -1)
        # exact-integer and border coordinates
        coords[0, 0, :6, 0] = torch.tensor([0.0, -1.0, Ws - 1.0, float(Ws), 5.0, -0.5])
        coords[0, 1, :6, 1] = torch.tensor([0.0, -1.0, Hs - 1.0, float(Hs), 3.0, Hs - 0.5])
        R, Rm = torch.randn(B, Ht, Wt, C, generator=g), torch.randn(B, Ht, Wt, 1, generator=g)
        put(name, imgs=imgs, coords=coords, R=R, Rm=Rm)
        for dt, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
            tf.set_float(dt)
            a, c = T(imgs, dt, True), T(coords, dt, True)
            out, wm = ref_lr.bilinear_sampler(a, c)
            gi, gc = torch.autograd.grad((out * R.to(dt)).sum() + (wm * Rm.to(dt)).sum(), [a, c])
            if tag == 'f32':
                put(name, out=plain(out), wmask=plain(wm))
            put(name, **{'g_imgs_' + tag: plain(gi), 'g_coords_' + tag: plain(gc)})
    tf.set_float(torch.float32)


def case_flow_and_consistency():
    g = torch.Generator().manual_seed(41)
    d = synth.make_snippets(2, 24, 32, S=1, V=1, seed=42)
    img = d['srcs'][0]
    fx, fy = 3 * torch.randn(2, 24, 32, 1, generator=g), 3 * torch.randn(2, 24, 32, 1, generator=g)
    out = ref_v1.optflow_warp(T(img, torch.float32), T(fx, torch.float32), T(fy, torch.float32))
    put('optflow', img=img, flowx=fx, flowy=fy, out=plain(out))
    depth = (1.0 / d['disp_pyr'][0]).squeeze(3)
    outs = ref_lr.projective_inverse_warp(T(img, torch.float32), T(depth, torch.float32),
                                          T(d['poses'][:, 0], torch.float32), T(d['K'], torch.float32), 'eular')
    ox, oy = ref_lr.depth_optflow(outs[1])
    src_depth = 1.0 / torch.clamp(d['disp_pyr'][0] + 0.05 * torch.randn(2, 24, 32, 1, generator=g), 0.05, 4)
    cons = ref_lr.consistent_depth_loss(T(src_depth, torch.float32), outs[3], outs[1])
    put('consist', coords=plain(outs[1]), z=plain(outs[3]), src_depth=src_depth,
        flowx=plain(ox), flowy=plain(oy), err=plain(cons))


def case_loss_terms():
    g = torch.Generator().manual_seed(51)
    disp = torch.rand(2, 10, 14, 1, generator=g) * 3 + 0.1
    logits = torch.randn(2, 10, 14, 2, generator=g)
    put('terms', disp=disp, logits=logits)
    for dt, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
        tf.set_float(dt)
        p, l = T(disp, dt, True), T(logits, dt, True)
        sm = losses['compute_smooth_loss'](p)
        smi = losses['compute_smooth_loss'](1.0 / p)

        class F(object):
            batch_size, resizedheight, resizedwidth = 2, 10, 14
        ref_mask = losses['get_reference_explain_mask'](0, F)
        ex = losses['compute_exp_reg_loss'](l, ref_mask)
        gs, = torch.autograd.grad(sm, p)
        gsi, = torch.autograd.grad(smi, p)
        gl, = torch.autograd.grad(ex, l)
        put('terms', **{'smooth_' + tag: plain(sm), 'smooth_inv_' + tag: plain(smi), 'exp_' + tag: plain(ex),
                        'g_smooth_' + tag: plain(gs), 'g_smooth_inv_' + tag: plain(gsi), 'g_exp_' + tag: plain(gl)})
    tf.set_float(torch.float32)
    # quadratic ramp KAT: q = x^2 + 2 y^2 + x y  ->  |dx2| = 2, |dy2| = 4, |dxdy| = |dydx| = 1  => 8
    ys, xs = torch.meshgrid(torch.arange(6.0), torch.arange(7.0), indexing='ij')
    q = (xs * xs + 2 * ys * ys + xs * ys).reshape(1, 6, 7, 1)
    put('terms', quad=q, smooth_quad=plain(losses['compute_smooth_loss'](T(q, torch.float32))))
    img = torch.rand(2, 16, 24, 3, generator=g)
    put('pyramid', img=img, **{'l%d' % s: plain(tf.image.resize_area(T(img, torch.float32), [16 >> s, 24 >> s]))
                               for s in (1, 2, 3)})
    K = synth.intrinsics(2, 128, 416)
    put('pyramid', K=K, K_pyr=plain(intr['get_multi_scale_intrinsics'](T(K, torch.float32), 4)))


def composite(tgt, srcs, x_pyr, poses, K_pyr, logits_pyr, flags, dt):
    """train.py:107-135 + train_depth_then_cam_lr.py:297-328, from reference functions only."""
    B, H, W, _ = tgt.shape
    pixel = smooth = exp = 0
    for s in range(flags['num_scales']):
        x = x_pyr[s]
        q = 1.0 / x if flags['smooth_on_inverse'] else x
        smooth += flags['smooth_weight'] / (2 ** s) * losses['compute_smooth_loss'](q)
        hs, ws = int(H / (2 ** s)), int(W / (2 ** s))
        tgt_s = tf.image.resize_area(tgt, [hs, ws])
        for v, src in enumerate(srcs):
            src_s = tf.image.resize_area(src, [hs, ws])
            warped = ref_lr.projective_inverse_warp(
                src_s, tf.squeeze(1.0 / x, axis=3), poses[:, v, :], K_pyr[:, s, :, :], flags['pose_format'])[0]
            err = tf.abs(warped - tgt_s)
            dw = flags['data_weight'] / (2 ** s) if flags['pixel_scale_norm'] else flags['data_weight']
            if logits_pyr is not None:
                class F(object):
                    batch_size, resizedheight, resizedwidth = B, H, W
                ref_mask = losses['get_reference_explain_mask'](s, F)
                lg = tf.slice(logits_pyr[s], [0, 0, 0, 2 * v], [-1, -1, -1, 2])
                exp += flags['explain_reg_weight'] * losses['compute_exp_reg_loss'](lg, ref_mask)
                m = tf.nn.softmax(lg)
                pixel += tf.reduce_mean(err * tf.expand_dims(m[:, :, :, 1], -1)) * dw
            else:
                pixel += tf.reduce_mean(err) * dw
    return pixel, smooth, exp


def case_composite():
    variants = {
        # SfMLearner-style: V=2, euler, exp mask, 1/2^s pixel weights, smoothness on the net output
        'loss_sfm': dict(num_scales=4, smooth_weight=0.5, data_weight=1.0, explain_reg_weight=0.2,
                         pose_format='eular', pixel_scale_norm=True, smooth_on_inverse=False, mask=True, V=2),
        # train_depth_then_cam_lr-style: angle-axis, no 1/2^s on the pixel term, smoothness on 1/x
        'loss_lr': dict(num_scales=4, smooth_weight=2.0, data_weight=10.0, explain_reg_weight=0.5,
                        pose_format='angleaxis', pixel_scale_norm=False, smooth_on_inverse=True, mask=True, V=1),
        # train.py-style: no mask at all
        'loss_nomask': dict(num_scales=3, smooth_weight=0.1, data_weight=1.0, explain_reg_weight=0.0,
                            pose_format='eular', pixel_scale_norm=True, smooth_on_inverse=False, mask=False, V=2),
    }
    for i, (case, fl) in enumerate(variants.items()):
        d = synth.make_snippets(2, 32, 48, S=fl['num_scales'], V=fl['V'], seed=60 + i, motion=2.0)
        put(case, tgt=d['tgt'], poses=d['poses'], K_pyr=d['K_pyr'],
            **{'src%d' % v: s for v, s in enumerate(d['srcs'])},
            **{'x%d' % s: x for s, x in enumerate(d['disp_pyr'])},
            **{'logits%d' % s: l for s, l in enumerate(d['logits_pyr'])})
        OUT[case + '/flags'] = np.array(repr(fl))
        for dt, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
            tf.set_float(dt)
            xs = [T(x, dt, True) for x in d['disp_pyr']]
            ps = T(d['poses'], dt, True)
            lgs = [T(l, dt, True) for l in d['logits_pyr']] if fl['mask'] else None
            pixel, smooth, exp = composite(T(d['tgt'], dt), [T(s, dt) for s in d['srcs']], xs, ps,
                                           T(d['K_pyr'], dt), lgs, fl, dt)
            total = pixel + smooth + exp
            wrt = xs + [ps] + (lgs or [])
            grads = torch.autograd.grad(total, wrt)
            S = fl['num_scales']
            put(case, **{'pixel_' + tag: plain(pixel), 'smooth_' + tag: plain(smooth),
                         'exp_' + tag: plain(exp) if fl['mask'] else np.zeros(())})
            put(case, **{'g_x%d_%s' % (s, tag): plain(grads[s]) for s in range(S)})
            put(case, **{'g_poses_' + tag: plain(grads[S])})
            if fl['mask']:
                put(case, **{'g_logits%d_%s' % (s, tag): plain(grads[S + 1 + s]) for s in range(S)})
        tf.set_float(torch.float32)


if __name__ == '__main__':
    torch.set_num_threads(1)
    case_warps()
    case_pose()
    case_sampler()
    case_flow_and_consistency()
    case_loss_terms()
    case_composite()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'vsl_golden.npz')
    np.savez_compressed(path, **OUT)
    print('wrote %s: %d arrays, %.1f KiB' % (path, len(OUT), os.path.getsize(path) / 1024.0))
