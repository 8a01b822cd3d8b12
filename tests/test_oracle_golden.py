"""CPU: the oracle restatement (oracle/vsl_oracle.py) against the golden vectors that were produced by
executing the reference's own source over the TF1 shim (tests/golden/make_golden.py).

Same op order on the same torch-CPU kernels => the fp32 comparisons are bit-exact except where libm-level
sin/cos enter (none here: both sides call torch.sin/cos)."""
import torch

from oracle import vsl_oracle as O
from tests.conftest import rel_err

WARP_CASES = ['warp_eular', 'warp_angleaxis', 'warp_v1_eular', 'warp_matrix_far', 'warp_identity',
              'warp_eular_wild']


def _functional(c, outs):
    Rs = [c.R_img, c.R_coords, c.R_wmask, c.R_z]
    n = 3 if 'z' not in c else 4
    return sum((o * r.to(o.dtype)).sum() for o, r in zip(outs[:n], Rs))


def test_warp_forward_bit_exact(golden):
    for name in WARP_CASES:
        c = golden[name]
        outs = O.projective_inverse_warp(c.img, c.depth, c.pose, c.K, c.format)
        assert torch.equal(outs[0], c.out), name
        assert torch.equal(outs[1], c.coords), name
        assert torch.equal(outs[2], c.wmask), name
        if 'z' in c:
            assert torch.equal(outs[3], c.z), name
            assert torch.equal(outs[4], c.pose_mat), name


def test_warp_gradients(golden):
    for name in WARP_CASES:
        c = golden[name]
        for dt, tag, tol in ((torch.float32, 'f32', 1e-6), (torch.float64, 'f64', 1e-12)):
            a = [c.img.to(dt).requires_grad_(), c.depth.to(dt).requires_grad_(), c.pose.to(dt).requires_grad_()]
            outs = O.projective_inverse_warp(a[0], a[1], a[2], c.K.to(dt), c.format)
            g = torch.autograd.grad(_functional(c, outs), a)
            for got, key in zip(g, ('g_img_', 'g_depth_', 'g_pose_')):
                assert rel_err(got, c[key + tag]) <= tol, (name, key, tag)


def test_identity_pose_kat(golden):
    c = golden['warp_identity']
    out, coords, wmask, z, _ = O.projective_inverse_warp(c.img, c.depth, c.pose, c.K, 'matrix')
    B, H, W, _ = c.img.shape
    grid = O.meshgrid(B, H, W, is_homogeneous=False).permute(0, 2, 3, 1)
    # K K^-1 p reproduces the fp32 grid to a few ulp; the warp reproduces the source to 1e-5
    assert (coords - grid).abs().max() < 2e-4
    assert (out - c.img).abs().max() < 2e-5 * 40
    assert (wmask[:, :-1, :-1] - 1).abs().max() < 1e-4
    assert torch.allclose(z.squeeze(3), c.depth)


def test_pose(golden):
    c = golden['pose']
    for fmt in ('eular', 'angleaxis'):
        assert torch.equal(O.pose_vec2mat(c.vec, fmt), c['mat_' + fmt])
        for dt, tag, tol in ((torch.float32, 'f32', 1e-6), (torch.float64, 'f64', 1e-12)):
            v = c.vec.to(dt).requires_grad_()
            g, = torch.autograd.grad((O.pose_vec2mat(v, fmt) * c.R.to(dt)).sum(), v)
            assert rel_err(g, c['g_%s_%s' % (fmt, tag)]) <= tol
    assert float(c['g_eular_f64'][0, 3]) == 0.0 and float(c['g_eular_f64'][0, 4]) == 0.0  # clipped at +-pi
    m0 = O.pose_vec2mat(torch.zeros(2, 6), 'angleaxis')
    assert torch.isnan(m0[:, :3, :3]).all() and torch.isnan(c.mat_angleaxis_zero[:, :3, :3]).all()


def test_sampler(golden):
    for name in ('sampler_c3', 'sampler_c1'):
        c = golden[name]
        out, wm = O.bilinear_sampler(c.imgs, c.coords)
        assert torch.equal(out, c.out) and torch.equal(wm, c.wmask)
        assert float(wm.min()) >= 0.0 and float(wm.max()) <= 1.0 + 1e-6
        for dt, tag, tol in ((torch.float32, 'f32', 1e-6), (torch.float64, 'f64', 1e-12)):
            a, co = c.imgs.to(dt).requires_grad_(), c.coords.to(dt).requires_grad_()
            o, w = O.bilinear_sampler(a, co)
            gi, gc = torch.autograd.grad((o * c.R.to(dt)).sum() + (w * c.Rm.to(dt)).sum(), [a, co])
            assert rel_err(gi, c['g_imgs_' + tag]) <= tol and rel_err(gc, c['g_coords_' + tag]) <= tol


def test_sampler_border_semantics(golden):
    """Zero padding, not clamping (utils.py:266-270): x = -1 and x = W give 0, x = -0.5 gives half."""
    img = torch.arange(1.0, 7.0).reshape(1, 1, 6, 1)

    def at(x):
        return float(O.bilinear_sampler(img, torch.tensor([[[[x, 0.0]]]]))[0])
    assert at(-1.0) == 0.0 and at(6.0) == 0.0 and at(-0.5) == 0.5 and at(5.5) == 3.0 and at(5.0) == 6.0


def test_flow_and_consistency(golden):
    c = golden['optflow']
    assert torch.equal(O.optflow_warp(c.img, c.flowx, c.flowy), c.out)
    c = golden['consist']
    fx, fy = O.depth_optflow(c.coords)
    assert torch.equal(fx, c.flowx) and torch.equal(fy, c.flowy)
    assert torch.equal(O.consistent_depth_loss(c.src_depth, c.z, c.coords), c.err)


def test_loss_terms(golden):
    c = golden['terms']
    for dt, tag, tol in ((torch.float32, 'f32', 1e-6), (torch.float64, 'f64', 1e-12)):
        p, l = c.disp.to(dt).requires_grad_(), c.logits.to(dt).requires_grad_()
        sm, smi = O.compute_smooth_loss(p), O.compute_smooth_loss(1.0 / p)
        ex = O.compute_exp_reg_loss(l, O.get_reference_explain_mask(0, 2, 10, 14, dt))
        assert rel_err(sm, c['smooth_' + tag]) <= tol and rel_err(smi, c['smooth_inv_' + tag]) <= tol
        assert rel_err(ex, c['exp_' + tag]) <= tol
        assert rel_err(torch.autograd.grad(sm, p)[0], c['g_smooth_' + tag]) <= tol
        assert rel_err(torch.autograd.grad(smi, p)[0], c['g_smooth_inv_' + tag]) <= tol
        assert rel_err(torch.autograd.grad(ex, l)[0], c['g_exp_' + tag]) <= tol
    assert float(c.smooth_quad) == 8.0 and float(O.compute_smooth_loss(c.quad)) == 8.0
    lin = torch.arange(20.0).reshape(1, 4, 5, 1) * 0.5 + 3
    assert float(O.compute_smooth_loss(lin)) == 0.0 and float(O.compute_smooth_loss(torch.ones(1, 5, 5, 1))) == 0.0


def test_pyramid_and_intrinsics(golden):
    c = golden['pyramid']
    for s in (1, 2, 3):
        assert torch.equal(O.resize_area(c.img, 16 >> s, 24 >> s), c['l%d' % s])
    assert torch.equal(O.multi_scale_intrinsics(c.K, 4), c.K_pyr)
    # the PRODUCT's own builders of the multi-scale intrinsics (what bench.py, the tests and the loader feed the kernels
    # with) against the matrices the reference's get_multi_scale_intrinsics produced (Demon_Data_loader.py:25-39)
    from tf_depth_estimation_b200 import data, synth
    assert torch.equal(synth.intrinsics_pyramid(c.K, 4), c.K_pyr)
    assert torch.equal(data.multi_scale_intrinsics(c.K, 4, 1.0, 1.0), c.K_pyr)


def test_composite_loss(golden):
    for name in ('loss_sfm', 'loss_lr', 'loss_nomask'):
        c = golden[name]
        fl = c.flags
        S, V = fl['num_scales'], fl['V']
        flags = O.LossFlags(**{k: v for k, v in fl.items() if k not in ('mask', 'V')})
        for dt, tag, tol in ((torch.float32, 'f32', 2e-6), (torch.float64, 'f64', 1e-12)):
            xs = [c['x%d' % s].to(dt).requires_grad_() for s in range(S)]
            ps = c.poses.to(dt).requires_grad_()
            lgs = [c['logits%d' % s].to(dt).requires_grad_() for s in range(S)] if fl['mask'] else None
            pixel, smooth, exp = O.view_synthesis_loss(
                c.tgt.to(dt), [c['src%d' % v].to(dt) for v in range(V)], xs, ps, c.K_pyr.to(dt), lgs, None, flags)
            assert rel_err(pixel, c['pixel_' + tag]) <= tol and rel_err(smooth, c['smooth_' + tag]) <= tol
            if fl['mask']:
                assert rel_err(exp, c['exp_' + tag]) <= tol
            grads = torch.autograd.grad(pixel + smooth + exp, xs + [ps] + (lgs or []))
            for s in range(S):
                assert rel_err(grads[s], c['g_x%d_%s' % (s, tag)]) <= tol, (name, s, tag)
                if fl['mask']:
                    assert rel_err(grads[S + 1 + s], c['g_logits%d_%s' % (s, tag)]) <= tol
            assert rel_err(grads[S], c['g_poses_' + tag]) <= tol


def test_lr_loss_with_consistency_term(golden_consist):
    """The oracle's step with the depth-consistency term, called once per warp direction, reproduces the
    reference-executed left-right loss of train_depth_then_cam_lr.py:211-340 (values and every gradient)."""
    from tests.parity_util import lr_flags, lr_two_directions
    c = golden_consist
    flags = lr_flags(O.LossFlags, c.flags)
    S = flags.num_scales
    for dt, tag, tol in ((torch.float32, 'f32', 2e-6), (torch.float64, 'f64', 1e-12)):
        fn = lambda tgt, srcs, xs, ps, K, lgs, sx, f: O.view_synthesis_loss(tgt, srcs, xs, ps, K, lgs, None, f, src_x_pyr=sx)
        terms, L = lr_two_directions(fn, c, lambda t, g: t.to(dt).requires_grad_(g), flags)
        for t, key in zip(terms, ('pixel', 'smooth', 'exp', 'consist')):
            assert rel_err(t, c['%s_%s' % (key, tag)]) <= tol, (key, tag)
        wrt = L['pl'] + L['pr'] + [L['po_r'], L['po_l']] + L['ll'] + L['lr']
        grads = torch.autograd.grad(sum(terms), wrt)
        names = (['g_pred_left%d' % s for s in range(S)] + ['g_pred_right%d' % s for s in range(S)] +
                 ['g_pose_right', 'g_pose_left'] + ['g_lg_left%d' % s for s in range(S)] +
                 ['g_lg_right%d' % s for s in range(S)])
        for g, n in zip(grads, names):
            assert rel_err(g, c['%s_%s' % (n, tag)]) <= tol, (n, tag)


def test_flow_depth_loss(golden_flow):
    """The oracle's flow-and-depth loss reproduces the reference-executed loop of train_optflow_combine.py:138-240:
    the four terms bit for bit in float32, every gradient within accumulation-order rounding."""
    c = golden_flow
    flags = O.FlowLossFlags(**c.flags)
    S = flags.num_scales
    for dt, tag, tol in ((torch.float32, 'f32', 2e-6), (torch.float64, 'f64', 1e-12)):
        pd = [c['pred_depth%d' % s].to(dt).requires_grad_() for s in range(S)]
        fx = [c['pred_flowx%d' % s].to(dt).requires_grad_() for s in range(S)]
        fy = [c['pred_flowy%d' % s].to(dt).requires_grad_() for s in range(S)]
        terms = O.flow_depth_loss(c.left.to(dt), c.right.to(dt), c.label.to(dt), pd, fx, fy, c.proj.to(dt),
                                  c.K_pyr.to(dt), flags)
        for t, key in zip(terms, ('depth', 'smooth', 'optflow', 'pixel')):
            if dt == torch.float32:
                assert float(t) == float(c['%s_%s' % (key, tag)]), key
            assert rel_err(t, c['%s_%s' % (key, tag)]) <= tol, (key, tag)
        grads = torch.autograd.grad(sum(terms), pd + fx + fy)
        names = ['g_pred_depth%d' % s for s in range(S)] + ['g_pred_flowx%d' % s for s in range(S)] + \
                ['g_pred_flowy%d' % s for s in range(S)]
        for g, n in zip(grads, names):
            assert rel_err(g, c['%s_%s' % (n, tag)]) <= tol, (n, tag)


def test_depth_loss_golden_file_is_consistent():
    """tests/golden/depth_losses_golden.npz (the reference's compute_loss_single_depth body executed over the shim)
    against a direct restatement from the oracle's building blocks: pins the fixture and oracle/demon_ops.py."""
    import ast
    import os
    import numpy as np
    from oracle import demon_ops as D
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'depth_losses_golden.npz'))
    F = ast.literal_eval(str(z['flags']))
    label = torch.from_numpy(z['single/label'])
    w = float(D.ease_out_quad(float(z['single/step'].reshape(-1)[0]), 0, F['depth_sig_weight'], float(F['max_steps'] // 3)))
    depth, sig = 0.0, 0.0
    for s in range(F['num_scales']):
        pred = torch.from_numpy(z['single/pred%d' % s])
        lab = O.resize_area(label, F['resizedheight'] >> s, F['resizedwidth'] >> s)
        sig = sig + w * D.pointwise_l2_loss(D.scale_invariant_gradient(pred.permute(0, 3, 1, 2), [2], [1]),
                                           D.scale_invariant_gradient(lab.permute(0, 3, 1, 2), [2], [1]), 1e-6)
        depth = depth + D.replace_nonfinite(lab - pred).abs().mean() * F['depth_weight'] / 2 ** s
    assert abs(float(depth) - float(z['single/depth_loss'].reshape(-1)[0])) <= 1e-6 * float(z['single/depth_loss'].reshape(-1)[0])
    assert abs(float(sig) - float(z['single/sig_loss'].reshape(-1)[0])) <= 1e-6 * float(z['single/sig_loss'].reshape(-1)[0])
    assert z['pair/zeros'].tolist() == [0.0, 0.0, 0.0]


def test_extension_and_optimiser_oracles_known_answers():
    """The oracle pieces that have NO reference implementation behind them (SSIM, edge-aware smoothness: absent from
    the reference; Adam and DeMoN's ops: un-vendored third parties) at least reproduce hand-derived values."""
    from oracle import demon_ops as D
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 6, 7, 2, generator=g, dtype=torch.float64)
    assert float(O.ssim_dissimilarity(x, x).abs().max()) <= 1e-12                       # SSIM(x, x) = 1
    one, zero = torch.ones(1, 4, 4, 1, dtype=torch.float64), torch.zeros(1, 4, 4, 1, dtype=torch.float64)
    want = 0.5 * (1 - 1e-4 / (1 + 1e-4))                                                # means 1 / 0, no variance
    assert float((O.ssim_dissimilarity(one, zero) - want).abs().max()) <= 1e-12
    # one window by hand
    a, b = x[:, :3, :3, :1], torch.rand(1, 3, 3, 1, generator=g, dtype=torch.float64)
    ma, mb = a.mean(), b.mean()
    va, vb, cab = ((a - ma) ** 2).mean(), ((b - mb) ** 2).mean(), ((a - ma) * (b - mb)).mean()
    s = (2 * ma * mb + 1e-4) * (2 * cab + 9e-4) / ((ma * ma + mb * mb + 1e-4) * (va + vb + 9e-4))
    assert abs(float(O.ssim_dissimilarity(a, b)) - float(torch.clamp((1 - s) / 2, 0, 1))) <= 1e-12
    # edge-aware smoothness: a unit disparity step across a flat image costs 1 / (B H (W - 1)) per crossing row
    disp = torch.zeros(1, 4, 6, 1, dtype=torch.float64)
    disp[:, :, 3:] = 1.0
    img = torch.full((1, 4, 6, 3), 0.5, dtype=torch.float64)
    assert abs(float(O.edge_aware_smooth_loss(disp, img)) - 4.0 / (4 * 5)) <= 1e-12
    img[:, :, 3:] += 0.2                                                                 # an image edge at the same place
    assert abs(float(O.edge_aware_smooth_loss(disp, img)) - 4.0 / (4 * 5) * float(torch.exp(torch.tensor(-0.2)))) <= 1e-7
    # Adam, first step: update = lr g / (|g| + eps / sqrt(1 - beta2)) with the float32-rounded hyper-parameters
    import numpy as np
    gg = torch.tensor([2.0, -0.5, 1e-9], dtype=torch.float64)
    p, m, v = O.adam_step_tf(torch.zeros(3, dtype=torch.float64), gg, torch.zeros(3, dtype=torch.float64),
                             torch.zeros(3, dtype=torch.float64), 1, 0.1)
    lr, b2, eps = float(np.float32(0.1)), float(np.float32(0.999)), float(np.float32(1e-8))
    assert torch.allclose(p, -lr * gg / (gg.abs() + eps / (1 - b2) ** 0.5), rtol=1e-12, atol=0)
    assert torch.allclose(m, (1 - float(np.float32(0.9))) * gg, rtol=1e-12, atol=0)
    # DeMoN ops: scale-invariant gradient of a horizontal ramp, hole handling, easing
    ramp = torch.arange(6, dtype=torch.float64).reshape(1, 1, 1, 6).repeat(1, 1, 5, 1)
    sig = D.scale_invariant_gradient(ramp, [2], [1], 0.001)
    assert tuple(sig.shape) == (1, 2, 5, 6)
    xs = torch.arange(4, dtype=torch.float64)
    assert torch.allclose(sig[0, 0, 0, :4], 2 / (xs + 2 + xs + 0.001)) and float(sig[0, 0, :, 4:].abs().max()) == 0
    assert float(sig[0, 1].abs().max()) == 0                                            # constant along y
    hole = torch.tensor([1.0, float('inf'), float('nan'), -2.0])
    assert D.replace_nonfinite(hole).tolist() == [1.0, 0.0, 0.0, -2.0]
    assert abs(float(D.pointwise_l2_loss(torch.full((1, 2, 1, 1), 3.0), torch.zeros(1, 2, 1, 1), 0.0)) - 18 ** 0.5) <= 1e-6
    assert float(D.ease_out_quad(0.0, 0, 2.0, 100.0)) == 0.0 and float(D.ease_out_quad(100.0, 0, 2.0, 100.0)) == 2.0
    assert abs(float(D.ease_out_quad(50.0, 0, 2.0, 100.0)) - 1.5) <= 1e-12 and float(D.ease_out_quad(1e9, 0, 2.0, 100.0)) == 2.0
