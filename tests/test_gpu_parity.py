"""GPU parity: the CUDA path, called through the C ABI (ctypes binding in tf_depth_estimation_b200/_lib.py),
against (a) the committed golden vectors produced by executing the reference's source and (b) the CPU oracle
on freshly seeded inputs.

Tolerances are BASELINE.json's: warped images 1e-5 abs, forward losses 1e-5 rel, gradients 1e-4 rel
(max-norm: max|a-b| / max|b|).  Where the pose is given as a matrix no transcendental enters and the sampled
coordinates are required to be BIT-EXACT.  d(img) is accumulated with float atomics: its summation order is
not deterministic, the tolerance covers it.
"""
import pytest
import torch

from oracle import vsl_oracle as O
from tests.conftest import rel_err
from tests.parity_util import masked_rel_err, smooth_pixels
from tf_depth_estimation_b200 import ops, synth

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'
WARP_CASES = ['warp_eular', 'warp_angleaxis', 'warp_v1_eular', 'warp_matrix_far', 'warp_identity']


def cu(t, grad=False):
    return t.to(DEV).float().contiguous().requires_grad_(grad)


def test_warp_forward_golden(golden):
    for name in WARP_CASES:
        c = golden[name]
        out, coords, wmask, z, pose_mat = ops.projective_inverse_warp(cu(c.img), cu(c.depth), cu(c.pose), cu(c.K), c.format)
        if c.format == 'matrix':
            assert torch.equal(coords.cpu(), c.coords), name + ': coords must be bit-exact'
            assert torch.equal(z.cpu(), c.z), name
            assert torch.equal(wmask.cpu(), c.wmask), name
            assert torch.equal(out.cpu(), c.out), name
        else:
            assert (coords.cpu() - c.coords).abs().max() < 2e-4, name
        assert (out.cpu() - c.out).abs().max() <= 1e-5, (name, float((out.cpu() - c.out).abs().max()))
        assert (wmask.cpu() - c.wmask).abs().max() <= 2e-4, name
        if 'pose_mat' in c:
            assert (pose_mat.cpu() - c.pose_mat).abs().max() <= 1e-6, name


def test_warp_wild_pose_golden(golden):
    """Near-90-degree rotation and a clipped Euler angle: points behind the camera / z ~ 0 are not special-cased
    (utils.py:136).  Compare where the projection is well conditioned; everything must stay finite or NaN/Inf in
    the same places."""
    c = golden['warp_eular_wild']
    out, coords, wmask, z, _ = ops.projective_inverse_warp(cu(c.img), cu(c.depth), cu(c.pose), cu(c.K), 'eular')
    ok = (z.cpu().abs() > 1e-2).squeeze(3) & (c.coords.abs().amax(3) < 1e4)
    assert ok.float().mean() > 0.9
    assert ((coords.cpu() - c.coords).abs().amax(3)[ok] / c.coords.abs().amax(3)[ok].clamp_min(1.0)).max() < 1e-5
    assert (out.cpu() - c.out).abs().amax(3)[ok].max() <= 1e-4


def _warp_grads(c, wrt_img=True):
    a = [cu(c.img, wrt_img), cu(c.depth, True), cu(c.pose, True)]
    outs = ops.projective_inverse_warp(a[0], a[1], a[2], cu(c.K), c.format)
    Rs = [c.R_img, c.R_coords, c.R_wmask, c.R_z]
    n = 4 if 'z' in c else 3
    L = sum((o * cu(r)).sum() for o, r in zip(outs[:n], Rs))
    L.backward()
    return a


def test_warp_backward_golden(golden):
    for name in WARP_CASES:
        c = golden[name]
        a = _warp_grads(c)
        for t, key in zip(a, ('g_img_', 'g_depth_', 'g_pose_')):
            e32, e64 = rel_err(t.grad, c[key + 'f32']), rel_err(t.grad, c[key + 'f64'])
            assert min(e32, e64) <= 1e-4, (name, key, e32, e64)


def test_pose_vec2mat_golden(golden):
    c = golden['pose']
    for fmt in ('eular', 'angleaxis'):
        v = cu(c.vec, True)
        m = ops.pose_vec2mat(v, fmt)
        assert (m.cpu() - c['mat_' + fmt]).abs().max() <= 1e-6
        (m * cu(c.R)).sum().backward()
        assert rel_err(v.grad, c['g_%s_f64' % fmt]) <= 1e-5
    assert float(v.grad[0, 3]) != 0  # angle-axis has no clip
    v = cu(c.vec, True)
    (ops.pose_vec2mat(v, 'eular') * cu(c.R)).sum().backward()
    assert float(v.grad[0, 3]) == 0.0 and float(v.grad[0, 4]) == 0.0  # clipped at +-pi (utils.py:40-42)
    m0 = ops.pose_vec2mat(torch.zeros(2, 6, device=DEV), 'angleaxis')
    assert torch.isnan(m0[:, :3, :3]).all()  # reference behaviour: axis / 0 (utils_lr.py:133)


def test_sampler_golden(golden):
    for name in ('sampler_c3', 'sampler_c1'):
        c = golden[name]
        imgs, coords = cu(c.imgs, True), cu(c.coords, True)
        out, wm = ops.bilinear_sampler(imgs, coords)
        assert torch.equal(out.detach().cpu(), c.out) and torch.equal(wm.detach().cpu(), c.wmask), name
        ((out * cu(c.R)).sum() + (wm * cu(c.Rm)).sum()).backward()
        assert rel_err(imgs.grad, c.g_imgs_f64) <= 1e-5 and rel_err(coords.grad, c.g_coords_f64) <= 1e-5


def test_flow_and_consistency_golden(golden):
    c = golden['optflow']
    assert torch.equal(ops.optflow_warp(cu(c.img), cu(c.flowx), cu(c.flowy)).cpu(), c.out)
    c = golden['consist']
    fx, fy = ops.depth_optflow(cu(c.coords))
    assert torch.equal(fx.cpu(), c.flowx) and torch.equal(fy.cpu(), c.flowy)
    err = ops.consistent_depth_loss(cu(c.src_depth), cu(c.z), cu(c.coords))
    assert torch.equal(err.cpu(), c.err)


def test_optflow_warp_gradients():
    g = torch.Generator().manual_seed(5)
    d = synth.make_snippets(2, 24, 32, S=1, V=1, seed=3)
    fx, fy = 2 * torch.randn(2, 24, 32, 1, generator=g), 2 * torch.randn(2, 24, 32, 1, generator=g)
    R = torch.randn(2, 24, 32, 3, generator=g)
    a = [d['srcs'][0].double().requires_grad_(), fx.double().requires_grad_(), fy.double().requires_grad_()]
    (O.optflow_warp(*a) * R.double()).sum().backward()
    b = [cu(d['srcs'][0], True), cu(fx, True), cu(fy, True)]
    (ops.optflow_warp(*b) * cu(R)).sum().backward()
    for x, y in zip(a, b):
        assert rel_err(y.grad, x.grad) <= 1e-4


def test_consistent_depth_loss_gradients():
    """utils_lr.py:369-458 as one kernel each way: d/d(src_depth) (atomics), d/d(pred), d/d(coords)."""
    g = torch.Generator().manual_seed(9)
    B, Hs, Ws, Ht, Wt = 2, 20, 28, 24, 32
    src = 1.0 + 4.0 * torch.rand(B, Hs, Ws, 1, generator=g)
    pred = 1.0 + 4.0 * torch.rand(B, Ht, Wt, 1, generator=g)
    # coordinates over and beyond the source frame (zero padding on every side)
    coords = torch.stack([torch.rand(B, Ht, Wt, generator=g) * (Ws + 3) - 2,
                          torch.rand(B, Ht, Wt, generator=g) * (Hs + 3) - 2], dim=3)
    R = torch.randn(B, Ht, Wt, 1, generator=g)
    a = [src.double().requires_grad_(), pred.double().requires_grad_(), coords.double().requires_grad_()]
    (O.consistent_depth_loss(*a) * R.double()).sum().backward()
    b = [cu(src, True), cu(pred, True), cu(coords, True)]
    err = ops.consistent_depth_loss(*b)
    assert float((err.detach().cpu().double() - O.consistent_depth_loss(*[t.detach() for t in a])).abs().max()) <= 1e-5
    (err * cu(R)).sum().backward()
    for x, y in zip(a, b):
        assert rel_err(y.grad, x.grad) <= 1e-4
    # only the prediction asks for a gradient: the other two buffers are not produced
    p2 = cu(pred, True)
    (ops.consistent_depth_loss(cu(src), p2, cu(coords)) * cu(R)).sum().backward()
    assert torch.equal(p2.grad, b[1].grad)


def test_loss_terms_golden(golden):
    c = golden['terms']
    for inverse, key in ((False, 'smooth'), (True, 'smooth_inv')):
        p = cu(c.disp, True)
        sm = ops.compute_smooth_loss(p, inverse=inverse)
        assert rel_err(sm, c[key + '_f64']) <= 1e-5
        sm.backward()
        assert rel_err(p.grad, c['g_' + key + '_f64']) <= 1e-4
    l = cu(c.logits, True)
    ex = ops.compute_exp_reg_loss(l)
    assert rel_err(ex, c.exp_f64) <= 1e-5
    (3.0 * ex).backward()
    assert rel_err(l.grad, 3.0 * c.g_exp_f64) <= 1e-4
    assert abs(float(ops.compute_smooth_loss(cu(c.quad))) - 8.0) <= 8e-6  # quadratic ramp KAT
    lin = torch.arange(20.0).reshape(1, 4, 5, 1) * 0.5 + 3
    assert float(ops.compute_smooth_loss(cu(lin))) == 0.0


def test_pyramid_golden_bit_exact(golden):
    c = golden['pyramid']
    lv = ops.image_pyramid(cu(c.img), 4)
    for s in (1, 2, 3):
        assert torch.equal(lv[s].cpu(), c['l%d' % s])


def _run_fused(c, fl, S, V, exact=False):
    flags = ops.LossFlags(exact_coords=exact, **{k: v for k, v in fl.items() if k not in ('mask', 'V')})
    xs = [cu(c['x%d' % s], True) for s in range(S)]
    ps = cu(c.poses, True)
    lgs = [cu(c['logits%d' % s], True) for s in range(S)] if fl['mask'] else None
    total, losses = ops.view_synthesis_loss(cu(c.tgt), [cu(c['src%d' % v]) for v in range(V)], xs, ps, cu(c.K_pyr),
                                            logits_pyr=lgs, flags=flags)
    total.backward()
    return flags, losses, xs, ps, lgs


@pytest.mark.parametrize('exact', [False, True])
def test_fused_loss_golden(golden, exact):
    """Both arithmetic modes of the fused kernel against the reference-executed composite loss.  Per-pixel
    gradients are compared away from the loss's kinks (tests/parity_util.py); pose gradients (sums over all
    pixels) and the loss values are compared everywhere."""
    for name in ('loss_sfm', 'loss_lr', 'loss_nomask'):
        c = golden[name]
        fl = c.flags
        S, V = fl['num_scales'], fl['V']
        flags, losses, xs, ps, lgs = _run_fused(c, fl, S, V, exact)
        for i, key in enumerate(('pixel', 'smooth', 'exp')):
            want = float(c[key + '_f64'])
            assert abs(float(losses[i]) - want) <= 1e-5 * abs(want) + 1e-9, (name, key, float(losses[i]), want)
        ok = smooth_pixels(c.tgt, [c['src%d' % v] for v in range(V)], [c['x%d' % s] for s in range(S)], c.poses,
                           c.K_pyr, flags)
        for s in range(S):
            all_views = torch.stack(ok[s]).all(0)
            assert all_views.float().mean() > 0.97, (name, s, float(all_views.float().mean()))
            e = masked_rel_err(xs[s].grad, c['g_x%d_f64' % s], all_views.unsqueeze(3))
            assert e <= 1e-4, (name, 'g_x', s, e)
            if fl['mask']:
                m = torch.stack([o for o in ok[s] for _ in (0, 1)], dim=3)
                assert masked_rel_err(lgs[s].grad, c['g_logits%d_f64' % s], m) <= 1e-4, (name, 'g_logits', s)
        e = rel_err(ps.grad, c.g_poses_f64)
        assert e <= 1e-4, (name, 'g_poses', e)


def _consist_fn(tgt, srcs, xs, ps, K, lgs, sx, f):
    total, losses = ops.view_synthesis_loss(tgt, srcs, xs, ps, K, logits_pyr=lgs, flags=f, src_x_pyr=sx)
    assert losses.shape == (4,)
    return [total, losses]


def test_fused_consistency_term_golden(golden_consist):
    """The left-right loss of train_depth_then_cam_lr.py:211-340 -- photometric + smoothness + mask regulariser + the
    left-right depth-consistency term (:336-340) -- as TWO fused steps (one per warp direction, each one's source-depth
    pyramid being the other one's prediction), against the reference-executed golden: loss terms 1e-5, pose
    gradients 1e-4, per-pixel gradients 1e-4 away from the kinks.  d/d(pred) sums the direct path of one direction
    and the scattered d/d(source depth) of the other."""
    from tests.parity_util import lr_flags, lr_two_directions
    c = golden_consist
    flags = lr_flags(ops.LossFlags, c.flags)
    S = flags.num_scales
    (total, losses), L = lr_two_directions(_consist_fn, c, cu, flags)
    total.backward()
    for i, key in enumerate(('pixel', 'smooth', 'exp', 'consist')):
        want = float(c[key + '_f64'])
        assert abs(float(losses[i]) - want) <= 1e-5 * abs(want), (key, float(losses[i]), want)
    assert abs(float(total.detach()) - sum(float(c[k + '_f64']) for k in ('pixel', 'smooth', 'exp', 'consist'))) <= 1e-4
    assert rel_err(L['po_r'].grad, c.g_pose_right_f64) <= 1e-4, rel_err(L['po_r'].grad, c.g_pose_right_f64)
    assert rel_err(L['po_l'].grad, c.g_pose_left_f64) <= 1e-4
    preds = {'left': [c['pred_left%d' % s] for s in range(S)], 'right': [c['pred_right%d' % s] for s in range(S)]}
    ok_l = smooth_pixels(c.image_left, [c.image_right], preds['left'], c.pose_right.unsqueeze(1), c.K_pyr, flags,
                         src_x_pyr=[preds['right']])
    ok_r = smooth_pixels(c.image_right, [c.image_left], preds['right'], c.pose_left.unsqueeze(1), c.K_pyr, flags,
                         src_x_pyr=[preds['left']])
    for s in range(S):
        for side, ok, xs, lgs in (('left', ok_l, L['pl'], L['ll']), ('right', ok_r, L['pr'], L['lr'])):
            m = ok[s][0]
            assert m.float().mean() > 0.95, (side, s, float(m.float().mean()))
            e = masked_rel_err(xs[s].grad, c['g_pred_%s%d_f64' % (side, s)], m.unsqueeze(3))
            assert e <= 1e-4, ('g_pred', side, s, e)
            e = masked_rel_err(lgs[s].grad, c['g_lg_%s%d_f64' % (side, s)], m.unsqueeze(3))
            assert e <= 1e-4, ('g_lg', side, s, e)


def test_fused_consistency_term_matches_unfused_ops_and_oracle():
    """The consistency term inside the fused step == the reference-signature ops composed by hand
    (projective_inverse_warp -> consistent_depth_loss -> mask -> mean) and == the float64 oracle, for two source views
    at a ragged shape, with the explainability mask, a constant mask and no mask; depth = x and depth = 1/x."""
    B, H, W, S, V = 2, 40, 104, 3, 2
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=41, motion=2.0)
    g = torch.Generator().manual_seed(5)
    src_x = [[(x + 0.1 * torch.rand(x.shape, generator=g)) for x in d['disp_pyr']] for _ in range(V)]
    masks = [torch.rand(B, H >> s, W >> s, 1, generator=g) for s in range(S)]
    for mode, inv in (('exp', True), ('const', False), ('none', True)):
        kw = dict(num_scales=S, consist_weight=1.5, depth_is_inverse=inv, smooth_on_inverse=inv, pose_format='eular')
        flags, of = ops.LossFlags(**kw), O.LossFlags(**kw)
        xs = [cu(x, True) for x in d['disp_pyr']]
        sx = [[cu(t, True) for t in p] for p in src_x]
        ps = cu(d['poses'], True)
        lgs = [cu(l, True) for l in d['logits_pyr']] if mode == 'exp' else None
        mk = [cu(m) for m in masks] if mode == 'const' else None
        tgt, srcs, Kp = cu(d['tgt']), [cu(s) for s in d['srcs']], cu(d['K_pyr'])
        total, losses = ops.view_synthesis_loss(tgt, srcs, xs, ps, Kp, logits_pyr=lgs, mask_pyr=mk, flags=flags, src_x_pyr=sx)
        total.backward()
        # (a) the float64 oracle
        oxs = [x.double().requires_grad_() for x in d['disp_pyr']]
        osx = [[t.double().requires_grad_() for t in p] for p in src_x]
        ops_ = d['poses'].double().requires_grad_()
        ol = [l.double().requires_grad_() for l in d['logits_pyr']] if mode == 'exp' else None
        ref = O.view_synthesis_loss(d['tgt'].double(), [s.double() for s in d['srcs']], oxs, ops_, d['K_pyr'].double(), ol,
                                    [m.double() for m in masks] if mode == 'const' else None, of, src_x_pyr=osx)
        sum(ref).backward()
        for got, want in zip(losses.tolist(), ref):
            assert abs(got - float(want)) <= 1e-5 * abs(float(want)) + 1e-9, (mode, got, float(want))
        assert rel_err(ps.grad, ops_.grad) <= 1e-4, (mode, rel_err(ps.grad, ops_.grad))
        ok = smooth_pixels(d['tgt'], d['srcs'], d['disp_pyr'], d['poses'], d['K_pyr'], flags, src_x_pyr=src_x)
        for s in range(S):
            m = torch.stack(ok[s]).all(0)
            assert m.float().mean() > 0.95
            assert masked_rel_err(xs[s].grad, oxs[s].grad, m.unsqueeze(3)) <= 1e-4, (mode, 'g_x', s)
            for v in range(V):
                # the scatter into the source depth has no per-pixel mask: a flipped sign lands on four source pixels
                e = rel_err(sx[v][s].grad, osx[v][s].grad)
                assert e <= 2e-4, (mode, 'g_src_x', v, s, e)
        # (b) the stand-alone ops, reference call signatures
        xs2 = [cu(x, True) for x in d['disp_pyr']]
        sx2 = [[cu(t, True) for t in p] for p in src_x]
        ps2 = cu(d['poses'], True)
        sp = [ops.image_pyramid(s_, S) for s_ in srcs]
        consist = 0
        for s in range(S):
            depth = (1.0 / xs2[s] if inv else xs2[s]).squeeze(3)
            for v in range(V):
                _, coords, _, z_u, _ = ops.projective_inverse_warp(sp[v][s], depth, ps2[:, v].contiguous(),
                                                                   Kp[:, s].contiguous(), 'eular')
                err = ops.consistent_depth_loss(1.0 / sx2[v][s] if inv else sx2[v][s], z_u, coords)
                if mode == 'exp':
                    err = err * torch.softmax(lgs[s].detach()[..., 2 * v:2 * v + 2], -1)[..., 1:2]
                elif mode == 'const':
                    err = err * mk[s]
                consist = consist + err.mean() * flags.consist_weight
        assert abs(float(losses[3]) - float(consist)) <= 1e-5 * abs(float(consist)), (mode, float(losses[3]), float(consist))


def test_consistency_entry_argument_checks():
    """src_x_pyr and flags.consist_weight go together; shapes are checked before the C call; the combinations the
    kernel does not implement are refused loudly."""
    B, H, W, S, V = 1, 16, 32, 2, 1
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=2)
    a = (cu(d['tgt']), [cu(s) for s in d['srcs']], [cu(x) for x in d['disp_pyr']], cu(d['poses']), cu(d['K_pyr']))
    sx = [[cu(x) for x in d['disp_pyr']]]
    with pytest.raises(ValueError):
        ops.view_synthesis_loss(*a, flags=ops.LossFlags(num_scales=S), src_x_pyr=sx)
    with pytest.raises(ValueError):
        ops.view_synthesis_loss(*a, flags=ops.LossFlags(num_scales=S, consist_weight=1.0))
    with pytest.raises(ValueError):
        ops.view_synthesis_loss(*a, flags=ops.LossFlags(num_scales=S, consist_weight=1.0), src_x_pyr=[sx[0][::-1]])
    with pytest.raises(ValueError):
        ops.view_synthesis_loss(*a, flags=ops.LossFlags(num_scales=S, consist_weight=1.0, exact_coords=True), src_x_pyr=sx)


@pytest.mark.parametrize('shape', [(2, 32, 64, 4, 2), (1, 48, 104, 4, 1), (3, 16, 416, 1, 2), (2, 20, 36, 3, 2),
                                   (1, 64, 96, 5, 3), (2, 18, 30, 2, 1), (1, 96, 96, 6, 1), (1, 40, 104, 4, 2), (2, 128, 416, 4, 2)])
def test_prep_launch_products_bit_exact(shape):
    """What launch 1 leaves in the workspace -- resize_area levels of the target (RGB) and of every source view
    (zero-bordered RGBA; the fourth channel zero, or the source view's depth map with the consistency term) -- is
    BIT-identical to the oracle's resize_area (TF's ComputePatchSum order), for the register form of the launch
    (W % 4 == 0, at most 5 scales), the staged form (the rest) and the uint8 ingest."""
    B, H, W, S, V = shape
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=11)
    for variant in ('f32', 'u8', 'consist'):
        if variant == 'consist' and S > 5:
            continue
        kw = dict(num_scales=S)
        imgs = [d['tgt']] + list(d['srcs'])
        if variant == 'u8':
            kw['img_format'] = 'u8_255'
            u8 = [(t * 255).round().clamp(0, 255).to(torch.uint8) for t in imgs]
            imgs = [O.images_from_uint8(t, 'u8_255') for t in u8]
            dev_imgs = [t.to(DEV) for t in u8]
        else:
            dev_imgs = [cu(t) for t in imgs]
        if variant == 'consist':
            kw.update(consist_weight=1.0)
        flags = ops.LossFlags(**kw)
        plan = ops.ViewSynthesisPlan(B, H, W, V, flags, 1, torch.device(DEV))
        sx = [[cu(x + 0.25 * (v + 1)) for x in d['disp_pyr']] for v in range(V)] if variant == 'consist' else None
        plan.run(dev_imgs[0], dev_imgs[1:], [cu(x) for x in d['disp_pyr']], cu(d['poses']), cu(d['K_pyr']),
                 logits_pyr=[cu(l) for l in d['logits_pyr']], src_x_pyr=sx)
        torch.cuda.synchronize()
        tgt_lv, src_lv = plan.prep_levels()
        for s in range(S):
            hs, ws = H >> s, W >> s
            if tgt_lv[s] is not None:
                assert torch.equal(tgt_lv[s].cpu(), O.resize_area(imgs[0], hs, ws)), (variant, 'tgt', s)
            else:
                assert s == 0 and variant != 'u8'
            for v in range(V):
                got = src_lv[v][s].cpu()
                assert torch.equal(got[:, 2:-2, 2:-2, :3], O.resize_area(imgs[1 + v], hs, ws)), (variant, 'src', v, s)
                border = got.clone()
                border[:, 2:-2, 2:-2] = 0
                assert float(border.abs().max()) == 0.0, (variant, 'border', v, s)
                want_w = (1.0 / (d['disp_pyr'][s] + 0.25 * (v + 1)))[..., 0] if variant == 'consist' else torch.zeros(B, hs, ws)
                assert torch.equal(got[:, 2:-2, 2:-2, 3], want_w), (variant, 'w', v, s)


def test_fused_exact_mode_matches_standalone_warp_bitwise(golden):
    """exact_coords=True: the fused kernel's photometric term is built from the same rounded operations as the
    stand-alone warp, so with smoothing / regulariser off its pixel loss equals mean|warp - tgt| computed from
    the bit-exact warp output to float32 summation accuracy, and the per-pixel d/dx agree to 1e-6."""
    c = golden['loss_nomask']
    fl = dict(c.flags, smooth_weight=0.0, num_scales=1)
    flags = ops.LossFlags(exact_coords=True, **{k: v for k, v in fl.items() if k not in ('mask', 'V')})
    x = cu(c.x0, True)
    ps = cu(c.poses, True)
    srcs = [cu(c.src0), cu(c.src1)]
    total, losses = ops.view_synthesis_loss(cu(c.tgt), srcs, [x], ps, cu(c.K_pyr[:, :1]), flags=flags)
    total.backward()
    x2, ps2 = cu(c.x0, True), cu(c.poses, True)
    pixel = 0
    for v in range(2):
        warped = ops.projective_inverse_warp(srcs[v], (1.0 / x2).squeeze(3), ps2[:, v].contiguous(),
                                             cu(c.K_pyr[:, 0]), 'eular')[0]
        pixel = pixel + (warped - cu(c.tgt)).abs().mean()
    pixel.backward()
    assert abs(float(losses[0]) - float(pixel.detach())) <= 2e-7 * float(pixel.detach())
    assert rel_err(x.grad, x2.grad) <= 1e-6 and rel_err(ps.grad, ps2.grad) <= 1e-5


def test_fused_matches_unfused_composition():
    """Fused kernel == the same loop assembled from the stand-alone ops (reference call signatures), at a
    shape with ragged tiles (W not a multiple of 32, H not a multiple of 8 at the coarse scales)."""
    B, H, W, S, V = 3, 40, 104, 3, 2
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=21, motion=2.0)
    flags = ops.LossFlags(num_scales=S, pose_format='angleaxis', smooth_weight=0.7, data_weight=2.0,
                          explain_reg_weight=0.3, pixel_scale_norm=False, smooth_on_inverse=True)
    xs = [cu(x, True) for x in d['disp_pyr']]
    ps = cu(d['poses'], True)
    lgs = [cu(l, True) for l in d['logits_pyr']]
    tgt, srcs, Kp = cu(d['tgt']), [cu(s) for s in d['srcs']], cu(d['K_pyr'])
    total, losses = ops.view_synthesis_loss(tgt, srcs, xs, ps, Kp, logits_pyr=lgs, flags=flags)
    total.backward()

    xs2 = [cu(x, True) for x in d['disp_pyr']]
    ps2 = cu(d['poses'], True)
    lgs2 = [cu(l, True) for l in d['logits_pyr']]
    tp, sp = ops.image_pyramid(tgt, S), [ops.image_pyramid(s, S) for s in srcs]
    pixel = smooth = exp = 0
    for s in range(S):
        smooth = smooth + flags.smooth_weight / 2 ** s * ops.compute_smooth_loss(xs2[s], inverse=True)
        for v in range(V):
            warped = ops.projective_inverse_warp(sp[v][s], (1.0 / xs2[s]).squeeze(3), ps2[:, v].contiguous(),
                                                 Kp[:, s].contiguous(), 'angleaxis')[0]
            lg = lgs2[s][..., 2 * v:2 * v + 2].contiguous()
            exp = exp + flags.explain_reg_weight * ops.compute_exp_reg_loss(lg)
            m = torch.softmax(lg, -1)[..., 1:2]
            pixel = pixel + ((warped - tp[s]).abs() * m).mean() * flags.data_weight
    (pixel + smooth + exp).backward()
    for got, want in zip(losses.tolist(), (float(pixel), float(smooth), float(exp))):
        assert abs(got - want) <= 1e-5 * abs(want)
    for s in range(S):
        assert rel_err(xs[s].grad, xs2[s].grad) <= 1e-4
        assert rel_err(lgs[s].grad, lgs2[s].grad) <= 1e-4
    assert rel_err(ps.grad, ps2.grad) <= 1e-4


def test_fused_against_oracle_fresh_inputs():
    """Oracle (float64 autograd) on a freshly seeded cfg-shaped crop, both mask modes."""
    B, H, W, S, V = 2, 32, 104, 4, 2
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=33)
    for mode in ('exp', 'const', 'none'):
        flags = ops.LossFlags(num_scales=S)
        of = O.LossFlags(num_scales=S)
        g = torch.Generator().manual_seed(3)
        masks = [torch.rand(B, H >> s, W >> s, 1, generator=g) for s in range(S)]
        xs = [cu(x, True) for x in d['disp_pyr']]
        ps = cu(d['poses'], True)
        lgs = [cu(l, True) for l in d['logits_pyr']] if mode == 'exp' else None
        total, losses = ops.view_synthesis_loss(cu(d['tgt']), [cu(s) for s in d['srcs']], xs, ps, cu(d['K_pyr']),
                                                logits_pyr=lgs, mask_pyr=[cu(m) for m in masks] if mode == 'const' else None,
                                                flags=flags)
        total.backward()
        oxs = [x.double().requires_grad_() for x in d['disp_pyr']]
        ops_ = d['poses'].double().requires_grad_()
        ol = [l.double().requires_grad_() for l in d['logits_pyr']] if mode == 'exp' else None
        ref = O.view_synthesis_loss(d['tgt'].double(), [s.double() for s in d['srcs']], oxs, ops_, d['K_pyr'].double(),
                                    ol, [m.double() for m in masks] if mode == 'const' else None, of)
        sum(ref).backward()
        for got, want in zip(losses.tolist(), ref):
            assert abs(got - float(want)) <= 1e-5 * abs(float(want)) + 1e-9, (mode, got, float(want))
        assert rel_err(ps.grad, ops_.grad) <= 1e-4, mode
        ok = smooth_pixels(d['tgt'], d['srcs'], d['disp_pyr'], d['poses'], d['K_pyr'], flags)
        for s in range(S):
            all_views = torch.stack(ok[s]).all(0)
            assert all_views.float().mean() > 0.97
            assert masked_rel_err(xs[s].grad, oxs[s].grad, all_views.unsqueeze(3)) <= 1e-4, (mode, s)
            if mode == 'exp':
                m = torch.stack([o for o in ok[s] for _ in (0, 1)], dim=3)
                assert masked_rel_err(lgs[s].grad, ol[s].grad, m) <= 1e-4


# ---------------------------------------------------------------------------------------- full size
def _cfg2_inputs(seed=1234, B=32):
    d = synth.make_snippets(B, 128, 416, S=4, V=2, seed=seed)
    return d


def test_full_size_properties_cfg2():
    """BASELINE.json config 2 (B=32, 128x416, 4 scales, 2 views): size-independent properties."""
    d = _cfg2_inputs()
    img, depth = cu(d['srcs'][0]), cu((1.0 / d['disp_pyr'][0]).squeeze(3))
    pose, K = cu(d['poses'][:, 0]), cu(d['K'])
    # identity pose: coords == the reference's fp32 grid up to K K^-1 rounding, warp == src, wmask == 1
    eye = torch.eye(4, device=DEV).repeat(32, 1, 1)
    out, coords, wmask, z, _ = ops.projective_inverse_warp(img, depth, eye, K, 'matrix')
    grid = O.meshgrid(1, 128, 416, is_homogeneous=False).permute(0, 2, 3, 1).to(DEV)
    assert (coords - grid).abs().max() < 5e-4
    assert (out - img).abs().max() <= 1e-4 and (wmask[:, :-1, :-1] - 1).abs().max() < 1e-3
    assert torch.equal(z.squeeze(3), depth)
    # linearity in the source image and wmask in [0, 1]
    img2 = cu(d['srcs'][1])
    o1 = ops.projective_inverse_warp(img, depth, pose, K)[0]
    o2 = ops.projective_inverse_warp(img2, depth, pose, K)[0]
    o12, _, wm, _, _ = ops.projective_inverse_warp(0.25 * img + 2.0 * img2, depth, pose, K)
    assert (o12 - (0.25 * o1 + 2.0 * o2)).abs().max() <= 1e-5
    assert float(wm.min()) >= 0.0 and float(wm.max()) <= 1.0 + 1e-6
    # pure x translation with constant depth: uniform shift of fx * tx / d pixels
    tx, dconst = 0.05, 2.0
    T = torch.eye(4, device=DEV).repeat(32, 1, 1)
    T[:, 0, 3] = tx
    c2 = ops.projective_inverse_warp(img, torch.full_like(depth, dconst), T, K, 'matrix')[1]
    shift = float(K[0, 0, 0]) * tx / dconst
    assert (c2[..., 0] - (grid[..., 0] + shift)).abs().max() < 1e-3 and (c2[..., 1] - grid[..., 1]).abs().max() < 1e-3


def test_full_size_fused_batch_shards_and_determinism():
    """The path shards over the batch: the fused loss of the whole batch equals the sum over two
    half-batch ranks of (loss * 1/2), per-sample gradients agree, and two runs are bit-identical."""
    d = _cfg2_inputs(seed=99)
    flags = ops.LossFlags()
    dev_in = dict(tgt=cu(d['tgt']), srcs=[cu(s) for s in d['srcs']], xs=[cu(x) for x in d['disp_pyr']],
                  poses=cu(d['poses']), Kp=cu(d['K_pyr']), lgs=[cu(l) for l in d['logits_pyr']])

    def run(lo, hi):
        plan = ops.ViewSynthesisPlan(hi - lo, 128, 416, 2, flags, 1, torch.device(DEV))
        sl = lambda t: t[lo:hi].contiguous()
        plan.run(sl(dev_in['tgt']), [sl(s) for s in dev_in['srcs']], [sl(x) for x in dev_in['xs']],
                 sl(dev_in['poses']), sl(dev_in['Kp']), [sl(l) for l in dev_in['lgs']])
        torch.cuda.synchronize()
        return plan

    whole, again = run(0, 32), run(0, 32)
    assert torch.equal(whole.losses, again.losses) and torch.equal(whole.g_poses, again.g_poses)
    assert all(torch.equal(a, b) for a, b in zip(whole.g_x, again.g_x))
    h0, h1 = run(0, 16), run(16, 32)
    assert rel_err(0.5 * (h0.losses + h1.losses), whole.losses) <= 1e-6
    # means are over the local batch, so a rank's gradients are 2x the global ones
    assert rel_err(torch.cat([h0.g_poses, h1.g_poses]) * 0.5, whole.g_poses) <= 1e-5
    assert rel_err(torch.cat([h0.g_x[0], h1.g_x[0]]) * 0.5, whole.g_x[0]) <= 1e-5
    assert rel_err(torch.cat([h0.g_logits[1], h1.g_logits[1]]) * 0.5, whole.g_logits[1]) <= 1e-5


def test_no_cpu_fallback():
    d = synth.make_snippets(1, 16, 32, S=1, V=1, seed=1)
    with pytest.raises(TypeError):
        ops.bilinear_sampler(d['srcs'][0], torch.zeros(1, 16, 32, 2))
    with pytest.raises(TypeError):
        ops.compute_smooth_loss(d['disp_pyr'][0].double().to(DEV))


@pytest.mark.parametrize('B,H,W,S,V,fmt,mode', [
    (2, 24, 36, 3, 1, 'eular', 'exp'),        # one view; 36 -> 18 -> 9: W % 4 != 0 at the coarse scales (unstaged path)
    (1, 32, 72, 4, 3, 'angleaxis', 'exp'),    # three views (butterfly reduction with N = 39)
    (2, 16, 48, 2, 4, 'matrix', 'none'),      # four views, matrix poses, no mask
    (1, 64, 96, 5, 2, 'eular', 'const'),      # five scales (F = 16), constant validity mask
    (3, 40, 44, 3, 2, 'eular', 'exp'),        # W = 44: ragged tiles, odd coarse widths (11)
    (2, 48, 64, 1, 2, 'eular', 'exp'),        # a single scale: no pyramid, the prep launch only re-lays the sources
    (1, 192, 256, 4, 1, 'angleaxis', 'none'), # BASELINE cfg4 frame size (DeMoN pairs), one view per direction
    (1, 480, 640, 4, 2, 'eular', 'exp'),      # BASELINE cfg5 frame size: 15 column strips x 30 row bands at scale 0
])
def test_fused_shape_sweep_against_oracle(B, H, W, S, V, fmt, mode):
    """Views 1..4, scales 2..5, widths that are not multiples of 32 / 4, every pose format and mask mode."""
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=100 + V + S, motion=1.5)
    g = torch.Generator().manual_seed(7)
    poses = d['poses']
    if fmt == 'matrix':
        poses = torch.stack([O.pose_vec2mat(d['poses'][:, v], 'eular') for v in range(V)], 1)
    masks = [torch.rand(B, H >> s, W >> s, 1, generator=g) for s in range(S)]
    for exact in (False, True):
        flags = ops.LossFlags(num_scales=S, pose_format=fmt, exact_coords=exact, smooth_weight=0.3)
        of = O.LossFlags(num_scales=S, pose_format=fmt, smooth_weight=0.3)
        xs = [cu(x, True) for x in d['disp_pyr']]
        ps = cu(poses, True)
        lgs = [cu(l, True) for l in d['logits_pyr']] if mode == 'exp' else None
        total, losses = ops.view_synthesis_loss(cu(d['tgt']), [cu(s) for s in d['srcs']], xs, ps, cu(d['K_pyr']),
                                                logits_pyr=lgs, mask_pyr=[cu(m) for m in masks] if mode == 'const' else None,
                                                flags=flags)
        total.backward()
        oxs = [x.double().requires_grad_() for x in d['disp_pyr']]
        op_ = poses.double().requires_grad_()
        ol = [l.double().requires_grad_() for l in d['logits_pyr']] if mode == 'exp' else None
        ref = O.view_synthesis_loss(d['tgt'].double(), [s.double() for s in d['srcs']], oxs, op_, d['K_pyr'].double(),
                                    ol, [m.double() for m in masks] if mode == 'const' else None, of)
        sum(ref).backward()
        for got, want in zip(losses.tolist(), ref):
            assert abs(got - float(want)) <= 1e-5 * abs(float(want)) + 1e-9, (exact, got, float(want))
        if fmt == 'matrix':
            assert rel_err(ps.grad[:, :, :3], op_.grad[:, :, :3]) <= 1e-4   # row 3 of T never reaches the coordinates
        else:
            assert rel_err(ps.grad, op_.grad) <= 1e-4
        ok = smooth_pixels(d['tgt'], d['srcs'], d['disp_pyr'], poses, d['K_pyr'], flags)
        for s in range(S):
            all_views = torch.stack(ok[s]).all(0)
            assert masked_rel_err(xs[s].grad, oxs[s].grad, all_views.unsqueeze(3)) <= 1e-4, (exact, s)
            if mode == 'exp':
                m = torch.stack([o for o in ok[s] for _ in (0, 1)], dim=3)
                e = masked_rel_err(lgs[s].grad, ol[s].grad, m)
                if e > 1e-4:   # 640-pixel rows: the reference's own float32 coordinates are this far from float64
                    e -= masked_rel_err(_oracle32_logit_grads(d, poses, of, mode)[s], ol[s].grad, m)
                assert e <= 1e-4, (exact, s, e)


def _oracle32_logit_grads(d, poses, of, mode):
    lg = [l.clone().requires_grad_() for l in d['logits_pyr']]
    r = O.view_synthesis_loss(d['tgt'], d['srcs'], [x.clone() for x in d['disp_pyr']], poses.clone(), d['K_pyr'], lg, None, of)
    sum(r).backward()
    return [l.grad for l in lg]


def test_pyramid_shapes_bit_exact():
    """resize_area levels for 1..4 channels and 2..6 scales against the oracle's summation order."""
    g = torch.Generator().manual_seed(11)
    for C, S, H, W in ((3, 4, 16, 40), (1, 3, 8, 12), (4, 2, 6, 10), (2, 5, 32, 48), (3, 6, 64, 96), (3, 2, 10, 14)):
        img = torch.rand(2, H, W, C, generator=g)
        lv = ops.image_pyramid(cu(img), S)
        for s in range(1, S):
            assert torch.equal(lv[s].cpu(), O.resize_area(img, H >> s, W >> s)), (C, S, s)


def test_fused_source_image_gradient_against_oracle():
    """d/d(source images) of the fused step (atomic 16-byte scatter into gradient levels + fold-back through the
    resize_area pyramid) against float64 autograd of the oracle.  A target pixel whose photometric error is within
    rounding of 0 has an undetermined sign(e) ("Discontinuities") and moves up to four source pixels, so
    the comparison is over the entries that agree to 1e-4 of the largest gradient, which must be >= 99.5 %."""
    B, H, W, S, V = 2, 32, 104, 4, 2
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=41, motion=1.5)
    flags = ops.LossFlags(num_scales=S)
    srcs = [cu(s, True) for s in d['srcs']]
    xs = [cu(x, True) for x in d['disp_pyr']]
    ps = cu(d['poses'], True)
    lgs = [cu(l, True) for l in d['logits_pyr']]
    total, losses = ops.view_synthesis_loss(cu(d['tgt']), srcs, xs, ps, cu(d['K_pyr']), logits_pyr=lgs, flags=flags)
    total.backward()
    osrc = [s.double().requires_grad_() for s in d['srcs']]
    oxs = [x.double().requires_grad_() for x in d['disp_pyr']]
    op_ = d['poses'].double().requires_grad_()
    ol = [l.double().requires_grad_() for l in d['logits_pyr']]
    ref = O.view_synthesis_loss(d['tgt'].double(), osrc, oxs, op_, d['K_pyr'].double(), ol, None, O.LossFlags(num_scales=S))
    sum(ref).backward()
    for got, want in zip(losses.tolist(), ref):
        assert abs(got - float(want)) <= 1e-5 * abs(float(want)) + 1e-9
    assert rel_err(ps.grad, op_.grad) <= 1e-4          # the other gradients are those of the plain step
    for v in range(V):
        g, w = srcs[v].grad.cpu().double(), osrc[v].grad
        assert float(w.abs().max()) > 0
        ok = (g - w).abs() <= 1e-4 * float(w.abs().max())
        assert float(ok.double().mean()) >= 0.995, (v, float(ok.double().mean()))
        assert abs(float(g.sum()) - float(w.sum())) <= 2e-3 * float(w.abs().sum())
    # the request is refused, not ignored, where it is not implemented
    with pytest.raises(Exception):
        ops.view_synthesis_loss(cu(d['tgt']), srcs, xs, ps, cu(d['K_pyr']), logits_pyr=lgs,
                                flags=ops.LossFlags(num_scales=S, exact_coords=True))


@pytest.mark.parametrize('smooth_inv,depth_inv', [(False, True), (True, False), (True, True)])
def test_fused_disparity_head_on_load(smooth_inv, depth_inv):
    """x_is_logit: the kernel consumes the disparity head's pre-activation output and applies
    DISP_SCALING * sigmoid + MIN_DISP (nets_optflow_depth.py:8-9,143-144) on load and its derivative on store;
    against the oracle fed with the activated disparity, gradient taken through the sigmoid by float64 autograd."""
    B, H, W, S, V = 2, 32, 104, 3, 2
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=55)
    g = torch.Generator().manual_seed(5)
    raw = [0.8 * torch.randn(B, H >> s, W >> s, 1, generator=g) for s in range(S)]
    scale, mn = 4.0, 0.01
    kw = dict(num_scales=S, smooth_on_inverse=smooth_inv, depth_is_inverse=depth_inv)
    for exact in (False, True):
        flags = ops.LossFlags(x_is_logit=True, disp_scaling=scale, min_disp=mn, exact_coords=exact, **kw)
        xs = [cu(x, True) for x in raw]
        ps = cu(d['poses'], True)
        lgs = [cu(l, True) for l in d['logits_pyr']]
        total, losses = ops.view_synthesis_loss(cu(d['tgt']), [cu(s) for s in d['srcs']], xs, ps, cu(d['K_pyr']),
                                                logits_pyr=lgs, flags=flags)
        total.backward()
        oraw = [x.double().requires_grad_() for x in raw]
        odisp = [scale * torch.sigmoid(x) + mn for x in oraw]
        op_ = d['poses'].double().requires_grad_()
        ol = [l.double().requires_grad_() for l in d['logits_pyr']]
        ref = O.view_synthesis_loss(d['tgt'].double(), [s.double() for s in d['srcs']], odisp, op_, d['K_pyr'].double(),
                                    ol, None, O.LossFlags(**kw))
        sum(ref).backward()
        for got, want in zip(losses.tolist(), ref):
            assert abs(got - float(want)) <= 1e-5 * abs(float(want)) + 1e-9, (exact, got, float(want))
        assert rel_err(ps.grad, op_.grad) <= 1e-4
        disp32 = [(scale * torch.sigmoid(x) + mn) for x in raw]
        ok = smooth_pixels(d['tgt'], d['srcs'], disp32, d['poses'], d['K_pyr'], ops.LossFlags(**kw))
        for s in range(S):
            all_views = torch.stack(ok[s]).all(0)
            assert masked_rel_err(xs[s].grad, oraw[s].grad, all_views.unsqueeze(3)) <= 1e-4, (exact, s)


# ---------------------------------------------------------------------------------------- full BASELINE sizes vs the oracle
def _oracle_one(d, b, of, dtype, mode, need_grad):
    sl = slice(b, b + 1)
    c = (lambda t: t.double()) if dtype == torch.float64 else (lambda t: t.float())
    xs = [c(x[sl]).clone().requires_grad_(need_grad) for x in d['disp_pyr']]
    ps = c(d['poses'][sl]).clone().requires_grad_(need_grad)
    lg = [c(l[sl]).clone().requires_grad_(need_grad) for l in d['logits_pyr']] if mode == 'exp' else None
    with torch.set_grad_enabled(need_grad):
        r = O.view_synthesis_loss(c(d['tgt'][sl]), [c(s[sl]) for s in d['srcs']], xs, ps, c(d['K_pyr'][sl]), lg, None, of)
    if need_grad:
        sum(r).backward()
        return r, ([x.grad for x in xs], ps.grad, [l.grad for l in lg] if lg else None)
    return r, None


def _oracle_per_sample(d, flags_kw, samples, S, V, mode='exp'):
    """One sample at a time (the batch mean of per-sample means IS the batch mean: every sample has the same pixel
    count): -> (float64 losses[3] averaged over ALL samples, {b: (float64 grads, float32 grads)}), grads = (g_x list,
    g_poses, g_logits list).  The float32 oracle is the reference's own arithmetic (bit-identical to the goldens
    produced by executing the reference's source, tests/test_oracle_golden.py); float64 is the mathematical truth."""
    B = d['tgt'].shape[0]
    of = O.LossFlags(**flags_kw)
    tot = torch.zeros(3, dtype=torch.float64)
    grads = {}
    for b in range(B):
        r, g64 = _oracle_one(d, b, of, torch.float64, mode, b in samples)
        tot += torch.stack([t.detach() for t in r])
        if b in samples:
            grads[b] = (g64, _oracle_one(d, b, of, torch.float32, mode, True)[1])
    return tot / B, grads


def _full_size_vs_oracle(d, B, H, W, S, V, samples, mode='exp', flags_kw=None, exact_too=False):
    """Losses of the whole batch against the float64 oracle (1e-5).  Gradients of the given samples (1e-4, max-norm):
    the bar is parity with the REFERENCE, whose arithmetic is float32 -- at 480x640 its own pose gradient is 3e-4
    away from float64 (coordinates near 640 px carry 4e-5 px of float32 rounding, and a sample within that of an
    integer coordinate lands in another bilinear cell; profiles/diag_fullsize.py prints the three-way comparison).
    A gradient therefore passes if it is within 1e-4 of the float32 oracle, OR within 1e-4 + n of the float64 one,
    n = the distance of the float32 oracle from float64 on the same entries (the fast arithmetic contracts FMAs and
    rounds a coordinate differently from the reference: it has to be as close to the truth as the reference's own
    float32 arithmetic is, not closer -- at 480x640 n reaches 1e-4 for d/dlogits where a sample falls between the
    last image column and the zero padding, the steepest part of the sampler); the exact mode must follow float32
    to 1e-5 in the pose gradient."""
    flags_kw = dict(flags_kw or {}, num_scales=S)
    want, grads = _oracle_per_sample(d, flags_kw, samples, S, V, mode)
    pose_noise = max(rel_err(g32[1], g64[1]) for g64, g32 in grads.values())
    for exact in ((False, True) if exact_too else (False,)):
        flags = ops.LossFlags(exact_coords=exact, **flags_kw)
        xs = [cu(x, True) for x in d['disp_pyr']]
        ps = cu(d['poses'], True)
        lgs = [cu(l, True) for l in d['logits_pyr']] if mode == 'exp' else None
        total, losses = ops.view_synthesis_loss(cu(d['tgt']), [cu(s) for s in d['srcs']], xs, ps, cu(d['K_pyr']),
                                                logits_pyr=lgs, flags=flags)
        total.backward()
        for i, key in enumerate(('pixel', 'smooth', 'exp')):
            assert abs(float(losses[i]) - float(want[i])) <= 1e-5 * abs(float(want[i])) + 1e-9, (key, float(losses[i]), float(want[i]))
        assert abs(float(total) - float(want.sum())) <= 1e-5 * float(want.sum())
        for b, ((gx64, gp64, gl64), (gx32, gp32, gl32)) in grads.items():
            sl = slice(b, b + 1)
            one = dict(tgt=d['tgt'][sl], srcs=[s[sl] for s in d['srcs']], disp=[x[sl] for x in d['disp_pyr']],
                       poses=d['poses'][sl], K=d['K_pyr'][sl])
            # the kernel's gradients are those of the B-sample mean: 1/B of the single-sample oracle's
            e32, e64 = rel_err(ps.grad[sl] * B, gp32), rel_err(ps.grad[sl] * B, gp64)
            # d/dpose sums a discontinuous function of the coordinates over every pixel: which samples fall into the
            # neighbouring bilinear cell under float32 rounding differs from sample to sample, so the yardstick is
            # the reference's largest own float32-vs-float64 gap over the tested samples (3e-4 at 480x640, 1e-6 at
            # 128x416), twice over
            assert e32 <= 1e-4 or e64 <= 1e-4 + 2.0 * pose_noise, ('g_poses', exact, b, e32, e64, pose_noise)
            if exact:
                assert e32 <= 1e-5, ('g_poses exact vs float32 reference arithmetic', b, e32)
            ok = smooth_pixels(one['tgt'], one['srcs'], one['disp'], one['poses'], one['K'], flags)
            for s in range(S):
                all_views = torch.stack(ok[s]).all(0)
                assert all_views.float().mean() > 0.97, (b, s, float(all_views.float().mean()))
                m = all_views.unsqueeze(3)
                e32 = masked_rel_err(xs[s].grad[sl] * B, gx32[s], m)
                e64 = masked_rel_err(xs[s].grad[sl] * B, gx64[s], m)
                assert e32 <= 1e-4 or e64 <= 1e-4 + masked_rel_err(gx32[s], gx64[s], m), ('g_x', exact, b, s, e32, e64)
                if mode == 'exp':
                    m = torch.stack([o for o in ok[s] for _ in (0, 1)], dim=3)
                    e32 = masked_rel_err(lgs[s].grad[sl] * B, gl32[s], m)
                    e64 = masked_rel_err(lgs[s].grad[sl] * B, gl64[s], m)
                    assert e32 <= 1e-4 or e64 <= 1e-4 + masked_rel_err(gl32[s], gl64[s], m), ('g_logits', exact, b, s, e32, e64)


def test_full_size_cfg2_against_oracle():
    """BASELINE.json configs[1] at FULL size (B=32, 128x416, 4 scales, 2 views, explainability mask): the three
    losses over the whole batch to 1e-5, pose and per-pixel gradients of samples 0, 13 and 31 to 1e-4 against the
    float64 oracle."""
    d = synth.make_snippets(32, 128, 416, S=4, V=2, seed=1234)
    _full_size_vs_oracle(d, 32, 128, 416, 4, 2, samples=(0, 13, 31))


def test_full_size_cfg5_beyond_2_pow_24_against_oracle():
    """BASELINE.json configs[4] at FULL size (B=64, 480x640: 19.66 M pixels at level 0, beyond the 2^24 elements where
    the reference's own float32 gather indices break, utils.py:273-294).  Samples 0, 31 and 63 -- the last
    one lives wholly above 2^24 -- pin the kernel's 32-bit wrapped offset arithmetic against the exact-index oracle."""
    d = synth.make_snippets(64, 480, 640, S=4, V=2, seed=1239)
    assert 63 * 480 * 640 > 2 ** 24
    _full_size_vs_oracle(d, 64, 480, 640, 4, 2, samples=(0, 31, 63), exact_too=True)


@pytest.mark.parametrize('direction', ['left_to_right', 'right_to_left'])
def test_full_size_cfg4_both_directions_against_oracle(direction):
    """BASELINE.json configs[3] at FULL size (DeMoN pairs, B=64, 192x256, one source view per direction, angle-axis
    poses, smoothness on 1/depth and unnormalised data weight as train_depth_then_cam_lr.py:217,310 use them)."""
    d = synth.make_snippets(64, 192, 256, S=4, V=1, seed=1238)
    if direction == 'right_to_left':       # the other image of the pair is the target, the motion is reversed
        d['tgt'], d['srcs'] = d['srcs'][0], [d['tgt']]
        d['poses'] = -d['poses']
    kw = dict(pose_format='angleaxis', smooth_on_inverse=True, depth_is_inverse=True, pixel_scale_norm=False,
              smooth_weight=0.3, data_weight=2.0, explain_reg_weight=0.4)
    _full_size_vs_oracle(d, 64, 192, 256, 4, 1, samples=(0, 40, 63), flags_kw=kw)


# ---------------------------------------------------------------------------------------- view-paired kernel
@pytest.mark.parametrize('V,mode,kw', [
    (2, 'exp', {}),
    (2, 'const', dict(smooth_on_inverse=True, depth_is_inverse=False)),
    (2, 'none', dict(smooth_on_inverse=True, depth_is_inverse=True, pixel_scale_norm=False)),
    (4, 'exp', dict(pose_format='angleaxis')),
    (2, 'exp', dict(x_is_logit=True, disp_scaling=4.0, min_disp=0.01)),
])
def test_paired_kernel_matches_scalar_fast_path(V, mode, kw):
    """An even number of views runs loss_fused_pair_kernel (packed fp32x2, folded projection); exact_coords = 2 asks
    for the scalar fast kernel on the same inputs.  The two differ by rounding only: losses to 2e-6, pose gradients
    to 2e-5, per-pixel gradients to 2e-5 of the largest one away from the kinks.  Ragged sizes: 44 -> 22 -> 11
    columns (partial strips), 40 rows (a 32-row band plus a partial one: the last image rows of the smoothness term
    fall inside a tile)."""
    B, H, W, S = 3, 40, 44, 3
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=300 + V, motion=1.5)
    g = torch.Generator().manual_seed(3)
    masks = [torch.rand(B, H >> s, W >> s, 1, generator=g) for s in range(S)]
    raw = [0.8 * torch.randn(B, H >> s, W >> s, 1, generator=g) for s in range(S)] if kw.get('x_is_logit') else d['disp_pyr']
    res = {}
    for which in (0, 2):
        flags = ops.LossFlags(num_scales=S, exact_coords=which, **kw)
        xs = [cu(x, True) for x in raw]
        ps = cu(d['poses'], True)
        lgs = [cu(l, True) for l in d['logits_pyr']] if mode == 'exp' else None
        total, losses = ops.view_synthesis_loss(cu(d['tgt']), [cu(s) for s in d['srcs']], xs, ps, cu(d['K_pyr']),
                                                logits_pyr=lgs, mask_pyr=[cu(m) for m in masks] if mode == 'const' else None,
                                                flags=flags)
        total.backward()
        res[which] = (losses.clone(), ps.grad.clone(), [x.grad.clone() for x in xs],
                      [l.grad.clone() for l in lgs] if lgs else [])
    a, b = res[0], res[2]
    assert rel_err(a[0], b[0]) <= 2e-6, (a[0], b[0])
    assert rel_err(a[1], b[1]) <= 2e-5
    disp = [4.0 * torch.sigmoid(x) + 0.01 for x in raw] if kw.get('x_is_logit') else raw
    fl = ops.LossFlags(num_scales=S, **{k: v for k, v in kw.items() if k not in ('x_is_logit', 'disp_scaling', 'min_disp')})
    ok = smooth_pixels(d['tgt'], d['srcs'], disp, d['poses'], d['K_pyr'], fl)
    for s in range(S):
        m = torch.stack(ok[s]).all(0).unsqueeze(3)
        # the smoothness term has kinks of its own (a second difference within rounding of 0): allow a stencil's worth
        diff = ((a[2][s] - b[2][s]).abs().cpu() * m) / b[2][s].abs().max().cpu()
        assert int((diff > 2e-5).sum()) <= 8, (s, int((diff > 2e-5).sum()), float(diff.max()))
        if mode == 'exp':
            ml = torch.stack([o for o in ok[s] for _ in (0, 1)], dim=3)
            assert masked_rel_err(a[3][s], b[3][s], ml) <= 2e-5, s


@pytest.mark.parametrize('B,H,W,S', [(2, 64, 96, 3), (4, 128, 416, 4)])
def test_exact_mode_gradients_everywhere_against_float32_oracle(B, H, W, S):
    """The strict reading of the gradient bar: NO pixel excluded.  With a matrix pose the exact mode's sample
    positions, weights and blend are bit-identical to the float32 oracle's (the reference's own arithmetic), so every
    sign(.) of the loss is the same and the per-pixel gradients must agree everywhere -- kink pixels included -- to
    float32 rounding: 1e-4 of the largest gradient, BASELINE.json's bar, with no mask.  (Against the float64 oracle a
    pixel whose float32 coordinate falls into another bilinear cell differs legitimately; that comparison is the
    masked one of the tests above.)"""
    V = 2
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=321, motion=2.0)
    poses = torch.stack([O.pose_vec2mat(d['poses'][:, v], 'eular') for v in range(V)], 1).contiguous()
    for mode in ('exp', 'none'):
        flags = ops.LossFlags(num_scales=S, pose_format='matrix', exact_coords=True, smooth_weight=0.3)
        of = O.LossFlags(num_scales=S, pose_format='matrix', smooth_weight=0.3)
        xs = [cu(x, True) for x in d['disp_pyr']]
        lgs = [cu(l, True) for l in d['logits_pyr']] if mode == 'exp' else None
        total, losses = ops.view_synthesis_loss(cu(d['tgt']), [cu(s) for s in d['srcs']], xs, cu(poses, True), cu(d['K_pyr']),
                                                logits_pyr=lgs, flags=flags)
        total.backward()
        oxs = [x.clone().requires_grad_() for x in d['disp_pyr']]
        ol = [l.clone().requires_grad_() for l in d['logits_pyr']] if mode == 'exp' else None
        ref = O.view_synthesis_loss(d['tgt'], d['srcs'], oxs, poses.clone(), d['K_pyr'], ol, None, of)
        sum(ref).backward()
        for s in range(S):
            e = rel_err(xs[s].grad, oxs[s].grad)
            assert e <= 1e-4, (mode, 'g_x', s, e)
            if mode == 'exp':
                e = rel_err(lgs[s].grad, ol[s].grad)
                assert e <= 1e-4, (mode, 'g_logits', s, e)
