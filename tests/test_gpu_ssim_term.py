"""The SSIM share of the photometric term inside the fused step (VslLossDesc.ssim_weight; csrc/vsl_loss_ssim.cu).
EXTENSION: SSIM is absent from the reference, so the checker is the float64 oracle's own definition
(oracle/vsl_oracle.py view_synthesis_loss with flags.ssim_weight) -- parity unpinned.  Bars as for the rest of the
step: losses 1e-5 relative, pose gradients 1e-4, per-pixel gradients 1e-4 away from the kinks of the L1 part and of
the bilinear footprint (tests/parity_util.py)."""
import pytest
import torch

from oracle import vsl_oracle as O
from tests.conftest import rel_err
from tests.parity_util import masked_rel_err, smooth_pixels
from tf_depth_estimation_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def cu(t, grad=False):
    return t.to(DEV).float().contiguous().requires_grad_(grad)


def _run(d, poses, masks, S, V, fmt, mode, alpha, **kw):
    flags = ops.LossFlags(num_scales=S, pose_format=fmt, smooth_weight=0.3, ssim_weight=alpha, **kw)
    of = O.LossFlags(num_scales=S, pose_format=fmt, smooth_weight=0.3, ssim_weight=alpha, **kw)
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
    ref = O.view_synthesis_loss(d['tgt'].double(), [s.double() for s in d['srcs']], oxs, op_, d['K_pyr'].double(), ol,
                                [m.double() for m in masks] if mode == 'const' else None, of)
    sum(ref).backward()
    return flags, losses, ref, xs, oxs, ps, op_, lgs, ol


def _oracle32(d, poses, masks, S, fmt, mode, alpha):
    """The same loss in the oracle's float32 (the arithmetic a float32 reference would run) -> (g_x, g_logits)."""
    of = O.LossFlags(num_scales=S, pose_format=fmt, smooth_weight=0.3, ssim_weight=alpha)
    xs = [x.clone().requires_grad_() for x in d['disp_pyr']]
    lg = [l.clone().requires_grad_() for l in d['logits_pyr']] if mode == 'exp' else None
    r = O.view_synthesis_loss(d['tgt'], d['srcs'], xs, poses.clone(), d['K_pyr'], lg, masks if mode == 'const' else None, of)
    sum(r).backward()
    return [x.grad for x in xs], ([l.grad for l in lg] if lg else None)


@pytest.mark.parametrize('B,H,W,S,V,fmt,mode,alpha', [
    (2, 32, 64, 3, 2, 'eular', 'exp', 0.85),
    (1, 24, 40, 2, 1, 'angleaxis', 'none', 1.0),       # SSIM only; ragged tiles (40 = 32 + 8, 24 = 3 x 8)
    (2, 16, 48, 1, 3, 'eular', 'const', 0.3),          # three views, the scalar fused kernel, a constant mask
    (1, 40, 72, 3, 2, 'matrix', 'exp', 0.5),           # coarsest level 10 x 18
    (2, 128, 416, 4, 2, 'eular', 'exp', 0.85),         # the BASELINE frame size
])
def test_ssim_term_against_oracle(B, H, W, S, V, fmt, mode, alpha):
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=500 + H + V, motion=1.5)
    g = torch.Generator().manual_seed(9)
    poses = d['poses']
    if fmt == 'matrix':
        poses = torch.stack([O.pose_vec2mat(d['poses'][:, v], 'eular') for v in range(V)], 1)
    masks = [torch.rand(B, H >> s, W >> s, 1, generator=g) for s in range(S)]
    flags, losses, ref, xs, oxs, ps, op_, lgs, ol = _run(d, poses, masks, S, V, fmt, mode, alpha)
    for got, want in zip(losses.tolist(), ref):
        assert abs(got - float(want)) <= 1e-5 * abs(float(want)) + 1e-9, (got, float(want))
    if fmt == 'matrix':
        assert rel_err(ps.grad[:, :, :3], op_.grad[:, :, :3]) <= 1e-4
    else:
        assert rel_err(ps.grad, op_.grad) <= 1e-4, rel_err(ps.grad, op_.grad)
    ok = smooth_pixels(d['tgt'], d['srcs'], d['disp_pyr'], poses, d['K_pyr'], flags)
    for s in range(S):
        all_views = torch.stack(ok[s]).all(0)
        if alpha < 1.0:
            assert all_views.float().mean() > 0.9
        e = masked_rel_err(xs[s].grad, oxs[s].grad, all_views.unsqueeze(3))
        if e > 1e-4:   # 416-pixel rows: the float32 oracle's own SSIM gradients are this far from float64
            e -= masked_rel_err(_oracle32(d, poses, masks, S, fmt, mode, alpha)[0][s], oxs[s].grad, all_views.unsqueeze(3))
        assert e <= 1e-4, ('g_x', s, e)
        if mode == 'exp':
            m = torch.stack([o for o in ok[s] for _ in (0, 1)], dim=3)
            e = masked_rel_err(lgs[s].grad, ol[s].grad, m)
            if e > 1e-4:
                e -= masked_rel_err(_oracle32(d, poses, masks, S, fmt, mode, alpha)[1][s], ol[s].grad, m)
            assert e <= 1e-4, ('g_logits', s, e)


def test_ssim_weight_zero_is_the_plain_step_and_one_has_no_l1():
    B, H, W, S, V = 2, 32, 64, 2, 2
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=3)
    base = _run(d, d['poses'], None, S, V, 'eular', 'exp', 0.0)
    plain_flags = ops.LossFlags(num_scales=S, smooth_weight=0.3)
    xs = [cu(x, True) for x in d['disp_pyr']]
    total, losses = ops.view_synthesis_loss(cu(d['tgt']), [cu(s) for s in d['srcs']], xs, cu(d['poses'], True),
                                            cu(d['K_pyr']), logits_pyr=[cu(l, True) for l in d['logits_pyr']], flags=plain_flags)
    assert torch.equal(losses, base[1])
    # identical views: SSIM(x, x) = 1 -> the SSIM term vanishes (identity pose, source = target)
    e = dict(d, srcs=[d['tgt'].clone() for _ in range(V)], poses=torch.zeros_like(d['poses']))
    only = _run(e, e['poses'], None, S, V, 'eular', 'none', 1.0)
    assert float(only[1][0]) <= 1e-6


def test_ssim_term_argument_checks():
    d = synth.make_snippets(1, 16, 32, S=2, V=2, seed=1)
    args = lambda: (cu(d['tgt']), [cu(s) for s in d['srcs']], [cu(x) for x in d['disp_pyr']], cu(d['poses']), cu(d['K_pyr']))
    for bad in (dict(ssim_weight=1.5), dict(ssim_weight=0.5, exact_coords=True), dict(ssim_weight=0.5, x_is_logit=True)):
        with pytest.raises(Exception):
            ops.view_synthesis_loss(*args(), flags=ops.LossFlags(num_scales=2, **bad))
