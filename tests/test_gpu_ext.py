"""GPU parity of the EXTENSION ops (SSIM dissimilarity, edge-aware smoothness) against the float64 oracle.

These terms are named by BASELINE.json's north_star but are absent from the reference: there
is no reference implementation to pin them to, so parity here is "against the build's own oracle" only.
Tolerances as for the path proper: values 1e-5 abs / losses 1e-5 rel, gradients 1e-4 rel (max-norm).
"""
import pytest
import torch

from oracle import vsl_oracle as O
from tests.conftest import rel_err
from tf_depth_estimation_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def cu(t, grad=False):
    return t.to(DEV).float().contiguous().requires_grad_(grad)


def pair(B, H, W, C, seed):
    g = torch.Generator().manual_seed(seed)
    if C == 3:
        d = synth.make_snippets(B, H, W, S=1, V=1, seed=seed)
        return d['tgt'].float(), d['srcs'][0].float()
    x = torch.rand(B, H, W, C, generator=g)
    return x, (x + 0.1 * torch.randn(B, H, W, C, generator=g)).clamp(0, 1)


@pytest.mark.parametrize('B,H,W,C', [(2, 32, 64, 3), (1, 3, 3, 1), (3, 11, 45, 3), (2, 16, 130, 1), (1, 9, 40, 4),
                                     (2, 17, 129, 2)])
def test_ssim_map_loss_and_gradients(B, H, W, C):
    x, y = pair(B, H, W, C, 100 + H)
    xd, yd = x.double().requires_grad_(), y.double().requires_grad_()
    ref = O.ssim_dissimilarity(xd, yd)
    g = torch.Generator().manual_seed(5)
    gm = torch.rand(ref.shape, generator=g)
    (ref * gm.double()).sum().backward()

    xc, yc = cu(x, True), cu(y, True)
    got = ops.ssim_dissimilarity(xc, yc)
    assert got.shape == ref.shape
    assert float((got.detach().cpu().double() - ref.detach()).abs().max()) <= 1e-5
    (got * gm.to(DEV)).sum().backward()
    assert rel_err(xc.grad, xd.grad) <= 1e-4
    assert rel_err(yc.grad, yd.grad) <= 1e-4

    # mean path: reduced in the kernel, no map written
    xd2, yd2 = x.double().requires_grad_(), y.double().requires_grad_()
    ref_l = O.ssim_dissimilarity(xd2, yd2).mean()
    (3.0 * ref_l).backward()
    xc2, yc2 = cu(x, True), cu(y, True)
    got_l = ops.ssim_loss(xc2, yc2)
    assert abs(float(got_l) - float(ref_l)) <= 1e-5 * abs(float(ref_l)) + 1e-9
    (3.0 * got_l).backward()
    assert rel_err(xc2.grad, xd2.grad) <= 1e-4
    assert rel_err(yc2.grad, yd2.grad) <= 1e-4


def test_ssim_known_answers_and_determinism():
    x = torch.rand(2, 20, 30, 3, generator=torch.Generator().manual_seed(1))
    xc = cu(x)
    assert float(ops.ssim_dissimilarity(xc, xc).abs().max()) <= 5e-7          # identical images: SSIM = 1
    one, zero = torch.ones(1, 5, 5, 1, device=DEV), torch.zeros(1, 5, 5, 1, device=DEV)
    # constant 1 against constant 0: SSIM = C1 / (1 + C1)
    want = 0.5 * (1.0 - 1e-4 / (1.0 + 1e-4))
    assert float((ops.ssim_dissimilarity(one, zero) - want).abs().max()) <= 1e-6
    y = cu(torch.rand(2, 20, 30, 3, generator=torch.Generator().manual_seed(2)), True)
    a = ops.ssim_loss(xc, y)
    a.backward()
    g1 = y.grad.clone()
    y.grad = None
    b = ops.ssim_loss(xc, y)
    b.backward()
    assert float(a) == float(b) and torch.equal(g1, y.grad)                   # gather form: bit-reproducible
    with pytest.raises(TypeError):
        ops.ssim_loss(x, x)                                                   # CPU tensors: no fallback


def test_photometric_alpha_blend_on_warped_image():
    """The SfMLearner-style use: alpha * SSIM-dissimilarity + (1 - alpha) * L1 between the warped source and
    the target, differentiated through the warp into depth and pose."""
    d = synth.make_snippets(2, 32, 64, S=1, V=1, seed=11)
    alpha = 0.85
    depth = (1.0 / d['disp_pyr'][0][..., 0]).float()
    K = d['K_pyr'][:, 0].float()
    pose = d['poses'][:, 0].float()

    dd, pd = depth.double().requires_grad_(), pose.double().requires_grad_()
    w = O.projective_inverse_warp(d['srcs'][0].double(), dd, pd, K.double())[0]
    tgt = d['tgt'].double()
    ref = alpha * O.ssim_dissimilarity(w, tgt).mean() + (1 - alpha) * (w - tgt).abs().mean()
    ref.backward()

    dc, pc = cu(depth, True), cu(pose, True)
    wc = ops.projective_inverse_warp(cu(d['srcs'][0]), dc, pc, cu(K))[0]
    got = alpha * ops.ssim_loss(wc, cu(d['tgt'])) + (1 - alpha) * (wc - cu(d['tgt'])).abs().mean()
    got.backward()
    assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref))
    assert rel_err(pc.grad, pd.grad) <= 1e-3
    # per-pixel depth gradients: compare away from the bilinear kinks (integer source coordinates)
    diff = (dc.grad.cpu().double() - dd.grad).abs()
    assert float((diff > 1e-4 * dd.grad.abs().max()).double().mean()) < 0.03


@pytest.mark.parametrize('B,H,W,C', [(2, 32, 64, 3), (1, 2, 2, 1), (3, 7, 33, 3)])
def test_edge_aware_smoothness(B, H, W, C):
    g = torch.Generator().manual_seed(H)
    disp = torch.rand(B, H, W, 1, generator=g) * 4
    img = torch.rand(B, H, W, C, generator=g)
    dd, im = disp.double().requires_grad_(), img.double().requires_grad_()
    ref = O.edge_aware_smooth_loss(dd, im)
    ref.backward()
    dc, ic = cu(disp, True), cu(img, True)
    got = ops.edge_aware_smooth_loss(dc, ic)
    assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref))
    got.backward()
    assert rel_err(dc.grad, dd.grad) <= 1e-4
    assert rel_err(ic.grad, im.grad) <= 1e-4
    # image as data: no gradient buffer is produced
    dc2 = cu(disp, True)
    ops.edge_aware_smooth_loss(dc2, cu(img)).backward()
    assert torch.equal(dc2.grad, dc.grad)
