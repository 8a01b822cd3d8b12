"""GPU: the reference-named drop-in modules (compat/utils.py, utils_lr.py, my_losses.py) and the stand-alone
geometry ops, against the oracle."""
import os
import sys

import pytest
import torch

from oracle import vsl_oracle as O
from tests.conftest import rel_err
from tf_depth_estimation_b200 import synth

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
COMPAT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tf_depth_estimation_b200', 'compat')


@pytest.fixture(scope='module')
def mods():
    """Import the way a reference script would: `from utils_lr import *` with compat/ on sys.path."""
    sys.path.insert(0, COMPAT)
    for name in ('utils', 'utils_lr', 'my_losses'):
        sys.modules.pop(name, None)
    import my_losses
    import utils
    import utils_lr
    yield utils, utils_lr, my_losses
    sys.path.remove(COMPAT)
    for name in ('utils', 'utils_lr', 'my_losses'):
        sys.modules.pop(name, None)


def cu(t, grad=False):
    return t.to(DEV).float().contiguous().requires_grad_(grad)


def test_reference_call_sites(mods, golden):
    utils, utils_lr, my_losses = mods
    c = golden['warp_eular']
    out3 = utils.projective_inverse_warp(cu(c.img), cu(c.depth), cu(c.pose), cu(c.K))                  # v1: 3-tuple
    assert len(out3) == 3 and (out3[0].cpu() - c.out).abs().max() <= 1e-5
    out5 = utils_lr.projective_inverse_warp(cu(c.img), cu(c.depth), cu(c.pose), cu(c.K), format='eular')  # v2: 5-tuple
    assert len(out5) == 5 and tuple(out5[3].shape) == tuple(c.z.shape) and tuple(out5[4].shape) == (2, 4, 4)
    c = golden['warp_matrix_far']
    out = utils.projective_inverse_warp(cu(c.img), cu(c.depth), cu(c.pose), cu(c.K))                   # v0 callers: 4x4
    assert torch.equal(out[1].cpu(), c.coords)
    p = golden['pose']
    assert (utils.pose_vec2mat(cu(p.vec)).cpu() - p.mat_eular).abs().max() <= 1e-6
    assert (utils_lr.pose_vec2mat(cu(p.vec), 'angleaxis').cpu() - p.mat_angleaxis).abs().max() <= 1e-6
    t = golden['terms']

    class FLAGS(object):
        batch_size, resizedheight, resizedwidth = 2, 10, 14
    ref = my_losses.get_reference_explain_mask(0, FLAGS)
    assert tuple(ref.shape) == (2, 10, 14, 2) and float(ref[..., 1].min()) == 1.0 and float(ref[..., 0].max()) == 0.0
    assert rel_err(my_losses.compute_exp_reg_loss(cu(t.logits), ref), t.exp_f64) <= 1e-5
    assert rel_err(my_losses.compute_smooth_loss(cu(t.disp)), t.smooth_f64) <= 1e-5
    with pytest.raises(ValueError):
        my_losses.compute_exp_reg_loss(cu(t.logits), 1 - ref)


def test_geometry_building_blocks(mods):
    _, L, _ = mods
    B, H, W = 2, 12, 20
    d = synth.make_snippets(B, H, W, S=1, V=1, seed=5)
    depth = (1.0 / d['disp_pyr'][0]).squeeze(3)
    grid = L.meshgrid(B, H, W)
    assert torch.equal(grid.cpu(), O.meshgrid(B, H, W))
    assert torch.equal(L.meshgrid(B, H, W, is_homogeneous=False).cpu(), O.meshgrid(B, H, W, False))
    dd = cu(depth, True)
    cam = L.pixel2cam(dd, grid, cu(d['K']))
    ocam = O.pixel2cam(depth, O.meshgrid(B, H, W), d['K'])
    assert torch.equal(cam.detach().cpu(), ocam)
    T = O.pose_vec2mat(d['poses'][:, 0], 'eular')
    K4 = torch.zeros(B, 4, 4)
    K4[:, :3, :3] = d['K']
    K4[:, 3, 3] = 1
    proj = O._mm(K4, T)
    pj = cu(proj, True)
    coords, z = L.cam2pixel(cam, pj)
    g = torch.Generator().manual_seed(1)
    R1, R2 = torch.randn(B, H, W, 2, generator=g), torch.randn(B, H, W, 1, generator=g)
    ((coords * cu(R1)).sum() + (z * cu(R2)).sum()).backward()
    od, op = depth.double().requires_grad_(), proj.double().requires_grad_()
    oc, oz = O.cam2pixel(O.pixel2cam(od, O.meshgrid(B, H, W, dtype=torch.float64), d['K'].double()), op)
    ((oc * R1.double()).sum() + (oz * R2.double()).sum()).backward()
    assert torch.equal(coords.detach().cpu(), O.cam2pixel(ocam, proj)[0])
    assert rel_err(dd.grad, od.grad) <= 1e-4 and rel_err(pj.grad[:, :3], op.grad[:, :3]) <= 1e-4
    # rotations
    ax = torch.nn.functional.normalize(torch.randn(5, 3, generator=g), dim=1)
    an = torch.rand(5, 1, 1, generator=g) * 2
    a1, a2 = cu(ax, True), cu(an, True)
    Rm = L.axis_angle_to_rotation_matrix(a1, a2)
    Rw = torch.randn(5, 3, 3, generator=g)
    (Rm * cu(Rw)).sum().backward()
    o1, o2 = ax.double().requires_grad_(), an.double().requires_grad_()
    (O.axis_angle_to_rotation_matrix(o1, o2) * Rw.double()).sum().backward()
    assert (Rm.detach().cpu() - O.axis_angle_to_rotation_matrix(ax, an)).abs().max() <= 1e-6
    assert rel_err(a1.grad, o1.grad) <= 1e-5 and rel_err(a2.grad, o2.grad) <= 1e-5
    e = torch.randn(4, 3, generator=g)
    Re = L.euler2mat(cu(e[:, 2:3]), cu(e[:, 1:2]), cu(e[:, 0:1]))
    assert tuple(Re.shape) == (4, 1, 3, 3)
    assert (Re.cpu() - O.euler2mat(e[:, 2:3], e[:, 1:2], e[:, 0:1])).abs().max() <= 1e-6


def _depth_golden():
    import ast
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'depth_losses_golden.npz'))
    cases = {'single': {}, 'pair': {}}
    for k in z.files:
        if '/' in k:
            c, n = k.split('/', 1)
            cases[c][n] = torch.from_numpy(np.array(z[k]))
    flags = ast.literal_eval(str(z['flags']))
    return cases, type('FLAGS', (object,), flags)


def test_composite_depth_losses_golden(mods):
    """my_losses.py:46 / :101 executed from the reference's own source over the TF1 shim (tests/golden/
    make_golden_depth_losses.py; DeMoN's un-vendored ops restated) against the drop-in on the GPU."""
    _, _, my_losses = mods
    cases, FLAGS = _depth_golden()
    S = FLAGS.num_scales
    c = cases['single']
    preds = [cu(c['pred%d' % s], True) for s in range(S)]
    depth_loss, smooth_loss, sig = my_losses.compute_loss_single_depth(preds, cu(c['label']), int(c['step']), FLAGS)
    assert smooth_loss == 0
    assert rel_err(depth_loss, c['depth_loss']) <= 1e-5 and rel_err(sig, c['sig_loss']) <= 1e-5
    (depth_loss + sig).backward()
    for s in range(S):
        assert rel_err(preds[s].grad, c['g_pred%d_f64' % s]) <= 1e-4

    c = cases['pair']
    n = S - 2
    pdl = [cu(c['pred_depth_left%d' % i], True) for i in range(n)]
    pdr = [cu(c['pred_depth_right%d' % i], True) for i in range(n)]
    p_r, p_l = cu(c['pred_poses_right'], True), cu(c['pred_poses_left'], True)
    out = my_losses.compute_loss_pairwise_depth(cu(c['image_left']), cu(c['image_right']), pdl, p_r, None, pdr, p_l, None,
                                                cu(c['gt_right_cam']), cu(c['intrinsics']), cu(c['label']), FLAGS,
                                                int(c['step']))
    assert len(out) == 11 and out[2] == 0 and out[3] == 0 and out[5] == 0
    assert rel_err(out[0], c['depth_loss']) <= 1e-5
    assert rel_err(out[1], c['cam_loss']) <= 1e-5
    assert rel_err(out[4], c['sig_loss']) <= 1e-5
    for name, lst in zip(('left_image', 'right_image', 'proj_image_left', 'proj_image_right', 'proj_error'), out[6:]):
        assert len(lst) == n
        for i, t in enumerate(lst):
            assert float((t.detach().cpu() - c['%s%d' % (name, i)]).abs().max()) <= 1e-5
    Rr = torch.Generator().manual_seed(7)
    extra = sum((w * torch.randn(w.shape, generator=Rr).to(DEV)).sum() for w in out[9]) * 1e-3
    (out[0] + out[1] + out[4] + extra).backward()
    for i in range(n):
        assert rel_err(pdl[i].grad, c['g_pdl%d_f64' % i]) <= 1e-4
        # through the warp: per-pixel gradients away from the bilinear kinks
        diff = (pdr[i].grad.cpu().double() - c['g_pdr%d_f64' % i]).abs()
        assert float((diff > 1e-4 * c['g_pdr%d_f64' % i].abs().max()).double().mean()) < 0.03
    assert rel_err(p_r.grad, c['g_poses_r_f64']) <= 1e-4
    assert rel_err(p_l.grad, c['g_poses_l_f64']) <= 1e-4
