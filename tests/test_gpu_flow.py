"""GPU parity of the fused flow-and-depth loss (vsl_flow_loss_fwd_bwd; train_optflow_combine.py:138-240, BASELINE
configs[3]) through the C ABI: against the golden produced by executing the reference's functions
(tests/golden/make_golden_flow.py), against the CPU oracle on fresh inputs, and against the stand-alone ops.

The kernel evaluates coordinates, weights and the blend in the reference's rounding sequence, so the signs of every
|.| are the float32 reference's own: gradients are compared with the FLOAT32 golden / oracle everywhere, without
excluding kink pixels (1e-4 relative, BASELINE.json); loss terms with the float64 values (1e-5)."""
import pytest
import torch

from oracle import vsl_oracle as O
from tests.conftest import rel_err
from tf_depth_estimation_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TERMS = ('depth', 'smooth', 'optflow', 'pixel')


def cu(t, grad=False):
    return t.to(DEV).float().contiguous().requires_grad_(grad)


def test_flow_depth_loss_golden(golden_flow):
    c = golden_flow
    flags = ops.FlowLossFlags(**c.flags)
    S = flags.num_scales
    pd = [cu(c['pred_depth%d' % s], True) for s in range(S)]
    fx = [cu(c['pred_flowx%d' % s], True) for s in range(S)]
    fy = [cu(c['pred_flowy%d' % s], True) for s in range(S)]
    total, losses = ops.flow_depth_loss(cu(c.left), cu(c.right), cu(c.label), pd, fx, fy, cu(c.proj), cu(c.K_pyr), flags)
    total.backward()
    for i, key in enumerate(TERMS):
        want = float(c[key + '_f64'])
        assert abs(float(losses[i]) - want) <= 1e-5 * abs(want), (key, float(losses[i]), want)
    assert abs(float(total.detach()) - sum(float(c[k + '_f64']) for k in TERMS)) <= 1e-5 * float(total.detach())
    for name, leaves in (('g_pred_depth', pd), ('g_pred_flowx', fx), ('g_pred_flowy', fy)):
        for s in range(S):
            e = rel_err(leaves[s].grad, c['%s%d_f32' % (name, s)])
            assert e <= 1e-4, (name, s, e)


def _oracle(d, flags, sel=None):
    pick = (lambda t: t) if sel is None else (lambda t: t[sel])
    pd = [pick(x).clone().requires_grad_() for x in d['depth_pyr']]
    fx = [pick(x).clone().requires_grad_() for x in d['flowx_pyr']]
    fy = [pick(x).clone().requires_grad_() for x in d['flowy_pyr']]
    terms = O.flow_depth_loss(pick(d['left']), pick(d['right']), pick(d['label']), pd, fx, fy, pick(d['proj']),
                              pick(d['K_pyr']), O.FlowLossFlags(**flags.__dict__))
    sum(terms).backward()
    return terms, pd, fx, fy


def _cuda(d, flags, loss_scale=1.0):
    pd = [cu(x, True) for x in d['depth_pyr']]
    fx = [cu(x, True) for x in d['flowx_pyr']]
    fy = [cu(x, True) for x in d['flowy_pyr']]
    total, losses = ops.flow_depth_loss(cu(d['left']), cu(d['right']), cu(d['label']), pd, fx, fy, cu(d['proj']),
                                        cu(d['K_pyr']), flags, loss_scale=loss_scale)
    total.backward()
    return total, losses, pd, fx, fy


@pytest.mark.parametrize('B,H,W,S,motion', [(2, 24, 40, 3, 1.0), (3, 48, 72, 4, 6.0), (1, 12, 20, 2, 1.0), (2, 16, 24, 1, 3.0),
                                              (2, 72, 104, 3, 2.0), (1, 40, 72, 2, 1.0)])
def test_flow_depth_loss_against_oracle(B, H, W, S, motion):
    """Ragged tiles (widths and heights that are no multiple of the 32 x 32 tile, levels that straddle one, two or three
    tile rows, levels smaller than a tile), 1-4 scales, a large-motion case with a third of the samples out of view."""
    d = synth.make_flow_pairs(B, H, W, S=S, seed=300 + H, motion=motion)
    flags = ops.FlowLossFlags(num_scales=S, smooth_weight=0.3, depth_weight=1.5, data_weight=2.0, optflow_weight=0.4)
    terms, opd, ofx, ofy = _oracle(d, flags)
    total, losses, pd, fx, fy = _cuda(d, flags)
    for i, key in enumerate(TERMS):
        want = float(terms[i])
        assert abs(float(losses[i]) - want) <= 1e-5 * abs(want), (key, float(losses[i]), want)
    for name, got, want in (('pd', pd, opd), ('fx', fx, ofx), ('fy', fy, ofy)):
        for s in range(S):
            e = rel_err(got[s].grad, want[s].grad)
            assert e <= 1e-4, (name, s, e)


def test_flow_depth_loss_matches_standalone_ops():
    """The same loss composed from the reference-named stand-alone ops of this library (warp, optflow_warp,
    depth_optflow, compute_smooth_loss, image_pyramid) through torch autograd."""
    B, H, W, S = 2, 32, 64, 3
    d = synth.make_flow_pairs(B, H, W, S=S, seed=5)
    flags = ops.FlowLossFlags(num_scales=S, smooth_weight=0.5, depth_weight=1.0, data_weight=1.0, optflow_weight=1.0)
    total, losses, pd, fx, fy = _cuda(d, flags)
    qd = [cu(x, True) for x in d['depth_pyr']]
    qx = [cu(x, True) for x in d['flowx_pyr']]
    qy = [cu(x, True) for x in d['flowy_pyr']]
    lefts, rights, labels = (ops.image_pyramid(cu(d[k]), S) for k in ('left', 'right', 'label'))
    proj, K = cu(d['proj']), cu(d['K_pyr'])
    want = 0
    for s in range(S):
        k = 1.0 / 2 ** s
        want = want + flags.smooth_weight * k * (ops.compute_smooth_loss(qd[s]) + ops.compute_smooth_loss(qx[s]) +
                                                 ops.compute_smooth_loss(qy[s]))
        want = want + (labels[s] - qd[s]).abs().mean() * flags.depth_weight * k
        _, coords_gt, wmask = ops.projective_inverse_warp(rights[s], (1.0 / labels[s]).squeeze(3), proj, K[:, s].contiguous(), 'matrix')[:3]
        warped = ops.projective_inverse_warp(rights[s], (1.0 / qd[s]).squeeze(3), proj, K[:, s].contiguous(), 'matrix')[0]
        want = want + ((warped - lefts[s]).abs() * wmask).mean() * flags.data_weight * k
        flowed = ops.optflow_warp(rights[s], qx[s], qy[s])
        want = want + ((flowed - lefts[s]).abs() * wmask).mean() * flags.data_weight * k
        gx, gy = ops.depth_optflow(coords_gt.detach())
        want = want + ((qx[s] - gx).abs().mean() + (qy[s] - gy).abs().mean()) * flags.optflow_weight * k
    want.backward()
    assert abs(float(total.detach()) - float(want.detach())) <= 1e-5 * float(want.detach())
    for got, ref in ((pd, qd), (fx, qx), (fy, qy)):
        for s in range(S):
            assert rel_err(got[s].grad, ref[s].grad) <= 1e-4


def test_flow_depth_loss_full_size_cfg4_and_batch_shards():
    """BASELINE configs[3] at full size (B=64, 192x256, 4 scales): gradients of samples 0, 31 and 63 against the oracle
    run on that three-sample batch (per-sample independence: same gradients up to the batch-size factor), and the
    loss as the mean of two batch shards run with loss_scale = B_local / B (the data-parallel contract)."""
    B, H, W, S = 64, 192, 256, 4
    d = synth.make_flow_pairs(B, H, W, S=S, seed=44)
    flags = ops.FlowLossFlags()
    total, losses, pd, fx, fy = _cuda(d, flags)
    sel = torch.tensor([0, 31, 63])
    terms, opd, ofx, ofy = _oracle(d, flags, sel)
    for got, want in ((pd, opd), (fx, ofx), (fy, ofy)):
        for s in range(S):
            e = rel_err(got[s].grad.cpu()[sel] * (B / 3.0), want[s].grad)
            assert e <= 1e-4, (s, e)
    halves = []
    for lo in (0, 32):
        h = {k: (v[lo:lo + 32] if torch.is_tensor(v) else [t[lo:lo + 32] for t in v]) for k, v in d.items()}
        halves.append(_cuda(h, flags, loss_scale=0.5))
    mean = 0.5 * (float(halves[0][0].detach()) + float(halves[1][0].detach()))
    assert abs(mean - float(total.detach())) <= 1e-5 * abs(mean)
    for s in range(S):
        both = torch.cat([halves[0][2][s].grad, halves[1][2][s].grad])
        assert rel_err(both, pd[s].grad) <= 1e-6
    # run to run the step is deterministic (no atomics anywhere)
    again = _cuda(d, flags)
    assert torch.equal(again[1], losses) and all(torch.equal(a.grad, b.grad) for a, b in zip(again[2], pd))


def test_flow_depth_loss_argument_checks():
    d = synth.make_flow_pairs(1, 16, 32, S=2, seed=1)
    flags = ops.FlowLossFlags(num_scales=2)
    args = lambda **kw: [kw.get('left', cu(d['left'])), cu(d['right']), kw.get('label', cu(d['label'])),
                         [cu(x) for x in d['depth_pyr']], kw.get('fx', [cu(x) for x in d['flowx_pyr']]),
                         [cu(x) for x in d['flowy_pyr']], kw.get('proj', cu(d['proj'])), cu(d['K_pyr'])]
    with pytest.raises(TypeError):
        ops.flow_depth_loss(*args(left=d['left']), flags)                       # CPU tensor: no fallback
    with pytest.raises(ValueError):
        ops.flow_depth_loss(*args(label=cu(d['label'])[..., 0]), flags)         # [B,H,W] instead of [B,H,W,1]
    with pytest.raises(ValueError):
        ops.flow_depth_loss(*args(fx=[cu(x) for x in d['flowx_pyr']][::-1]), flags)   # coarsest first
    with pytest.raises(ValueError):
        ops.flow_depth_loss(*args(proj=cu(d['proj'])[:, :3]), flags)            # 3x4 pose
    with pytest.raises(ValueError):
        ops.flow_depth_loss(*args(), ops.FlowLossFlags(num_scales=3))           # pyramids hold 2 levels
