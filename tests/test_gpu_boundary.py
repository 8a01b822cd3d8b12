"""GPU: the host boundary around the C ABI -- foreign-framework ingress (DLPack / __cuda_array_interface__), the
layout policy for non-contiguous inputs, shape validation in front of the fused step, the per-call gradient arenas
(any number of forwards before their backwards, backward twice) and the upstream gradient applied by ONE libvsl launch
(no eager torch multiplies on the path)."""
import warnings

import pytest
import torch

from tests.conftest import rel_err
from tf_depth_estimation_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def cu(t, grad=False):
    return t.to(DEV).float().contiguous().requires_grad_(grad)


class DLPackOnly(object):
    """A foreign tensor: nothing but the DLPack protocol (what a TF / JAX / CuPy array offers)."""

    def __init__(self, t):
        self._t = t

    def __dlpack__(self, stream=None):
        return self._t.__dlpack__()

    def __dlpack_device__(self):
        return self._t.__dlpack_device__()


class CaiOnly(object):
    """A foreign buffer: nothing but __cuda_array_interface__ (CuPy, Numba)."""

    def __init__(self, t):
        self._keep = t
        self.__cuda_array_interface__ = t.__cuda_array_interface__


def _small(seed=3, B=2, H=32, W=64, S=3, V=2):
    return synth.make_snippets(B, H, W, S=S, V=V, seed=seed), ops.LossFlags(num_scales=S)


def test_dlpack_and_cuda_array_interface_ingress():
    d, flags = _small()
    img, depth = cu(d['srcs'][0]), cu((1.0 / d['disp_pyr'][0]).squeeze(3))
    pose, K = cu(d['poses'][:, 0]), cu(d['K'])
    want = ops.projective_inverse_warp(img, depth, pose, K)
    got = ops.projective_inverse_warp(DLPackOnly(img), CaiOnly(depth), torch.utils.dlpack.to_dlpack(pose), DLPackOnly(K))
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    # the fused entry, every input foreign; results leave as DLPack capsules
    args = dict(tgt=cu(d['tgt']), srcs=[cu(s) for s in d['srcs']], xs=[cu(x) for x in d['disp_pyr']],
                poses=cu(d['poses']), Kp=cu(d['K_pyr']), lgs=[cu(l) for l in d['logits_pyr']])
    t0, l0 = ops.view_synthesis_loss(args['tgt'], args['srcs'], args['xs'], args['poses'], args['Kp'],
                                     logits_pyr=args['lgs'], flags=flags)
    t1, l1 = ops.view_synthesis_loss(DLPackOnly(args['tgt']), [CaiOnly(s) for s in args['srcs']],
                                     [DLPackOnly(x) for x in args['xs']], CaiOnly(args['poses']), DLPackOnly(args['Kp']),
                                     logits_pyr=[DLPackOnly(l) for l in args['lgs']], flags=flags)
    assert torch.equal(l0, l1) and torch.equal(t0, t1)
    back = torch.utils.dlpack.from_dlpack(ops.to_dlpack(l1))
    assert back.data_ptr() == l1.data_ptr()
    with pytest.raises(TypeError):
        ops.bilinear_sampler(DLPackOnly(d['srcs'][0]), torch.zeros(2, 32, 64, 2, device=DEV))    # a CPU producer
    with pytest.raises(TypeError):
        ops.bilinear_sampler([[1.0]], torch.zeros(2, 32, 64, 2, device=DEV))


def test_layout_policy_for_non_contiguous_inputs():
    d, _ = _small()
    stack = cu(torch.cat(d['srcs'], dim=3))                      # [B,H,W,6]: the reference's src_image_stack
    depth = cu((1.0 / d['disp_pyr'][0]).squeeze(3))
    pose, K = cu(d['poses'][:, 0]), cu(d['K'])
    view = stack[:, :, :, 3:6]                                   # train.py:127 passes exactly such a slice
    want = ops.projective_inverse_warp(view.contiguous(), depth, pose, K)[0]
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter('always')
        got = ops.projective_inverse_warp(view, depth, pose, K)[0]
    assert torch.equal(got, want)
    assert any(issubclass(w.category, ops.VslLayoutWarning) and 'img' in str(w.message) for w in rec)
    prev = ops.set_layout_policy('strict')
    try:
        with pytest.raises(ValueError, match='not contiguous'):
            ops.projective_inverse_warp(view, depth, pose, K)
    finally:
        ops.set_layout_policy(prev)


def test_fused_step_validates_every_shape():
    d, flags = _small()
    a = dict(tgt=cu(d['tgt']), srcs=[cu(s) for s in d['srcs']], xs=[cu(x) for x in d['disp_pyr']],
             poses=cu(d['poses']), Kp=cu(d['K_pyr']), lgs=[cu(l) for l in d['logits_pyr']])
    call = lambda **kw: ops.view_synthesis_loss(kw.get('tgt', a['tgt']), kw.get('srcs', a['srcs']), kw.get('xs', a['xs']),
                                                kw.get('poses', a['poses']), kw.get('Kp', a['Kp']),
                                                logits_pyr=kw.get('lgs', a['lgs']), flags=flags)
    call()
    with pytest.raises(ValueError, match='K_pyr'):
        call(Kp=a['Kp'][:, 0].contiguous())                                     # [B,3,3] instead of [B,S,3,3]
    with pytest.raises(ValueError, match='x_pyr'):
        call(xs=[x.permute(0, 3, 1, 2).contiguous() for x in a['xs']])          # NCHW pyramid
    with pytest.raises(ValueError, match='x_pyr'):
        call(xs=list(reversed(a['xs'])))                                        # coarsest first
    with pytest.raises(ValueError, match='poses'):
        call(poses=a['poses'][:, 0].contiguous())                               # [B,6] instead of [B,V,6]
    with pytest.raises(ValueError, match='logits'):
        call(lgs=[l[..., :2].contiguous() for l in a['lgs']])                   # 2 instead of 2V channels
    with pytest.raises(ValueError, match='srcs'):
        call(srcs=[a['srcs'][0][:, :-1].contiguous(), a['srcs'][1]])
    ceil = [torch.zeros(2, -(-32 // 2 ** s) + (1 if s else 0), 64 >> s, 1, device=DEV) for s in range(3)]
    with pytest.raises(ValueError, match='x_pyr'):
        call(xs=ceil)                                                           # ceil-sized levels
    plan = ops.ViewSynthesisPlan(2, 32, 64, 2, flags, 1, torch.device(DEV))
    with pytest.raises(ValueError):
        plan.bind(a['tgt'], a['srcs'], a['xs'], a['poses'], a['Kp'][:, 0].contiguous(), a['lgs'])
    with pytest.raises(TypeError):
        plan.bind(a['tgt'], a['srcs'], a['xs'], a['poses'].double(), a['Kp'], a['lgs'])


def test_two_forwards_before_their_backwards_and_loss_scale():
    """Left-to-right and right-to-left losses of the same shape summed before one backward
    (train_depth_then_cam_lr.py:253-273): each forward owns its gradient arena."""
    d, flags = _small(seed=9)
    tgt, srcs, Kp = cu(d['tgt']), [cu(s) for s in d['srcs']], cu(d['K_pyr'])

    def leaves():
        return [cu(x, True) for x in d['disp_pyr']], cu(d['poses'], True), [cu(l, True) for l in d['logits_pyr']]
    xs, ps, lgs = leaves()
    ta, _ = ops.view_synthesis_loss(tgt, srcs, xs, ps, Kp, logits_pyr=lgs, flags=flags)
    tb, _ = ops.view_synthesis_loss(srcs[0], [tgt, srcs[1]], xs, ps, Kp, logits_pyr=lgs, flags=flags)
    (ta + 0.5 * tb).backward()
    xs2, ps2, lgs2 = leaves()
    ta2, _ = ops.view_synthesis_loss(tgt, srcs, xs2, ps2, Kp, logits_pyr=lgs2, flags=flags)
    ta2.backward()
    tb2, _ = ops.view_synthesis_loss(srcs[0], [tgt, srcs[1]], xs2, ps2, Kp, logits_pyr=lgs2, flags=flags, loss_scale=0.5)
    tb2.backward()                                   # the 0.5 folded into the kernel instead of into autograd
    assert rel_err(ps.grad, ps2.grad) <= 1e-6
    for s in range(3):
        assert rel_err(xs[s].grad, xs2[s].grad) <= 1e-6 and rel_err(lgs[s].grad, lgs2[s].grad) <= 1e-6


def test_backward_twice_with_different_upstream_gradients():
    d, flags = _small(seed=11)
    xs, ps = [cu(x, True) for x in d['disp_pyr']], cu(d['poses'], True)
    lgs = [cu(l, True) for l in d['logits_pyr']]
    total, _ = ops.view_synthesis_loss(cu(d['tgt']), [cu(s) for s in d['srcs']], xs, ps, cu(d['K_pyr']), logits_pyr=lgs, flags=flags)
    (3.0 * total).backward(retain_graph=True)
    g3 = [ps.grad.clone()] + [x.grad.clone() for x in xs]
    ps.grad = None
    for x in xs:
        x.grad = None
    (0.25 * total).backward()
    g025 = [ps.grad] + [x.grad for x in xs]
    for a, b in zip(g3, g025):
        assert rel_err(a / 12.0, b) <= 1e-6


def test_fused_loss_launches_no_eager_torch_kernels():
    """total.backward() after the fused step: the profiler must see libvsl's kernels only -- no at::native elementwise
    multiply / copy / reduce (the round-1 wrapper re-touched every gradient with eager torch launches)."""
    from torch.profiler import ProfilerActivity, profile
    d, flags = _small(seed=13)
    tgt, srcs, Kp = cu(d['tgt']), [cu(s) for s in d['srcs']], cu(d['K_pyr'])
    xs, ps = [cu(x, True) for x in d['disp_pyr']], cu(d['poses'], True)
    lgs = [cu(l, True) for l in d['logits_pyr']]
    for _ in range(2):                                # warm: plan creation, first-use allocations
        t, _l = ops.view_synthesis_loss(tgt, srcs, xs, ps, Kp, logits_pyr=lgs, flags=flags)
        t.backward()
    for x in xs + lgs + [ps]:
        x.grad = None
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        t, _l = ops.view_synthesis_loss(tgt, srcs, xs, ps, Kp, logits_pyr=lgs, flags=flags)
        t.backward()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA or 'kernel' in e.key.lower()]
    kernels = [n for n in names if 'memcpy' not in n.lower() and 'memset' not in n.lower()]
    assert any('loss_fused' in n for n in kernels), kernels
    bad = [n for n in kernels if 'at::native' in n or 'elementwise' in n]
    # autograd materialises the upstream gradient of `total` (one fill of a single float) -- nothing else may be eager
    assert all('fill' in n.lower() or 'FillFunctor' in n for n in bad), bad


# ---------------------------------------------------------------------------------------- uint8 images
@pytest.mark.parametrize('B,H,W,S,V,fmt', [
    (2, 32, 64, 3, 2, 'u8_255'),            # rows of 64 x 3 bytes: the 16-byte staged path
    (2, 24, 36, 3, 1, 'u8_255_centred'),    # W % 16 != 0: byte loads; one view (scalar kernel)
    (1, 40, 44, 3, 2, 'u8_raw'),            # ragged tiles, raw [0, 255] values
    (4, 128, 416, 4, 2, 'u8_255'),          # BASELINE frame size
])
def test_uint8_images_bit_identical_to_converted_float32(B, H, W, S, V, fmt):
    """vsl_loss_fwd_bwd_u8 (the loader's uint8 frames, converted on load) == vsl_loss_fwd_bwd on the images the
    reference's loader would have produced from them (imageselect_Dataloader.py:93 and siblings, restated in
    oracle.images_from_uint8): losses and every gradient bit for bit."""
    from oracle import vsl_oracle as O
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=17)
    g = torch.Generator().manual_seed(9)
    u8 = [torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8) for _ in range(V + 1)]
    f32 = [O.images_from_uint8(t, fmt) for t in u8]
    scale = 1.0 / 255.0 if fmt == 'u8_raw' else 1.0          # keep the raw-valued loss in a sane range
    res = []
    for imgs, kind in ((f32, 'f32'), (u8, fmt)):
        flags = ops.LossFlags(num_scales=S, img_format=kind, data_weight=scale)
        xs = [cu(x).requires_grad_() for x in d['disp_pyr']]
        ps = cu(d['poses']).requires_grad_()
        lgs = [cu(l).requires_grad_() for l in d['logits_pyr']]
        dev_imgs = [t.to(DEV) for t in imgs]
        total, losses = ops.view_synthesis_loss(dev_imgs[0], dev_imgs[1:], xs, ps, cu(d['K_pyr']), logits_pyr=lgs, flags=flags)
        total.backward()
        res.append([losses.clone(), ps.grad.clone()] + [x.grad.clone() for x in xs] + [l.grad.clone() for l in lgs])
    assert float(res[0][0].abs().sum()) > 0
    for a, b in zip(*res):
        assert torch.equal(a, b)
    # dtype / format mismatches are refused before anything reaches the library
    with pytest.raises(TypeError):
        ops.view_synthesis_loss(u8[0].to(DEV), [t.to(DEV) for t in u8[1:]], xs, ps, cu(d['K_pyr']), logits_pyr=lgs,
                                flags=ops.LossFlags(num_scales=S))
    with pytest.raises(TypeError):
        ops.view_synthesis_loss(f32[0].to(DEV), [t.to(DEV) for t in f32[1:]], xs, ps, cu(d['K_pyr']), logits_pyr=lgs,
                                flags=ops.LossFlags(num_scales=S, img_format=fmt))


def test_host_pipeline_uint8_frames():
    """HostPipeline fed with uint8 frames (a quarter of the image bytes over PCIe) returns what the float32 pipeline
    returns for the converted images."""
    from oracle import vsl_oracle as O
    from tf_depth_estimation_b200 import _lib
    B, H, W, S, V = 2, 32, 64, 3, 2
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=23)
    g = torch.Generator().manual_seed(4)
    u8 = [torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8) for _ in range(V + 1)]
    outs = []
    for kind in ('f32', 'u8_255'):
        pipe = ops.HostPipeline(B, H, W, V, ops.LossFlags(num_scales=S, img_format=kind), _lib.MASK_EXP, torch.device(DEV))
        h = pipe.host_inputs()
        imgs = u8 if kind != 'f32' else [O.images_from_uint8(t) for t in u8]
        h['tgt'].copy_(imgs[0]); h['poses'].copy_(d['poses']); h['Kp'].copy_(d['K_pyr'])
        for dst, src in zip(h['srcs'] + h['xs'] + h['lgs'], imgs[1:] + d['disp_pyr'] + d['logits_pyr']):
            dst.copy_(src)
        l, gx, gp, gl = pipe.result(pipe.submit(h))
        outs.append([l.clone(), gp.clone()] + [t.clone() for t in gx] + [t.clone() for t in gl])
        if kind != 'f32':
            assert pipe.bytes_per_step()[0] < 0.6 * h2d_f32
        else:
            h2d_f32 = pipe.bytes_per_step()[0]
    for a, b in zip(*outs):
        assert torch.equal(a, b)
