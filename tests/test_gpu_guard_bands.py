"""Memory-safety checks of the fused steps without a memory checker: every buffer a kernel may touch lies inside a larger
allocation whose surroundings hold a canary.

* Outputs and the workspace: the bands before and after them, and the padding between the views of the output arena,
  must still hold the canary after the step (no stray write), and every element of every output view must have been
  written (no element the caller would read uninitialised).
* Inputs: the bands around them hold NaNs (0xFF bytes for uint8 frames); the step's results must be bit-identical to the
  same step on inputs in ordinary allocations -- an out-of-bounds read that reaches a result shows up as a difference.

Shapes are chosen off the kernels' tile sizes (widths that are not a multiple of 32, levels smaller than a tile, odd
level sizes) and the poses throw a third of the pixels out of view, where the gather indices are clamped."""
import ctypes

import pytest
import torch

from tf_depth_estimation_b200 import _lib, ops, synth
from tf_depth_estimation_b200._lib import check, ptr_array

pytestmark = pytest.mark.gpu

CANARY = 0x7FC0BEEF          # a quiet NaN with a payload no kernel produces
BAND = 64 * 1024             # floats (256 KiB) on each side; a multiple of 64 keeps the 256-byte alignment


def _canary(n, dev):
    return torch.full((n,), CANARY, dtype=torch.int32, device=dev).view(torch.float32)


class Guarded(object):
    """A tensor inside a larger buffer filled with the canary (float32) or 0xFF (uint8)."""

    def __init__(self, shape, dtype, dev, like=None):
        cnt = 1
        for d in shape:
            cnt *= d
        if dtype == torch.uint8:
            band = BAND * 4
            self.big = torch.full((band + cnt + band,), 0xFF, dtype=torch.uint8, device=dev)
        else:
            band = BAND
            self.big = _canary(band + cnt + band, dev)
        self.band, self.cnt = band, cnt
        self.t = self.big[band:band + cnt].view(tuple(shape))
        if like is not None:
            self.t.copy_(like)

    def bands_intact(self):
        raw = self.big.view(torch.int32) if self.big.dtype == torch.float32 else self.big
        want = CANARY if self.big.dtype == torch.float32 else 0xFF
        return bool((raw[:self.band] == want).all()) and bool((raw[self.band + self.cnt:] == want).all())


def _guard_inputs(d, dev, img_dtype=torch.float32):
    """Every tensor of a synth batch moved into its own guarded buffer -> (dict of views, [Guarded])."""
    keep, out = [], {}

    def one(t, dtype=torch.float32):
        g = Guarded(tuple(t.shape), dtype, dev, like=t.to(dev))
        keep.append(g)
        return g.t
    for k, v in d.items():
        is_img = k in ('tgt', 'srcs', 'left', 'right')
        dt = img_dtype if is_img else torch.float32
        out[k] = [one(t, dt) for t in v] if isinstance(v, (list, tuple)) else one(v, dt)
    return out, keep


def _guarded_outputs(plan):
    """A LossOutputs of the plan whose arena sits between two canary bands and is itself pre-filled with the canary."""
    keep = {}
    orig = ops._arena

    def arena(shapes, device=None, pinned=False, dtypes=None):
        buf, views = orig(shapes, device=device, pinned=pinned, dtypes=dtypes)
        big = _canary(BAND + buf.numel() + BAND, device)
        nbuf = big[BAND:BAND + buf.numel()]
        nviews = []
        for v in views:
            o = (v.data_ptr() - buf.data_ptr()) // 4
            nviews.append(nbuf[o:o + v.numel()].view(*v.shape))
        keep['big'], keep['views'], keep['n'] = big, nviews, buf.numel()
        return nbuf, nviews
    ops._arena = arena
    try:
        out = plan.new_outputs()
    finally:
        ops._arena = orig
    return out, keep


def _check_arena(keep, written_views):
    """Bands and inter-view padding untouched; every element of the views the step owns written."""
    big = keep['big'].view(torch.int32)
    untouched = torch.ones(big.numel(), dtype=torch.bool, device=big.device)
    base = keep['big'].data_ptr()
    o = (keep['views'][0].data_ptr() - base) // 4
    untouched[o:o + 8] = False               # losses[8]: the terms, their sum, reserved slots
    for v in written_views:
        o = (v.data_ptr() - base) // 4
        untouched[o:o + v.numel()] = False
        assert not bool((v.view(torch.int32) == CANARY).any()), 'an output element was never written'
    assert bool((big[untouched] == CANARY).all()), 'a write landed outside the output views'


def _guard_ws(plan, dev):
    n = plan.ws.numel()
    assert n % 4 == 0
    big = _canary(BAND + n // 4 + BAND, dev)
    plan.ws = big.view(torch.uint8)[BAND * 4:BAND * 4 + n]
    assert plan.ws.data_ptr() % 256 == 0
    return big, n // 4


def _ws_bands_intact(big, n):
    raw = big.view(torch.int32)
    return bool((raw[:BAND] == CANARY).all()) and bool((raw[BAND + n:] == CANARY).all())


FUSED_CASES = [
    # B, H, W, S, V, flags, mask mode, uint8 frames, d(source images)
    (3, 40, 72, 4, 2, {}, _lib.MASK_EXP, False, False),                       # the paired kernel, ragged tiles
    (1, 128, 416, 4, 2, {}, _lib.MASK_EXP, False, False),                     # one cfg2 image
    (2, 48, 104, 3, 1, {'pose_format': 'angleaxis'}, _lib.MASK_EXP, False, False),   # one view: the scalar kernel
    (2, 16, 24, 2, 3, {}, _lib.MASK_NONE, False, False),                      # odd V, levels smaller than a tile
    (2, 32, 40, 3, 2, {'exact_coords': True}, _lib.MASK_EXP, False, False),   # the reference's rounding sequence
    (2, 24, 56, 2, 2, {}, _lib.MASK_EXP, False, True),                        # + the scatter into d(source images)
    (2, 40, 72, 4, 2, {'img_format': 'u8_255'}, _lib.MASK_EXP, True, False),  # the loader's uint8 frames
    (2, 32, 64, 3, 2, {'ssim_weight': 0.85}, _lib.MASK_EXP, False, False),    # + the SSIM-term launch
    (2, 32, 72, 3, 4, {}, _lib.MASK_EXP, False, False),                       # two view pairs
]


@pytest.mark.parametrize('B,H,W,S,V,fl,mask_mode,u8,dsrc', FUSED_CASES)
def test_fused_step_stays_inside_its_buffers(B, H, W, S, V, fl, mask_mode, u8, dsrc):
    dev = torch.device('cuda:0')
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=31, motion=6.0)
    d = {k: d[k] for k in ('tgt', 'srcs', 'disp_pyr', 'poses', 'K_pyr', 'logits_pyr')}
    if u8:
        d['tgt'] = (d['tgt'] * 255).round().clamp(0, 255).to(torch.uint8)
        d['srcs'] = [(s * 255).round().clamp(0, 255).to(torch.uint8) for s in d['srcs']]
    if fl.get('pose_format') == 'angleaxis':
        d['poses'] = d['poses'].clone()
    flags = ops.LossFlags(num_scales=S, **fl)
    img_dtype = torch.uint8 if u8 else torch.float32

    def run(guarded):
        plan = ops.ViewSynthesisPlan(B, H, W, V, flags, mask_mode, dev, want_src_grad=dsrc)
        if guarded:
            g, keep_in = _guard_inputs(d, dev, img_dtype)
            out, keep_out = _guarded_outputs(plan)
            ws_big, ws_n = _guard_ws(plan, dev)
            g_srcs = [Guarded((B, H, W, 3), torch.float32, dev) for _ in range(V)] if dsrc else None
        else:
            g = {k: ([t.to(dev).contiguous() for t in v] if isinstance(v, list) else v.to(dev).contiguous())
                 for k, v in d.items()}
            out, g_srcs = plan.out, None
        logits = g['logits_pyr'] if mask_mode == _lib.MASK_EXP else None
        args = plan.bind(g['tgt'], g['srcs'], g['disp_pyr'], g['poses'], g['K_pyr'], logits, out=out,
                         g_srcs=[x.t for x in g_srcs] if g_srcs else None)
        plan.run_bound(args)
        plan.run_bound(args)                 # a second step over the same workspace (stale contents must not matter)
        torch.cuda.synchronize()
        res = [out.losses.clone()] + [t.clone() for t in out.g_x] + [out.g_poses.clone()]
        if out.g_logits:
            res += [t.clone() for t in out.g_logits]
        if dsrc:
            res += [x.t.clone() for x in g_srcs] if guarded else [t.clone() for t in plan.g_srcs]
        if guarded:
            written = [out.losses] + list(out.g_x) + [out.g_poses] + list(out.g_logits or [])
            _check_arena(keep_out, written)
            assert _ws_bands_intact(ws_big, ws_n), 'a write landed outside the workspace'
            for x in keep_in + (g_srcs or []):
                assert x.bands_intact(), 'a write landed next to an input'
            for x in g_srcs or []:
                assert not bool((x.t.view(torch.int32) == CANARY).any())
        return res

    plain, guarded = run(False), run(True)
    # the two paths that accumulate with floating-point atomics (the scatter into d(source images), the SSIM term's
    # pose partials) differ from run to run in the last bits; a NaN from a band would still poison the result
    atomics = dsrc or fl.get('ssim_weight', 0.0) > 0.0
    for a, b in zip(plain, guarded):
        if atomics:
            assert bool(torch.isfinite(b).all()) and torch.allclose(a, b, rtol=1e-5, atol=1e-7)
        else:
            assert torch.equal(a.view(torch.int32), b.view(torch.int32)), 'results depend on what lies around the inputs'


@pytest.mark.parametrize('B,H,W,S', [(3, 40, 72, 4), (2, 192, 256, 4), (2, 24, 40, 2), (1, 32, 32, 1)])
def test_flow_step_stays_inside_its_buffers(B, H, W, S):
    dev = torch.device('cuda:0')
    lib = _lib.load()
    base = synth.make_flow_pairs(B, H, W, S=S, seed=5)
    base = {k: v for k, v in base.items() if k != 'K'}
    desc = _lib.VslFlowLossDesc(B=B, H=H, W=W, S=S, smooth_weight=0.5, depth_weight=1.0, data_weight=1.0,
                                optflow_weight=1.0, loss_scale=1.0)
    nws = lib.vsl_flow_loss_ws_bytes(ctypes.byref(desc))
    assert nws > 0 and nws % 4 == 0
    st = torch.cuda.current_stream(dev).cuda_stream
    P = lambda ts: ptr_array([t.data_ptr() for t in ts])

    def run(guarded):
        if guarded:
            g, keep = _guard_inputs(base, dev)
            ws = Guarded((nws // 4,), torch.float32, dev)
            losses = Guarded((8,), torch.float32, dev)
            grads = [[Guarded((B, H >> s, W >> s, 1), torch.float32, dev) for s in range(S)] for _ in range(3)]
            keep += [ws, losses] + [x for gs in grads for x in gs]
            wst, lt, gt = ws.t, losses.t, [[x.t for x in gs] for gs in grads]
        else:
            g = {k: ([t.to(dev).contiguous() for t in v] if isinstance(v, list) else v.to(dev).contiguous())
                 for k, v in base.items()}
            wst = torch.empty(nws // 4, device=dev)
            lt = torch.zeros(8, device=dev)
            gt = [[torch.empty(B, H >> s, W >> s, 1, device=dev) for s in range(S)] for _ in range(3)]
        for _ in range(2):
            check(lib.vsl_flow_loss_fwd_bwd(ctypes.byref(desc), g['left'].data_ptr(), g['right'].data_ptr(),
                                            g['label'].data_ptr(), P(g['depth_pyr']), P(g['flowx_pyr']),
                                            P(g['flowy_pyr']), g['proj'].data_ptr(), g['K_pyr'].data_ptr(),
                                            lt.data_ptr(), P(gt[0]), P(gt[1]), P(gt[2]), wst.data_ptr(), st))
        torch.cuda.synchronize()
        if guarded:
            for x in keep:
                assert x.bands_intact(), 'a write landed outside a buffer of the flow step'
            for gs in gt:
                for t in gs:
                    assert not bool((t.view(torch.int32) == CANARY).any()), 'a gradient element was never written'
            assert not bool((lt[:5].view(torch.int32) == CANARY).any())
        return [lt[:5].clone()] + [t.clone() for gs in gt for t in gs]

    plain, guarded = run(False), run(True)
    for a, b in zip(plain, guarded):
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), 'results depend on what lies around the inputs'


def test_adam_and_strip_kernels_stay_inside_their_buffers():
    dev = torch.device('cuda:0')
    lib = _lib.load()
    st = torch.cuda.current_stream(dev).cuda_stream
    # the loader's strip kernel: 2 frames side by side, resized on the way
    B, h, w, H, W = 2, 37, 2 * 53, 24, 40
    gen = torch.Generator().manual_seed(3)
    strip = Guarded((B, h, w, 3), torch.uint8, dev,
                    like=torch.randint(0, 256, (B, h, w, 3), generator=gen, dtype=torch.uint8))
    tgt, src = Guarded((B, H, W, 3), torch.float32, dev), Guarded((B, H, W, 3), torch.float32, dev)
    check(lib.vsl_unpack_strip(strip.t.data_ptr(), B, h, w, H, W, tgt.t.data_ptr(), src.t.data_ptr(), st))
    torch.cuda.synchronize()
    for x in (strip, tgt, src):
        assert x.bands_intact()
    for x in (tgt, src):
        assert not bool((x.t.view(torch.int32) == CANARY).any())
        assert bool(((x.t >= 0) & (x.t <= 255)).all())      # only in-bounds bytes were blended
    # the optimiser step over a range whose length is not a multiple of the vector width, at a 16-byte aligned offset
    n = 4 * 1000 + 3
    bufs = [Guarded((n,), torch.float32, dev, like=torch.randn(n, generator=gen)) for _ in range(2)]
    bufs += [Guarded((n,), torch.float32, dev, like=torch.zeros(n)) for _ in range(2)]
    check(lib.vsl_adam_step(bufs[0].t.data_ptr(), bufs[1].t.data_ptr(), bufs[2].t.data_ptr(), bufs[3].t.data_ptr(), n,
                            2e-4, 0.9, 0.999, 1e-8, 1, 1.0, st))
    torch.cuda.synchronize()
    for x in bufs:
        assert x.bands_intact()
        assert bool(torch.isfinite(x.t).all())


class _GuardedTorch(object):
    """Stands in for the `torch` module inside tf_depth_estimation_b200.ops: everything the wrappers allocate on the
    device (outputs, gradients, workspaces -- all through torch.empty / torch.empty_like) lands between canary bands."""

    def __init__(self):
        self.guards = []

    def __getattr__(self, name):
        return getattr(torch, name)

    def _guarded(self, shape, dtype, dev):
        g = Guarded(tuple(shape), dtype, dev)
        self.guards.append(g)
        return g.t

    def empty(self, *size, **kw):
        dev, dtype = kw.get('device'), kw.get('dtype', torch.float32)
        if dev is None or torch.device(dev).type != 'cuda' or dtype not in (torch.float32, torch.uint8):
            return torch.empty(*size, **kw)
        shape = size[0] if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else size
        return self._guarded(shape, dtype, torch.device(dev))

    def empty_like(self, t, **kw):
        if kw or not t.is_cuda or t.dtype != torch.float32:
            return torch.empty_like(t, **kw)
        return self._guarded(t.shape, t.dtype, t.device)


def test_standalone_ops_stay_inside_their_buffers(monkeypatch):
    """What an unedited reference loop calls (`from utils_lr import *`): every op forward and backward, with the
    inputs between NaN bands and every buffer the wrapper allocates between canary bands."""
    dev = torch.device('cuda:0')
    B, H, W = 2, 24, 40           # off every tile size; the coarse level (12 x 20) is smaller than a tile
    d = synth.make_snippets(B, H, W, S=2, V=2, seed=11, motion=6.0)
    f = synth.make_flow_pairs(B, H, W, S=2, seed=12)

    def run(guarded):
        gen = torch.Generator().manual_seed(13)
        rnd = lambda *shape: torch.rand(*shape, generator=gen)
        keep = []
        gt = _GuardedTorch()
        if guarded:
            monkeypatch.setattr(ops, 'torch', gt)

        def inp(t, grad=False):
            if guarded:
                g = Guarded(tuple(t.shape), torch.float32, dev, like=t.to(dev))
                keep.append(g)
                x = g.t
            else:
                x = t.to(dev).contiguous()
            return x.requires_grad_() if grad else x
        res = []

        def done(outs, leaves):
            outs = [o for o in (outs if isinstance(outs, (tuple, list)) else [outs]) if torch.is_tensor(o)]
            res.extend(o.detach().clone() for o in outs)
            diff = [o for o in outs if o.requires_grad]
            if diff and leaves:
                gs = torch.autograd.grad([o.sum() for o in diff], leaves, allow_unused=True)
                res.extend(g.clone() for g in gs if g is not None)
        try:
            img, depth = inp(d['srcs'][0], True), inp(1.0 / d['disp_pyr'][0][..., 0], True)
            pose, K = inp(d['poses'][:, 0], True), inp(d['K'])
            done(ops.projective_inverse_warp(img, depth, pose, K, 'eular'), [img, depth, pose])
            vec = inp(d['poses'][:, 1], True)
            done(ops.pose_vec2mat(vec, 'angleaxis'), [vec])
            coords = inp(rnd(B, H, W, 2) * torch.tensor([W + 6.0, H + 6.0]) - 3.0, True)
            im2 = inp(d['tgt'], True)
            done(ops.bilinear_sampler(im2, coords), [im2, coords])
            fx, fy = inp(f['flowx_pyr'][0], True), inp(f['flowy_pyr'][0], True)
            im3 = inp(f['right'], True)
            done(ops.optflow_warp(im3, fx, fy), [im3, fx, fy])
            done(ops.depth_optflow(coords.detach()), [])
            sd, pr = inp(rnd(B, H, W, 1) + 0.5, True), inp(rnd(B, H, W, 1) + 0.5, True)
            done(ops.consistent_depth_loss(sd, pr, coords), [sd, pr, coords])
            grid = ops.meshgrid(B, H, W, True, device=dev)
            done(grid, [])
            dep2 = inp(rnd(B, H, W) + 0.5, True)
            cam = ops.pixel2cam(dep2, grid, K, True)
            done(cam, [dep2])
            proj = inp(f['proj'], True)
            cam2 = inp(cam.detach().cpu(), True)
            done(ops.cam2pixel(cam2, proj), [cam2, proj])
            axis, ang = inp(rnd(B, 3) - 0.5, True), inp(rnd(B, 1, 1) + 0.1, True)
            done(ops.axis_angle_to_rotation_matrix(axis, ang), [axis, ang])
            for inverse in (False, True):
                x = inp(d['disp_pyr'][1], True)
                done(ops.compute_smooth_loss(x, inverse), [x])
            lg = inp(d['logits_pyr'][1][..., :2], True)
            done(ops.compute_exp_reg_loss(lg), [lg])
            a, b = inp(d['tgt'], True), inp(d['srcs'][1], True)
            done(ops.ssim_loss(a, b), [a, b])
            done(ops.ssim_dissimilarity(a, b), [a, b])
            dsp, im4 = inp(d['disp_pyr'][0], True), inp(d['tgt'], True)
            done(ops.edge_aware_smooth_loss(dsp, im4), [dsp, im4])
            done(ops.image_pyramid(inp(d['tgt']), 3)[1:], [])
        finally:
            if guarded:
                monkeypatch.undo()
        torch.cuda.synchronize()
        if guarded:
            assert len(gt.guards) > 40            # the wrappers' allocations did go through the stand-in
            for g in keep + gt.guards:
                assert g.bands_intact(), 'a write landed outside a buffer of a stand-alone op'
        return res

    plain, guarded = run(False), run(True)
    assert len(plain) == len(guarded) and len(plain) > 40
    for a, b in zip(plain, guarded):
        assert a.shape == b.shape and bool(torch.isfinite(b).all())
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), 'results depend on what lies around the inputs'
