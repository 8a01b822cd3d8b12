"""Randomised parity sweep of the fused step against the float64 oracle (not part of the default test run: ~1 min).
Shapes, view counts, scales, pose formats, mask modes, depth / smoothness conventions, disparity head, both
arithmetic modes.  Prints one line per case and exits non-zero on the first violation of the BASELINE bars."""
import os, random, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vsl_oracle as O
from tests.conftest import rel_err
from tests.parity_util import smooth_pixels, masked_rel_err
from tf_depth_estimation_b200 import ops, synth

dev = torch.device('cuda:0')
cu = lambda t, g=False: t.to(dev).contiguous().requires_grad_(g)
def run(seed=0, n_cases=40, only=-1, verbose=True):
    """-> number of failing cases (only >= 0: examine that one case and print mismatch counts)."""
    rng = random.Random(seed)
    bad = 0
    for case in range(n_cases):
        S = rng.choice([1, 2, 3, 4])
        F = 1 << (S - 1)
        H = F * rng.randint(max(1, (3 + F - 1) // F * 1), 12) * rng.choice([1, 1, 2])
        W = F * rng.randint(max(1, (3 + F - 1) // F * 1), 20) * rng.choice([1, 1, 2])
        H, W = max(H, 3 * F), max(W, 3 * F)
        B, V = rng.randint(1, 3), rng.randint(1, 4)
        fmt = rng.choice(['eular', 'angleaxis', 'matrix'])
        mode = rng.choice(['exp', 'const', 'none'])
        kw = dict(num_scales=S, pose_format=fmt, smooth_on_inverse=rng.random() < 0.4, depth_is_inverse=rng.random() < 0.6,
                  pixel_scale_norm=rng.random() < 0.5, smooth_weight=rng.choice([0.2, 0.5, 3.0]),
                  data_weight=rng.choice([1.0, 10.0]), explain_reg_weight=rng.choice([0.2, 1.0]))
        logit = rng.random() < 0.3
        ar = rng.random()
        exact = 1 if ar < 0.4 else (2 if ar < 0.6 else 0)   # reference rounding / fast scalar kernel / fast (view-paired for even V)
        motion = rng.choice([0.5, 1.5])
        if only >= 0 and case != only:
            continue
        d = synth.make_snippets(B, H, W, S=S, V=V, seed=1000 + case, motion=motion)
        g = torch.Generator().manual_seed(case)
        poses = d['poses']
        if fmt == 'matrix':
            poses = torch.stack([O.pose_vec2mat(d['poses'][:, v], 'eular') for v in range(V)], 1)
        masks = [torch.rand(B, H >> s, W >> s, 1, generator=g) for s in range(S)]
        raw = [0.8 * torch.randn(B, H >> s, W >> s, 1, generator=g) for s in range(S)] if logit else d['disp_pyr']
        flags = ops.LossFlags(exact_coords=exact, x_is_logit=logit, disp_scaling=4.0, min_disp=0.02, **kw)
        xs = [cu(x, True) for x in raw]
        ps = cu(poses, True)
        lgs = [cu(l, True) for l in d['logits_pyr']] if mode == 'exp' else None
        total, losses = ops.view_synthesis_loss(cu(d['tgt']), [cu(s) for s in d['srcs']], xs, ps, cu(d['K_pyr']), logits_pyr=lgs,
                                                mask_pyr=[cu(m) for m in masks] if mode == 'const' else None, flags=flags)
        total.backward()
        oraw = [x.double().requires_grad_() for x in raw]
        ox = [4.0 * torch.sigmoid(x) + 0.02 for x in oraw] if logit else oraw
        op_ = poses.double().requires_grad_()
        ol = [l.double().requires_grad_() for l in d['logits_pyr']] if mode == 'exp' else None
        ref = O.view_synthesis_loss(d['tgt'].double(), [s.double() for s in d['srcs']], ox, op_, d['K_pyr'].double(), ol,
                                    [m.double() for m in masks] if mode == 'const' else None, O.LossFlags(**kw))
        sum(ref).backward()
        el = max(abs(got - float(want)) / (abs(float(want)) + 1e-9) for got, want in zip(losses.tolist(), ref))
        ep = rel_err(ps.grad[:, :, :3], op_.grad[:, :, :3]) if fmt == 'matrix' else rel_err(ps.grad, op_.grad)
        if ep > 1e-4:
            # d/dpose sums over ALL pixels, kinks included: a coordinate within float32 rounding of an integer falls in
            # another bilinear cell in the reference's float32 arithmetic than in float64.  The exact mode reproduces
            # the float32 decisions, so judge it against the float32 oracle before calling it a failure.
            o32 = [x.float().detach().requires_grad_() for x in (ox if not logit else oraw)]
            x32 = [4.0 * torch.sigmoid(x) + 0.02 for x in o32] if logit else o32
            p32 = poses.float().clone().detach().requires_grad_()
            l32 = [l.float().detach().requires_grad_() for l in d['logits_pyr']] if mode == 'exp' else None
            r32 = O.view_synthesis_loss(d['tgt'], d['srcs'], x32, p32, d['K_pyr'], l32, masks if mode == 'const' else None,
                                        O.LossFlags(**kw))
            sum(r32).backward()
            ep = min(ep, rel_err(ps.grad[:, :, :3], p32.grad[:, :, :3]) if fmt == 'matrix' else rel_err(ps.grad, p32.grad))
        disp32 = [(4.0 * torch.sigmoid(x) + 0.02) for x in raw] if logit else d['disp_pyr']
        ok = smooth_pixels(d['tgt'], d['srcs'], disp32, poses, d['K_pyr'], ops.LossFlags(**kw))
        # per-pixel gradients: away from the photometric kinks (mask) AND tolerating the rare smoothness kink (a second
        # difference within float32 rounding of 0 flips its sign and moves the 3-4 pixels of that stencil): the share of
        # pixels off by more than 1e-4 of the largest gradient must stay below 0.5 %
        def frac_off(a, b, m):
            a, b = a.detach().cpu().double(), b.detach().double()
            diff = ((a - b).abs() * m) / b.abs().max().clamp_min(1e-30)
            n_off = int((diff > 1e-4).sum())
            return 0.0 if n_off <= 8 else n_off / diff.numel()   # up to two stencils' worth of pixels in a tiny image
        ex = max(frac_off(xs[s].grad, oraw[s].grad, torch.stack(ok[s]).all(0).unsqueeze(3)) for s in range(S))
        eg = 0.0
        if mode == 'exp':
            eg = max(masked_rel_err(lgs[s].grad, ol[s].grad, torch.stack([o for o in ok[s] for _ in (0, 1)], dim=3)) for s in range(S))
        if only >= 0:
            for sc in range(S):
                ga, gb = xs[sc].grad.cpu().double(), oraw[sc].grad
                m = torch.stack(ok[sc]).all(0).unsqueeze(3)
                diff = ((ga - gb).abs() * m) / gb.abs().max()
                print('scale', sc, 'pixels off by > 1e-4 of max:', int((diff > 1e-4).sum()), 'of', diff.numel(), 'worst', float(diff.max()),
                      'at', [int(i) for i in torch.nonzero(diff == diff.max())[0]])
        fail = el > 1e-5 or ep > 1e-4 or ex > 5e-3 or eg > 1e-4 or not all(torch.isfinite(t.grad).all() for t in xs)
        bad += fail
        if verbose or fail:
            print('%s case %2d B=%d %3dx%-3d S=%d V=%d %-9s %-5s inv(s/d)=%d/%d logit=%d exact=%d | loss %.1e pose %.1e x-off %.4f lg %.1e' % (
                'FAIL' if fail else 'ok  ', case, B, H, W, S, V, fmt, mode, kw['smooth_on_inverse'], kw['depth_is_inverse'], logit, exact,
                el, ep, ex, eg))

    return bad


if __name__ == '__main__':
    a = [int(v) for v in sys.argv[1:]]
    sys.exit(1 if run(*(a + [0, 40, -1][len(a):])) else 0)
