"""Diagnosis of the full-size parity cases (not a test): for given samples of a BASELINE configuration, how far are
the fused kernel's gradients (fast and exact arithmetic) from the float64 oracle and from the float32 oracle, how far
are the two oracles from each other, and are the per-pixel outliers explained by smoothness kinks (a second
difference within float32 rounding of 0)?   python profiles/diag_fullsize.py cfg5|cfg4|cfg2"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vsl_oracle as O
from tests.conftest import rel_err
from tests.parity_util import smooth_pixels, masked_rel_err
from tf_depth_estimation_b200 import ops, synth

dev = torch.device('cuda:0')
cu = lambda t, g=False: t.to(dev).float().contiguous().requires_grad_(g)

CFG = {
    'cfg2': (32, 128, 416, 4, 2, 1234, (0, 13, 31), {}),
    'cfg5': (64, 480, 640, 4, 2, 1239, (0, 31, 63), {}),
    'cfg4': (64, 192, 256, 4, 1, 1238, (0, 40, 63),
             dict(pose_format='angleaxis', smooth_on_inverse=True, depth_is_inverse=True, pixel_scale_norm=False,
                  smooth_weight=0.3, data_weight=2.0, explain_reg_weight=0.4)),
}


def oracle_one(d, b, kw, dtype):
    sl = slice(b, b + 1)
    c = (lambda t: t.double()) if dtype == torch.float64 else (lambda t: t.float())
    xs = [c(x[sl]).clone().requires_grad_() for x in d['disp_pyr']]
    ps = c(d['poses'][sl]).clone().requires_grad_()
    lg = [c(l[sl]).clone().requires_grad_() for l in d['logits_pyr']]
    r = O.view_synthesis_loss(c(d['tgt'][sl]), [c(s[sl]) for s in d['srcs']], xs, ps, c(d['K_pyr'][sl]), lg, None,
                              O.LossFlags(**kw))
    sum(r).backward()
    return [x.grad for x in xs], ps.grad, [l.grad for l in lg]


def main(name):
    B, H, W, S, V, seed, samples, kw = CFG[name]
    kw = dict(kw, num_scales=S)
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=seed)
    got = {}
    for exact in (False, True):
        flags = ops.LossFlags(exact_coords=exact, **kw)
        xs = [cu(x, True) for x in d['disp_pyr']]
        ps = cu(d['poses'], True)
        lgs = [cu(l, True) for l in d['logits_pyr']]
        total, losses = ops.view_synthesis_loss(cu(d['tgt']), [cu(s) for s in d['srcs']], xs, ps, cu(d['K_pyr']),
                                                logits_pyr=lgs, flags=flags)
        total.backward()
        got[exact] = ([x.grad.cpu() for x in xs], ps.grad.cpu(), [l.grad.cpu() for l in lgs])
    for b in samples:
        sl = slice(b, b + 1)
        gx64, gp64, gl64 = oracle_one(d, b, kw, torch.float64)
        gx32, gp32, gl32 = oracle_one(d, b, kw, torch.float32)
        print('--- %s sample %d' % (name, b))
        print('  g_poses  o32 vs o64 %.2e | fast vs o64 %.2e  fast vs o32 %.2e | exact vs o64 %.2e  exact vs o32 %.2e' % (
            rel_err(gp32, gp64), rel_err(got[False][1][sl] * B, gp64), rel_err(got[False][1][sl] * B, gp32),
            rel_err(got[True][1][sl] * B, gp64), rel_err(got[True][1][sl] * B, gp32)))
        one = dict(tgt=d['tgt'][sl], srcs=[s[sl] for s in d['srcs']], disp=[x[sl] for x in d['disp_pyr']],
                   poses=d['poses'][sl], K=d['K_pyr'][sl])
        ok = smooth_pixels(one['tgt'], one['srcs'], one['disp'], one['poses'], one['K'], ops.LossFlags(**kw))
        for s in range(S):
            m = torch.stack(ok[s]).all(0).unsqueeze(3)
            # smoothness kinks: second differences of q within rounding of 0
            q = one['disp'][s].double()
            if kw.get('smooth_on_inverse'):
                q = 1.0 / q
            q = q[0, :, :, 0]
            eps = 8e-7 * float(q.abs().max())
            near = torch.zeros_like(q, dtype=torch.bool)
            dx = q[:, 1:] - q[:, :-1]; dy = q[1:] - q[:-1]
            dxx = dx[:, 1:] - dx[:, :-1]; dyy = dy[1:] - dy[:-1]; dxy = dx[1:] - dx[:-1]
            kxx = dxx.abs() < eps; kyy = dyy.abs() < eps; kxy = dxy.abs() < eps
            for o in range(3):
                near[:, o:o + kxx.shape[1]] |= kxx
                near[o:o + kyy.shape[0], :] |= kyy
            for oy in range(2):
                for ox in range(2):
                    near[oy:oy + kxy.shape[0], ox:ox + kxy.shape[1]] |= kxy
            line = '  s=%d masked %.3f%% smooth-kink px %d |' % (s, 100 * (1 - float(m.float().mean())), int(near.sum()))
            for exact in (False, True):
                a = got[exact][0][s][sl].double() * B
                for nm, ref in (('o64', gx64[s]), ('o32', gx32[s].double())):
                    diff = ((a - ref).abs() * m) / ref.abs().max()
                    off = diff[0, :, :, 0] > 1e-4
                    line += ' %s/%s off %d (unexplained %d) worst %.1e |' % ('exact' if exact else 'fast', nm, int(off.sum()),
                                                                         int((off & ~near).sum()), float(diff.max()))
            e32 = ((gx32[s].double() - gx64[s]).abs() * m / gx64[s].abs().max())
            line += ' o32/o64 off %d worst %.1e' % (int((e32 > 1e-4).sum()), float(e32.max()))
            print(line)
            ml = torch.stack([o for o in ok[s] for _ in (0, 1)], dim=3)
            print('       g_logits fast vs o64 %.2e exact vs o64 %.2e' % (
                masked_rel_err(got[False][2][s][sl] * B, gl64[s], ml), masked_rel_err(got[True][2][s][sl] * B, gl64[s], ml)))




def worst_logit_pixels(name='cfg5', b=0, s=0, arith=0):
    """Where does d/dlogits differ most from the float64 oracle (sample b, scale s), and what does the pixel look like?"""
    B, H, W, S, V, seed, samples, kw = CFG[name]
    kw = dict(kw, num_scales=S)
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=seed)
    sl = slice(b, b + 1)
    one = {k: ([t[sl] for t in v] if isinstance(v, list) else v[sl]) for k, v in d.items() if k in ('tgt', 'srcs', 'disp_pyr', 'poses', 'K_pyr', 'logits_pyr')}
    flags = ops.LossFlags(exact_coords=arith, **kw)
    lgs = [cu(l, True) for l in one['logits_pyr']]
    total, _ = ops.view_synthesis_loss(cu(one['tgt']), [cu(t) for t in one['srcs']], [cu(x, True) for x in one['disp_pyr']], cu(one['poses'], True),
                                       cu(one['K_pyr']), logits_pyr=lgs, flags=flags)
    total.backward()
    gx64, gp64, gl64 = oracle_one(d, b, kw, torch.float64)
    ok = smooth_pixels(one['tgt'], one['srcs'], one['disp_pyr'], one['poses'], one['K_pyr'], ops.LossFlags(**kw))
    m = torch.stack([o for o in ok[s] for _ in (0, 1)], dim=3)
    got = lgs[s].grad.cpu().double()
    diff = ((got - gl64[s]).abs() * m) / gl64[s].abs().max()
    print('%s b=%d s=%d arith=%d: d/dlogits off by > 1e-4: %d px, > 5e-5: %d px, worst %.2e' % (name, b, s, arith, int((diff > 1e-4).sum()), int((diff > 5e-5).sum()), float(diff.max())))
    hs, ws = H >> s, W >> s
    tgt_s = O.resize_area(one['tgt'].double(), hs, ws)
    for idx in torch.nonzero(diff > 0.8 * diff.max())[:6]:
        _, y, x, c = [int(i) for i in idx]
        v = c // 2
        src_s = O.resize_area(one['srcs'][v].double(), hs, ws)
        depth = (1.0 / one['disp_pyr'][s].double()).squeeze(3)
        warped, coords, _, z, _ = O.projective_inverse_warp(src_s, depth, one['poses'][:, v].double(), one['K_pyr'][:, s].double(), 'eular')
        w32, c32, _, _, _ = O.projective_inverse_warp(src_s.float(), depth.float(), one['poses'][:, v], one['K_pyr'][:, s], 'eular')
        e = (warped - tgt_s)[0, y, x]
        print('  (y=%d x=%d ch=%d) got %.6e want %.6e | coords64 (%.5f, %.5f) coords32-64 (%.1e, %.1e) z %.3f | e = %s' % (
            y, x, c, float(got[0, y, x, c]), float(gl64[s][0, y, x, c]), float(coords[0, y, x, 0]), float(coords[0, y, x, 1]),
            float(c32[0, y, x, 0]) - float(coords[0, y, x, 0]), float(c32[0, y, x, 1]) - float(coords[0, y, x, 1]), float(z[0, y, x, 0]),
            ['%.2e' % float(t) for t in e]))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'logits':
        for arith in (0, 1, 2):
            worst_logit_pixels('cfg5', 0, 0, arith)
    else:
        for n in sys.argv[1:] or ['cfg5', 'cfg4']:
            main(n)
