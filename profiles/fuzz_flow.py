"""Randomised parity sweep of the fused flow-and-depth step (vsl_flow_loss_fwd_bwd) against the float32 oracle (gradients:
the kernel follows the reference's float32 rounding sequence, so every pixel is compared, kinks included) and the
float64-accumulated loss terms.  Shapes off the 32 x 32 tile, 1-4 scales, coarse levels down to 3 x 3, random term
weights, motions up to a third of the samples out of view.  python profiles/fuzz_flow.py [seed] [cases]"""
import os, random, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vsl_oracle as O
from tests.conftest import rel_err
from tf_depth_estimation_b200 import ops, synth

dev = torch.device('cuda:0')
cu = lambda t, g=False: t.to(dev).float().contiguous().requires_grad_(g)
TERMS = ('depth', 'smooth', 'optflow', 'pixel')


def run(seed=0, n_cases=40, verbose=True):
    """-> number of cases that break a BASELINE bar (loss terms 1e-5 relative, gradients 1e-4 of the largest)."""
    rng = random.Random(seed)
    bad = 0
    for case in range(n_cases):
        S = rng.choice([1, 2, 3, 4])
        F = 1 << (S - 1)
        H = max(F * rng.randint(1, 14) * rng.choice([1, 1, 2]), 3 * F)
        W = max(F * rng.randint(1, 20) * rng.choice([1, 1, 2]), 3 * F)
        B = rng.randint(1, 3)
        motion = rng.choice([0.5, 1.0, 3.0, 6.0])
        kw = dict(num_scales=S, smooth_weight=rng.choice([0.1, 0.5, 2.0]), depth_weight=rng.choice([0.5, 1.0, 3.0]),
                  data_weight=rng.choice([1.0, 5.0]), optflow_weight=rng.choice([0.2, 1.0]))
        d = synth.make_flow_pairs(B, H, W, S=S, seed=seed * 1000 + case, motion=motion)
        flags = ops.FlowLossFlags(**kw)
        opd = [x.clone().requires_grad_() for x in d['depth_pyr']]
        ofx = [x.clone().requires_grad_() for x in d['flowx_pyr']]
        ofy = [x.clone().requires_grad_() for x in d['flowy_pyr']]
        terms = O.flow_depth_loss(d['left'], d['right'], d['label'], opd, ofx, ofy, d['proj'], d['K_pyr'],
                                  O.FlowLossFlags(**kw))
        sum(terms).backward()
        pd = [cu(x, True) for x in d['depth_pyr']]
        fx = [cu(x, True) for x in d['flowx_pyr']]
        fy = [cu(x, True) for x in d['flowy_pyr']]
        total, losses = ops.flow_depth_loss(cu(d['left']), cu(d['right']), cu(d['label']), pd, fx, fy, cu(d['proj']),
                                            cu(d['K_pyr']), flags)
        total.backward()
        tv = [float(t.detach()) for t in terms]
        worst_l = max(abs(float(losses[i]) - tv[i]) / max(abs(tv[i]), 1e-30) for i in range(4))
        worst_g = max(rel_err(g[s].grad, o[s].grad) for g, o in ((pd, opd), (fx, ofx), (fy, ofy)) for s in range(S))
        ok = worst_l <= 1e-5 and worst_g <= 1e-4
        bad += 0 if ok else 1
        if verbose or not ok:
            print('%s case %3d B=%d %dx%d S=%d motion=%.1f %s: loss %.2e grad %.2e' % (
                'ok  ' if ok else 'FAIL', case, B, H, W, S, motion, {k: v for k, v in kw.items() if k != 'num_scales'},
                worst_l, worst_g))
    return bad


if __name__ == '__main__':
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    bad = run(seed, n)
    print('%d of %d cases outside the bars' % (bad, n))
    sys.exit(1 if bad else 0)
