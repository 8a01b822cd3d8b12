"""Times of the stand-alone reference-named ops at the cfg2 frame size (B=32, 128x416); not a bench line."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_depth_estimation_b200 import ops, synth
dev = torch.device('cuda:0')
d = synth.make_snippets(32, 128, 416, S=4, V=2, seed=7)
cu = lambda t: t.to(dev).contiguous()
img, depth = cu(d['srcs'][0]), cu((1.0 / d['disp_pyr'][0]).squeeze(3))
pose, K = cu(d['poses'][:, 0]), cu(d['K'])
disp, lg = cu(d['disp_pyr'][0]), cu(d['logits_pyr'][0][..., :2])


def timeit(name, fn, nbytes, n=100):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / n
    print('%-44s %7.1f us   %6.0f GB/s of compulsory traffic' % (name, us, nbytes / us / 1e3))


npx = 32 * 128 * 416
timeit('projective_inverse_warp fwd (5 outputs)', lambda: ops.projective_inverse_warp(img, depth, pose, K, 'eular'), npx * (12 + 4 + 12 + 12 + 8 + 4 + 4))
dr, pr = depth.clone().requires_grad_(), pose.clone().requires_grad_()
def fb():
    out = ops.projective_inverse_warp(img, dr, pr, K, 'eular')[0]
    out.backward(out)
timeit('projective_inverse_warp fwd+bwd (d depth, d pose)', fb, npx * (12 + 4 + 12 + 12 + 8 + 4 + 4 + 12 + 12 + 4 + 4))
timeit('compute_smooth_loss fwd', lambda: ops.compute_smooth_loss(disp), npx * 4)
timeit('compute_exp_reg_loss fwd', lambda: ops.compute_exp_reg_loss(lg), npx * 8)
timeit('image_pyramid (3 levels)', lambda: ops.image_pyramid(img, 4), npx * 12 * 1.328)
