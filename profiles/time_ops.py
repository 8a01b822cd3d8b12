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
ig = img.clone().requires_grad_()
def fbi():
    out = ops.projective_inverse_warp(ig, dr, pr, K, 'eular')[0]
    out.backward(out)
timeit('projective_inverse_warp fwd+bwd (+ d img: aggregated vector reductions)', fbi, npx * (12 + 4 + 12 + 12 + 8 + 4 + 4 + 12 + 12 + 4 + 4 + 24))
def graphed(fn):
    """The same calls captured into a CUDA graph (the library never synchronises or allocates): what the GPU needs once
    the Python / autograd dispatch of ~10 us per call is out of the way."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    return g.replay
timeit('  ... fwd (5 outputs), CUDA graph replay', graphed(lambda: ops.projective_inverse_warp(img, depth, pose, K, 'eular')), npx * (12 + 4 + 12 + 12 + 8 + 4 + 4))
timeit('  ... fwd+bwd (d depth, d pose), CUDA graph replay', graphed(fb), npx * (12 + 4 + 12 + 12 + 8 + 4 + 4 + 12 + 12 + 4 + 4))
timeit('  ... fwd+bwd (+ d img), CUDA graph replay', graphed(fbi), npx * (12 + 4 + 12 + 12 + 8 + 4 + 4 + 12 + 12 + 4 + 4 + 24))
timeit('compute_smooth_loss fwd', lambda: ops.compute_smooth_loss(disp), npx * 4)
timeit('compute_exp_reg_loss fwd', lambda: ops.compute_exp_reg_loss(lg), npx * 8)
timeit('image_pyramid (3 levels)', lambda: ops.image_pyramid(img, 4), npx * 12 * 1.328)

# ---- consistent_depth_loss as one kernel each way; the extension ops; the flat Adam kernel
tgt = cu(d['tgt'])
coords = ops.projective_inverse_warp(img, depth, pose, K, 'eular')[1].detach()
sd, pd = cu(1.0 / d['disp_pyr'][0]), cu(1.0 / d['disp_pyr'][0]) * 1.01
timeit('consistent_depth_loss fwd', lambda: ops.consistent_depth_loss(sd, pd, coords), npx * (4 + 4 + 8 + 4))
pdr = pd.clone().requires_grad_()
def cfb():
    e = ops.consistent_depth_loss(sd, pdr, coords)
    e.backward(e)
timeit('consistent_depth_loss fwd+bwd (d pred)', cfb, npx * (2 * (4 + 4 + 8) + 4 + 4 + 4))
timeit('ssim_loss fwd (mean, no map)', lambda: ops.ssim_loss(img, tgt), npx * 24)
ir = img.clone().requires_grad_()
def sfb():
    ops.ssim_loss(ir, tgt).backward()
timeit('ssim_loss fwd+bwd (d x)', sfb, npx * (24 + 24 + 12))
timeit('edge_aware_smooth_loss fwd', lambda: ops.edge_aware_smooth_loss(disp, tgt), npx * 16)
dpr = disp.clone().requires_grad_()
def efb():
    ops.edge_aware_smooth_loss(dpr, tgt).backward()
timeit('edge_aware_smooth_loss fwd+bwd (d disp)', efb, npx * (16 + 16 + 4))
NP = 33_200_000
p_, g_, m_, v_ = (torch.zeros(NP, device=dev) for _ in range(4))
g_.normal_()
step = [0]
def adam():
    step[0] += 1
    ops.adam_step(p_, g_, m_, v_, step[0], lr=2e-4)
timeit('adam_step, 33.2 M parameters', adam, NP * 28)
