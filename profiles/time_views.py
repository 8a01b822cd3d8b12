import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tf_depth_estimation_b200 import ops, synth, _lib
dev = torch.device('cuda:0')
cu = lambda t: t.to(dev).contiguous()
for V in (1, 2, 3, 4):
    B, H, W, S = 32, 128, 416, 4
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=7)
    args = (cu(d['tgt']), [cu(s) for s in d['srcs']], [cu(x) for x in d['disp_pyr']], cu(d['poses']), cu(d['K_pyr']), [cu(l) for l in d['logits_pyr']])
    plan = ops.ViewSynthesisPlan(B, H, W, V, ops.LossFlags(), _lib.MASK_EXP, dev)
    bound = plan.bind(*args)
    for _ in range(5): plan.run_bound(bound)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): plan.run_bound(bound)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 10
    pv = B * H * W * 1.328125 * V
    print('V=%d: %.1f us/step, %.1f Gpix-views/s' % (V, us, pv / us / 1e3))
