import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tf_depth_estimation_b200 import ops, synth, _lib
dev = torch.device('cuda:0')
cu = lambda t: t.to(dev).contiguous()
for name, (B, H, W, S, V) in (('cfg2', (32, 128, 416, 4, 2)), ('cfg4 (one direction)', (64, 192, 256, 4, 1)), ('cfg5', (64, 480, 640, 4, 2))):
    d = synth.make_snippets(min(B, 8), H, W, S=S, V=V, seed=7)
    rep = lambda t: t.repeat(B // min(B, 8), *([1] * (t.dim() - 1)))
    args = (cu(rep(d['tgt'])), [cu(rep(s)) for s in d['srcs']], [cu(rep(x)) for x in d['disp_pyr']], cu(rep(d['poses'])),
            cu(rep(d['K_pyr'])), [cu(rep(l)) for l in d['logits_pyr']])
    plan = ops.ViewSynthesisPlan(B, H, W, V, ops.LossFlags(), _lib.MASK_EXP, dev)
    bound = plan.bind(*args)
    for _ in range(5): plan.run_bound(bound)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n): plan.run_bound(bound)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / n
    pv = B * H * W * sum(0.25 ** s for s in range(S)) * V
    algo = B * H * W * sum(0.25 ** s for s in range(S)) * (20 + 28 * V)
    print('%-22s B=%d %dx%d V=%d: %.1f us/step, %.1f Gpix-views/s, step algorithmic %.0f MB -> %.2f TB/s, ws %.0f MB, losses %s' % (
        name, B, H, W, V, us, pv / us / 1e3, algo / 1e6, algo / us / 1e6, plan.ws.numel() / 1e6, [round(float(x), 4) for x in plan.losses.cpu()]))
