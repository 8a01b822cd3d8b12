import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tf_depth_estimation_b200 import ops, synth, _lib
dev = torch.device('cuda:0')
B, H, W, S, V = 32, 128, 416, 4, 2
d = synth.make_snippets(B, H, W, S=S, V=V, seed=7)
cu = lambda t: t.to(dev).contiguous()
args = (cu(d['tgt']), [cu(s) for s in d['srcs']], [cu(x) for x in d['disp_pyr']], cu(d['poses']), cu(d['K_pyr']), [cu(l) for l in d['logits_pyr']])
for want in (False, True):
    plan = ops.ViewSynthesisPlan(B, H, W, V, ops.LossFlags(), _lib.MASK_EXP, dev, want_src_grad=want)
    bound = plan.bind(*args)
    for _ in range(10): plan.run_bound(bound)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): plan.run_bound(bound)
    e1.record(); torch.cuda.synchronize()
    print('want_src_grad', want, 'us/step %.1f' % (e0.elapsed_time(e1) * 1000 / 200), 'ws MB %.1f' % (plan.ws.numel() / 1e6))
