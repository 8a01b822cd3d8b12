import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tf_depth_estimation_b200 import ops, synth, _lib
dev = torch.device('cuda:0')
B, H, W, S, V = 32, 128, 416, 4, 2
d = synth.make_snippets(B, H, W, S=S, V=V, seed=7)
cu = lambda t: t.to(dev).contiguous()
args = (cu(d['tgt']), [cu(s) for s in d['srcs']], [cu(x) for x in d['disp_pyr']], cu(d['poses']), cu(d['K_pyr']), [cu(l) for l in d['logits_pyr']])
for ex in (False, True):
    plan = ops.ViewSynthesisPlan(B, H, W, V, ops.LossFlags(exact_coords=ex), _lib.MASK_EXP, dev)
    bound = plan.bind(*args)
    for _ in range(5): plan.run_bound(bound)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): plan.run_bound(bound)
    e1.record(); torch.cuda.synchronize()
    print('exact_coords', ex, 'us/step %.1f' % (e0.elapsed_time(e1) * 10), [round(float(x), 6) for x in plan.losses.cpu()])
