import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tf_depth_estimation_b200 import ops, synth, _lib
dev = torch.device('cuda:0')
B, H, W, S, V = 32, 128, 416, 4, 2
host = synth.make_snippets(B, H, W, S=S, V=V, seed=1234)
NS = 6
plan = ops.ViewSynthesisPlan(B, H, W, V, ops.LossFlags(), _lib.MASK_EXP, dev)
def to_dev(roll):
    r = lambda t: torch.roll(t, roll, dims=0).to(dev).contiguous()
    return (r(host['tgt']), [r(s) for s in host['srcs']], [r(x) for x in host['disp_pyr']], r(host['poses']), r(host['K_pyr']), [r(l) for l in host['logits_pyr']])
sets = [to_dev(i) for i in range(NS)]
bound = [plan.bind(*s) for s in sets]
for i in range(10): plan.run_bound(bound[i % NS])
torch.cuda.synchronize()
def timeit(fn, n=600):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / n
print('direct  us/step %.1f' % timeit(lambda i: plan.run_bound(bound[i % NS])))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for k in range(NS): plan.run_bound(bound[k], torch.cuda.current_stream().cuda_stream)
for _ in range(3): g.replay()
torch.cuda.synchronize()
print('graph of %d steps us/step %.1f' % (NS, timeit(lambda i: g.replay(), 100) / NS))
