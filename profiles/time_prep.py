"""Times launch 1 (pyramids + RGBA re-layout) in situ: step time with and without the fused kernel's own time, and
(under ncu: -k regex:loss_prep) gives the launch list something to wrap.
   python profiles/time_prep.py cfg2|cfg4|cfg5 [S] [steps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_depth_estimation_b200 import ops, synth, _lib
if os.environ.get('VSL_LIB_VARIANT'):
    _lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_exp', 'libvsl_%s.so' % os.environ['VSL_LIB_VARIANT'])
CFG = {'cfg2': (32, 128, 416, 4, 2), 'cfg4': (64, 192, 256, 4, 1), 'cfg5': (64, 480, 640, 4, 2)}
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
B, H, W, S, V = CFG[name]
S = int(sys.argv[2]) if len(sys.argv) > 2 else S
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 60
dev = torch.device('cuda:0')
cu = lambda t: t.to(dev).contiguous()
d = synth.make_snippets(min(B, 8), H, W, S=S, V=V, seed=7)
rep = lambda t: t.repeat(B // min(B, 8), *([1] * (t.dim() - 1)))
NSETS = 6 if name != 'cfg5' else 2
mk = lambda k: (cu(torch.roll(rep(d['tgt']), k, 0)), [cu(torch.roll(rep(s), k, 0)) for s in d['srcs']],
                [cu(torch.roll(rep(x), k, 0)) for x in d['disp_pyr']], cu(torch.roll(rep(d['poses']), k, 0)),
                cu(torch.roll(rep(d['K_pyr']), k, 0)), [cu(torch.roll(rep(l), k, 0)) for l in d['logits_pyr']])
sets = [mk(k) for k in range(NSETS)]
plan = ops.ViewSynthesisPlan(B, H, W, V, ops.LossFlags(num_scales=S), _lib.MASK_EXP, dev)
bounds = [plan.bind(*a) for a in sets]
for i in range(5):
    plan.run_bound(bounds[i % NSETS])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    plan.run_bound(bounds[i % NSETS])
e1.record(); torch.cuda.synchronize()
step_us = e0.elapsed_time(e1) * 1e3 / steps
b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
b.record(); e.record(); torch.cuda.synchronize()   # torch only reads events it has recorded itself once
pre, main, post = [], [], []
for i in range(20):
    plan.set_profile_events(b.cuda_event, e.cuda_event)
    s0.record()
    plan.run_bound(bounds[i % NSETS])
    s1.record()
    torch.cuda.synchronize()
    pre.append(s0.elapsed_time(b) * 1e3); main.append(b.elapsed_time(e) * 1e3); post.append(e.elapsed_time(s1) * 1e3)
plan.set_profile_events(None, None)
med = lambda v: sorted(v)[len(v) // 2]
print('%s S=%d: step %.1f us back to back; isolated: prep %.1f + fused %.1f + finalize %.1f us (medians of 20, events inside the step)'
      % (name, S, step_us, med(pre), med(main), med(post)))
