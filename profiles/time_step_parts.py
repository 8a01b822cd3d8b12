"""Where the cfg2 step's time goes IN SITU (no profiler): CUDA events before the step, around the fused kernel (the
library records those two itself: VslLossDesc.ev_main_begin / ev_main_end) and after the step, over back-to-back steps on
rotating input sets.  prep = step begin -> fused begin, finalize = fused end -> step end (launch gaps included).
   python profiles/time_step_parts.py [cfg2|cfg4|cfg5] [u8]"""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_depth_estimation_b200 import ops, synth, _lib
CFG = {'cfg2': (32, 128, 416, 4, 2), 'cfg4': (64, 192, 256, 4, 1), 'cfg5': (64, 480, 640, 4, 2)}
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
u8 = len(sys.argv) > 2 and sys.argv[2] == 'u8'
B, H, W, S, V = CFG[name]
dev = torch.device('cuda:0')
d = synth.make_snippets(min(B, 8), H, W, S=S, V=V, seed=7)
rep = lambda t: t.repeat(B // min(B, 8), *([1] * (t.dim() - 1)))
img = (lambda t: (t * 255).round().clamp(0, 255).to(torch.uint8)) if u8 else (lambda t: t)
cu = lambda t: t.to(dev).contiguous()
NSETS = 6
mk = lambda k: (cu(img(torch.roll(rep(d['tgt']), k, 0))), [cu(img(torch.roll(rep(s), k, 0))) for s in d['srcs']],
                [cu(torch.roll(rep(x), k, 0)) for x in d['disp_pyr']], cu(torch.roll(rep(d['poses']), k, 0)),
                cu(torch.roll(rep(d['K_pyr']), k, 0)), [cu(torch.roll(rep(l), k, 0)) for l in d['logits_pyr']])
flags = ops.LossFlags(img_format='u8_255') if u8 else ops.LossFlags()
plan = ops.ViewSynthesisPlan(B, H, W, V, flags, _lib.MASK_EXP, dev)
bounds = [plan.bind(*mk(k)) for k in range(NSETS)]
for i in range(10):
    plan.run_bound(bounds[i % NSETS])
torch.cuda.synchronize()
N = 60
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(N)]
for e4 in ev:
    for e in e4:
        e.record()
torch.cuda.synchronize()
for i in range(N):
    ev[i][0].record()
    plan.set_profile_events(ev[i][1].cuda_event, ev[i][2].cuda_event)
    plan.run_bound(bounds[i % NSETS])
    ev[i][3].record()
torch.cuda.synchronize()
plan.set_profile_events(None, None)
us = lambda a, b: a.elapsed_time(b) * 1e3
parts = [(us(e[0], e[1]), us(e[1], e[2]), us(e[2], e[3]), us(e[0], e[3])) for e in ev[5:]]
med = [statistics.median(p[k] for p in parts) for k in range(4)]
print('%s%s in situ (events between the launches add their own ~1-2 us each): prep %.1f us, fused %.1f us, finalize %.1f us, step %.1f us'
      % (name, ' u8' if u8 else '', med[0], med[1], med[2], med[3]))
