"""Reads the per-warp stage stamps a VSL_EXP_TIMELINE build leaves in the workspace (timing experiment).
   VSL_LIB_VARIANT=timeline python profiles/timeline.py [cfg2]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_depth_estimation_b200 import ops, synth, _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_exp', 'libvsl_%s.so' % os.environ.get('VSL_LIB_VARIANT', 'timeline'))
B, H, W, S, V = {'cfg2': (32, 128, 416, 4, 2), 'cfg5': (64, 480, 640, 4, 2)}[sys.argv[1] if len(sys.argv) > 1 else 'cfg2']
dev = torch.device('cuda:0')
cu = lambda t: t.to(dev).contiguous()
d = synth.make_snippets(min(B, 8), H, W, S=S, V=V, seed=7)
rep = lambda t: t.repeat(B // min(B, 8), *([1] * (t.dim() - 1)))
sets = [(cu(torch.roll(rep(d['tgt']), k, 0)), [cu(torch.roll(rep(s), k, 0)) for s in d['srcs']], [cu(torch.roll(rep(x), k, 0)) for x in d['disp_pyr']],
         cu(torch.roll(rep(d['poses']), k, 0)), cu(torch.roll(rep(d['K_pyr']), k, 0)), [cu(torch.roll(rep(l), k, 0)) for l in d['logits_pyr']]) for k in range(6)]
plan = ops.ViewSynthesisPlan(B, H, W, V, ops.LossFlags(), _lib.MASK_EXP, dev)
bounds = [plan.bind(*a) for a in sets]
for i in range(7):
    plan.run_bound(bounds[i % 6])
torch.cuda.synchronize()
ru = lambda n, a: (n + a - 1) // a * a
off = ru(84 * S * V * B, 256) + ru(48 * S * V * B, 256)
n_items = sum(B * ((H >> s) + 31) // 32 * 0 + B * (((H >> s) + 31) // 32) * (((W >> s) + 31) // 32) for s in range(S))
N = 3 + 12 * V
raw = plan.ws[off:off + 4 * n_items * N].view(torch.int32).view(n_items, N).cpu().to(torch.int64) & 0xffffffff
st = raw[:, 0:16:2] + (raw[:, 1:16:2] << 32)          # [tiles, 8] ns
smid, scale = raw[:, 16], raw[:, 17]
t0 = int(st[:, 0].min())
names = ['entry', 'x tile staged', 'halo signs done', 'after griddepcontrol.wait', 'transforms + owners', 'first gathers issued', 'rows done', 'epilogue done']
print('tiles', n_items, 'kernel span %.1f us (first entry -> last exit)' % ((int(st[:, 7].max()) - t0) / 1e3))
for sc in range(S):
    m = scale == sc
    if int(m.sum()) == 0:
        continue
    rel = (st[m] - t0).double() / 1e3
    dur = (st[m][:, 1:] - st[m][:, :-1]).double() / 1e3
    print('scale %d: %d tiles; entry at %.1f..%.1f us; stage durations (mean us): %s; exit at %.1f..%.1f' % (
        sc, int(m.sum()), float(rel[:, 0].min()), float(rel[:, 0].max()),
        ', '.join('%s %.2f' % (n, float(v)) for n, v in zip(names[1:], dur.mean(0))), float(rel[:, 7].min()), float(rel[:, 7].max())))
# per-SM busy: how many SMs, last exit per SM
last = {}
for i in range(n_items):
    k = int(smid[i]); last[k] = max(last.get(k, 0), int(st[i, 7]) - t0)
v = sorted(last.values())
print('SMs used %d; last exit per SM: min %.1f median %.1f max %.1f us' % (len(v), v[0] / 1e3, v[len(v) // 2] / 1e3, v[-1] / 1e3))
