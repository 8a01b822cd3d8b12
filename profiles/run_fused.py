"""Runs the fused step of a BASELINE configuration a few times (the command ncu wraps; also prints in-situ times).
   python profiles/run_fused.py cfg2|cfg4|cfg5 [arith: 0 paired/fast, 1 exact, 2 scalar fast] [steps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_depth_estimation_b200 import ops, synth, _lib
if os.environ.get('VSL_LIB_PATH'):
    pass                                # _lib honours VSL_LIB_PATH itself
elif os.environ.get('VSL_LIB_VARIANT'):   # timing experiments: profiles/build_variant.sh
    _lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_exp', 'libvsl_%s.so' % os.environ['VSL_LIB_VARIANT'])
CFG = {'cfg2': (32, 128, 416, 4, 2), 'cfg3': (256, 128, 416, 4, 2), 'cfg4': (64, 192, 256, 4, 1), 'cfg5': (64, 480, 640, 4, 2)}
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
arith = int(sys.argv[2]) if len(sys.argv) > 2 else 0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
B, H, W, S, V = CFG[name]
dev = torch.device('cuda:0')
cu = lambda t: t.to(dev).contiguous()
d = synth.make_snippets(min(B, 8), H, W, S=S, V=V, seed=7)
rep = lambda t: t.repeat(B // min(B, 8), *([1] * (t.dim() - 1)))
NSETS = int(os.environ.get('VSL_SETS', '6'))      # rotating input sets: streamed operands come from DRAM, as in bench.py
mk = lambda k: (cu(torch.roll(rep(d['tgt']), k, 0)), [cu(torch.roll(rep(s), k, 0)) for s in d['srcs']],
                [cu(torch.roll(rep(x), k, 0)) for x in d['disp_pyr']], cu(torch.roll(rep(d['poses']), k, 0)),
                cu(torch.roll(rep(d['K_pyr']), k, 0)), [cu(torch.roll(rep(l), k, 0)) for l in d['logits_pyr']])
sets = [mk(k) for k in range(NSETS)]
plan = ops.ViewSynthesisPlan(B, H, W, V, ops.LossFlags(exact_coords=arith, ssim_weight=float(os.environ.get('VSL_SSIM', '0'))),
                             _lib.MASK_EXP, dev)   # VSL_SSIM=0.85: the step with the SSIM share of the photometric term
bounds = [plan.bind(*a) for a in sets]
for i in range(3):
    plan.run_bound(bounds[i % NSETS])
torch.cuda.synchronize()
b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
b.record(); e.record(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    plan.run_bound(bounds[i % NSETS])
e1.record(); torch.cuda.synchronize()
ks = []
for i in range(12):
    plan.set_profile_events(b.cuda_event, e.cuda_event)
    plan.run_bound(bounds[i % NSETS])
    torch.cuda.synchronize()
    ks.append(b.elapsed_time(e) * 1e3)
plan.set_profile_events(None, None)
print('%s arith=%d lib=%s sets=%d: %.1f us/step, fused kernel %.1f us (min %.1f), losses %s' % (
    name, arith, os.environ.get('VSL_LIB_VARIANT', 'shipped'), NSETS, e0.elapsed_time(e1) * 1e3 / steps, sum(ks) / len(ks), min(ks),
    [round(float(x), 6) for x in plan.losses.cpu()]))
