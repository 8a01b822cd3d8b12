import sys, torch
sys.path.insert(0, '/root/repo')
from tf_depth_estimation_b200 import ops, synth, _lib
dev = torch.device('cuda:0')
for (B, H, W, S, V) in [(2,32,48,2,2),(2,24,64,1,2),(2,32,64,1,1),(1,32,72,4,3),(2,16,48,2,4)]:
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=7)
    flags = ops.LossFlags(num_scales=S)
    plan = ops.ViewSynthesisPlan(B, H, W, V, flags, _lib.MASK_EXP, dev)
    cu = lambda t: t.to(dev).contiguous()
    plan.run(cu(d['tgt']), [cu(s) for s in d['srcs']], [cu(x) for x in d['disp_pyr']], cu(d['poses']), cu(d['K_pyr']),
             [cu(l) for l in d['logits_pyr']])
    torch.cuda.synchronize()
    print('losses', plan.losses.tolist())
    for s in range(S):
        print('g_x', s, torch.isnan(plan.g_x[s]).sum().item(), 'g_lg', torch.isnan(plan.g_logits[s]).sum().item())
    print((B,H,W,S,V), "g_poses nan", torch.isnan(plan.g_poses).sum().item())
