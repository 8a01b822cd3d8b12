"""GPU: ops.HostPipeline (host buffers in / out, three overlapped streams) returns exactly what the plain plan
computes, for every step of a run in which consecutive steps carry DIFFERENT inputs (buffer reuse hazards)."""
import pytest
import torch

from tf_depth_estimation_b200 import _lib, ops, synth

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


def _host_inputs(d):
    pin = lambda t: t.contiguous().pin_memory()
    return dict(tgt=pin(d['tgt']), srcs=[pin(s) for s in d['srcs']], xs=[pin(x) for x in d['disp_pyr']],
                poses=pin(d['poses']), Kp=pin(d['K_pyr']), lgs=[pin(l) for l in d['logits_pyr']])


def test_pipeline_matches_plan_over_many_steps():
    B, H, W, S, V = 4, 32, 64, 3, 2
    flags = ops.LossFlags(num_scales=S)
    pipe = ops.HostPipeline(B, H, W, V, flags, _lib.MASK_EXP, DEV)
    plan = ops.ViewSynthesisPlan(B, H, W, V, flags, _lib.MASK_EXP, DEV)
    datasets = [synth.make_snippets(B, H, W, S=S, V=V, seed=200 + i) for i in range(5)]
    hosts = [_host_inputs(d) for d in datasets]
    slots, got = [], []
    for i in range(7):                       # more steps than buffer sets: every slot is reused
        slots.append(pipe.submit(hosts[i % 5]))
        if i >= 1:                           # collect step i-1 while step i is in flight
            l, gx, gp, gl = pipe.result(slots[i - 1])
            got.append((l.clone(), [g.clone() for g in gx], gp.clone(), [g.clone() for g in gl]))
    l, gx, gp, gl = pipe.result(slots[-1])
    got.append((l.clone(), [g.clone() for g in gx], gp.clone(), [g.clone() for g in gl]))
    for i, (l, gx, gp, gl) in enumerate(got):
        d = datasets[i % 5]
        cu = lambda t: t.to(DEV).contiguous()
        plan.run(cu(d['tgt']), [cu(s) for s in d['srcs']], [cu(x) for x in d['disp_pyr']], cu(d['poses']),
                 cu(d['K_pyr']), [cu(t) for t in d['logits_pyr']])
        torch.cuda.synchronize()
        assert torch.equal(l, plan.losses.cpu()), i
        assert torch.equal(gp, plan.g_poses.cpu()), i
        assert all(torch.equal(a, b.cpu()) for a, b in zip(gx, plan.g_x)), i
        assert all(torch.equal(a, b.cpu()) for a, b in zip(gl, plan.g_logits)), i
    h2d, d2h = pipe.bytes_per_step()
    assert h2d == 4 * sum(t.numel() for t in pipe._flat(hosts[0])) and d2h > 0


def test_pipeline_single_copy_arena_matches_per_tensor_copies():
    """host_inputs(): the same step fed from one pinned arena (one H2D copy) and from separate pinned tensors."""
    B, H, W, S, V = 2, 32, 64, 3, 2
    flags = ops.LossFlags(num_scales=S)
    pipe = ops.HostPipeline(B, H, W, V, flags, _lib.MASK_EXP, DEV)
    d = synth.make_snippets(B, H, W, S=S, V=V, seed=321)
    sep = _host_inputs(d)
    arena = pipe.host_inputs()
    arena['tgt'].copy_(d['tgt']); arena['poses'].copy_(d['poses']); arena['Kp'].copy_(d['K_pyr'])
    for dst, src in zip(arena['srcs'] + arena['xs'] + arena['lgs'], d['srcs'] + d['disp_pyr'] + d['logits_pyr']):
        dst.copy_(src)
    outs = []
    for h in (sep, arena):
        l, gx, gp, gl = pipe.result(pipe.submit(h))
        outs.append((l.clone(), [g.clone() for g in gx], gp.clone(), [g.clone() for g in gl]))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][2], outs[1][2])
    assert all(torch.equal(a, b) for a, b in zip(outs[0][1] + outs[0][3], outs[1][1] + outs[1][3]))
    assert float(outs[0][0].abs().sum()) > 0


def test_fused_step_is_cuda_graph_capturable():
    """The three launches of a step (programmatic dependent launches included) capture into a CUDA graph; replaying
    the graph on new input values reproduces the directly launched step bit for bit."""
    B, H, W, S, V = 2, 32, 64, 3, 2
    flags = ops.LossFlags(num_scales=S)
    plan = ops.ViewSynthesisPlan(B, H, W, V, flags, _lib.MASK_EXP, DEV)
    d0 = synth.make_snippets(B, H, W, S=S, V=V, seed=11)
    d1 = synth.make_snippets(B, H, W, S=S, V=V, seed=12)
    cu = lambda t: t.to(DEV).contiguous()
    buf = dict(tgt=cu(d0['tgt']), srcs=[cu(s) for s in d0['srcs']], xs=[cu(x) for x in d0['disp_pyr']],
               poses=cu(d0['poses']), Kp=cu(d0['K_pyr']), lgs=[cu(l) for l in d0['logits_pyr']])
    bound = plan.bind(buf['tgt'], buf['srcs'], buf['xs'], buf['poses'], buf['Kp'], buf['lgs'])
    side = torch.cuda.Stream(device=DEV)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        plan.run_bound(bound, side.cuda_stream)          # warm-up outside capture (attribute opt-ins, lazy loads)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        plan.run_bound(bound, torch.cuda.current_stream().cuda_stream)
    # new values in the same buffers, then replay
    buf['tgt'].copy_(d1['tgt']); buf['poses'].copy_(d1['poses']); buf['Kp'].copy_(d1['K_pyr'])
    for dst, src in zip(buf['srcs'] + buf['xs'] + buf['lgs'], d1['srcs'] + d1['disp_pyr'] + d1['logits_pyr']):
        dst.copy_(src)
    graph.replay()
    torch.cuda.synchronize()
    got = (plan.losses.clone(), [g.clone() for g in plan.g_x], plan.g_poses.clone(), [g.clone() for g in plan.g_logits])
    plan.run_bound(bound)
    torch.cuda.synchronize()
    assert torch.equal(got[0], plan.losses) and torch.equal(got[2], plan.g_poses)
    assert all(torch.equal(a, b) for a, b in zip(got[1] + got[3], plan.g_x + plan.g_logits))
    assert float(got[0].abs().sum()) > 0
