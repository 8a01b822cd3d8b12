"""Data-parallel optimiser step (dist.DataParallelAdam) on N GPUs: bucketed NCCL all-reduce of a 130 MB gradient
arena + the flat Adam kernel, alone and behind the cfg2 loss step.  Run under torchrun (or plain python for N=1):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/time_dp_step.py
Times are CUDA events on the compute stream, max over ranks.  33.2 M parameters ~ DispNet + PoseExpNet."""
import os, sys, json, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_depth_estimation_b200 import ops, synth, _lib
from tf_depth_estimation_b200 import dist as vdist

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)

NPAR = 33_200_000
B, H, W, S, V = 32, 128, 416, 4, 2


def timeit(fn, n=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1000 / n], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

out = {'world': world, 'params': NPAR}
host = synth.make_snippets(B, H, W, S=S, V=V, seed=1234 + rank)
plan = ops.ViewSynthesisPlan(B, H, W, V, ops.LossFlags(), _lib.MASK_EXP, dev, loss_scale=vdist.local_loss_scale(B, B * world))
c = lambda t: t.to(dev).contiguous()
bound = plan.bind(c(host['tgt']), [c(s) for s in host['srcs']], [c(x) for x in host['disp_pyr']], c(host['poses']),
                  c(host['K_pyr']), [c(l) for l in host['logits_pyr']])
out['loss_step_us'] = timeit(lambda: plan.run_bound(bound), n=200)
for mb in (32, 128):
    dp = vdist.DataParallelAdam([(NPAR,)], dev, lr=2e-4, bucket_bytes=mb << 20)
    dp.grad_flat.normal_()
    out['allreduce_adam_us_bucket%dMB' % mb] = timeit(dp.step)
    if mb == 32:
        def adam_only():
            dp.t += 1
            for lo, hi in dp.buckets:
                dp.adam_fn(dp.param_flat[lo:hi], dp.grad_flat[lo:hi], dp.m_flat[lo:hi], dp.v_flat[lo:hi], dp.t)
        out['adam_only_us'] = timeit(adam_only)
        out['adam_GBps'] = 28.0 * NPAR / out['adam_only_us'] / 1e3
        if world > 1:
            out['allreduce_only_us'] = timeit(lambda: dist.all_reduce(dp.grad_flat))
        out['loss_plus_dp_step_us'] = timeit(lambda: (plan.run_bound(bound), dp.step()))
    del dp
# the same step as ONE kernel over NVLink peer memory (reduce-scatter + Adam + all-gather), checked against NCCL + Adam
ref = vdist.DataParallelAdam([(NPAR,)], dev, lr=2e-4, bucket_bytes=1 << 30)
peer = vdist.PeerDataParallelAdam([(NPAR,)], dev, lr=2e-4)
# second reference, independent of NCCL's summation order: gather every rank's gradient, add them in rank order
# (the order the fused kernel uses), plain Adam kernel => must be BIT-identical to the peer path
ordp, ordm, ordv = (torch.zeros(NPAR, device=dev) for _ in range(3))
gen = torch.Generator(device=dev).manual_seed(5 + rank)
for t in range(3):
    g = torch.randn(NPAR, device=dev, generator=gen)
    ref.grad_flat[:NPAR].copy_(g); peer.grad_flat[:NPAR].copy_(g)
    if world > 1:
        allg = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(allg, g)
        acc = torch.zeros_like(g)
        for a in allg:
            acc += a
        del allg
    else:
        acc = g.clone()
    ops.adam_step(ordp, acc, ordm, ordv, t + 1, lr=2e-4)
    ref.step(); peer.step()
torch.cuda.synchronize()
peer.check_peers()
out['peer_vs_rank_ordered_sum_max_abs_diff'] = float((ordp - peer.param_flat[:NPAR]).abs().max())
d = (ref.param_flat - peer.param_flat).abs()
out['peer_vs_nccl_max_abs_diff'] = float(d.max())
out['peer_vs_nccl_frac_elements_differing_gt_1e-7'] = float((d > 1e-7).float().mean())
out['peer_param_absmax'] = float(peer.param_flat.abs().max())
del d, ordp, ordm, ordv, acc
if world > 1:
    chk = peer.param_flat.double().sum().reshape(1).clone()
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    out['replicas_identical'] = bool(all(float(x) == float(lst[0]) for x in lst))
out['peer_fused_step_us'] = timeit(peer.step)
for dbg in ('1', '2', '3'):
    os.environ['VSL_DP_DEBUG'] = dbg
    out['peer_step_us_debug%s' % dbg] = timeit(peer.step)
os.environ.pop('VSL_DP_DEBUG')
if world > 1:
    st = torch.cuda.current_stream().cuda_stream
    out['peer_two_barriers_us'] = timeit(lambda: (peer._barrier(st), peer._barrier(st)))
out['loss_plus_peer_step_us'] = timeit(lambda: (plan.run_bound(bound), peer.step()))
peer.check_peers()
peer.close()
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
