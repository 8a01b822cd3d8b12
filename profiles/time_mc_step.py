"""The fused data-parallel optimiser step, peer-to-peer form against the NVSwitch-multicast (NVLS) form, on N GPUs:
correctness of the multicast form (against "gather, add in rank order, plain Adam" and replica identity) and times.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 profiles/time_mc_step.py
CUDA events on the compute stream, max over ranks.  33.2 M parameters ~ DispNet + PoseExpNet."""
import os, sys, json, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_depth_estimation_b200 import ops
from tf_depth_estimation_b200 import dist as vdist

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
NPAR = int(os.environ.get('NPAR', 33_200_000))


def timeit(fn, n=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1000 / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


out = {'world': world, 'params': NPAR, 'multicast_available': vdist.MulticastArena.available(dev)}
peer = vdist.PeerDataParallelAdam([(NPAR,)], dev, lr=2e-4)
out['peer_fused_step_us'] = timeit(peer.step)
if out['multicast_available']:
    mc = vdist.PeerDataParallelAdam([(NPAR,)], dev, lr=2e-4, multicast=True)
    ordp, ordm, ordv = (torch.zeros(NPAR, device=dev) for _ in range(3))
    gen = torch.Generator(device=dev).manual_seed(5 + rank)
    for t in range(3):
        g = torch.randn(NPAR, device=dev, generator=gen)
        mc.grad_flat[:NPAR].copy_(g)
        allg = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(allg, g)
        acc = torch.zeros_like(g)
        for a in allg:
            acc += a
        del allg
        ops.adam_step(ordp, acc, ordm, ordv, t + 1, lr=2e-4)
        mc.step()
    torch.cuda.synchronize()
    mc.check_peers()
    d = (ordp - mc.param_flat[:NPAR]).abs()
    out['mc_vs_rank_ordered_sum_max_abs_diff'] = float(d.max())
    out['mc_vs_rank_ordered_sum_frac_gt_1e-7'] = float((d > 1e-7).float().mean())
    out['mc_param_absmax'] = float(mc.param_flat.abs().max())
    chk = mc.param_flat.double().sum().reshape(1).clone()
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    out['mc_replicas_identical'] = bool(all(float(x) == float(lst[0]) for x in lst))
    out['mc_fused_step_us'] = timeit(mc.step)
    if os.environ.get('VSL_MC_SWEEP'):          # timing-experiment build (-DVSL_DP_TIMING_EXPERIMENTS) only
        for per_sm in os.environ['VSL_MC_SWEEP'].split(','):
            os.environ['VSL_MC_BLOCKS_PER_SM'] = per_sm
            out['mc_fused_step_us_blocks_per_sm_%s' % per_sm] = timeit(mc.step)
        os.environ.pop('VSL_MC_BLOCKS_PER_SM')
    mc.check_peers()
    mc.close()
peer.check_peers()
peer.close()
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
