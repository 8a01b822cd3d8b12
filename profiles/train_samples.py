"""End-to-end TRAINING samples/s at N GPUs (BASELINE.json: "train samples/s at 1/2/4/8 B200", configs[2]:
train.py DispNet + PoseExpNet with explainability mask, 128x416, global batch 256 data-parallel).

A measurement harness, not a product component: the networks are OUT of the hot path and run "through the
framework's own GPU convolutions" (north_star) -- plain torch.nn modules shaped like nets.py:16-150 (disp_net,
pose_exp_net; conv + batch-norm + ReLU, cuDNN fp32 with TF32 allowed), random init, synthetic snippets.  What this
repository contributes to the step is everything after the networks:
  fused view-synthesis loss forward+backward (libvsl, disparity head fused: x_is_logit, DISP_SCALING = 10,
  MIN_DISP = 0.01, nets.py:8-9)  ->  network backward (torch)  ->  dist.PeerDataParallelAdam.step(): gradient
  sum + tf.train.AdamOptimizer + parameter broadcast as ONE kernel over NVLink peer memory.
BatchNorm statistics stay per rank.

  python profiles/train_samples.py [--global-batch 256] [--steps 10] [--optim peer|nccl]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
      profiles/train_samples.py
Prints one JSON line (rank 0): samples/s = global batch / max-over-ranks step time (CUDA events).
"""
import argparse, json, os, sys
import torch
import torch.nn as nn
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_depth_estimation_b200 import ops, synth
from tf_depth_estimation_b200 import dist as vdist


def cbr(i, o, k, s):
    return nn.Sequential(nn.Conv2d(i, o, k, s, k // 2, bias=False), nn.BatchNorm2d(o), nn.ReLU(inplace=True))


def up(i, o, k=3):
    return nn.Sequential(nn.ConvTranspose2d(i, o, k, 2, k // 2, output_padding=1, bias=False), nn.BatchNorm2d(o),
                         nn.ReLU(inplace=True))


class DispNet(nn.Module):       # nets.py:86-150
    def __init__(self):
        super().__init__()
        ch = [32, 64, 128, 256, 512, 512, 512]
        ks = [7, 5, 3, 3, 3, 3, 3]
        self.enc = nn.ModuleList()
        i = 3
        for o, k in zip(ch, ks):
            self.enc.append(nn.Sequential(cbr(i, o, k, 2), cbr(o, o, k, 1)))
            i = o
        self.up7, self.i7 = up(512, 512), cbr(1024, 512, 3, 1)
        self.up6, self.i6 = up(512, 512), cbr(1024, 512, 3, 1)
        self.up5, self.i5 = up(512, 256), cbr(512, 256, 3, 1)
        self.up4, self.i4 = up(256, 128), cbr(256, 128, 3, 1)
        self.d4 = nn.Conv2d(128, 1, 3, 1, 1)
        self.up3, self.i3 = up(128, 64), cbr(64 + 64 + 1, 64, 3, 1)
        self.d3 = nn.Conv2d(64, 1, 3, 1, 1)
        self.up2, self.i2 = up(64, 32), cbr(32 + 32 + 1, 32, 3, 1)
        self.d2 = nn.Conv2d(32, 1, 3, 1, 1)
        self.up1, self.i1 = up(32, 16), cbr(16 + 1, 16, 3, 1)
        self.d1 = nn.Conv2d(16, 1, 3, 1, 1)

    def forward(self, x):
        f = []
        for e in self.enc:
            x = e(x)
            f.append(x)
        rs = lambda a, b: a if a.shape[2:] == b.shape[2:] else nn.functional.interpolate(a, size=b.shape[2:])
        x = self.i7(torch.cat([rs(self.up7(f[6]), f[5]), f[5]], 1))
        x = self.i6(torch.cat([rs(self.up6(x), f[4]), f[4]], 1))
        x = self.i5(torch.cat([rs(self.up5(x), f[3]), f[3]], 1))
        x = self.i4(torch.cat([self.up4(x), f[2]], 1))
        d4 = self.d4(x)
        bl = lambda a: nn.functional.interpolate(a, scale_factor=2, mode='bilinear', align_corners=False)
        x = self.i3(torch.cat([self.up3(x), f[1], bl(d4)], 1))
        d3 = self.d3(x)
        x = self.i2(torch.cat([self.up2(x), f[0], bl(d3)], 1))
        d2 = self.d2(x)
        x = self.i1(torch.cat([self.up1(x), bl(d2)], 1))
        d1 = self.d1(x)
        return [d1, d2, d3, d4]          # PRE-activation: the loss kernel applies 10 * sigmoid + 0.01 on load


class PoseExpNet(nn.Module):    # nets.py:16-84
    def __init__(self, V=2):
        super().__init__()
        self.V = V
        ch, ks = [16, 32, 64, 128, 256], [7, 5, 3, 3, 3]
        self.enc = nn.ModuleList()
        i = 3 * (V + 1)
        for o, k in zip(ch, ks):
            self.enc.append(cbr(i, o, k, 2))
            i = o
        self.p6, self.p7, self.pp = cbr(256, 256, 3, 2), cbr(256, 256, 3, 2), nn.Conv2d(256, 6 * V, 1)
        self.u5, self.u4, self.u3, self.u2, self.u1 = up(256, 256), up(256, 128), up(128, 64), up(64, 32, 5), up(32, 16, 7)
        self.m4, self.m3 = nn.Conv2d(128, 2 * V, 3, 1, 1), nn.Conv2d(64, 2 * V, 3, 1, 1)
        self.m2, self.m1 = nn.Conv2d(32, 2 * V, 5, 1, 2), nn.Conv2d(16, 2 * V, 7, 1, 3)

    def forward(self, x):
        for e in self.enc:
            x = e(x)
        pose = 0.01 * self.pp(self.p7(self.p6(x))).mean((2, 3)).reshape(-1, self.V, 6)
        u4 = self.u4(self.u5(x))
        u3 = self.u3(u4)
        u2 = self.u2(u3)
        u1 = self.u1(u2)
        return pose, [self.m1(u1), self.m2(u2), self.m3(u3), self.m4(u4)]


def measure(rank, world, dev, global_batch=256, steps=10, warmup=3, optim='peer', graph=1, timeout_s=120.0):
    """One measurement on an initialised process group (or a single process) -> dict (identical on every rank except
    that only rank 0's is meant to be printed).  The caller owns the process group."""
    H, W, S, V = 128, 416, 4, 2
    # the framework's own convolutions (outside the path): let cuDNN pick its algorithm per layer shape once, in the warm-up
    torch.backends.cudnn.benchmark = os.environ.get('VSL_CUDNN_BENCHMARK', '1') != '0'
    lo, hi = vdist.shard_range(global_batch, rank, world)
    B = hi - lo
    torch.manual_seed(0)        # identical initial weights on every rank
    disp_net, pose_net = DispNet().to(dev).to(memory_format=torch.channels_last), \
        PoseExpNet(V).to(dev).to(memory_format=torch.channels_last)
    params = [p for p in list(disp_net.parameters()) + list(pose_net.parameters())]
    if optim == 'peer':
        dp = vdist.PeerDataParallelAdam([p.shape for p in params], dev, lr=2e-4, beta1=0.9, timeout_s=timeout_s,
                                        multicast=os.environ.get('VSL_MULTICAST', 'auto') if os.environ.get('VSL_MULTICAST', 'auto') != '0' else False)
    else:
        dp = vdist.DataParallelAdam([p.shape for p in params], dev, lr=2e-4, beta1=0.9, bucket_bytes=1 << 30)
    for p, pv, gv in zip(params, dp.params, dp.grads):
        pv.copy_(p.data)
        p.data = pv             # the networks compute on / into the flat arenas: no copies around the optimiser
        p.grad = gv
    d = synth.make_snippets(min(B, 8), H, W, S=S, V=V, seed=100 + rank)
    rep = lambda t: t.repeat((B + t.shape[0] - 1) // t.shape[0], *([1] * (t.dim() - 1)))[:B].to(dev).contiguous()
    tgt, srcs, K_pyr = rep(d['tgt']), [rep(s) for s in d['srcs']], rep(d['K_pyr'])
    flags = ops.LossFlags(num_scales=S, x_is_logit=True, disp_scaling=10.0, min_disp=0.01)
    nchw = lambda t: t.permute(0, 3, 1, 2)      # NHWC storage seen as NCHW / channels_last: no copy
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
    scale = vdist.local_loss_scale(B, global_batch)
    loss_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def fwd_bwd(time_loss=False):
        dp.grad_flat.zero_()
        x_pyr = [nhwc(x) for x in disp_net(nchw(tgt))]
        pose, masks = pose_net(nchw(torch.cat([tgt] + srcs, dim=3)))
        lg = [nhwc(m) for m in masks]
        pose = pose.contiguous()
        if time_loss:
            loss_ev[0].record()
        total, losses = ops.view_synthesis_loss(tgt, srcs, x_pyr, pose, K_pyr, logits_pyr=lg, flags=flags, loss_scale=scale)
        if time_loss:
            loss_ev[1].record()
        total.backward()        # the rank's B_local / B_global share is folded into the kernel's gradients
        return losses

    def step():
        losses = fwd_bwd()
        dp.step()
        return losses

    for _ in range(warmup):
        losses = step()
    # this repository's share of the step, timed on their own (eager, before any graph capture)
    fwd_bwd(time_loss=True)
    oe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    oe[0].record()
    for _ in range(5):
        dp.step()
    oe[1].record()
    torch.cuda.synchronize()
    loss_us, optim_us = loss_ev[0].elapsed_time(loss_ev[1]) * 1e3, oe[0].elapsed_time(oe[1]) * 1e3 / 5
    graphed = False
    if graph:
        # networks + loss, forward and backward, as ONE CUDA graph (the step is launch-bound at 32 samples per GPU);
        # the optimiser step is launched right behind every replay (its epoch / step count live on the device)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    fwd_bwd()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_losses = fwd_bwd()

            def step():     # noqa: F811
                g.replay()
                dp.step()
                return static_losses
            graphed = True
            for _ in range(2):
                losses = step()
        except Exception as e:      # a measurement harness: fall back to eager launches and say so
            sys.stderr.write('CUDA graph capture failed, running eagerly: %r\n' % (e,))
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        losses = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps, optim_us, loss_us], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if hasattr(dp, 'check_peers'):
        dp.check_peers()
    ok = bool(torch.isfinite(losses).all())
    out = {'metric': 'train samples/s (DispNet + PoseExpNet torch/cuDNN fp32-TF32, fused loss, %s optimiser step)' % optim,
           'value': global_batch / (float(t[0]) * 1e-3), 'unit': 'samples/s', 'n_gpus': world,
           'global_batch': global_batch, 'per_gpu_batch': B, 'ms_per_step': float(t[0]), 'steps': steps,
           'optim_us': float(t[1]), 'loss_us': float(t[2]), 'optimiser': optim,
           'params': int(sum(p.numel() for p in params)), 'scaling': 'strong', 'cuda_graph': graphed, 'losses_finite': ok,
           'losses': [float(v) for v in losses]}
    if hasattr(dp, 'close'):
        dp.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--global-batch', type=int, default=256)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--optim', default='peer', choices=['peer', 'nccl'])
    ap.add_argument('--graph', type=int, default=1, help='replay networks + loss (forward and backward) as one CUDA graph')
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    out = measure(rank, world, dev, a.global_batch, a.steps, a.warmup, a.optim, a.graph)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
