#!/usr/bin/env python
"""bench.py -- view-synthesis loss forward+backward throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic frame snippets: K_s^-1 / P tables, the
resize_area pyramids of target + source images, and the fused multi-scale loss forward+backward (3 launches).
Workload at every N: BASELINE.json configs[1] per GPU -- B=32, 128x416, 4 scales, 2 source views, fp32,
explainability mask on (weak scaling: the path shards over the batch, no data-path collective).

Unit: Mpix/s, where one "pix" is one PIXEL-VIEW (one target pixel x one scale x one source view;
SURVEY.md 8d): a step processes B*H*W*1.328125*V of them.

JSON keys beyond the base contract:
  roofline      the fused loss kernel, timed in situ with CUDA events recorded by the library immediately
                around its launch in every timed step; achieved = algorithmic bytes / mean duration, against
                MEASURED_PEAKS.json hbm_gbs (fallback 6650 GB/s, B200_PROFILING.md).
  cpu_baseline  the CPU oracle (op-for-op torch-CPU restatement of the reference, autograd backward) on this
                box's host cores, on a bounded sample of the same workload.
  e2e           the same step through the public API with HOST buffers (ops.HostPipeline): pinned H2D of every
                input, the step, D2H of the losses and every gradient, all inside the timed region; copies and
                kernels of neighbouring steps overlap (3 streams, 2 buffer sets).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'view-synthesis loss fwd+bwd throughput (pixel-views/s)'
UNIT = 'Mpix/s'
WORKLOAD = dict(B=32, H=128, W=416, S=4, V=2)
PYR = sum(0.25 ** s for s in range(WORKLOAD['S']))  # 1.328125


def pixel_views(B):
    return B * WORKLOAD['H'] * WORKLOAD['W'] * PYR * WORKLOAD['V']


def fused_kernel_bytes(B):
    """Algorithmic (compulsory) HBM bytes of ONE launch of loss_fused_kernel ('bytes'): per target
    pixel read x 4 + target 12, write g_x 4; per view read gathered source 12 + logits 8, write g_logits 8."""
    V = WORKLOAD['V']
    return B * WORKLOAD['H'] * WORKLOAD['W'] * PYR * (20 + 28 * V)


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def recorded_traffic():
    """dram bytes per launch of the fused kernel from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            t = json.load(fh)
        return t.get('loss_fused_kernel_cfg2_dram_bytes_per_launch')
    except Exception:
        return None


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, uuid):
        self.rows, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', uuid, '--query-gpu=' + self.FIELDS, '--format=csv,noheader,nounits', '-lms', '50'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.12)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for t, line in self.rows:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                clk, mx, pw = float(parts[0]), float(parts[1]), float(parts[2])
            except ValueError:
                continue
            smax = mx
            if t0 - 0.05 <= t <= t1 + 0.05:
                sm.append(clk)
                power.append(pw)
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons),
                'samples': len(sm), 'power_w_max': max(power) if power else None}


# ------------------------------------------------------------------------------------------ CPU legs
def oracle_step_fn(B, seed=4321):
    import torch
    from oracle import vsl_oracle as O
    from tf_depth_estimation_b200 import synth
    d = synth.make_snippets(B, WORKLOAD['H'], WORKLOAD['W'], S=WORKLOAD['S'], V=WORKLOAD['V'], seed=seed)
    flags = O.LossFlags()

    def step():
        xs = [x.clone().requires_grad_() for x in d['disp_pyr']]
        ps = d['poses'].clone().requires_grad_()
        lg = [l.clone().requires_grad_() for l in d['logits_pyr']]
        r = O.view_synthesis_loss(d['tgt'], d['srcs'], xs, ps, d['K_pyr'], lg, None, flags)
        sum(r).backward()
        return float(r[0])
    return step


def cpu_baseline(B_sample, reps, budget_s=25.0):
    """Times the CPU oracle (kind 'port': TF is not installable) on B_sample snippets of the
    workload with every host core."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = oracle_step_fn(B_sample)
    step()
    times, t_begin = [], time.time()
    for _ in range(reps):
        t = time.time()
        step()
        times.append(time.time() - t)
        if time.time() - t_begin > budget_s:
            break
    dt = statistics.median(times)
    return {'value': pixel_views(B_sample) / dt / 1e6, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'B=%d of the %d snippets per step (128x416, 4 scales, 2 views, fwd+autograd bwd), median of %d '
                      'runs, %.3f s each' % (B_sample, WORKLOAD['B'], len(times), dt)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  TensorFlow 1.x cannot run here, so
    this is the oracle port (an op-for-op restatement validated against the reference's own source, see
    oracle/); each step is a bounded sample (B=8 of the 32 snippets)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = 8
    step = oracle_step_fn(Bs)
    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = (time.time() - t0) / args.steps
    val = pixel_views(Bs) / dt / 1e6
    args.out.emit(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': dict(workload='cfg2: view-synthesis loss fwd+bwd, B=32 128x416 4 scales 2 views, exp mask '
                                '(each reference step = B=8 sample of it)', **WORKLOAD),
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': 'B=8 of 32 snippets per step, torch-CPU oracle, all host threads'},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from tf_depth_estimation_b200 import _lib, ops, synth
    from tf_depth_estimation_b200 import dist as vdist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the hot path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    _lib.load()

    B, H, W, S, V = (WORKLOAD[k] for k in 'BHWSV')
    flags = ops.LossFlags()
    # weak scaling: every rank holds B snippets of a global batch of B * world
    plan = ops.ViewSynthesisPlan(B, H, W, V, flags, _lib.MASK_EXP, dev, loss_scale=vdist.local_loss_scale(B, B * world))

    # rotating input sets so that consecutive steps never find their inputs in the 126 MB L2
    NSETS = 6
    host = synth.make_snippets(B, H, W, S=S, V=V, seed=1234 + rank)

    def to_dev(d, roll):
        r = lambda t: torch.roll(t, roll, dims=0).to(dev).contiguous()
        return dict(tgt=r(d['tgt']), srcs=[r(s) for s in d['srcs']], xs=[r(x) for x in d['disp_pyr']],
                    poses=r(d['poses']), Kp=r(d['K_pyr']), lgs=[r(l) for l in d['logits_pyr']])
    sets = [to_dev(host, i) for i in range(NSETS)]
    bound = [plan.bind(s['tgt'], s['srcs'], s['xs'], s['poses'], s['Kp'], s['lgs']) for s in sets]
    set_bytes = sum(t.numel() * 4 for t in [sets[0]['tgt']] + sets[0]['srcs'] + sets[0]['xs'] + sets[0]['lgs'])
    stream = torch.cuda.current_stream().cuda_stream

    K, Wm = args.steps, args.warmup
    # the fused kernel is timed in situ on every 4th timed step: an event between two launches keeps the second
    # from starting under the first one's tail (programmatic dependent launch), which the other steps do
    timed = [i for i in range(K) if i % 4 == 0]
    begins = [torch.cuda.Event(enable_timing=True) for _ in timed]
    ends = [torch.cuda.Event(enable_timing=True) for _ in timed]
    for e in begins + ends:
        e.record()  # materialise the cudaEvent_t handles
    torch.cuda.synchronize()

    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    sampler = ClockSampler(uuid if uuid.startswith('GPU-') else 'GPU-' + uuid)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(Wm):
        plan.run_bound(bound[i % NSETS], stream)
    barrier()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        if i % 4 == 0:
            plan.set_profile_events(begins[i // 4].cuda_event, ends[i // 4].cuda_event)
        else:
            plan.set_profile_events(None, None)
        plan.run_bound(bound[(Wm + i) % NSETS], stream)
    ev1.record()
    barrier()
    plan.set_profile_events(None, None)
    ms = ev0.elapsed_time(ev1)
    kern_ms = [b.elapsed_time(e) for b, e in zip(begins, ends)]
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    losses = vdist.reduce_losses(plan.losses, B, B * world).cpu().tolist()  # 12-byte all-reduce, outside the timed region

    # ---- e2e: host buffers in, host results out, through the public API (ops.HostPipeline): every step copies
    # all its inputs from pinned host memory and all its losses + gradients back; H2D / kernels / D2H of
    # neighbouring steps overlap on three streams
    pipe = ops.HostPipeline(B, H, W, V, flags, _lib.MASK_EXP, dev, loss_scale=vdist.local_loss_scale(B, B * world))
    h_in = pipe.host_inputs()          # pinned host tensors carved from one arena: a step's inputs move as ONE copy
    if os.environ.get('VSL_E2E_SEPARATE'):   # experiment switch: one pinned tensor and one copy per input
        h_in = dict(tgt=host['tgt'].pin_memory(), srcs=[t.pin_memory() for t in host['srcs']],
                    xs=[t.pin_memory() for t in host['disp_pyr']], poses=host['poses'].pin_memory(),
                    Kp=host['K_pyr'].pin_memory(), lgs=[t.pin_memory() for t in host['logits_pyr']])
    if h_in.get('_arena') is not None:
        h_in['tgt'].copy_(host['tgt']); h_in['poses'].copy_(host['poses']); h_in['Kp'].copy_(host['K_pyr'])
        for dst, src in zip(h_in['srcs'] + h_in['xs'] + h_in['lgs'], host['srcs'] + host['disp_pyr'] + host['logits_pyr']):
            dst.copy_(src)
    h2d, d2h = pipe.bytes_per_step()
    Ke = max(6, min(K, 150))
    for _ in range(4):
        slot = pipe.submit(h_in)
    e2e_losses = pipe.result(slot)[0].tolist()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(pipe.s_in)
    for _ in range(Ke):
        slot = pipe.submit(h_in)
    e1.record(pipe.s_out)
    pipe.result(slot)
    barrier()
    t_wall1 = time.time()
    te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item()) / Ke
    clocks = sampler.stop(t_wall0, t_wall1)

    if rank == 0:
        peak, peak_src = measured_peak()
        kmean = statistics.mean(kern_ms)
        achieved = fused_kernel_bytes(B) / (kmean * 1e-3) / 1e9
        out = {
            'metric': METRIC, 'value': pixel_views(B) * world * K / (ms_max * 1e-3) / 1e6, 'unit': UNIT,
            'n_gpus': world, 'steps': K, 'warmup': Wm, 'ms_per_step': ms_max / K, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(workload='cfg2: view-synthesis loss fwd+bwd (pyramids + fused multi-scale loss), per GPU '
                                    'B=32 128x416 4 scales 2 source views, explainability mask, euler poses',
                           l2='%d rotating input sets of %.0f MB each (> 126 MB L2 between reuses)' % (NSETS, set_bytes / 1e6),
                           pix='1 pix = 1 pixel-view = target pixel x scale x view; %.0f per step per GPU' % pixel_views(B),
                           parallelism='dp%d (batch shards, no data-path collective)' % world, **WORKLOAD),
            'roofline': {'bound': 'hbm', 'kernel': 'loss_fused_kernel<2>', 'achieved': achieved, 'peak': peak,
                         'unit': 'GB/s', 'frac': achieved / peak, 'traffic': recorded_traffic(),
                         'algorithmic_bytes_per_launch': fused_kernel_bytes(B), 'kernel_ms_mean': kmean,
                         'kernel_ms_min': min(kern_ms), 'kernel_share_of_step': kmean / (ms_max / K), 'kernel_timed_steps': len(kern_ms),
                         'peak_source': peak_src},
            'e2e': {'value': pixel_views(B) * world / (e2e_ms * 1e-3) / 1e6, 'unit': UNIT, 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': Ke,
                    'how': 'ops.HostPipeline: pinned host inputs -> H2D (one copy) -> 3 launches -> D2H of losses and all '
                           'gradients (one copy), double-buffered over 3 streams'},
            'gpu_launches': 3 * K, 'launches_per_step': 3, 'clocks': clocks,
            'losses': {'pixel': losses[0], 'smooth': losses[1], 'exp': losses[2]},
        }
        if world == 1 and not args.no_cpu_baseline:
            out['cpu_baseline'] = cpu_baseline(B_sample=8, reps=12)
        args.out.emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


class _QuietStdout(object):
    """stdout of the process (C level too: NCCL prints its version banner there) goes to stderr until emit(), so
    that the ONE JSON line is all the driver reads on stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line)
        sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=600)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.out = _QuietStdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
