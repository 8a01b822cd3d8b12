#!/usr/bin/env python
"""bench.py -- view-synthesis loss forward+backward throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg4|cfg5]

A *step* is one pass of the hot path over one batch of synthetic frame snippets: K_s^-1 / P tables, the
resize_area pyramids of target + source images, and the fused multi-scale loss forward+backward (3 launches).
Default workload (the BENCH / SCALE line): BASELINE.json configs[1] per GPU -- B=32, 128x416, 4 scales, 2 source
views, fp32, explainability mask on (weak scaling: the path shards over the batch, no data-path collective).
--config selects the other BASELINE shapes: cfg3 (global batch 256 split over the ranks, strong scaling), cfg4
(DeMoN pairs 192x256, B=64, one view per direction, angle-axis poses), cfg5 (480x640, B=64, the HBM-stress shape).

Unit: Mpix/s, where one "pix" is one PIXEL-VIEW (one target pixel x one scale x one source view;
SURVEY.md 8d): a step processes B*H*W*(sum_s 4^-s)*V of them.

JSON keys beyond the base contract:
  roofline      the fused loss kernel, timed in situ with CUDA events recorded by the library immediately
                around its launch in every 4th timed step; achieved = algorithmic bytes / mean duration, against
                MEASURED_PEAKS.json hbm_gbs (fallback 6650 GB/s, B200_PROFILING.md).
  cpu_baseline  the CPU oracle (op-for-op torch-CPU restatement of the reference, autograd backward) on this
                box's host cores, on a bounded sample of the same workload.
  e2e           the same step through the public API with HOST buffers (ops.HostPipeline): the frames as the
                reference's loader holds them (uint8, converted on load exactly as imageselect_Dataloader.py:93 does),
                network outputs as float32; pinned H2D of every input, the step, D2H of the losses and every gradient,
                all inside the timed region; copies and kernels of neighbouring steps overlap (3 streams, 2 slots).
  flow          (--config cfg4 only) the DeMoN-pair family's own loss loop (train_optflow_combine.py:138-240) as the
                fused flow-and-depth step vsl_flow_loss_fwd_bwd: device-resident step time, per rank.
  other_configs (default config, one GPU) the same device-resident step at the other BASELINE shapes -- cfg3 (B=256 on this
                GPU), cfg4, cfg5 -- each with its step time, throughput and the fused kernel's in-situ roofline fraction.
  train         (default config only) BASELINE's second metric, end-to-end training samples/s at N GPUs:
                configs[2] (DispNet + PoseExpNet, global batch 256 split over the ranks) with this repository's
                fused loss and its fused reduce-scatter + Adam + all-gather optimiser step over NVLink peer memory
                (profiles/train_samples.py; the networks are torch/cuDNN, outside the hot path).
"""
import argparse
import importlib.util
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
DEMON = dict(pose_format='angleaxis', smooth_on_inverse=True, depth_is_inverse=True, pixel_scale_norm=False)
CONFIGS = {
    'cfg2': dict(B=32, H=128, W=416, S=4, V=2, flags={}, sets=6, cpu_B=32, scaling='weak',
                 text='cfg2: view-synthesis loss fwd+bwd (pyramids + fused multi-scale loss), per GPU B=32 128x416 4 '
                      'scales 2 source views, explainability mask, euler poses'),
    'cfg3': dict(B=256, H=128, W=416, S=4, V=2, flags={}, sets=3, cpu_B=32, scaling='strong',
                 text='cfg3: the loss step of train.py at global batch 256 (split over the ranks), 128x416 4 scales 2 '
                      'source views, explainability mask, euler poses'),
    'cfg4': dict(B=64, H=192, W=256, S=4, V=1, flags=DEMON, sets=4, cpu_B=64, scaling='weak',
                 text='cfg4: DeMoN pairs 192x256, per GPU B=64, one source view per direction, angle-axis poses, '
                      'smoothness on 1/depth, explainability mask'),
    'cfg5': dict(B=64, H=480, W=640, S=4, V=2, flags={}, sets=2, cpu_B=1, scaling='weak',
                 text='cfg5: 480x640 multi-scale refinement shape, per GPU B=64, 4 scales 2 source views, '
                      'explainability mask (HBM stress: 2.9 GB workspace)'),
}


def pyr(S):
    return sum(0.25 ** s for s in range(S))


def pixel_views(c, B):
    return B * c['H'] * c['W'] * pyr(c['S']) * c['V']


def fused_kernel_bytes(c, B):
    """Algorithmic (compulsory) HBM bytes of ONE launch of the fused loss kernel: per target
    pixel read x 4 + target 12, write g_x 4; per view read gathered source 12 + logits 8, write g_logits 8."""
    return B * c['H'] * c['W'] * pyr(c['S']) * (20 + 28 * c['V'])


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def recorded_traffic(name):
    """dram bytes per launch of the fused kernel from the committed ncu --set full capture of this config, if any."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            t = json.load(fh)
        return t.get('loss_fused_kernel_%s_dram_bytes_per_launch' % name)
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
def oracle_step_fn(c, B, seed=4321):
    import torch
    from oracle import vsl_oracle as O
    from tf_depth_estimation_b200 import synth
    d = synth.make_snippets(B, c['H'], c['W'], S=c['S'], V=c['V'], seed=seed)
    flags = O.LossFlags(num_scales=c['S'], **c['flags'])

    def step():
        xs = [x.clone().requires_grad_() for x in d['disp_pyr']]
        ps = d['poses'].clone().requires_grad_()
        lg = [l.clone().requires_grad_() for l in d['logits_pyr']]
        r = O.view_synthesis_loss(d['tgt'], d['srcs'], xs, ps, d['K_pyr'], lg, None, flags)
        sum(r).backward()
        return float(r[0])
    return step


def cpu_baseline(c, reps, budget_s=25.0):
    """Times the CPU oracle (kind 'port': TF is not installable) on cpu_B snippets of the workload
    with every host core."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = c['cpu_B']
    step = oracle_step_fn(c, Bs)
    step()
    times, t_begin = [], time.time()
    for _ in range(reps):
        t = time.time()
        step()
        times.append(time.time() - t)
        if time.time() - t_begin > budget_s:
            break
    dt = statistics.median(times)
    return {'value': pixel_views(c, Bs) / dt / 1e6, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'B=%d snippets of the step (%dx%d, %d scales, %d views, fwd+autograd bwd), median of %d '
                      'runs, %.3f s each' % (Bs, c['H'], c['W'], c['S'], c['V'], len(times), dt)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  TensorFlow 1.x cannot run here, so
    this is the oracle port (an op-for-op restatement validated against the reference's own source, see
    oracle/); each step is cpu_B snippets of the batch: the whole per-GPU batch at cfg2 / cfg4 (0.3-1 s per step on
    the host cores), a bounded sample at cfg3 / cfg5."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    c = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = c['cpu_B']
    step = oracle_step_fn(c, Bs)
    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = (time.time() - t0) / args.steps
    val = pixel_views(c, Bs) / dt / 1e6
    args.out.emit(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': c['scaling'],
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': dict(workload=c['text'] + (' (each reference step = the whole per-GPU batch)' if Bs == c['B'] else
                                             ' (each reference step = B=%d sample of it)' % Bs),
                       B=c['B'], H=c['H'], W=c['W'], S=c['S'], V=c['V']),
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': 'B=%d snippets per step, torch-CPU oracle, all host threads' % Bs},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# ------------------------------------------------------------------------------------------ GPU arm
def train_leg(rank, world, dev):
    """BASELINE's second metric through profiles/train_samples.py; never lets a failure take the bench line down.
    Strong scaling as configs[2] words it (global batch 256 split over the ranks) and, beside it, the weak-scaling
    reading (32 snippets per GPU at every N, i.e. the per-GPU work of the 8-GPU run of configs[2])."""
    try:
        spec = importlib.util.spec_from_file_location('train_samples', os.path.join(ROOT, 'profiles', 'train_samples.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        r = mod.measure(rank, world, dev, global_batch=256, steps=6, warmup=3, optim='peer', graph=1, timeout_s=30.0)
        if 32 * world == 256:
            w = r
        else:
            w = mod.measure(rank, world, dev, global_batch=32 * world, steps=10, warmup=3, optim='peer', graph=1, timeout_s=30.0)
        weak = {'samples_per_s': w['value'], 'ms_per_step': w['ms_per_step'], 'global_batch': w['global_batch'],
                'per_gpu_batch': w['per_gpu_batch'], 'scaling': 'weak', 'optim_us': w['optim_us'], 'loss_us': w['loss_us']}
        return {'weak': weak, 'samples_per_s': r['value'], 'ms_per_step': r['ms_per_step'], 'optim_us': r['optim_us'], 'loss_us': r['loss_us'],
                'global_batch': r['global_batch'], 'per_gpu_batch': r['per_gpu_batch'], 'scaling': 'strong',
                'optimiser': 'fused reduce-scatter + Adam + all-gather over NVLink peer memory (dp_adam_kernel)',
                'cuda_graph': r['cuda_graph'], 'params': r['params'], 'losses_finite': r['losses_finite'],
                'workload': 'configs[2]: DispNet + PoseExpNet (torch/cuDNN fp32, TF32 allowed, cudnn.benchmark), explainability mask, 128x416'}
    except Exception as e:   # noqa: BLE001
        return {'error': '%s: %s' % (type(e).__name__, e)}


def flow_leg(dev, steps):
    """configs[3]'s own loss (train_optflow_combine.py:138-240) as the fused flow-and-depth step, through
    profiles/time_flow.py; never lets a failure take the bench line down."""
    try:
        spec = importlib.util.spec_from_file_location('time_flow', os.path.join(ROOT, 'profiles', 'time_flow.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        r = mod.measure(dev, steps=max(20, min(steps, 200)))
        peak, _ = measured_peak()     # whole step (4 launches) against the 52 compulsory bytes per pixel and scale
        r['roofline_frac_of_step'] = r['achieved_gbs'] / peak
        return r
    except Exception as e:   # noqa: BLE001
        return {'error': '%s: %s' % (type(e).__name__, e)}


def quick_config(name, dev, rank, steps=40, warmup=5, extra_flags=None, note=None):
    """One more BASELINE shape through the same device-resident step (rotating input sets, in-situ kernel events on
    every 4th step): the short form of the main measurement, for the `other_configs` object of the default line."""
    import torch
    from tf_depth_estimation_b200 import _lib, ops, synth
    try:
        import gc
        gc.collect()
        torch.cuda.empty_cache()      # the larger shapes allocate GBs: start from the driver's pool, not from cached fragments
        c = CONFIGS[name]
        B, H, W, S, V = (c[k] for k in 'BHWSV')
        flags = ops.LossFlags(num_scales=S, **dict(c['flags'], **(extra_flags or {})))
        plan = ops.ViewSynthesisPlan(B, H, W, V, flags, _lib.MASK_EXP, dev)
        host = synth.make_snippets(min(B, 16), H, W, S=S, V=V, seed=4321 + rank)
        rep = lambda t, k: torch.roll(t.repeat((B + t.shape[0] - 1) // t.shape[0], *([1] * (t.dim() - 1)))[:B], k, 0).to(dev).contiguous()
        bound = []
        for k in range(c['sets']):
            bound.append(plan.bind(rep(host['tgt'], k), [rep(x, k) for x in host['srcs']], [rep(x, k) for x in host['disp_pyr']],
                                   rep(host['poses'], k), rep(host['K_pyr'], k), [rep(x, k) for x in host['logits_pyr']]))
        stream = torch.cuda.current_stream().cuda_stream
        timed = list(range(0, steps, 4))
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in timed]
        for a, b in ev:
            a.record(); b.record()
        torch.cuda.synchronize()
        for i in range(warmup):
            plan.run_bound(bound[i % len(bound)], stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            if i % 4 == 0:
                plan.set_profile_events(ev[i // 4][0].cuda_event, ev[i // 4][1].cuda_event)
            else:
                plan.set_profile_events(None, None)
            plan.run_bound(bound[(warmup + i) % len(bound)], stream)
        e1.record()
        torch.cuda.synchronize()
        plan.set_profile_events(None, None)
        ms = e0.elapsed_time(e1) / steps
        kms = statistics.mean(a.elapsed_time(b) for a, b in ev)
        peak, _ = measured_peak()
        algo = fused_kernel_bytes(c, B)
        out = {'workload': c['text'], 'B': B, 'H': H, 'W': W, 'S': S, 'V': V, 'steps': steps, 'ms_per_step': ms,
               'value': pixel_views(c, B) / (ms * 1e-3) / 1e6, 'unit': UNIT, 'kernel_ms_mean': kms,
               'algorithmic_bytes_per_launch': algo, 'roofline_frac': algo / (kms * 1e-3) / 1e9 / peak,
               'losses_finite': bool(torch.isfinite(plan.losses).all())}
        if note:                   # a variant of the step: the roofline model of the plain fused kernel does not describe it
            out['workload'] = c['text'] + '; ' + note
            del out['algorithmic_bytes_per_launch'], out['roofline_frac']
        del plan, bound
        torch.cuda.empty_cache()
        return out
    except Exception as e:   # noqa: BLE001
        return {'error': '%s: %s' % (type(e).__name__, e)}


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
    to_u8 = lambda t: (t * 255.0).round().clamp(0, 255).to(dtype=torch.uint8)   # what a decoded frame is

    c = CONFIGS[args.config]
    H, W, S, V = (c[k] for k in 'HWSV')
    if c['scaling'] == 'strong':            # a fixed global batch, split over the ranks
        lo, hi = vdist.shard_range(c['B'], rank, world)
        B, B_global = hi - lo, c['B']
    else:                                    # every rank holds B snippets of a global batch of B * world
        B, B_global = c['B'], c['B'] * world
    flags = ops.LossFlags(num_scales=S, **c['flags'])
    scale = vdist.local_loss_scale(B, B_global)
    plan = ops.ViewSynthesisPlan(B, H, W, V, flags, _lib.MASK_EXP, dev, loss_scale=scale)

    # rotating input sets so that consecutive steps never find their inputs in the 126 MB L2
    NSETS = c['sets']
    host = synth.make_snippets(min(B, 32), H, W, S=S, V=V, seed=1234 + rank)
    if B > 32:
        host = {k: ([t.repeat(B // 32, *([1] * (t.dim() - 1))) for t in v] if isinstance(v, list) else
                    (v.repeat(B // 32, *([1] * (v.dim() - 1))) if hasattr(v, 'repeat') else v)) for k, v in host.items()}

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
    losses = vdist.reduce_losses(plan.losses, B, B_global).cpu().tolist()  # 12-byte all-reduce, outside the timed region

    # ---- the same device-resident step fed with the loader's uint8 frames (converted on load)
    flags8 = ops.LossFlags(num_scales=S, img_format='u8_255', **c['flags'])
    plan8 = ops.ViewSynthesisPlan(B, H, W, V, flags8, _lib.MASK_EXP, dev, loss_scale=scale)
    sets8 = [dict(s, tgt=to_u8(s['tgt']), srcs=[to_u8(x) for x in s['srcs']]) for s in sets]
    bound8 = [plan8.bind(s['tgt'], s['srcs'], s['xs'], s['poses'], s['Kp'], s['lgs']) for s in sets8]
    for i in range(Wm):
        plan8.run_bound(bound8[i % NSETS], stream)
    barrier()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    for i in range(K):
        plan8.run_bound(bound8[(Wm + i) % NSETS], stream)
    u1.record()
    barrier()
    t8 = torch.tensor([u0.elapsed_time(u1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t8, op=dist.ReduceOp.MAX)
    ms8 = float(t8.item())
    del sets8, bound8, plan8

    # ---- e2e: host buffers in, host results out, through the public API (ops.HostPipeline): every step copies
    # all its inputs from pinned host memory and all its losses + gradients back; H2D / kernels / D2H of
    # neighbouring steps overlap on three streams.  Frames travel as the loader's uint8.
    numa_cpus = vdist.bind_host_to_gpu(dev) if world > 1 else None   # pinned arenas on the GPU's own NUMA node
    pipe = ops.HostPipeline(B, H, W, V, flags8, _lib.MASK_EXP, dev, loss_scale=scale)
    h_in = pipe.host_inputs()          # pinned host tensors carved from one arena: a step's inputs move as ONE copy
    h_in['tgt'].copy_(to_u8(host['tgt'])); h_in['poses'].copy_(host['poses']); h_in['Kp'].copy_(host['K_pyr'])
    for dst, src in zip(h_in['srcs'], host['srcs']):
        dst.copy_(to_u8(src))
    for dst, src in zip(h_in['xs'] + h_in['lgs'], host['disp_pyr'] + host['logits_pyr']):
        dst.copy_(src)
    h2d, d2h = pipe.bytes_per_step()
    Ke = max(6, min(K, 150))
    for _ in range(4):
        slot = pipe.submit(h_in)
    pipe.result(slot)
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
    del pipe, h_in, sets, bound
    torch.cuda.empty_cache()

    train = None
    if args.config == 'cfg2' and not args.no_train:
        train = train_leg(rank, world, dev)

    flow = flow_leg(dev, K) if args.config == 'cfg4' else None
    others = None
    if args.config == 'cfg2' and world == 1 and not args.no_other_configs:
        # the other BASELINE shapes on this GPU: cfg3 = configs[2]'s loss step at its global batch 256 (the same frame
        # size in 8 waves instead of one), cfg4 = configs[3], cfg5 = configs[4] (the HBM-stress shape)
        others = {n: quick_config(n, dev, rank) for n in ('cfg3', 'cfg4', 'cfg5')}
        others['cfg4_flow'] = flow_leg(dev, 60)      # configs[3]'s own loss loop as the fused flow-and-depth step
        # configs[1] words the photometric term as L1/SSIM; the reference itself has no SSIM, so this is the
        # library's extension (VslLossDesc.ssim_weight: one more launch between the fused kernel and the finalize)
        others['cfg2_l1_ssim'] = quick_config('cfg2', dev, rank, extra_flags={'ssim_weight': 0.85},
                                              note='photometric term = 0.15 L1 + 0.85 SSIM(3x3) (extension, 4 launches per step; kernel_ms_mean is the fused L1 kernel alone)')

    if rank == 0:
        peak, peak_src = measured_peak()
        kmean = statistics.mean(kern_ms)
        algo = fused_kernel_bytes(c, B)
        achieved = algo / (kmean * 1e-3) / 1e9
        total_pv = pixel_views(c, B_global if c['scaling'] == 'strong' else B) * (1 if c['scaling'] == 'strong' else world)
        kname = 'loss_fused_pair_kernel<%d>' % V if V % 2 == 0 else 'loss_fused_kernel<%d, false, false>' % V
        out = {
            'metric': METRIC, 'value': total_pv * K / (ms_max * 1e-3) / 1e6, 'unit': UNIT,
            'n_gpus': world, 'steps': K, 'warmup': Wm, 'ms_per_step': ms_max / K, 'higher_is_better': True,
            'scaling': c['scaling'], 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(workload=c['text'], name=args.config,
                           l2='%d rotating input sets of %.0f MB each (> 126 MB L2 between reuses)' % (NSETS, set_bytes / 1e6),
                           pix='1 pix = 1 pixel-view = target pixel x scale x view; %.0f per step per GPU' % pixel_views(c, B),
                           parallelism='dp%d (batch shards, no data-path collective)' % world,
                           B=B, H=H, W=W, S=S, V=V),
            'roofline': {'bound': 'hbm', 'kernel': kname, 'achieved': achieved, 'peak': peak,
                         'unit': 'GB/s', 'frac': achieved / peak, 'traffic': recorded_traffic(args.config),
                         'algorithmic_bytes_per_launch': algo, 'kernel_ms_mean': kmean,
                         'kernel_ms_min': min(kern_ms), 'kernel_share_of_step': kmean / (ms_max / K), 'kernel_timed_steps': len(kern_ms),
                         'peak_source': peak_src},
            'value_u8_frames': {'value': total_pv * K / (ms8 * 1e-3) / 1e6, 'unit': UNIT, 'ms_per_step': ms8 / K,
                                'how': 'the same device-resident step with the frames as uint8 (converted on load, results bit-identical)'},
            'e2e': {'value': total_pv / (e2e_ms * 1e-3) / 1e6, 'unit': UNIT, 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': Ke,
                    'host_numa_bound_cpus': len(numa_cpus) if numa_cpus else None,
                    'how': 'ops.HostPipeline: pinned host inputs (frames uint8 as the loader holds them, network outputs float32) '
                           '-> H2D (one copy) -> 3 launches -> D2H of losses and all gradients (one copy), double-buffered over 3 streams'},
            'gpu_launches': 3 * K, 'launches_per_step': 3, 'clocks': clocks,
            'losses': {'pixel': losses[0], 'smooth': losses[1], 'exp': losses[2]},
        }
        if train is not None:
            out['train'] = train
        if flow is not None:
            out['flow'] = flow
        if others is not None:
            out['other_configs'] = others
        if world == 1 and not args.no_cpu_baseline:
            out['cpu_baseline'] = cpu_baseline(c, reps=12)
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
    ap.add_argument('--config', default='cfg2', choices=sorted(CONFIGS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train', action='store_true')
    ap.add_argument('--no-other-configs', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.out = _QuietStdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
