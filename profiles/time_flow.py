"""Step time of the fused flow-and-depth loss (vsl_flow_loss_fwd_bwd; train_optflow_combine.py:138-240) at BASELINE
configs[3]: B=64, 192x256, 4 scales.  Calls the C ABI directly on pre-allocated buffers (no autograd, no allocation in
the loop), rotating input sets larger than L2, CUDA events.  bench.py --config cfg4 reports measure() as `flow`."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(dev, B=64, H=192, W=256, S=4, steps=200, warmup=8, nsets=4, seed=10):
    from tf_depth_estimation_b200 import _lib, synth
    from tf_depth_estimation_b200._lib import check, ptr_array
    lib = _lib.load()
    base = synth.make_flow_pairs(min(B, 16), H, W, S=S, seed=seed)
    rep = lambda t: t.repeat((B + t.shape[0] - 1) // t.shape[0], *([1] * (t.dim() - 1)))[:B]
    sets = []
    for k in range(nsets):      # rotating input sets: consecutive steps never find their inputs in L2
        r = lambda t: torch.roll(rep(t), k, dims=0).to(dev).contiguous()
        sets.append({n: (r(v) if torch.is_tensor(v) else [r(t) for t in v]) for n, v in base.items()})
    desc = _lib.VslFlowLossDesc(B=B, H=H, W=W, S=S, smooth_weight=0.5, depth_weight=1.0, data_weight=1.0,
                                optflow_weight=1.0, loss_scale=1.0)
    ws = torch.empty(lib.vsl_flow_loss_ws_bytes(ctypes.byref(desc)), dtype=torch.uint8, device=dev)
    losses = torch.zeros(8, device=dev)
    g = [[torch.empty(B, H >> s, W >> s, 1, device=dev) for s in range(S)] for _ in range(3)]
    P = lambda ts: ptr_array([t.data_ptr() for t in ts])
    st = torch.cuda.current_stream(dev).cuda_stream
    calls = [(d['left'].data_ptr(), d['right'].data_ptr(), d['label'].data_ptr(), P(d['depth_pyr']), P(d['flowx_pyr']),
              P(d['flowy_pyr']), d['proj'].data_ptr(), d['K_pyr'].data_ptr()) for d in sets]
    gp = (P(g[0]), P(g[1]), P(g[2]))

    def step(c):
        check(lib.vsl_flow_loss_fwd_bwd(ctypes.byref(desc), c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7],
                                        losses.data_ptr(), gp[0], gp[1], gp[2], ws.data_ptr(), st))
    for i in range(max(warmup, 3)):
        step(calls[i % nsets])
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(calls[i % nsets])
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    px = B * H * W * sum(0.25 ** s for s in range(S))
    set_bytes = sum(t.numel() * 4 for n, v in sets[0].items() for t in ([v] if torch.is_tensor(v) else v))
    # compulsory bytes per pixel and scale: left 12 + right 12 (each byte once) + label 4 + 3 predictions 12 + 3 gradients 12
    return {'ms_per_step': ms, 'value': px / (ms * 1e-3) / 1e6, 'unit': 'Mpix/s (1 pix = target pixel x scale)',
            'launches_per_step': 4, 'algorithmic_bytes_per_step': px * 52, 'achieved_gbs': px * 52 / (ms * 1e-3) / 1e9,
            'l2': '%d rotating input sets of %.0f MB' % (nsets, set_bytes / 1e6), 'steps': steps,
            'losses': dict(zip(('depth', 'smooth', 'optflow', 'pixel', 'total'), losses[:5].tolist())),
            'workload': 'configs[3]: flow-and-depth loss of train_optflow_combine.py:138-240 (three smoothness terms, '
                        'supervised inverse depth, depth warp + flow warp under the ground-truth validity mask, flow vs '
                        'depth_optflow), B=%d %dx%d %d scales, forward+backward in one pass' % (B, H, W, S)}


if __name__ == '__main__':
    r = measure(torch.device('cuda:0'))
    print('flow+depth loss step cfg4: %.1f us  (%.1f Gpix/s; %.0f GB/s of 52 compulsory bytes/pixel)'
          % (r['ms_per_step'] * 1e3, r['value'] / 1e3, r['achieved_gbs']))
    print('losses', r['losses'])
