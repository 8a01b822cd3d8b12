"""Step time of the fused flow-and-depth loss (vsl_flow_loss_fwd_bwd) at BASELINE configs[3]: B=64, 192x256, 4 scales.
Calls the C ABI directly on pre-allocated buffers (no autograd, no allocation in the loop); CUDA events."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_depth_estimation_b200 import _lib, synth
from tf_depth_estimation_b200._lib import ptr_array, check
dev = torch.device('cuda:0')
B, H, W, S = 64, 192, 256, 4
lib = _lib.load()
sets = []
for k in range(4):          # rotating input sets: 4 x ~190 MB > L2
    d = synth.make_flow_pairs(B, H, W, S=S, seed=10 + k)
    sets.append({n: (v.to(dev) if torch.is_tensor(v) else [t.to(dev) for t in v]) for n, v in d.items()})
desc = _lib.VslFlowLossDesc(B=B, H=H, W=W, S=S, smooth_weight=0.5, depth_weight=1.0, data_weight=1.0, optflow_weight=1.0, loss_scale=1.0)
ws = torch.empty(lib.vsl_flow_loss_ws_bytes(ctypes.byref(desc)), dtype=torch.uint8, device=dev)
losses = torch.zeros(8, device=dev)
g = [[torch.empty(B, H >> s, W >> s, 1, device=dev) for s in range(S)] for _ in range(3)]
P = lambda ts: ptr_array([t.data_ptr() for t in ts])
st = torch.cuda.current_stream().cuda_stream
def step(d):
    check(lib.vsl_flow_loss_fwd_bwd(ctypes.byref(desc), d['left'].data_ptr(), d['right'].data_ptr(), d['label'].data_ptr(),
                                    P(d['depth_pyr']), P(d['flowx_pyr']), P(d['flowy_pyr']), d['proj'].data_ptr(),
                                    d['K_pyr'].data_ptr(), losses.data_ptr(), P(g[0]), P(g[1]), P(g[2]), ws.data_ptr(), st))
for i in range(8): step(sets[i % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 200
e0.record()
for i in range(n): step(sets[i % 4])
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1000 / n
px = B * H * W * sum(0.25 ** s for s in range(S))
# compulsory bytes per pixel and scale: left 12 + right 12 (each byte once) + label 4 + 3 predictions 12 + 3 gradients 12
print('flow+depth loss step cfg4: %.1f us  (%.1f Gpix/s; %.0f GB/s of %d compulsory bytes/pixel)' % (us, px / us / 1e3, px * 52 / us / 1e3, 52))
print('losses', losses[:5].tolist())
