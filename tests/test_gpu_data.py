"""GPU side of the DeMoN-pair loader: vsl_unpack_strip against the oracle's restatement of
tf.image.resize_images + unpack_image_sequence, bit for bit, and the whole chain files -> loader -> fused flow-and-depth
loss against the oracle on the same files."""
import pytest
import torch

from oracle import vsl_oracle as O
from tf_depth_estimation_b200 import data, ops

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('B,h,w,H,W', [(2, 24, 64, 24, 32), (3, 37, 90, 16, 24), (1, 16, 40, 48, 56), (2, 240, 1440, 192, 256)])
def test_unpack_strip_bit_exact(B, h, w, H, W):
    """identity size, shrink by non-integer factors, enlarge, and a loader-sized strip down to the cfg4 frame size"""
    g = torch.Generator().manual_seed(h * 1000 + w)
    strip = torch.randint(0, 256, (B, h, w, 3), generator=g, dtype=torch.uint8)
    tgt, src = data.unpack_strip(strip.to(DEV), H, W)
    want_t, want_s = O.unpack_strip(strip, H, W)
    assert torch.equal(tgt.cpu(), want_t) and torch.equal(src.cpu(), want_s)


def test_files_to_flow_loss(tmp_path):
    pytest.importorskip('PIL')
    h, w, S = 48, 64, 3
    data.write_synthetic_dataset(str(tmp_path), 4, h, w, seed=11)
    ds = data.PairDataset(str(tmp_path), h, w, num_scales=S, resizedheight=32, resizedwidth=48)
    samples = next(ds.batches(4, seed=0))
    tgt, src, label, Kp, projs, m_scale = data.load_batch(ds, samples, DEV)
    assert tuple(tgt.shape) == (4, 32, 48, 3) and tuple(Kp.shape) == (4, S, 3, 3) and tuple(projs.shape) == (4, 2, 4, 4)
    # the script feeds the label at the resized size (:143 set_shape); here: the oracle's area resize of the file's map
    lab = torch.nn.functional.interpolate(label.permute(0, 3, 1, 2), size=(32, 48), mode='area').permute(0, 2, 3, 1).contiguous()
    g = torch.Generator().manual_seed(3)
    mk = lambda s, a: (a * torch.randn(4, 32 >> s, 48 >> s, 1, generator=g))
    pd = [(lab.cpu()[:, ::2 ** s, ::2 ** s] * (1 + mk(s, 0.05))).clamp(0.05, 4.0).contiguous() for s in range(S)]
    fx, fy = [mk(s, 1.0) for s in range(S)], [mk(s, 1.0) for s in range(S)]
    flags = ops.FlowLossFlags(num_scales=S)
    leaf = lambda t: t.to(DEV).requires_grad_()
    cpd, cfx, cfy = [leaf(t) for t in pd], [leaf(t) for t in fx], [leaf(t) for t in fy]
    total, losses = ops.flow_depth_loss(tgt, src, lab, cpd, cfx, cfy, projs[:, 0].contiguous(), Kp, flags)
    total.backward()
    opd, ofx, ofy = ([t.clone().requires_grad_() for t in p] for p in (pd, fx, fy))
    strips = torch.stack([torch.from_numpy(s['strip']) for s in samples])
    ot, os_ = O.unpack_strip(strips, 32, 48)
    terms = O.flow_depth_loss(ot, os_, lab.cpu(), opd, ofx, ofy, projs[:, 0].cpu(), Kp.cpu(), O.FlowLossFlags(num_scales=S))
    sum(terms).backward()
    for i in range(4):
        want = float(terms[i].detach())
        assert abs(float(losses[i]) - want) <= 1e-5 * abs(want), (i, float(losses[i]), want)
    for got, want in ((cpd, opd), (cfx, ofx), (cfy, ofy)):
        for s in range(S):
            e = float((got[s].grad.cpu() - want[s].grad).abs().max() / want[s].grad.abs().max())
            assert e <= 1e-4, (s, e)
