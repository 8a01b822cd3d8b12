"""GPU: the flat-arena Adam kernel (vsl_adam_step) against the oracle's restatement of
tf.train.AdamOptimizer (train_depth_then_cam_lr.py:413).  TensorFlow is un-vendored and unpinned in the reference:
parity is against the documented algorithm only.  Tolerance 1e-6 relative (fp32 arithmetic against float64)."""
import pytest
import torch

from oracle import vsl_oracle as O
from tf_depth_estimation_b200 import dist as vdist
from tf_depth_estimation_b200 import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('n,offset', [(1, 0), (3, 1), (1000, 0), (1003, 3), (1 << 20, 2), (4097, 0)])
def test_adam_kernel_against_oracle(n, offset):
    g = torch.Generator().manual_seed(n)
    arena = lambda: torch.zeros(n + 8, device=DEV)
    P, G, M, V = arena(), arena(), arena(), arena()
    p, gr, m, v = (t[offset:offset + n] for t in (P, G, M, V))       # ranges off the 16-byte boundary too
    p0 = torch.randn(n, generator=g)
    p.copy_(p0)
    rp, rm, rv = p0.double(), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    for t in (1, 2, 3, 50):
        gt = torch.randn(n, generator=g) * (10.0 ** float(torch.randint(-3, 2, (1,), generator=g)))
        gr.copy_(gt)
        ops.adam_step(p, gr, m, v, t, lr=2e-4, beta1=0.9)
        rp, rm, rv = O.adam_step_tf(rp, gt.double(), rm, rv, t, 2e-4, 0.9)
        assert float((p.cpu().double() - rp).abs().max()) <= 1e-6 * float(rp.abs().max()) + 1e-9
        assert float((m.cpu().double() - rm).abs().max()) <= 1e-6 * float(rm.abs().max()) + 1e-12
        assert float((v.cpu().double() - rv).abs().max()) <= 1e-6 * float(rv.abs().max()) + 1e-12
    # nothing outside the range was touched
    for T in (P, M, V):
        assert float(T[:offset].abs().sum()) == 0.0 and float(T[offset + n:].abs().sum()) == 0.0


def test_adam_grad_scale_and_errors():
    n = 257
    p, g, m, v = (torch.zeros(n, device=DEV) for _ in range(4))
    g.fill_(4.0)
    ops.adam_step(p, g, m, v, 1, lr=0.1, grad_scale=0.25)
    # t = 1: m_hat / sqrt(v_hat) = sign(g) => p = -lr (up to eps)
    assert float((p + 0.1).abs().max()) <= 1e-6
    assert float((m - 0.1).abs().max()) <= 1e-7                       # (1 - beta1) * (g * 0.25)
    with pytest.raises(TypeError):
        ops.adam_step(p.cpu(), g.cpu(), m.cpu(), v.cpu(), 1, lr=0.1)  # no CPU fallback
    with pytest.raises(Exception):
        ops.adam_step(p, g, m, v, 0, lr=0.1)                          # t >= 1


def test_data_parallel_adam_single_rank_on_gpu():
    shapes = [(64, 3, 7, 7), (64,), (1000, 33), (5,)]
    dp = vdist.DataParallelAdam(shapes, DEV, lr=1e-3, bucket_bytes=4096)
    gen = torch.Generator().manual_seed(0)
    ref = [(torch.zeros(s, dtype=torch.float64), torch.zeros(s, dtype=torch.float64), torch.zeros(s, dtype=torch.float64))
           for s in shapes]
    for t in (1, 2):
        for i, s in enumerate(shapes):
            gi = torch.randn(s, generator=gen)
            dp.grads[i].copy_(gi)
            ref[i] = O.adam_step_tf(ref[i][0], gi.double(), ref[i][1], ref[i][2], t, 1e-3)
        dp.step()
    for i in range(len(shapes)):
        assert float((dp.params[i].cpu().double() - ref[i][0]).abs().max()) <= 1e-6 * float(ref[i][0].abs().max()) + 1e-9


def test_peer_data_parallel_adam_single_rank():
    """world = 1: the fused reduce-scatter + Adam + all-gather kernel degenerates to Adam on the whole arena."""
    shapes = [(64, 3, 7, 7), (64,), (1000, 33), (5,)]
    dp = vdist.PeerDataParallelAdam(shapes, DEV, lr=1e-3)
    gen = torch.Generator().manual_seed(0)
    ref = [(torch.zeros(s, dtype=torch.float64), torch.zeros(s, dtype=torch.float64), torch.zeros(s, dtype=torch.float64))
           for s in shapes]
    for t in (1, 2, 3):
        for i, s in enumerate(shapes):
            gi = torch.randn(s, generator=gen)
            dp.grads[i].copy_(gi)
            ref[i] = O.adam_step_tf(ref[i][0], gi.double(), ref[i][1], ref[i][2], t, 1e-3)
        dp.step()
    dp.check_peers()
    got = [p.cpu() for p in dp.params]
    dp.close()
    for i in range(len(shapes)):
        assert float((got[i].double() - ref[i][0]).abs().max()) <= 1e-6 * float(ref[i][0].abs().max()) + 1e-9


def _peer_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    shapes = [(301, 7), (64,), (5,), (4099,)]
    peer = vdist.PeerDataParallelAdam(shapes, dev, lr=1e-2)
    assert peer.hi - peer.lo > 0 and (peer.lo % 4, peer.hi % 4) == (0, 0)
    for t in (1, 2, 3):
        for r in range(world):                                   # every rank can regenerate every rank's gradient
            gen = torch.Generator().manual_seed(100 * t + r)
            gs = [torch.randn(s, generator=gen) for s in shapes]
            if r == rank:
                for gv, g in zip(peer.grads, gs):
                    gv.copy_(g)
        peer.step()
    torch.cuda.synchronize()
    peer.check_peers()
    torch.save([p.cpu() for p in peer.params], out % rank)
    peer.close()
    dist.barrier()
    dist.destroy_process_group()


def _two_peer_gpus():
    try:
        return torch.cuda.device_count() >= 2 and torch.cuda.can_device_access_peer(0, 1) and \
            torch.cuda.can_device_access_peer(1, 0)
    except Exception:
        return False


@pytest.mark.skipif(not _two_peer_gpus(), reason='needs 2 GPUs of one node with peer access (NVLink peer memory)')
def test_peer_data_parallel_adam_two_ranks(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / 'rank%d.pt')
    mp.spawn(_peer_worker, args=(2, port, out), nprocs=2, join=True)
    a, b = torch.load(out % 0), torch.load(out % 1)
    shapes = [(301, 7), (64,), (5,), (4099,)]
    ref = [(torch.zeros(s, dtype=torch.float64), torch.zeros(s, dtype=torch.float64), torch.zeros(s, dtype=torch.float64))
           for s in shapes]
    for t in (1, 2, 3):
        gsum = [torch.zeros(s, dtype=torch.float64) for s in shapes]
        for r in range(2):
            gen = torch.Generator().manual_seed(100 * t + r)
            for acc, s in zip(gsum, shapes):
                acc += torch.randn(s, generator=gen).double()
        for i in range(len(shapes)):
            ref[i] = O.adam_step_tf(ref[i][0], gsum[i].float().double(), ref[i][1], ref[i][2], t, 1e-2)
    for i in range(len(shapes)):
        assert torch.equal(a[i], b[i])                           # replicas bit-identical
        assert float((a[i].double() - ref[i][0]).abs().max()) <= 2e-6 * float(ref[i][0].abs().max()) + 1e-9
