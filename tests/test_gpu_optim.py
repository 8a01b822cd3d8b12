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


def _peer_worker(rank, world, port, out, multicast=False):
    import os
    import torch.distributed as dist
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    shapes = [(301, 7), (64,), (5,), (4099,)]
    if multicast and not vdist.MulticastArena.available(dev):
        torch.save('no multicast', out % rank)
        dist.destroy_process_group()
        return
    peer = vdist.PeerDataParallelAdam(shapes, dev, lr=1e-2, multicast=multicast)
    assert bool(peer.mc) == bool(multicast)
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


@pytest.mark.skipif(not _two_peer_gpus(), reason='needs 2 GPUs of one node behind an NVSwitch (multicast / NVLS)')
def test_peer_data_parallel_adam_two_ranks_multicast(tmp_path):
    """The same step through the switch's multicast engine (dp_adam_mc_kernel: multimem.ld_reduce / multimem.st on a
    symmetric allocation): replicas bit-identical, parameters equal to the oracle's Adam on the summed gradients (two
    addends: the in-switch sum has only one association order)."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / 'rank%d.pt')
    mp.spawn(_peer_worker, args=(2, port, out, True), nprocs=2, join=True)
    a, b = torch.load(out % 0), torch.load(out % 1)
    if isinstance(a, str):
        pytest.skip('this node has no multicast support')
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
        assert torch.equal(a[i], b[i])
        assert float((a[i].double() - ref[i][0]).abs().max()) <= 2e-6 * float(ref[i][0].abs().max()) + 1e-9


# ------------------------------------------------------------------ the fused peer step on ONE GPU: ranks = streams
def _adam_reference_on_gpu(shapes, grads_per_step, lr):
    """'gather, add in rank order, vsl_adam_step' on one flat arena -> (param_flat, m_flat, v_flat)."""
    _, _, _, numel = vdist._flat_layout(shapes)
    p, g, m, v = (torch.zeros(numel, device=DEV) for _ in range(4))
    for t, per_rank in enumerate(grads_per_step, start=1):
        g.zero_()
        for gr in per_rank:                 # rank order, float32 adds: what dp_adam_kernel does element by element
            g += gr
        ops.adam_step(p, g, m, v, t, lr=lr)
    return p, m, v


@pytest.mark.parametrize('world', [2, 4])
def test_peer_step_ranks_in_one_process_single_gpu(world):
    """dp_adam_kernel<2>/<4> + peer_barrier_kernel proven on a ONE-GPU box: `world` ranks live in this process, each
    with its own arena (InProcessArena) and its own stream; their barrier kernels run concurrently and meet through
    the flag words exactly as ranks on different GPUs do.  Result: every replica bit-identical to 'sum the gradients
    in rank order, then vsl_adam_step', moments included."""
    shapes = [(301, 7), (64,), (5,), (4099,), (33, 1031)]
    ranks = vdist.PeerDataParallelAdam.in_process(shapes, [DEV] * world, lr=1e-2, timeout_s=20.0)
    streams = [torch.cuda.Stream() for _ in ranks]
    numel = ranks[0].numel
    gen = torch.Generator().manual_seed(5)
    history = []
    for t in (1, 2, 3, 4):
        per_rank = [torch.randn(numel, generator=gen).to(DEV) * (0.1 if r else 1.0) for r in range(world)]
        history.append(per_rank)
        for dp, gr in zip(ranks, per_rank):
            dp.grad_flat.copy_(gr)
        torch.cuda.synchronize()
        for dp, st in zip(ranks, streams):                      # rank r's whole step is queued before rank r+1's:
            dp.step(stream=st.cuda_stream)                      # its first barrier spins until the others arrive
        torch.cuda.synchronize()                                # (the next gradients are written on another stream)
    for dp in ranks:
        dp.check_peers()
    p_ref, m_ref, v_ref = _adam_reference_on_gpu(shapes, history, 1e-2)
    for dp in ranks:
        assert torch.equal(dp.param_flat, p_ref), 'rank %d parameters differ from the reference order' % dp.rank
        assert torch.equal(dp.m_shard[:dp.hi - dp.lo], m_ref[dp.lo:dp.hi])
        assert torch.equal(dp.v_shard[:dp.hi - dp.lo], v_ref[dp.lo:dp.hi])
        assert dp.state[:2].tolist() == [8, 4]                  # 2 barriers per step; Adam's t lives on the device
    for dp in ranks:
        dp.close()


def test_peer_step_timeout_fails_the_step_instead_of_corrupting():
    """A rank whose peer never arrives: the barrier gives up after timeout_s, the update is NOT applied (no
    half-written gradients are ever summed), and the host-side step raises."""
    shapes = [(1000,)]
    ranks = vdist.PeerDataParallelAdam.in_process(shapes, [DEV, DEV], lr=1e-2, timeout_s=0.3)
    a = ranks[0]
    a.param_flat.fill_(1.0)
    a.grad_flat.fill_(3.0)
    st = torch.cuda.Stream()
    a.step(stream=st.cuda_stream)                                # rank 1 never steps
    with pytest.raises(vdist.PeerTimeout):
        a.check_peers()
    assert float((a.param_flat - 1.0).abs().max()) == 0.0        # untouched
    assert float(a.m_shard.abs().max()) == 0.0
    with pytest.raises(vdist.PeerTimeout):
        a.step(stream=st.cuda_stream)
    torch.cuda.synchronize()
    for dp in ranks:
        dp.close()


def test_peer_step_replays_from_a_cuda_graph():
    """The barrier epoch and Adam's step count live in device memory, so the step takes no per-step host argument:
    captured once, replayed three times == three eager steps of the oracle's Adam."""
    shapes = [(257, 3), (1000,)]
    dp = vdist.PeerDataParallelAdam(shapes, DEV, lr=1e-3)
    gen = torch.Generator().manual_seed(1)
    gs = [torch.randn(dp.numel, generator=gen).to(DEV) for _ in range(3)]
    static_g = torch.zeros(dp.numel, device=DEV)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            dp.grad_flat.copy_(static_g)
            dp.step()
    torch.cuda.current_stream().wait_stream(side)
    rp = torch.zeros(dp.numel, dtype=torch.float64)
    rm, rv = torch.zeros_like(rp), torch.zeros_like(rp)
    for t, g in enumerate(gs, start=1):
        static_g.copy_(g)
        graph.replay()
        rp, rm, rv = O.adam_step_tf(rp, g.cpu().double(), rm, rv, t, 1e-3)
    dp.check_peers()
    assert dp.state[:2].tolist() == [3, 3]
    assert float((dp.param_flat.cpu().double() - rp).abs().max()) <= 1e-6 * float(rp.abs().max()) + 1e-9
    dp.close()
