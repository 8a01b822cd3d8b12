"""CPU, world_size 2 over gloo: the batch-shard host logic (tf_depth_estimation_b200/dist.py).  The local
compute on each rank is the CPU oracle (the CUDA path needs a GPU); what is under test is that shard ranges,
loss_scale weights and the loss / gradient reductions reproduce the single-process global-batch result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vsl_oracle as O
from tf_depth_estimation_b200 import dist as vdist
from tf_depth_estimation_b200 import synth

B, H, W, S, V = 5, 16, 24, 2, 2  # odd batch: ranks get 3 and 2 samples


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _local(d, scale):
    xs = [x.double().requires_grad_() for x in d['disp_pyr']]
    ps = d['poses'].double().requires_grad_()
    lg = [l.double().requires_grad_() for l in d['logits_pyr']]
    flags = O.LossFlags(num_scales=S)
    losses = O.view_synthesis_loss(d['tgt'].double(), [s.double() for s in d['srcs']], xs, ps, d['K_pyr'].double(),
                                   lg, None, flags)
    (sum(losses) * scale).backward()
    return torch.stack([l.detach() for l in losses]), xs, ps, lg


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    full = synth.make_snippets(B, H, W, S=S, V=V, seed=77)
    mine = vdist.shard_snippets(full, rank, world)
    b_local = mine['tgt'].shape[0]
    losses, xs, ps, lg = _local(mine, vdist.local_loss_scale(b_local, B))
    glob = vdist.reduce_losses(losses, b_local, B)
    # a "network gradient": something every rank holds a replica of, here d(total)/d(a shared scalar gain on K)
    shared = [torch.stack([x.grad.sum() for x in xs]), ps.grad.sum(0)]
    vdist.reduce_sum_(shared)
    if rank == 0:
        torch.save({'glob': glob, 'shared': shared, 'ps': ps.grad, 'lo_hi': vdist.shard_range(B, rank, world)}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges():
    assert [vdist.shard_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [vdist.shard_range(256, r, 8) for r in range(8)][-1] == (224, 256)
    covered = [i for r in range(3) for i in range(*vdist.shard_range(7, r, 3))]
    assert covered == list(range(7))
    with pytest.raises(ValueError):
        vdist.shard_range(4, 2, 2)
    assert vdist.local_loss_scale(32, 256) == 0.125


def test_two_rank_gloo_matches_global_batch(tmp_path):
    out = str(tmp_path / 'rank0.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    full = synth.make_snippets(B, H, W, S=S, V=V, seed=77)
    losses, xs, ps, lg = _local(full, 1.0)
    assert torch.allclose(got['glob'], losses, rtol=1e-12, atol=0)
    assert torch.allclose(got['shared'][0], torch.stack([x.grad.sum() for x in xs]), rtol=1e-10, atol=1e-14)
    assert torch.allclose(got['shared'][1], ps.grad.sum(0), rtol=1e-10, atol=1e-14)
    lo, hi = got['lo_hi']
    # per-sample gradients of a rank are exactly that rank's rows of the global gradient
    assert torch.allclose(got['ps'], ps.grad[lo:hi], rtol=1e-10, atol=1e-14)


# ---- the data-parallel optimiser step (dist.DataParallelAdam): bucketing + reduction logic over gloo, with the
# oracle's Adam restatement injected as the update (the CUDA kernel needs a GPU; tests/test_gpu_optim.py covers it)
SHAPES = [(7, 3), (5,), (2, 3, 4), (1,), (33,)]


def _oracle_adam(hyper):
    def fn(p, g, m, v, step):
        np_, nm, nv = O.adam_step_tf(p.double(), g.double() * hyper.get('grad_scale', 1.0), m.double(), v.double(), step,
                                     hyper['lr'], hyper['beta1'], hyper['beta2'], hyper['eps'])
        p.copy_(np_); m.copy_(nm); v.copy_(nv)
    return fn


def _rank_grads(rank, step):
    g = torch.Generator().manual_seed(1000 * step + rank)
    return [torch.randn(s, generator=g) for s in SHAPES]


def _dp_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    hyper = dict(lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8)
    dp = vdist.DataParallelAdam(SHAPES, 'cpu', bucket_bytes=64, adam_fn=_oracle_adam(hyper), **hyper)  # 16-float buckets
    assert len(dp.buckets) > 3
    for p in dp.params:
        p.fill_(0.5)
    for step in (1, 2, 3):
        for gv, g in zip(dp.grads, _rank_grads(rank, step)):
            gv.copy_(g)
        assert dp.step() == step
    if rank == 0:
        torch.save([p.clone() for p in dp.params], out)
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_adam_two_ranks(tmp_path):
    out = str(tmp_path / 'params.pt')
    mp.spawn(_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    # single process: Adam on the summed gradients, tensor by tensor
    hyper = dict(lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8)
    ps = [torch.full(s, 0.5, dtype=torch.float64) for s in SHAPES]
    ms = [torch.zeros(s, dtype=torch.float64) for s in SHAPES]
    vs = [torch.zeros(s, dtype=torch.float64) for s in SHAPES]
    for step in (1, 2, 3):
        gs = [(a + b).double() for a, b in zip(_rank_grads(0, step), _rank_grads(1, step))]
        for i in range(len(SHAPES)):
            ps[i], ms[i], vs[i] = O.adam_step_tf(ps[i], gs[i], ms[i], vs[i], step, **hyper)
    for a, b in zip(got, ps):
        assert a.shape == b.shape
        assert torch.allclose(a.double(), b, rtol=1e-5, atol=1e-7)


def test_data_parallel_adam_layout():
    dp = vdist.DataParallelAdam(SHAPES, 'cpu', lr=0.1, adam_fn=lambda *a: None)
    assert all(o % 4 == 0 for o in dp.offsets)                      # 16-byte aligned tensors
    assert dp.buckets[0][0] == 0 and dp.buckets[-1][1] == dp.numel
    assert all(a[1] == b[0] for a, b in zip(dp.buckets, dp.buckets[1:]))
    assert [tuple(p.shape) for p in dp.params] == SHAPES
    dp.grads[2].fill_(3.0)
    assert float(dp.grad_flat.sum()) == 3.0 * 24                    # views alias the arena
    with pytest.raises(TypeError):
        vdist.DataParallelAdam(SHAPES, 'cpu', lr=0.1).step()        # default update = the CUDA kernel: no CPU fallback


def test_numa_binding_helper(tmp_path, monkeypatch):
    """bind_host_to_gpu: cpulist parsing, and no effect where the topology is not exposed (this CPU container)."""
    import os
    from tf_depth_estimation_b200 import dist as vdist
    assert vdist._parse_cpulist('0-3,8,10-11\n') == {0, 1, 2, 3, 8, 10, 11}
    assert vdist._parse_cpulist('') == set()
    before = os.sched_getaffinity(0)
    assert vdist.bind_host_to_gpu('cuda:0', sysfs=str(tmp_path)) is None
    assert os.sched_getaffinity(0) == before
