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
