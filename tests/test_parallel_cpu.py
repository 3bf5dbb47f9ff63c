"""Data-parallel gradient sync (cliora_b200/parallel.py) on CPU: world_size 2 over gloo.

Each rank owns its own sentence shard end to end; the only collective is one all-reduce of the flat gradient
buffer.  The check: averaged per-rank grads == grads of a single process that computes
every shard's loss separately and averages."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(seed=0):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))


def _shard(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(4, 6, generator=g), torch.randn(4, 3, generator=g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from cliora_b200.parallel import GradSync
    m = _model(seed=rank)      # every rank draws its OWN weights (the reference never seeds torch) ...
    sync = GradSync(list(m.parameters()), world)     # ... and the wrap broadcasts rank 0's, like DDP
    assert sync.in_sync()
    ref = _model(seed=0)
    for p, q in zip(m.parameters(), ref.parameters()):
        assert torch.equal(p, q)
    x, y = _shard(rank)
    ((m(x) - y) ** 2).mean().backward()
    sync()
    if rank == 0:
        torch.save([p.grad.clone() for p in m.parameters()], out)
    with torch.no_grad():      # a rank-dependent update must be noticed
        if rank == 1:
            next(m.parameters()).add_(1e-3)
    assert not sync.in_sync()
    dist.barrier()
    dist.destroy_process_group()


def test_gradsync_world2_gloo(tmp_path):
    world, port, out = 2, _free_port(), str(tmp_path / 'g.pt')
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = torch.load(out)
    m = _model()
    for r in range(world):
        x, y = _shard(r)
        (((m(x) - y) ** 2).mean() / world).backward()
    for a, p in zip(got, m.parameters()):
        assert torch.allclose(a, p.grad, rtol=1e-5, atol=1e-7)


def test_gradsync_world1_is_noop():
    from cliora_b200.parallel import GradSync
    m = _model()
    sync = GradSync(list(m.parameters()), 1)
    ((m(torch.randn(2, 6))) ** 2).mean().backward()
    before = [p.grad.clone() for p in m.parameters()]
    sync()
    for a, p in zip(before, m.parameters()):
        assert torch.equal(a, p.grad)


def _worker_frozen(rank, world, port):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from cliora_b200.parallel import GradSync
    m = _model(seed=rank)
    m[0].weight.requires_grad = False            # a frozen tensor (the word-embedding table in the real net)
    m.register_buffer('stat', torch.full((3,), float(rank)))
    GradSync.for_module(m, world)
    ref = _model(seed=0)
    assert torch.equal(m[0].weight, ref[0].weight)           # frozen parameters are broadcast too
    assert torch.equal(m.stat, torch.zeros(3))               # and buffers
    dist.barrier()
    dist.destroy_process_group()


def test_for_module_broadcasts_frozen_parameters_and_buffers():
    world, port = 2, _free_port()
    mp.spawn(_worker_frozen, args=(world, port), nprocs=world, join=True)
