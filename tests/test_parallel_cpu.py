"""Host-side logic of the data-parallel gradient exchange, exercised on CPU with the gloo backend (world size 2):
bucket construction in backward order, zero-copy views, staging of foreign tensors, SUM (not mean) semantics,
parameters without gradients, and the work sharding helpers."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from prob_unet_mds_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.a = torch.nn.Parameter(torch.randn(300, 7, generator=g))
        self.b = torch.nn.Parameter(torch.randn(13, generator=g))
        self.c = torch.nn.Parameter(torch.randn(64, 5, 3, 3, generator=g))
        self.unused = torch.nn.Parameter(torch.randn(4, generator=g))   # like unet.map_layer*: never gets a grad
        self.d = torch.nn.Parameter(torch.randn(1001, generator=g))


def _fake_backward(model, sink, rank, step):
    """Mimics engine.backward: gradients appear in reverse order; big ones are written into sink.alloc() memory,
    small ones are produced elsewhere and handed over (staged by the sink)."""
    def val(p, k):
        return torch.full_like(p, float(rank + 1) * (k + 1) + step)
    g = sink.alloc(model.d)
    g.copy_(val(model.d, 0))
    sink[id(model.d)] = g
    sink[id(model.c)] = val(model.c, 1)           # foreign tensor -> staged
    g = sink.alloc(model.b)
    g.copy_(val(model.b, 2))
    sink[id(model.b)] = g
    g = sink.alloc(model.a)
    g.copy_(val(model.a, 3))
    sink[id(model.a)] = g
    sink.finish()


def _worker(rank, world, port):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        model = _Toy()
        if rank == 1:
            with torch.no_grad():
                model.a.add_(1.0)       # must be overwritten by the rank-0 broadcast
        dp = parallel.DataParallel(model, bucket_bytes=8 << 10)
        ref = _Toy()
        assert torch.equal(model.a, ref.a)
        for step in range(3):
            sink = model._grad_sink_factory()
            _fake_backward(model, sink, rank, step)
            for k, p in enumerate([model.d, model.c, model.b, model.a]):
                want = sum(float(r + 1) * (k + 1) + step for r in range(world))
                got = sink[id(p)]
                assert got.shape == p.shape
                assert torch.allclose(got, torch.full_like(p, want)), (step, k, got.flatten()[:3], want)
            assert id(model.unused) not in sink
            if step >= 1:
                # bucketed steps: gradients are views into the flat buckets (zero copy)
                b, off = dp.slots[id(model.a)]
                assert sink[id(model.a)].data_ptr() == dp.flat[b][off:].data_ptr()
        assert len(dp.flat) >= 2                       # 8 KiB buckets -> several buckets
        # gradient accumulation: two backward passes without zero_grad(set_to_none=True).  param.grad of the first pass
        # IS the bucket view; the second pass must not overwrite it before autograd's AccumulateGrad adds to it.
        params = [model.d, model.c, model.b, model.a]
        for p in params:
            p.grad = None
        for step in (5, 6):
            sink = model._grad_sink_factory()
            _fake_backward(model, sink, rank, step)
            for p in params:                               # what AccumulateGrad does
                g = sink[id(p)]
                if p.grad is None:
                    p.grad = g
                else:
                    p.grad.add_(g)
        for k, p in enumerate(params):
            want = sum(float(r + 1) * (k + 1) + 5 for r in range(world)) + sum(float(r + 1) * (k + 1) + 6 for r in range(world))
            assert torch.allclose(p.grad, torch.full_like(p, want)), (k, p.grad.flatten()[:3], want)
        assert [id(p) for p in dp.order] == [id(model.d), id(model.c), id(model.b), id(model.a)]
        # no_sync(): micro-batches whose gradients stay local, folded into the all-reduce of the next synchronised step
        for p in params:
            p.grad = None
        with dp.no_sync():
            for step in (7, 8):
                sink = model._grad_sink_factory()
                assert type(sink) is parallel.GradSink
                _fake_backward(model, sink, rank, step)
                for p in params:
                    g = sink[id(p)]
                    p.grad = g if p.grad is None else p.grad.add_(g)
        sink = model._grad_sink_factory()
        _fake_backward(model, sink, rank, 9)
        for p in params:
            g = sink[id(p)]
            p.grad = g if p.grad is None else p.grad.add_(g)
        for k, p in enumerate(params):
            want = sum(float(r + 1) * (k + 1) + st for r in range(world) for st in (7, 8, 9))
            assert torch.allclose(p.grad, torch.full_like(p, want)), ('no_sync', k, p.grad.flatten()[:3], want)
        t, = parallel.allreduce_losses(torch.tensor(float(rank + 1)))
        assert t.item() == 3.0
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_shard_range():
    assert [parallel.shard_range(100, r, 8) for r in range(8)] == [(0, 13), (13, 26), (26, 39), (39, 52), (52, 65),
                                                                  (65, 78), (78, 91), (91, 100)]
    assert parallel.shard_range(3, 7, 8) == (3, 3)
    covered = []
    for r in range(4):
        lo, hi = parallel.shard_range(64, r, 4)
        covered += list(range(lo, hi))
    assert covered == list(range(64))
