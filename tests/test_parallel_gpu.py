"""Data-parallel ELBO gradients over NCCL on two GPUs (skipped on single-GPU boxes): each rank runs the real model on
half of the batch with parallel.DataParallel attached; the summed gradients must equal the single-process gradients of
the whole batch (losses are sums over samples, prob_unet.py:227,230, so the exchange is a SUM without division)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth

pytestmark = pytest.mark.gpu
L, B, H = 6, 4, 32


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_model(dev):
    from prob_unet_mds_b200 import ProbabilisticUNet
    m = ProbabilisticUNet(3, 3, latent_dim=L).to(dev)
    m.load_state_dict(synth.make_weights(synth.load_schema(f'schema_probunet_L{L}.json'), seed=0))
    m.set_precision('fp32')
    for blk in m.unet.modules():
        if hasattr(blk, 'dropout'):
            blk.dropout = 0
    m.train()
    return m


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from prob_unet_mds_b200 import parallel
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        m = _make_model(dev)
        dp = parallel.DataParallel(m)          # noqa: F841  (attaches the gradient sink)
        x, t = synth.make_inputs(B, H, H, seed=1)
        eps = synth.make_eps(B, L, seed=2)
        lo, hi = parallel.shard_range(B, rank, world)
        out = {}
        for step in range(2):                  # step 0 learns the gradient order, step 1 uses the overlapped buckets
            for p in m.parameters():
                p.grad = None
            m.eps_override = eps[lo:hi]
            total, recon, kl = m.elbo(x[lo:hi].to(dev), t[lo:hi].to(dev))
            total.backward()
            tot, = parallel.allreduce_losses(total.detach())
            out[step] = (tot.item(), {k: p.grad.detach().cpu() for k, p in m.named_parameters() if p.grad is not None})
        if rank == 0:
            torch.save(out, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_gpu_gradients_equal_full_batch(tmp_path):
    out_path = str(tmp_path / 'dp.pt')
    mp.spawn(_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    got = torch.load(out_path)
    dev = torch.device('cuda', 0)
    m = _make_model(dev)
    x, t = synth.make_inputs(B, H, H, seed=1)
    m.eps_override = synth.make_eps(B, L, seed=2)
    total, _, _ = m.elbo(x.to(dev), t.to(dev))
    total.backward()
    ref = {k: p.grad.detach().cpu() for k, p in m.named_parameters() if p.grad is not None}
    for step in (0, 1):
        tot, grads = got[step]
        assert abs(tot - total.item()) <= 1e-5 * abs(total.item())
        assert grads.keys() == ref.keys()
        errs = sorted(((g - ref[k]).norm().item() / (ref[k].norm().item() + 1e-30), k) for k, g in grads.items())
        # fp32 mode: both sides are ~1e-6 from the exact gradient; a ReLU unit within rounding noise of zero can differ
        # between the two runs (see test_model_gpu.py), hence the looser bound on the maximum
        assert errs[len(errs) // 2][0] <= 2e-5, (step, errs[-3:])
        assert errs[-1][0] <= 2e-2, (step, errs[-3:])


def _ens_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from prob_unet_mds_b200 import parallel
        m = _make_model(dev)
        m.eval()
        x, _ = synth.make_inputs(B, H, H, seed=1)
        S = 6
        g = torch.Generator().manual_seed(3)
        eps = torch.randn(B, S, L, generator=g).to(dev)
        out, (s_lo, s_hi) = parallel.ensemble_sharded(m, x.to(dev), S, eps=eps)
        torch.save({'out': out.cpu(), 'range': (s_lo, s_hi)}, f'{out_path}.{rank}')
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_gpu_sharded_ensemble_equals_single_process(tmp_path):
    """SURVEY 8e, ensemble row: inputs sharded for the encode, features all-gathered, members sharded for the decode.
    The two ranks' member blocks together must equal the single-process ensemble for the same eps."""
    out_path = str(tmp_path / 'ens.pt')
    mp.spawn(_ens_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    dev = torch.device('cuda', 0)
    m = _make_model(dev)
    m.eval()
    x, _ = synth.make_inputs(B, H, H, seed=1)
    S = 6
    g = torch.Generator().manual_seed(3)
    eps = torch.randn(B, S, L, generator=g).to(dev)
    ref = m.sample_ensemble(x.to(dev), S, eps=eps).cpu()
    covered = []
    for rank in range(2):
        blk = torch.load(f'{out_path}.{rank}')
        lo, hi = blk['range']
        covered += list(range(lo, hi))
        assert blk['out'].shape == (B, hi - lo) + tuple(ref.shape[2:])
        err = (blk['out'] - ref[:, lo:hi]).norm().item() / ref[:, lo:hi].norm().item()
        assert err <= 1e-5, (rank, err)
    assert covered == list(range(S))
