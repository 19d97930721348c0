"""Fused multi-tensor AdamW against torch.optim.AdamW (GPU only)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    flat = torch.randn(70000, generator=g).to(DEV)
    shapes = [(64, 3, 3, 3), (128,), (7,), (1, 1), (300, 257), (65536,), (65537,), (512, 512, 3, 3)]
    ps = [torch.randn(*s, generator=g).to(DEV).requires_grad_(True) for s in shapes]
    # a parameter whose storage is a 4-byte-aligned (not 16-byte-aligned) view, like a slice of a flat bucket
    ps.append(flat[3:3 + 4099].detach().requires_grad_(True))
    return ps


def _grads(ps, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(p.shape, generator=g).to(DEV) * (0.1 + i) for i, p in enumerate(ps)]


@pytest.mark.parametrize('wd', [1e-2, 0.0])
def test_adamw_matches_torch(wd):
    from prob_unet_mds_b200 import AdamW
    a = _params(0)
    b = [p.detach().clone().requires_grad_(True) for p in a]
    ours = AdamW(a, lr=1e-3, weight_decay=wd)
    ref = torch.optim.AdamW(b, lr=1e-3, weight_decay=wd, foreach=False, fused=False)
    for step in range(6):
        gs = _grads(a, 100 + step)
        for i, (p, q, g) in enumerate(zip(a, b, gs)):
            skip = (step == 2 and i == 1)          # a parameter without a gradient keeps its own step count
            p.grad = None if skip else g.clone()
            q.grad = None if skip else g.clone()
        ours.step()
        ref.step()
    for i, (p, q) in enumerate(zip(a, b)):
        err = (p - q).abs().max().item() / (q.abs().max().item() + 1e-12)
        assert err < 2e-6, (i, tuple(p.shape), err)
        so, sr = ours.state[p], ref.state[q]
        assert int(so['step']) == int(sr['step'])
        assert torch.allclose(so['exp_avg'], sr['exp_avg'], rtol=1e-5, atol=1e-8)
        assert torch.allclose(so['exp_avg_sq'], sr['exp_avg_sq'], rtol=1e-5, atol=1e-10)
    # state_dict round trip into torch's optimizer (same layout)
    ref2 = torch.optim.AdamW([p.detach().clone().requires_grad_(True) for p in a], lr=1e-3, weight_decay=wd)
    ref2.load_state_dict(ours.state_dict())


def test_adamw_one_launch_per_step():
    from prob_unet_mds_b200 import AdamW, _lib
    ps = _params(1)
    opt = AdamW(ps, lr=1e-3)
    for p, g in zip(ps, _grads(ps, 5)):
        p.grad = g
    n0 = int(_lib._raw_lib().pu_launch_count(0))
    opt.step()
    assert int(_lib._raw_lib().pu_launch_count(0)) == n0 + 1


def test_adamw_rejects_unsupported():
    from prob_unet_mds_b200 import AdamW
    with pytest.raises(ValueError):
        AdamW([torch.zeros(3, device=DEV, requires_grad=True)], amsgrad=True)
    p = torch.zeros(3, requires_grad=True)
    p.grad = torch.zeros(3)
    with pytest.raises(RuntimeError):
        AdamW([p]).step()
