"""Device-side ClimEx sample preparation against the reference's golden vectors and the oracle (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import climex_oracle as CO

pytestmark = pytest.mark.gpu
DEV = 'cuda'
G = os.path.join(os.path.dirname(__file__), 'golden', 'climex_prepare.npz')


@pytest.mark.parametrize('mode', ['none', 'perpixel', 'pertimestep', 'minmax'])
def test_prepare_and_inverse_match_reference(mode):
    from prob_unet_mds_b200 import data
    fx = np.load(G)
    hr = torch.from_numpy(fx['hr_all']).to(DEV)
    stats = data.compute_stats(hr, mode)
    if mode != 'none':
        np.testing.assert_allclose(stats[0].cpu().numpy().reshape(fx[f'{mode}_s0'].shape), fx[f'{mode}_s0'], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(stats[1].cpu().numpy().reshape(fx[f'{mode}_s1'].shape), fx[f'{mode}_s1'], rtol=1e-5, atol=1e-5)
        # use the reference's own statistics for the element-wise comparison below
        stats = (torch.from_numpy(fx[f'{mode}_s0']).to(DEV), torch.from_numpy(fx[f'{mode}_s1']).to(DEV))
    out = data.prepare_batch(hr, mode, stats)
    for k in ('inputs', 'targets', 'lr', 'lrinterp'):
        np.testing.assert_allclose(out[k].cpu().numpy(), fx[f'{mode}_{k}'], rtol=2e-6, atol=2e-6, err_msg=k)
    res = torch.from_numpy(fx['residual']).to(DEV)
    hp = data.residual_to_hr(res, out['lrinterp'], mode, stats)
    np.testing.assert_allclose(hp.cpu().numpy(), fx[f'{mode}_hr_pred'], rtol=2e-6, atol=2e-5)
    # ensemble form [B, S, C, H, W]
    ens = torch.stack([res, 2 * res], dim=1).contiguous()
    hp2 = data.residual_to_hr(ens, out['lrinterp'], mode, stats)
    assert torch.equal(hp2[:, 0], hp)


def test_prepare_full_size_against_oracle():
    """BASELINE-sized tiles (128x128, the bench batch of 64): oracle on the CPU, kernels on the GPU."""
    from prob_unet_mds_b200 import data
    g = torch.Generator().manual_seed(3)
    hr = torch.randn(64, 3, 128, 128, generator=g) * 3 + 1
    stats = CO.compute_stats(hr, 'perpixel')
    ref = CO.prepare_batch(hr, 'perpixel', stats)
    out = data.prepare_batch(hr.to(DEV), 'perpixel', [s.to(DEV) for s in stats])
    for k in ('inputs', 'targets', 'lr', 'lrinterp'):
        assert torch.allclose(out[k].cpu(), ref[k], rtol=2e-6, atol=2e-6), k


def test_prepare_rejects_cpu_tensors():
    from prob_unet_mds_b200 import data
    with pytest.raises(RuntimeError):
        data.prepare_batch(torch.zeros(1, 3, 8, 8), 'none')
