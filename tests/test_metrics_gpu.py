"""Empirical CRPS kernel against the reference's golden vectors and the oracle (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MO

pytestmark = pytest.mark.gpu
DEV = 'cuda'
G = os.path.join(os.path.dirname(__file__), 'golden', 'crps.npz')


@pytest.mark.parametrize('case', ['s100', 's7', 's1', 's130'])
def test_crps_matches_reference(case):
    from prob_unet_mds_b200 import metrics
    fx = np.load(G)
    pred, truth = torch.from_numpy(fx[case + '_pred']).to(DEV), torch.from_numpy(fx[case + '_truth']).to(DEV)
    got = metrics.crps_empirical(pred, truth)
    # the pair-sum form and the reference's sorted form are the same number up to fp32 summation order
    np.testing.assert_allclose(got.cpu().numpy(), fx[case + '_crps'], rtol=2e-5, atol=2e-6)
    assert metrics.CRPSLoss()(pred, truth).shape == truth.shape


def test_crps_ensemble_layout_and_properties():
    from prob_unet_mds_b200 import metrics
    g = torch.Generator().manual_seed(4)
    B, S, Cc, H, W = 3, 100, 3, 16, 24
    ens = torch.randn(B, S, Cc, H, W, generator=g)
    truth = torch.randn(B, Cc, H, W, generator=g)
    ref = MO.crps_empirical(ens.transpose(0, 1).contiguous(), truth)
    got = metrics.crps_ensemble(ens.to(DEV), truth.to(DEV)).cpu()
    assert torch.allclose(got, ref, rtol=2e-5, atol=2e-6)
    # size-independent properties: invariant to the order of the members; equals the absolute error for S = 1;
    # non-negative; shifting ensemble and truth together changes nothing
    perm = torch.randperm(S, generator=g)
    got_p = metrics.crps_ensemble(ens[:, perm].contiguous().to(DEV), truth.to(DEV)).cpu()
    assert torch.allclose(got_p, got, rtol=1e-5, atol=1e-6)
    one = metrics.crps_ensemble(ens[:, :1].contiguous().to(DEV), truth.to(DEV)).cpu()
    assert torch.allclose(one, (ens[:, 0] - truth).abs(), atol=1e-7)
    assert (got > -1e-6).all()
    got_s = metrics.crps_ensemble((ens + 3.0).to(DEV), (truth + 3.0).to(DEV)).cpu()
    assert torch.allclose(got_s, got, rtol=1e-4, atol=1e-5)


def test_crps_full_size():
    """The bench ensemble (64 inputs x 100 members x 3 x 128 x 128) against the oracle on a slice."""
    from prob_unet_mds_b200 import metrics
    g = torch.Generator(device=DEV).manual_seed(1)
    ens = torch.randn(64, 100, 3, 128, 128, generator=g, device=DEV)
    truth = torch.randn(64, 3, 128, 128, generator=g, device=DEV)
    got = metrics.crps_ensemble(ens, truth)
    ref = MO.crps_empirical(ens[5:7, :, :, 40:56].transpose(0, 1).contiguous().cpu(), truth[5:7, :, 40:56].cpu())
    assert torch.allclose(got[5:7, :, 40:56].cpu(), ref, rtol=2e-5, atol=2e-6)
    assert torch.isfinite(got).all()


def test_crps_shape_errors():
    from prob_unet_mds_b200 import metrics
    with pytest.raises(ValueError):
        metrics.crps_empirical(torch.zeros(4, 3, device=DEV), torch.zeros(2, device=DEV))
    with pytest.raises(RuntimeError):
        metrics.crps_empirical(torch.zeros(4, 3), torch.zeros(3))
