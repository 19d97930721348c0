"""oracle/climex_oracle.py against the golden vectors produced by the unmodified reference methods (CPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import climex_oracle as CO

G = os.path.join(os.path.dirname(__file__), 'golden', 'climex_prepare.npz')


@pytest.mark.parametrize('mode', ['none', 'perpixel', 'pertimestep', 'minmax'])
def test_oracle_matches_reference(mode):
    fx = np.load(G)
    hr = torch.from_numpy(fx['hr_all'])
    stats = CO.compute_stats(hr, mode)
    if mode != 'none':
        np.testing.assert_allclose(stats[0].numpy(), fx[f'{mode}_s0'], rtol=0, atol=0)
        np.testing.assert_allclose(stats[1].numpy(), fx[f'{mode}_s1'], rtol=0, atol=0)
    out = CO.prepare_batch(hr, mode, stats)
    for k in ('inputs', 'targets', 'lr', 'lrinterp'):
        np.testing.assert_allclose(out[k].numpy(), fx[f'{mode}_{k}'], rtol=1e-6, atol=1e-6, err_msg=k)
    hp = CO.residual_to_hr(torch.from_numpy(fx['residual']), out['lrinterp'], mode, stats)
    np.testing.assert_allclose(hp.numpy(), fx[f'{mode}_hr_pred'], rtol=1e-6, atol=1e-6)
