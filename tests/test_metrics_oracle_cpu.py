"""oracle/metrics_oracle.py against golden vectors from the unmodified reference crps_empirical (CPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MO

G = os.path.join(os.path.dirname(__file__), 'golden', 'crps.npz')


@pytest.mark.parametrize('case', ['s100', 's7', 's1', 's130'])
def test_crps_oracle_matches_reference(case):
    fx = np.load(G)
    got = MO.crps_empirical(torch.from_numpy(fx[case + '_pred']), torch.from_numpy(fx[case + '_truth']))
    np.testing.assert_allclose(got.numpy(), fx[case + '_crps'], rtol=0, atol=0)
