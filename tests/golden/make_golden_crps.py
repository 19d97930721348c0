"""Golden vectors for the empirical CRPS from the UNMODIFIED reference function `trainmodel.crps_empirical`
(trainmodel.py:66-110).  `trainmodel` imports the plotting / NetCDF stack at module level; the stub finder of
make_golden_climex.py stands in for the packages that are not installed here.

    python tests/golden/make_golden_crps.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_climex import load_reference  # noqa: E402


def main():
    load_reference()
    import trainmodel
    g = torch.Generator().manual_seed(9)
    out = {}
    for name, S, shape in (('s100', 100, (2, 3, 8, 12)), ('s7', 7, (5, 3)), ('s1', 1, (4, 4)), ('s130', 130, (33,))):
        pred = torch.randn(S, *shape, generator=g) * 2 + 0.5
        truth = torch.randn(*shape, generator=g)
        out[name + '_pred'] = pred.numpy()
        out[name + '_truth'] = truth.numpy()
        out[name + '_crps'] = trainmodel.crps_empirical(pred, truth).numpy()
    path = os.path.join(HERE, 'crps.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
