"""CPU oracle for the ensemble metric.  TEST INFRASTRUCTURE ONLY.
Restates trainmodel.crps_empirical (trainmodel.py:66-110, itself borrowed from pyro); pinned by
tests/golden/crps.npz, produced by the unmodified reference function (tests/golden/make_golden_crps.py)."""
import torch


def crps_empirical(pred, truth):
    n = pred.size(0)
    if n == 1:
        return (pred[0] - truth).abs()
    pred = pred.sort(dim=0).values
    diff = pred[1:] - pred[:-1]
    weight = torch.arange(1, n, dtype=pred.dtype) * torch.arange(n - 1, 0, -1, dtype=pred.dtype)
    weight = weight.reshape(weight.shape + (1,) * (diff.dim() - 1))
    return (pred - truth).abs().mean(0) - (diff * weight).sum(0) / n ** 2
