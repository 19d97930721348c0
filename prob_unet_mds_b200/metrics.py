"""Ensemble metrics on the device: ``trainmodel.crps_empirical`` / ``CRPSLoss`` (trainmodel.py:66-117)."""
import torch

from . import _lib as L
from .ops import check, ptr, stream_ptr


def _require(t, name):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise RuntimeError(f'prob_unet_mds_b200.metrics: {name} must be a contiguous fp32 CUDA tensor (there is no CPU path)')


def crps_empirical(pred, truth):
    """Reference signature: pred [S, *truth.shape] -> tensor of truth.shape."""
    if pred.shape[1:] != (1,) * (pred.dim() - truth.dim() - 1) + truth.shape:
        raise ValueError('Expected pred to have one extra sample dim on left. '
                         'Actual shapes: {} versus {}'.format(pred.shape, truth.shape))
    _require(pred, 'pred')
    _require(truth, 'truth')
    out = torch.empty_like(truth)
    M = truth.numel()
    check(L.lib().pu_crps_empirical(ptr(pred), ptr(truth), ptr(out), pred.shape[0], 1, M, M, 0, stream_ptr()), 'crps_empirical')
    return out


def crps_ensemble(ens, truth):
    """ens [B, S, C, H, W] as ``ProbabilisticUNet.sample_ensemble`` returns it, truth [B, C, H, W] -> [B, C, H, W]."""
    if ens.dim() != truth.dim() + 1 or ens.shape[0] != truth.shape[0] or ens.shape[2:] != truth.shape[1:]:
        raise ValueError(f'crps_ensemble: shapes {tuple(ens.shape)} and {tuple(truth.shape)} do not match')
    _require(ens, 'ens')
    _require(truth, 'truth')
    out = torch.empty_like(truth)
    B, S = ens.shape[0], ens.shape[1]
    inner = truth.numel() // B
    check(L.lib().pu_crps_empirical(ptr(ens), ptr(truth), ptr(out), S, B, inner, inner, S * inner, stream_ptr()),
          'crps_empirical')
    return out


class CRPSLoss(torch.nn.Module):
    """trainmodel.CRPSLoss (forward only: the reference never back-propagates through it on this path)."""

    def forward(self, pred, truth):
        return crps_empirical(pred, truth)
