"""Device-side ClimEx sample preparation: the batched, on-GPU equivalent of
``climex_utils.climex2torch.__getitem__`` (climex_utils.py:122-162), ``compute_stats`` (:165-195) and
``residual_to_hr`` (:198-211).  High-resolution fields go in as fp32 NCHW CUDA tensors; nothing runs on the CPU."""
import torch

from . import _lib as L
from .ops import check, ptr, stream_ptr

MODES = {'none': 0, 'perpixel': 1, 'pertimestep': 2, 'minmax': 3}
EPSILON = 1e-10          # climex_utils.py:70


def _require(t, name):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise RuntimeError(f'prob_unet_mds_b200.data: {name} must be a contiguous fp32 CUDA tensor (there is no CPU path)')


def compute_stats(hr_all, standardization='perpixel', lowres_scale=4):
    """Statistics of the low-resolution data (climex_utils.py:165-195).  hr_all: [T, C, H, W] on the device.
    perpixel -> (mean, std) as [C, H, W] (already repeated to the high-resolution grid); pertimestep -> (mean, std)
    and minmax -> (min, max) as [T, C, 1, 1].  Run once per dataset: plain torch ops, not a hot path."""
    lr = torch.nn.functional.avg_pool2d(hr_all, lowres_scale)
    if standardization == 'perpixel':
        mean, std = lr.mean(dim=0), lr.std(dim=0)
        rep = lambda t: t.repeat_interleave(lowres_scale, dim=1).repeat_interleave(lowres_scale, dim=2).contiguous()  # noqa: E731
        return rep(mean), rep(std)
    if standardization == 'pertimestep':
        return lr.mean(dim=(2, 3), keepdim=True), lr.std(dim=(2, 3), keepdim=True)
    if standardization == 'minmax':
        return lr.amin(dim=(2, 3), keepdim=True), lr.amax(dim=(2, 3), keepdim=True)
    if standardization == 'none':
        return None
    raise ValueError(f'unknown standardization {standardization!r}')


def prepare_batch(hr, standardization='perpixel', stats=None, lowres_scale=4, epsilon=EPSILON):
    """hr: [B, C, H, W].  stats: compute_stats() output (for pertimestep / minmax: the rows of this batch, [B, C, 1, 1]).
    Returns the dict of ``__getitem__`` with a leading batch dimension: inputs, targets, hr, lr, lrinterp, stand_stats."""
    _require(hr, 'hr')
    mode = MODES[standardization]
    B, Cc, H, W = hr.shape
    s0 = s1 = None
    if mode:
        s0, s1 = [t.contiguous().float() for t in stats]
        _require(s0, 'stats[0]')
        want = Cc * H * W if mode == 1 else B * Cc
        if s0.numel() != want or s1.numel() != want:
            raise ValueError(f'prepare_batch: {standardization} statistics must have {want} elements each')
    lr = torch.empty((B, Cc, H // lowres_scale, W // lowres_scale), dtype=torch.float32, device=hr.device)
    lrinterp, inputs, targets = torch.empty_like(hr), torch.empty_like(hr), torch.empty_like(hr)
    check(L.lib().pu_climex_prepare(ptr(hr), ptr(s0), ptr(s1), mode, float(epsilon), B, Cc, H, W, lowres_scale, ptr(lr),
                                    ptr(lrinterp), ptr(inputs), ptr(targets), stream_ptr()), 'climex_prepare')
    return {'inputs': inputs, 'targets': targets, 'hr': hr, 'lr': lr, 'lrinterp': lrinterp,
            'stand_stats': (s0, s1) if mode in (2, 3) else 0}


def residual_to_hr(residual, lrinterp, standardization='perpixel', stats=None, epsilon=EPSILON):
    """climex_utils.py:198-211 on the device; residual may be [B, C, H, W] or an ensemble [B, S, C, H, W]."""
    _require(residual, 'residual')
    _require(lrinterp, 'lrinterp')
    mode = MODES[standardization]
    s0 = s1 = None
    if mode:
        s0, s1 = [t.contiguous().float() for t in stats]
    if residual.dim() == 5:
        B, S, Cc, H, W = residual.shape
        out = torch.empty_like(residual)
        for s in range(S):          # members are strided views of one buffer: one launch per member, no copies of lrinterp
            r = residual[:, s].contiguous()
            o = residual_to_hr(r, lrinterp, standardization, stats, epsilon)
            out[:, s] = o
        return out
    B, Cc, H, W = residual.shape
    out = torch.empty_like(residual)
    check(L.lib().pu_climex_residual_to_hr(ptr(residual), ptr(lrinterp), ptr(s0), ptr(s1), mode, float(epsilon), B, Cc, H, W,
                                           ptr(out), stream_ptr()), 'climex_residual_to_hr')
    return out
