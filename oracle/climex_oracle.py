"""CPU oracle for the ClimEx sample preparation.  TEST INFRASTRUCTURE ONLY.

Restates, batched, the arithmetic of ``climex2torch.__getitem__`` (climex_utils.py:122-162), ``compute_stats``
(:165-195) and ``residual_to_hr`` / ``invstand_residual`` (:198-211) with the same torch calls the reference makes.
Pinned by ``tests/golden/climex_prepare.npz``, which ``tests/golden/make_golden_climex.py`` produces by calling the
UNMODIFIED reference methods (the module is imported with stub packages for xarray / dask / cartopy, which those methods
never touch)."""
import torch
import torch.nn as nn

EPSILON = 1e-10


def compute_stats(hr_all, standardization, lowres_scale=4):
    lr = nn.AvgPool2d(kernel_size=lowres_scale)(hr_all)
    if standardization == 'perpixel':
        mean, std = lr.mean(dim=0), lr.std(dim=0)
        mean = mean.repeat_interleave(repeats=lowres_scale, dim=1).repeat_interleave(repeats=lowres_scale, dim=2)
        std = std.repeat_interleave(repeats=lowres_scale, dim=1).repeat_interleave(repeats=lowres_scale, dim=2)
        return mean, std
    if standardization == 'pertimestep':
        return lr.mean(dim=(2, 3)).unsqueeze(2).unsqueeze(3), lr.std(dim=(2, 3)).unsqueeze(2).unsqueeze(3)
    if standardization == 'minmax':
        return (lr.min(dim=2)[0].min(dim=2)[0].unsqueeze(2).unsqueeze(3),
                lr.max(dim=2)[0].max(dim=2)[0].unsqueeze(2).unsqueeze(3))
    return None


def prepare_batch(hr, standardization, stats=None, lowres_scale=4, epsilon=EPSILON):
    """hr [B, C, H, W]; stats as compute_stats returns them (rows of this batch for pertimestep / minmax)."""
    lr = nn.AvgPool2d(kernel_size=lowres_scale)(hr)
    lrinterp = nn.functional.interpolate(input=lr, scale_factor=lowres_scale, mode='bilinear')
    if standardization == 'none':
        a, b = lrinterp, hr
    elif standardization == 'minmax':
        a = (lrinterp - stats[0]) / (stats[1] - stats[0] + epsilon)
        b = (hr - stats[0]) / (stats[1] - stats[0] + epsilon)
    else:
        a = (lrinterp - stats[0]) / (stats[1] + epsilon)
        b = (hr - stats[0]) / (stats[1] + epsilon)
    return {'inputs': a, 'targets': b - a, 'lr': lr, 'lrinterp': lrinterp}


def residual_to_hr(residual, lrinterp, standardization, stats=None, epsilon=EPSILON):
    if standardization == 'none':
        return lrinterp + residual
    if standardization == 'minmax':
        return lrinterp + residual * (stats[1] - stats[0] + epsilon)
    return lrinterp + residual * (stats[1] + epsilon)
