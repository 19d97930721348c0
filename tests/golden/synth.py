"""Seeded synthetic ClimEx-shaped inputs and weights shared by the golden generator, the tests and
bench.py.  Pure torch, CPU generators only, imports neither the reference, the oracle nor the product.

Inputs mirror what climex_utils.py:124-128,155 hands the model: a standardised
high-resolution field, its 4x average-pooled / bilinearly re-interpolated low-resolution version as
`inputs`, and the residual as `targets`.
"""
import json
import math
import os

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))


def make_inputs(B, H, W, seed=1, channels=3):
    g = torch.Generator().manual_seed(seed)
    hr = torch.randn(B, channels, H, W, generator=g)
    lr = F.avg_pool2d(hr, 4)
    inputs = F.interpolate(lr, scale_factor=4, mode='bilinear')
    targets = hr - inputs
    return inputs.contiguous(), targets.contiguous()


def make_eps(B, L, seed=2):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, L, generator=g)


def load_schema(name):
    """[(key, shape), ...] in the reference's state_dict order (dumped by make_golden.py)."""
    with open(os.path.join(HERE, name)) as f:
        return [(k, tuple(s)) for k, s in json.load(f)]


def make_weights(schema, seed=0):
    """Deterministic, well-conditioned weights for every entry of a state_dict schema.

    The reference zero-initialises conv1 / proj / out_conv (networks.py:152,162,298), which would make a
    parity test vacuous, so every tensor gets a non-trivial value here.
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in schema:
        leaf = key.rsplit('.', 1)[-1]
        if leaf == 'resample_filter':
            sd[key] = torch.full(shape, 0.25)
        elif len(shape) == 1:
            is_norm_weight = leaf == 'weight'
            t = torch.randn(shape, generator=g) * 0.1
            sd[key] = t + 1.0 if is_norm_weight else t
        else:
            fan_in = math.prod(shape[1:])
            gain = 1.0
            # the residual branches end in conv1/proj: keep them a bit smaller so 29 blocks stay O(1)
            if key.endswith(('conv1.weight', 'proj.weight')):
                gain = 0.5
            # prior/posterior heads: spread mu so that KL is O(10) (5.6 at 32x32/L6, 12.7 at 64x64/L16 -- SURVEY section 4
            # item 5: a KL of 0.3 is a difference of nearly equal numbers and tests little); log_sigma stays modest
            if 'conv_mu' in key:
                gain = 8.0
            if 'conv_log_sigma' in key:
                gain = 1.0
            sd[key] = torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
    return sd
