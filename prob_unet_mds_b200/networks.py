"""U-Net backbone (ADM / DDPM++ style) of the Probabilistic U-Net, B200-native.

Mirrors the module tree, constructor arguments, initialisation and ``state_dict`` schema of the reference's
``networks.py`` (UNet :224-333, UNetBlock :132-185, Conv2d :49-90, GroupNorm :95-105, Linear :31-44), so that
``load_state_dict(reference.state_dict())`` works and checkpoints interchange.  The modules below are *parameter
containers*: all device work is done by ``engine.UNetEngine`` through the C ABI (hand-written sm_100a kernels),
with an explicit, hand-derived backward pass.  ``UNet.forward`` is provided for drop-in use of the bare backbone
(``baseline/deterministic_unet.py`` usage, trainmodel.py:157) and routes through the same engine.
"""
import math

import numpy as np
import torch

from . import engine


def weight_init(shape, mode, fan_in, fan_out):
    """Same four initialisers as networks.py:21-26."""
    if mode == 'xavier_uniform':
        return math.sqrt(6 / (fan_in + fan_out)) * (torch.rand(*shape) * 2 - 1)
    if mode == 'xavier_normal':
        return math.sqrt(2 / (fan_in + fan_out)) * torch.randn(*shape)
    if mode == 'kaiming_uniform':
        return math.sqrt(3 / fan_in) * (torch.rand(*shape) * 2 - 1)
    if mode == 'kaiming_normal':
        return math.sqrt(1 / fan_in) * torch.randn(*shape)
    raise ValueError(f'Invalid init mode "{mode}"')


class Linear(torch.nn.Module):
    """networks.py:31-44.  In this model it only appears as `affine` with a zero embedding (and the unused
    map_layer0/1), so it never needs a GEMM: affine(emb) == bias."""

    def __init__(self, in_features, out_features, bias=True, init_mode='kaiming_normal', init_weight=1, init_bias=0):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        kw = dict(mode=init_mode, fan_in=in_features, fan_out=out_features)
        self.weight = torch.nn.Parameter(weight_init([out_features, in_features], **kw) * init_weight)
        self.bias = torch.nn.Parameter(weight_init([out_features], **kw) * init_bias) if bias else None


class Conv2d(torch.nn.Module):
    """networks.py:49-90 (non fused_resample path): optional 2x up/down with a [1,1] filter, then a dense conv."""

    def __init__(self, in_channels, out_channels, kernel, bias=True, up=False, down=False, resample_filter=(1, 1),
                 fused_resample=False, init_mode='kaiming_normal', init_weight=1, init_bias=0):
        assert not (up and down)
        if fused_resample:
            raise NotImplementedError('fused_resample is never used by the Probabilistic U-Net path')
        if list(resample_filter) != [1, 1]:
            raise NotImplementedError('only the [1,1] (nearest / 2x2 mean) resample filter is supported')
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel = kernel
        self.up = up
        self.down = down
        kw = dict(mode=init_mode, fan_in=in_channels * kernel * kernel, fan_out=out_channels * kernel * kernel)
        self.weight = torch.nn.Parameter(weight_init([out_channels, in_channels, kernel, kernel], **kw) * init_weight) \
            if kernel else None
        self.bias = torch.nn.Parameter(weight_init([out_channels], **kw) * init_bias) if kernel and bias else None
        f = torch.as_tensor(list(resample_filter), dtype=torch.float32)
        f = f.ger(f).unsqueeze(0).unsqueeze(1) / f.sum().square()
        self.register_buffer('resample_filter', f if up or down else None)


class GroupNorm(torch.nn.Module):
    """networks.py:95-105."""

    def __init__(self, num_channels, num_groups=32, min_channels_per_group=4, eps=1e-5):
        super().__init__()
        self.num_groups = min(num_groups, num_channels // min_channels_per_group)
        self.eps = eps
        self.weight = torch.nn.Parameter(torch.ones(num_channels))
        self.bias = torch.nn.Parameter(torch.zeros(num_channels))


class UNetBlock(torch.nn.Module):
    """networks.py:132-185."""

    def __init__(self, in_channels, out_channels, emb_channels, up=False, down=False, attention=False,
                 num_heads=None, channels_per_head=64, dropout=0, skip_scale=1, eps=1e-5, resample_filter=(1, 1),
                 resample_proj=False, adaptive_scale=True, init=dict(), init_zero=dict(init_weight=0), init_attn=None):
        super().__init__()
        if not adaptive_scale or skip_scale != 1 or resample_proj:
            raise NotImplementedError('only adaptive_scale=True, skip_scale=1, resample_proj=False are on the path')
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.emb_channels = emb_channels
        self.num_heads = 0 if not attention else num_heads if num_heads is not None else out_channels // channels_per_head
        if self.num_heads and (channels_per_head != 64 or out_channels % 64):
            raise NotImplementedError('attention kernels are built for 64 channels per head')
        self.dropout = dropout
        self.skip_scale = skip_scale
        self.adaptive_scale = adaptive_scale
        self.up = up
        self.down = down

        self.norm0 = GroupNorm(num_channels=in_channels, eps=eps)
        self.conv0 = Conv2d(in_channels=in_channels, out_channels=out_channels, kernel=3, up=up, down=down,
                            resample_filter=resample_filter, **init)
        self.affine = Linear(in_features=emb_channels, out_features=out_channels * 2, **init)
        self.norm1 = GroupNorm(num_channels=out_channels, eps=eps)
        self.conv1 = Conv2d(in_channels=out_channels, out_channels=out_channels, kernel=3, **init_zero)
        self.skip = None
        if out_channels != in_channels or up or down:
            kernel = 1 if out_channels != in_channels else 0
            self.skip = Conv2d(in_channels=in_channels, out_channels=out_channels, kernel=kernel, up=up, down=down,
                               resample_filter=resample_filter, **init)
        if self.num_heads:
            self.norm2 = GroupNorm(num_channels=out_channels, eps=eps)
            self.qkv = Conv2d(in_channels=out_channels, out_channels=out_channels * 3, kernel=1,
                              **(init_attn if init_attn is not None else init))
            self.proj = Conv2d(in_channels=out_channels, out_channels=out_channels, kernel=1, **init_zero)


class UNet(torch.nn.Module):
    """networks.py:224-333 with label_dim=0, augment_dim=0, use_diffuse=False (the only configuration on the path).

    `attention=False` gives baseline/deterministic_unet.py's variant (no self-attention anywhere)."""

    def __init__(self, img_resolution, in_channels, out_channels, label_dim=0, augment_dim=0, model_channels=128,
                 channel_mult=(1, 2, 3, 4), channel_mult_emb=4, num_blocks=2, attn_resolutions=(32, 16, 8),
                 dropout=0.10, label_dropout=0, use_diffuse=True, attention=True):
        super().__init__()
        if label_dim or augment_dim or use_diffuse:
            raise NotImplementedError('the Probabilistic U-Net path uses label_dim=0, augment_dim=0, use_diffuse=False')
        assert len(img_resolution) == 2
        self.label_dropout = label_dropout
        self.model_channels = model_channels
        emb_channels = model_channels * channel_mult_emb
        init = dict(init_mode='kaiming_uniform', init_weight=np.sqrt(1 / 3), init_bias=np.sqrt(1 / 3))
        init_zero = dict(init_mode='kaiming_uniform', init_weight=0, init_bias=0)
        block_kwargs = dict(emb_channels=emb_channels, channels_per_head=64, dropout=dropout, init=init,
                            init_zero=init_zero)

        self.map_noise = None
        self.map_augment = None
        self.map_layer0 = Linear(in_features=model_channels, out_features=emb_channels, **init)   # never used
        self.map_layer1 = Linear(in_features=emb_channels, out_features=emb_channels, **init)     # never used
        self.map_label = None

        self.enc = torch.nn.ModuleDict()
        cout = in_channels
        for level, mult in enumerate(channel_mult):
            resx, resy = img_resolution[0] >> level, img_resolution[1] >> level
            if level == 0:
                cin, cout = cout, model_channels * mult
                self.enc[f'{resx}x{resy}_conv'] = Conv2d(in_channels=cin, out_channels=cout, kernel=3, **init)
            else:
                self.enc[f'{resx}x{resy}_down'] = UNetBlock(in_channels=cout, out_channels=cout, down=True, **block_kwargs)
            for idx in range(num_blocks):
                cin, cout = cout, model_channels * mult
                self.enc[f'{resx}x{resy}_block{idx}'] = UNetBlock(
                    in_channels=cin, out_channels=cout, attention=attention and (resx in attn_resolutions), **block_kwargs)
        skips = [block.out_channels for block in self.enc.values()]

        self.dec = torch.nn.ModuleDict()
        for level, mult in reversed(list(enumerate(channel_mult))):
            resx, resy = img_resolution[0] >> level, img_resolution[1] >> level
            if level == len(channel_mult) - 1:
                self.dec[f'{resx}x{resy}_in0'] = UNetBlock(in_channels=cout, out_channels=cout, attention=attention,
                                                          **block_kwargs)
                self.dec[f'{resx}x{resy}_in1'] = UNetBlock(in_channels=cout, out_channels=cout, **block_kwargs)
            else:
                self.dec[f'{resx}x{resy}_up'] = UNetBlock(in_channels=cout, out_channels=cout, up=True, **block_kwargs)
            for idx in range(num_blocks + 1):
                cin = cout + skips.pop()
                cout = model_channels * mult
                self.dec[f'{resx}x{resy}_block{idx}'] = UNetBlock(
                    in_channels=cin, out_channels=cout, attention=attention and (resx in attn_resolutions), **block_kwargs)
        self.out_norm = GroupNorm(num_channels=cout)
        self.out_conv = Conv2d(in_channels=cout, out_channels=out_channels, kernel=3, **init_zero)

        self.compute_dtype = engine.default_compute_dtype()
        self._engine = None

    def engine(self):
        if self._engine is None or self._engine.dtype != self.compute_dtype:
            self._engine = engine.UNetEngine(self, self.compute_dtype)
        return self._engine

    def forward(self, x, noise_labels=None, class_labels=None, augment_labels=None):
        """[N, in_channels, H, W] fp32 -> [N, out_channels, H, W] fp32 (differentiable wrt the parameters)."""
        return engine.unet_apply(self, x)
