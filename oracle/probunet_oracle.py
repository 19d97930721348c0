"""CPU oracle for the Probabilistic U-Net hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch fp32 *restatement* of the reference algorithm
(pierrelouislemaire/prob-unet-mds: prob_unet.py / networks.py), written as pure
functions over a ``state_dict``.  It exists so that the CUDA path can be checked
on a box where ``/root/reference`` is not mounted.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; the product package never does.

Pinning status: the reference ships no tests and no golden vectors (SURVEY.md
section 4), so the oracle is pinned against outputs of the reference itself run
in the authoring container: ``tests/golden/make_golden.py`` imports the
unmodified reference from ``/root/reference``, runs it on seeded synthetic
inputs/weights and commits the results under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays them through this file.

Reference citations (file:line into /root/reference):
  conv2d_resample    networks.py:68-90   (Conv2d.forward, non fused_resample branch)
  group_norm         networks.py:95-105
  attention          networks.py:112-125 (AttentionOp) + :179-184
  unet_block         networks.py:164-185
  unet_plan/forward  networks.py:225-333
  gaussian_encoder   prob_unet.py:44-78
  fcomb              prob_unet.py:100-121
  forward / elbo     prob_unet.py:168-234
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


class ReluProbe:
    """Test hook for the three ReLU sites of the path (prob_unet.py:56 encoder, :112-116 Fcomb).

    ReLU is not differentiable at 0 and a fp32 implementation may land on either side of it for a unit whose
    pre-activation is within rounding noise of zero; both are valid sub-gradients.  With ``record`` set the probe
    stores every pre-activation; with ``flips = {call_index: flat_indices}`` it evaluates the backward with the
    mask of exactly those units inverted (the forward value moves by |pre-activation| <= the probe threshold).
    """

    def __init__(self, flips=None, record=False):
        self.flips = flips or {}
        self.record = [] if record else None
        self.calls = 0


RELU_PROBE: Optional[ReluProbe] = None


def _relu(x: Tensor) -> Tensor:
    probe = RELU_PROBE
    if probe is None:
        return F.relu(x)
    idx = probe.calls
    probe.calls += 1
    if probe.record is not None:
        probe.record.append(x.detach())
    flip = probe.flips.get(idx)
    if flip is None:
        return F.relu(x)
    mask = (x > 0).reshape(-1).clone()
    mask[flip] = ~mask[flip]
    return x * mask.reshape(x.shape).to(x.dtype)


# --------------------------------------------------------------------------------------
# network plan (restates the constructor loops of networks.py:258-298)
# --------------------------------------------------------------------------------------
def unet_plan(in_channels: int, out_channels: int, model_channels: int = 128,
              channel_mult=(1, 2, 3, 4), num_blocks: int = 2,
              attn_resolutions=(32, 16, 8), img_resolution: int = 64,
              attention: bool = True) -> Dict[str, List[dict]]:
    """Returns {'enc': [...], 'dec': [...]} lists of layer specs, in execution order."""
    enc, dec = [], []
    cout = in_channels
    for level, mult in enumerate(channel_mult):
        res = img_resolution >> level
        if level == 0:
            cin, cout = cout, model_channels * mult
            enc.append(dict(name=f'{res}x{res}_conv', kind='conv', cin=cin, cout=cout))
        else:
            enc.append(dict(name=f'{res}x{res}_down', kind='block', cin=cout, cout=cout,
                            up=False, down=True, attn=False))
        for idx in range(num_blocks):
            cin, cout = cout, model_channels * mult
            enc.append(dict(name=f'{res}x{res}_block{idx}', kind='block', cin=cin, cout=cout,
                            up=False, down=False, attn=attention and (res in attn_resolutions)))
    skips = [l['cout'] for l in enc]
    for level, mult in reversed(list(enumerate(channel_mult))):
        res = img_resolution >> level
        if level == len(channel_mult) - 1:
            dec.append(dict(name=f'{res}x{res}_in0', kind='block', cin=cout, cout=cout,
                            up=False, down=False, attn=attention))
            dec.append(dict(name=f'{res}x{res}_in1', kind='block', cin=cout, cout=cout,
                            up=False, down=False, attn=False))
        else:
            dec.append(dict(name=f'{res}x{res}_up', kind='block', cin=cout, cout=cout,
                            up=True, down=False, attn=False))
        for idx in range(num_blocks + 1):
            cin = cout + skips.pop()
            cout = model_channels * mult
            dec.append(dict(name=f'{res}x{res}_block{idx}', kind='block', cin=cin, cout=cout,
                            up=False, down=False, attn=attention and (res in attn_resolutions)))
    return dict(enc=enc, dec=dec, final=cout, out_channels=out_channels)


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def conv2d_resample(x: Tensor, w: Optional[Tensor], b: Optional[Tensor],
                    up: bool = False, down: bool = False) -> Tensor:
    """networks.py:82-89: optional 2x resample with a [1,1] box filter, then dense conv, then bias.

    The depthwise transposed conv with an all-ones 2x2 filter and stride 2 is nearest-neighbour
    upsampling; the depthwise stride-2 conv with a 0.25 filter is 2x2 average pooling.
    """
    if up:
        x = F.interpolate(x, scale_factor=2, mode='nearest')
    if down:
        x = F.avg_pool2d(x, 2)
    if w is not None:
        x = F.conv2d(x, w, padding=w.shape[-1] // 2)
    if b is not None:
        x = x + b.reshape(1, -1, 1, 1)
    return x


def group_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """networks.py:97-104: num_groups = min(32, C // 4)."""
    return F.group_norm(x, min(32, x.shape[1] // 4), w, b, eps)


def attention(qkv: Tensor, num_heads: int) -> Tensor:
    """networks.py:180-182 with AttentionOp (:114-116): softmax_k(q^T k / sqrt(d)) then w . v.

    qkv is [B, 3C, H, W]; the reference views channels as (head, d, {q,k,v}).
    """
    B, C3, H, W = qkv.shape
    C = C3 // 3
    q, k, v = qkv.reshape(B * num_heads, C // num_heads, 3, H * W).unbind(2)
    logits = torch.einsum('ncq,nck->nqk', q, k / math.sqrt(k.shape[1]))
    w = logits.softmax(dim=2)
    a = torch.einsum('nqk,nck->ncq', w, v)
    return a.reshape(B, C, H, W)


def unet_block(sd: Dict[str, Tensor], p: str, x: Tensor, spec: dict, emb: Tensor,
               dropout_mask: Optional[Tensor] = None, dropout_p: float = 0.0) -> Tensor:
    """networks.py:164-185."""
    orig = x
    up, down = spec['up'], spec['down']
    h = F.silu(group_norm(x, sd[p + 'norm0.weight'], sd[p + 'norm0.bias']))
    h = conv2d_resample(h, sd[p + 'conv0.weight'], sd[p + 'conv0.bias'], up, down)
    params = emb @ sd[p + 'affine.weight'].t() + sd[p + 'affine.bias']
    scale, shift = params[:, :, None, None].chunk(2, dim=1)
    h = F.silu(torch.addcmul(shift, group_norm(h, sd[p + 'norm1.weight'], sd[p + 'norm1.bias']), scale + 1))
    if dropout_mask is not None:
        h = h * dropout_mask / (1.0 - dropout_p)
    h = conv2d_resample(h, sd[p + 'conv1.weight'], sd[p + 'conv1.bias'])
    if (p + 'skip.weight') in sd:
        s = conv2d_resample(orig, sd[p + 'skip.weight'], sd[p + 'skip.bias'], up, down)
    elif up or down:
        s = conv2d_resample(orig, None, None, up, down)
    else:
        s = orig
    x = h + s
    if spec['attn']:
        heads = spec['cout'] // 64
        qkv = conv2d_resample(group_norm(x, sd[p + 'norm2.weight'], sd[p + 'norm2.bias']),
                              sd[p + 'qkv.weight'], sd[p + 'qkv.bias'])
        a = attention(qkv, heads)
        x = conv2d_resample(a, sd[p + 'proj.weight'], sd[p + 'proj.bias']) + x
    return x


def unet_forward(sd: Dict[str, Tensor], x: Tensor, plan: dict, prefix: str = 'unet.',
                 dropout_masks: Optional[Dict[str, Tensor]] = None, dropout_p: float = 0.10) -> Tensor:
    """networks.py:300-333 with label_dim=0, use_diffuse=False: emb = silu(zeros) = 0."""
    emb_dim = sd[prefix + 'map_layer1.weight'].shape[1]
    emb = F.silu(torch.zeros(1, emb_dim, dtype=x.dtype, device=x.device))
    skips = []
    for spec in plan['enc']:
        p = f"{prefix}enc.{spec['name']}."
        if spec['kind'] == 'conv':
            x = conv2d_resample(x, sd[p + 'weight'], sd[p + 'bias'])
        else:
            m = dropout_masks.get(p) if dropout_masks else None
            x = unet_block(sd, p, x, spec, emb, m, dropout_p)
        skips.append(x)
    for spec in plan['dec']:
        p = f"{prefix}dec.{spec['name']}."
        if x.shape[1] != spec['cin']:
            x = torch.cat([x, skips.pop()], dim=1)
        m = dropout_masks.get(p) if dropout_masks else None
        x = unet_block(sd, p, x, spec, emb, m, dropout_p)
    x = F.silu(group_norm(x, sd[prefix + 'out_norm.weight'], sd[prefix + 'out_norm.bias']))
    return conv2d_resample(x, sd[prefix + 'out_conv.weight'], sd[prefix + 'out_conv.bias'])


def gaussian_encoder(sd: Dict[str, Tensor], prefix: str, x: Tensor, target: Optional[Tensor] = None):
    """prob_unet.py:44-78 -> (mu, log_sigma), both [B, L]."""
    if target is not None:
        x = torch.cat([x, target], dim=1)
    i = 0
    while f'{prefix}encoder.{i}.weight' in sd:
        x = F.conv2d(x, sd[f'{prefix}encoder.{i}.weight'], sd[f'{prefix}encoder.{i}.bias'], padding=1)
        x = F.avg_pool2d(_relu(x), 2)
        i += 3
    h = x.mean(dim=[2, 3], keepdim=True)
    mu = F.conv2d(h, sd[prefix + 'conv_mu.weight'], sd[prefix + 'conv_mu.bias'])
    ls = F.conv2d(h, sd[prefix + 'conv_log_sigma.weight'], sd[prefix + 'conv_log_sigma.bias'])
    return mu[:, :, 0, 0], ls[:, :, 0, 0]


def fcomb(sd: Dict[str, Tensor], feat: Tensor, z: Tensor, prefix: str = 'fcomb.') -> Tensor:
    """prob_unet.py:100-121: tile z over HxW, concat, three 1x1 convs with ReLU between."""
    zt = z[:, :, None, None].expand(-1, -1, feat.shape[2], feat.shape[3])
    h = torch.cat([feat, zt], dim=1)
    h = _relu(F.conv2d(h, sd[prefix + 'layers.0.weight'], sd[prefix + 'layers.0.bias']))
    h = _relu(F.conv2d(h, sd[prefix + 'layers.2.weight'], sd[prefix + 'layers.2.bias']))
    return F.conv2d(h, sd[prefix + 'layers.4.weight'], sd[prefix + 'layers.4.bias'])


def rsample(mu: Tensor, log_sigma: Tensor, eps: Tensor) -> Tensor:
    """torch Normal.rsample: loc + eps * scale with scale = exp(log_sigma) (prob_unet.py:77)."""
    return mu + eps * torch.exp(log_sigma)


def kl_normal(mu_q: Tensor, ls_q: Tensor, mu_p: Tensor, ls_p: Tensor) -> Tensor:
    """kl(Independent(Normal(q)) || Independent(Normal(p))) per sample (torch kl.py _kl_normal_normal)."""
    sq, sp = torch.exp(ls_q), torch.exp(ls_p)
    var_ratio = (sq / sp).pow(2)
    t1 = ((mu_q - mu_p) / sp).pow(2)
    return (0.5 * (var_ratio + t1 - 1 - var_ratio.log())).sum(-1)


# --------------------------------------------------------------------------------------
# the two public entry points of the path
# --------------------------------------------------------------------------------------
def _plan_from_sd(sd, input_channels):
    mc = sd['unet.enc.64x64_conv.weight'].shape[0]
    out_ch = sd['unet.out_conv.weight'].shape[0]
    attn = any(k.endswith('qkv.weight') for k in sd)
    return unet_plan(input_channels, out_ch, model_channels=mc, attention=attn)


def forward(sd, x, eps, target=None, training=True, dropout_masks=None):
    """prob_unet.py:168-196.  Returns (output, mu, log_sigma, z)."""
    plan = _plan_from_sd(sd, x.shape[1])
    feat = unet_forward(sd, x, plan, dropout_masks=dropout_masks)
    if training and target is not None:
        mu, ls = gaussian_encoder(sd, 'posterior.', x, target)
    else:
        mu, ls = gaussian_encoder(sd, 'prior.', x)
    z = rsample(mu, ls, eps)
    return fcomb(sd, feat, z), mu, ls, z


def elbo(sd, x, target, eps, beta=1.0, dropout_masks=None):
    """prob_unet.py:198-234.  Returns dict with total/recon/kl and the intermediates."""
    plan = _plan_from_sd(sd, x.shape[1])
    feat = unet_forward(sd, x, plan, dropout_masks=dropout_masks)
    mu_p, ls_p = gaussian_encoder(sd, 'prior.', x)
    mu_q, ls_q = gaussian_encoder(sd, 'posterior.', x, target)
    z = rsample(mu_q, ls_q, eps)
    out = fcomb(sd, feat, z)
    recon = ((out - target) ** 2).sum()
    kl = kl_normal(mu_q, ls_q, mu_p, ls_p).sum()
    total = recon + beta * kl
    return dict(total=total, recon=recon, kl=kl, output=out, feat=feat, z=z,
                mu_p=mu_p, ls_p=ls_p, mu_q=mu_q, ls_q=ls_q)


def det_unet_forward(sd, x, prefix=''):
    """baseline/deterministic_unet.py:224-331: the same U-Net with model_channels=64, no attention."""
    mc = sd[prefix + 'enc.64x64_conv.weight'].shape[0]
    out_ch = sd[prefix + 'out_conv.weight'].shape[0]
    plan = unet_plan(x.shape[1], out_ch, model_channels=mc, attention=False)
    return unet_forward(sd, x, plan, prefix=prefix)
