"""Probabilistic U-Net -- drop-in for the reference module ``prob_unet`` (prob_unet.py:8-234).

Same classes, constructor arguments, attributes and ``state_dict`` keys:
    ProbabilisticUNet(input_channels, num_classes, latent_dim=6, num_filters=[64,128,256,512], beta=1.0)
        .forward(x, target=None, training=True) -> [B, num_classes, H, W]
        .elbo(x, target) -> (total, recon, kl)   0-dim fp32 tensors supporting .backward() / .item()
        .prior / .posterior : AxisAlignedConvGaussian     .fcomb : Fcomb     .unet : networks.UNet
so train_prob_unet_model.py's loops run unchanged against it.  All device work is hand-written sm_100a CUDA
behind the C ABI (include/probunet_b200.h); there is no cuDNN/cuBLAS call and no CPU fallback.

Extensions (not in the reference): ``sample_ensemble(x, num_samples)`` encodes each input once and re-runs only
Fcomb per latent sample; ``model.compute_dtype`` selects bf16 (default) or fp32 arithmetic; ``model.eps_override``
injects the standard-normal draw of ``rsample`` for bit-exact sampling tests.
"""
import torch
import torch.nn as nn
from torch.distributions import Independent, Normal

from . import engine, ops
from .networks import UNet


class AxisAlignedConvGaussian(nn.Module):
    """prob_unet.py:8-78: conv encoder -> global mean -> 1x1 heads for mu and log_sigma."""

    def __init__(self, input_channels, num_filters, latent_dim, posterior=False):
        super().__init__()
        self.input_channels = input_channels
        self.num_filters = num_filters
        self.latent_dim = latent_dim
        self.posterior = posterior
        if posterior:
            self.input_channels += input_channels
        layers = []
        cin = self.input_channels
        for cout in num_filters:
            layers += [nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.ReLU(inplace=True),
                       nn.AvgPool2d(kernel_size=2, stride=2)]
            cin = cout
        self.encoder = nn.Sequential(*layers)   # parameter containers; executed by engine.GaussianEngine
        self.conv_mu = nn.Conv2d(num_filters[-1], latent_dim, kernel_size=1)
        self.conv_log_sigma = nn.Conv2d(num_filters[-1], latent_dim, kernel_size=1)
        self._engine = None

    def engine(self, dtype):
        if self._engine is None or self._engine.dtype != dtype:
            self._engine = engine.GaussianEngine(self, dtype)
        return self._engine

    def _input(self, x, target, dtype):
        if self.posterior:
            if target is None:
                raise ValueError('the posterior net needs the target (prob_unet.py:57-58)')
            return engine.input_nhwc(x, dtype, extra=target)
        return engine.input_nhwc(x, dtype)

    def forward(self, x, target=None, dtype=None):
        """Inference-only convenience (no autograd): returns Independent(Normal(mu, exp(log_sigma)), 1)."""
        dtype = dtype or engine.default_compute_dtype()
        mu, ls, _ = self.engine(dtype).forward(self._input(x, target, dtype), save=False)
        z, sigma = ops.rsample(mu, ls, torch.zeros_like(mu))
        return Independent(Normal(loc=mu, scale=sigma, validate_args=False), 1)


class Fcomb(nn.Module):
    """prob_unet.py:80-121: three 1x1 convs over [features ; tiled z] (executed as one fused kernel)."""

    def __init__(self, unet_output_channels, latent_dim, num_classes):
        super().__init__()
        if unet_output_channels != 64:
            raise NotImplementedError('the fused Fcomb kernel is built for 64 feature channels (num_filters[0])')
        self.latent_dim = latent_dim
        self.num_classes = num_classes
        self.layers = nn.Sequential(
            nn.Conv2d(unet_output_channels + latent_dim, unet_output_channels, kernel_size=1),
            nn.ReLU(inplace=True),
            nn.Conv2d(unet_output_channels, unet_output_channels, kernel_size=1),
            nn.ReLU(inplace=True),
            nn.Conv2d(unet_output_channels, num_classes, kernel_size=1),
        )

    def params(self):
        l0, l1, l2 = self.layers[0], self.layers[2], self.layers[4]
        return l0.weight, l0.bias, l1.weight, l1.bias, l2.weight, l2.bias

    def forward(self, feature_map, z):
        """feature_map: [B, 64, H, W] fp32 NCHW, z: [B, L] -> [B, num_classes, H, W] (inference convenience)."""
        dtype = engine.default_compute_dtype()
        feat = ops.nchw_to_nhwc(feature_map.contiguous(), dtype)
        out, _, _ = ops.fcomb_fwd(feat, z.contiguous(), *[p.detach() for p in self.params()])
        return out


class ProbabilisticUNet(nn.Module):
    """prob_unet.py:123-234."""

    def __init__(self, input_channels, num_classes, latent_dim=6, num_filters=[64, 128, 256, 512], beta=1.0):
        super().__init__()
        self.input_channels = input_channels
        self.num_classes = num_classes
        self.latent_dim = latent_dim
        self.beta = beta
        self.unet = UNet(img_resolution=(64, 64), in_channels=input_channels, out_channels=num_filters[0],
                         label_dim=0, use_diffuse=False)
        self.prior = AxisAlignedConvGaussian(input_channels, num_filters, latent_dim, posterior=False)
        self.posterior = AxisAlignedConvGaussian(input_channels, num_filters, latent_dim, posterior=True)
        self.fcomb = Fcomb(num_filters[0], latent_dim, num_classes)
        self.compute_dtype = engine.default_compute_dtype()
        self.eps_override = None          # optional [B, L] standard-normal draw used by the next rsample
        self.validate_args = True         # keep the reference's Normal(validate_args) ValueError semantics
        self._zero_cache = {}
        self._flag = None
        self._flag_host = None
        self._flag_evt = None
        if torch.cuda.is_available():     # the reference pins sub-modules to cuda at construction (prob_unet.py:6)
            self.to('cuda')

    # ---- helpers ----------------------------------------------------------------------------------------------------
    def set_precision(self, name):
        self.compute_dtype = {'bf16': torch.bfloat16, 'fp32': torch.float32}[name]
        self.unet.compute_dtype = self.compute_dtype
        return self

    def _draw_eps(self, B, device):
        if self.eps_override is not None:
            eps = self.eps_override.to(device=device, dtype=torch.float32).contiguous()
            self.eps_override = None
            return eps
        # same draw as torch.distributions.utils._standard_normal (prob_unet.py:188,193,221 via Normal.rsample)
        return torch.empty([B, self.latent_dim], dtype=torch.float32, device=device).normal_()

    def _flag_for(self, device):
        if self._flag is None or self._flag.device != device:
            self._flag = torch.zeros(1, dtype=torch.int32, device=device)
            self._flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._flag_evt = None
        return self._flag

    def _check_flag(self):
        """Normal(validate_args): loc must be real and scale > 0, else ValueError (prob_unet.py:77).

        The reference pays four device->host syncs per step for this.  Here the kernels raise a device flag; it
        is copied back asynchronously and examined without blocking, so a violation surfaces as a ValueError on
        the first elbo()/forward() call after the offending step has finished on the device."""
        if not self.validate_args or self._flag is None:
            return
        if self._flag_evt is not None and self._flag_evt.query():
            bad = int(self._flag_host[0]) != 0
            self._flag_evt = None
            if bad:
                self._flag.zero_()
                self._flag_host.zero_()
                raise ValueError('Expected parameter loc to be real and scale to be positive (Normal validate_args)')
        if self._flag_evt is None:
            self._flag_host.copy_(self._flag, non_blocking=True)
            self._flag_evt = torch.cuda.Event()
            self._flag_evt.record()

    def _dist(self, mu, sigma):
        return Independent(Normal(loc=mu, scale=sigma, validate_args=False), 1)

    def _named(self):
        return list(self.named_parameters())

    # ---- forward (prob_unet.py:168-196) -------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, target=None, training=True):
        engine._require_cuda(x, 'input')
        dt = self.compute_dtype
        self.unet.compute_dtype = dt
        B = x.shape[0]
        ue = self.unet.engine()
        ue._step += 1
        feat, _ = ue.forward(engine.input_nhwc(x, dt), self.unet.training, False,
                             seed_base=engine._seed_base(ue._step))
        self._flag_for(x.device)
        if training and target is not None:
            net = self.posterior
            mu, ls, _ = net.engine(dt).forward(net._input(x, target, dt), save=False)
        else:
            net = self.prior
            mu, ls, _ = net.engine(dt).forward(net._input(x, None, dt), save=False)
        eps = self._draw_eps(B, x.device)
        z, sigma = ops.rsample(mu, ls, eps, self._flag)
        if training and target is not None:
            self.posterior_latent_space = self._dist(mu, sigma)
        else:
            self.prior_latent_space = self._dist(mu, sigma)
        self.last_z = z
        out, _, _ = ops.fcomb_fwd(feat, z, *[p.detach() for p in self.fcomb.params()])
        self._check_flag()
        return out

    # ---- sample() ------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, x, num_samples=1, eps=None):
        """BASELINE.json's north_star names a ``sample()`` method; the reference has none -- its sampling is
        ``model(inputs, training=False)`` called num_samples times (train_prob_unet_model.py:176-182), i.e. z ~ prior(x)
        decoded by Fcomb.  ``sample(x)`` returns one such draw [B, num_classes, H, W] (same result as
        ``forward(x, training=False)`` for the same eps); ``sample(x, S)`` returns [B, S, num_classes, H, W] with the
        U-Net and the prior evaluated once per input (``sample_ensemble``).  eps: optional [B, L] / [B, S, L]."""
        if num_samples == 1 and (eps is None or eps.dim() == 2):
            if eps is not None:
                self.eps_override = eps
            return self.forward(x, training=False)
        return self.sample_ensemble(x, num_samples, eps=eps)

    def invalidate_packed_weights(self):
        """Drop the packed (bf16 / re-laid-out) copies of the conv weights.  They are refreshed automatically when a
        parameter's version counter changes (optimizer steps, load_state_dict, in-place ops); call this after writing
        parameters through ``.data`` or raw pointers, which bypass the version counter."""
        for eng in (getattr(self.unet, '_engine', None), self.prior._engine, self.posterior._engine):
            if eng is not None:
                eng.cache.clear()

    # ---- ensemble sampling (encode once, S x Fcomb) ---------------------------------------------------
    @torch.no_grad()
    def sample_ensemble(self, x, num_samples, eps=None):
        """Returns [B, num_samples, num_classes, H, W]: z_s ~ prior(x), U-Net and prior evaluated once per input."""
        engine._require_cuda(x, 'input')
        dt = self.compute_dtype
        self.unet.compute_dtype = dt
        B = x.shape[0]
        ue = self.unet.engine()
        feat, _ = ue.forward(engine.input_nhwc(x, dt), False, False)
        mu, ls, _ = self.prior.engine(dt).forward(self.prior._input(x, None, dt), save=False)
        return self.decode_ensemble(feat, mu, ls, num_samples, eps)

    @torch.no_grad()
    def decode_ensemble(self, feat, mu, ls, num_samples, eps=None):
        B = feat.shape[0]
        if eps is None:
            eps = torch.empty([B, num_samples, self.latent_dim], dtype=torch.float32, device=feat.device).normal_()
        mu_s = mu[:, None, :].expand(B, num_samples, self.latent_dim).contiguous()
        ls_s = ls[:, None, :].expand(B, num_samples, self.latent_dim).contiguous()
        z, _ = ops.rsample(mu_s, ls_s, eps.contiguous())
        out, _, _ = ops.fcomb_fwd(feat, z, *[p.detach() for p in self.fcomb.params()], S=num_samples)
        return out

    # ---- elbo (prob_unet.py:198-234) ----------------------------------------------------------------------------------
    def elbo(self, x, target):
        engine._require_cuda(x, 'input')
        self.unet.compute_dtype = self.compute_dtype
        named = self._named()
        params = [p for _, p in named]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            total, recon, kl = _ElboFunction.apply(self, x, target, *params)
        else:       # evaluation (eval_probunet_model runs elbo under @torch.no_grad): no tape, no saved activations
            total, recon, kl, _ = _elbo_forward(self, x, target, save=False)
        self._check_flag()
        return total, recon, kl


def _elbo_forward(model, x, target, save):
    """The forward half of elbo (prob_unet.py:214-232).  save=False (eval under torch.no_grad, as the reference's
    eval_probunet_model does, train_prob_unet_model.py:109) keeps no backward tape and no Fcomb hidden activations."""
    dt = model.compute_dtype
    dev = x.device
    B, _, H, W = x.shape
    ue = model.unet.engine()
    ue._step += 1
    x = x.contiguous()
    target = target.contiguous()
    feat, utape = ue.forward(engine.input_nhwc(x, dt), model.unet.training, save, seed_base=engine._seed_base(ue._step))
    pe, qe = model.prior.engine(dt), model.posterior.engine(dt)
    mu_p, ls_p, ptape = pe.forward(model.prior._input(x, None, dt), save=save)
    mu_q, ls_q, qtape = qe.forward(model.posterior._input(x, target, dt), save=save)
    model._flag_for(dev)
    eps = model._draw_eps(B, dev)
    z, sigma_q = ops.rsample(mu_q, ls_q, eps, model._flag)
    _, sigma_p = ops.rsample(mu_p, ls_p, eps, model._flag)
    w0, b0, w1, b1, w2, b2 = [p.detach() for p in model.fcomb.params()]
    out, h1, h2 = ops.fcomb_fwd(feat, z, w0, b0, w1, b1, w2, b2, save_hidden=save)
    acc = ops.zeros_f64(2, dev)
    ops.mse_fwd_bwd(out, target, acc[0:1])
    ops.kl_fwd_bwd(mu_q, ls_q, mu_p, ls_p, acc[1:2], want_grads=False)
    total, recon, kl = ops.loss_finalize(acc, model.beta)
    # side-effect attributes of the reference (prob_unet.py:217-218)
    model.prior_latent_space = model._dist(mu_p, sigma_p)
    model.posterior_latent_space = model._dist(mu_q, sigma_q)
    model.last_output = out
    model.last_z = z
    saved = None
    if save:
        saved = dict(utape=utape, ptape=ptape, qtape=qtape, feat=feat, z=z, eps=eps, sigma_q=sigma_q, out=out,
                     h1=h1, h2=h2, target=target, mu_p=mu_p, ls_p=ls_p, mu_q=mu_q, ls_q=ls_q, HW=H * W)
    return total, recon, kl, saved


class _ElboFunction(torch.autograd.Function):
    """One autograd node for the whole ELBO: forward runs the kernels and keeps the tape, backward is hand-derived."""

    @staticmethod
    def forward(ctx, model, x, target, *params):
        total, recon, kl, saved = _elbo_forward(model, x, target, save=True)
        ctx.model = model
        ctx.saved = saved
        return total, recon, kl

    @staticmethod
    def backward(ctx, g_total, g_recon, g_kl):
        model = ctx.model
        s = ctx.saved
        ctx.saved = None
        dt = model.compute_dtype
        dev = s['feat'].device
        grads = engine.new_grad_sink(model)
        fc = model.fcomb
        w0p, b0p, w1p, b1p, w2p, b2p = fc.params()
        w0, w1, w2 = w0p.detach(), w1p.detach(), w2p.detach()
        Lz = model.latent_dim
        scales = ops.loss_bwd_scales(g_total, g_recon, g_kl, model.beta, dev)
        # reconstruction branch: d recon / d logits, then Fcomb backward (prob_unet.py:224-227)
        scratch = ops.zeros_f64(2, dev)
        feat, h1, h2 = s['feat'], s['h1'], s['h2']
        nc = w2p.shape[0]
        # the num_classes (3) logit channels are carried in a zero-padded 64-channel tile where the tensor-core kernels
        # apply, so that layer 2's weight and data gradients do not run on the CUDA-core kernels
        Cp = 64 if (nc < 64 and ops.conv_tc_applies(feat, 0, 64)) else nc
        dlogits = ops.mse_fwd_bwd(s['out'], s['target'], scratch[0:1], dtype=dt, gscale=scales[0:1], Cdst=Cp)
        # layer 2: 64 -> num_classes
        g = grads.alloc(w2p)
        ops.unpack_wgrad(ops.conv2d_wgrad(h2, dlogits, 1), g)          # rows beyond num_classes are padding and ignored
        grads[id(w2p)] = g
        if Cp != nc:
            grads[id(b2p)] = ops.clone(ops.bias_grad(dlogits)[:nc].contiguous(), out=grads.alloc(b2p))
            w2pad = ops.zeros((Cp, 64, 1, 1), torch.float32, dev)
            ops.clone(w2.reshape(-1), out=w2pad.reshape(-1)[:w2.numel()])
            dh2 = ops.conv2d(dlogits, ops.pack_weight(w2pad, 1, dt), 64, 1)
        else:
            grads[id(b2p)] = ops.bias_grad(dlogits, db=grads.alloc(b2p))
            dh2 = ops.conv2d(dlogits, ops.pack_weight(w2, 1, dt), 64, 1)
        dpre2 = ops.relu_mask(dh2, h2, out=dh2)
        # layer 1: 64 -> 64
        g = grads.alloc(w1p)
        ops.unpack_wgrad(ops.conv2d_wgrad(h1, dpre2, 1), g)
        grads[id(w1p)] = g
        grads[id(b1p)] = ops.bias_grad(dpre2, db=grads.alloc(b1p))
        dh1 = ops.conv2d(dpre2, ops.pack_weight(w1, 1, dt), 64, 1)
        dpre1 = ops.relu_mask(dh1, h1, out=dh1)
        # layer 0: [feat ; z] -> 64, split into the feature half (a 1x1 conv) and the z half (per-sample bias)
        g0 = grads.alloc(w0p)
        gb0 = grads.alloc(b0p)
        ops.unpack_wgrad(ops.conv2d_wgrad(feat, dpre1, 1), g0, Ci=64, dst_co_stride=64 + Lz)
        rmean = ops.global_mean(dpre1)
        dz = ops.fcomb_z_bwd(rmean, float(s['HW']), s['z'], w0, g0, gb0)
        grads[id(w0p)] = g0
        grads[id(b0p)] = gb0
        dfeat = ops.conv2d(dpre1, ops.pack_weight(w0, 1, dt, Ci=64, src_co_stride=64 + Lz), 64, 1)
        # KL branch (prob_unet.py:230) and the reparameterisation (prob_unet.py:221)
        dmu_q, dls_q, dmu_p, dls_p = ops.kl_fwd_bwd(s['mu_q'], s['ls_q'], s['mu_p'], s['ls_p'], scratch[1:2],
                                                   gscale=scales[1:2])
        ops.rsample_bwd(dz, s['eps'], s['sigma_q'], dmu_q, dls_q)
        model.posterior.engine(dt).backward(s['qtape'], dmu_q, dls_q, grads)
        model.prior.engine(dt).backward(s['ptape'], dmu_p, dls_p, grads)
        model.unet.engine().backward(s['utape'], dfeat, grads)
        grads.finish()
        named = model._named()
        return (None, None, None) + tuple(engine._collect_grads(named, grads, model._zero_cache))
