"""Pins oracle/probunet_oracle.py against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

import synth
from make_golden import grad_digest_indices
from oracle import probunet_oracle as O

G = os.path.join(os.path.dirname(__file__), 'golden')


def _leaf_sd(schema, seed):
    sd = synth.make_weights(schema, seed=seed)
    return {k: (v.requires_grad_(True) if v.dtype.is_floating_point and 'resample_filter' not in k else v)
            for k, v in sd.items()}


def _check_digest(sd, fx, rtol):
    names = [str(n) for n in fx['grad_names']]
    rows = fx['grad_digest']
    for name, row in zip(names, rows):
        g = sd[name].grad
        assert g is not None, name
        flat = g.reshape(-1).double()
        scale = row[0] / max(1.0, flat.numel() ** 0.5) + 1e-12   # rms of the reference gradient
        assert abs(flat.norm().item() - row[0]) <= rtol * row[0] + 1e-10, name
        for i, idx in enumerate(grad_digest_indices(name, flat.numel())):
            assert abs(flat[idx].item() - row[3 + i]) <= 50 * rtol * scale + 1e-9, (name, idx)


@pytest.mark.parametrize('tag,B,H,L,grads', [
    ('probunet_32_L6_B2', 2, 32, 6, True),
    ('probunet_64_L16_B1', 1, 64, 16, True),
    ('probunet_128_L16_B1', 1, 128, 16, False),
])
def test_probunet_oracle_matches_reference(tag, B, H, L, grads):
    if tag.startswith('probunet_128') and os.environ.get('PU_SKIP_SLOW'):
        pytest.skip('slow')
    fx = np.load(os.path.join(G, tag + '.npz'))
    sd = _leaf_sd(synth.load_schema(f'schema_probunet_L{L}.json'), 0)
    x, t = synth.make_inputs(B, H, H, seed=1)
    with torch.set_grad_enabled(grads):
        r = O.elbo(sd, x, t, torch.from_numpy(fx['eps']))
    for k in ('total', 'recon', 'kl'):
        assert abs(r[k].item() - float(fx[k])) <= 2e-6 * abs(float(fx[k])) + 1e-6, k
    np.testing.assert_allclose(r['mu_p'].detach().numpy(), fx['mu_p'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r['mu_q'].detach().numpy(), fx['mu_q'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(torch.exp(r['ls_p']).detach().numpy(), fx['sigma_p'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(torch.exp(r['ls_q']).detach().numpy(), fx['sigma_q'], rtol=1e-5, atol=1e-6)
    if grads:
        r['total'].backward()
        _check_digest(sd, fx, 2e-5)
        none_names = set(str(n) for n in fx['none_grad_names'])
        assert none_names == {'unet.map_layer0.weight', 'unet.map_layer0.bias',
                              'unet.map_layer1.weight', 'unet.map_layer1.bias'}
    with torch.no_grad():
        y, mu, ls, z = O.forward(sd, x, torch.from_numpy(fx['sample_eps']), training=False)
        np.testing.assert_allclose(y.numpy(), fx['sample_output'], rtol=1e-4, atol=2e-5)
        # z is bit-exact given identical mu, sigma, eps
        assert torch.equal(z, mu + torch.from_numpy(fx['sample_eps']) * torch.exp(ls))
        y2, *_ = O.forward(sd, x, torch.from_numpy(fx['post_eps']), target=t, training=True)
        np.testing.assert_allclose(y2.numpy(), fx['post_output'], rtol=1e-4, atol=2e-5)


def test_detunet_oracle_matches_reference():
    fx = np.load(os.path.join(G, 'detunet_64_B1.npz'))
    sd = _leaf_sd(synth.load_schema('schema_detunet.json'), 3)
    x, t = synth.make_inputs(1, 64, 64, seed=5)
    y = O.det_unet_forward(sd, x)
    np.testing.assert_allclose(y.detach().numpy(), fx['output'], rtol=1e-4, atol=2e-5)
    loss = ((y - t) ** 2).mean()
    assert abs(loss.item() - float(fx['loss'])) < 2e-6 * float(fx['loss'])
    loss.backward()
    _check_digest(sd, fx, 2e-5)
