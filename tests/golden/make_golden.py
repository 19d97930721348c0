"""Generates the golden fixtures in this directory from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden.py

It imports prob_unet.ProbabilisticUNet and baseline/deterministic_unet.UNet from /root/reference, loads
seeded weights (tests/golden/synth.py), switches dropout off (attribute networks.py:144) and records
outputs, losses and a digest of every gradient.  The reference has no tests of its own (SURVEY.md
section 4), so these files are what pins oracle/probunet_oracle.py.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import synth  # noqa: E402

REF = '/root/reference'


def grad_digest_indices(name, numel, k=6):
    h = int(hashlib.sha256(name.encode()).hexdigest()[:12], 16)
    return [(h * (i + 1) * 2654435761) % numel for i in range(k)]


def grad_digest(named_grads):
    names, rows = [], []
    for name, g in named_grads:
        if g is None:
            continue
        flat = g.detach().reshape(-1).double()
        idx = grad_digest_indices(name, flat.numel())
        rows.append([flat.norm().item(), flat.sum().item(), flat.abs().sum().item()] + [flat[i].item() for i in idx])
        names.append(name)
    return names, np.asarray(rows, dtype=np.float64)


def dump_schema(model, fname):
    schema = [(k, list(v.shape)) for k, v in model.state_dict().items()]
    with open(os.path.join(HERE, fname), 'w') as f:
        json.dump(schema, f)
    return [(k, tuple(s)) for k, s in schema]


def disable_dropout(model):
    for m in model.modules():
        if hasattr(m, 'dropout'):
            m.dropout = 0


def _autocast_bf16(model, x, t, seed):
    """The reference's OWN bf16 behaviour: the unmodified model under torch.autocast(bfloat16) on the CPU, same weights,
    same inputs, same eps.  Its distance from the fp32 run is the noise floor that a bf16 pipeline is measured against
    (tests/test_model_gpu.py: the product's bf16 logits error must not exceed the reference's own autocast error).
    Under autocast mu is bf16, and torch's CPU normal_() produces a different stream for bf16 tensors than for fp32
    ones, so the fp32 draw of the same seed is injected through torch.distributions' _standard_normal hook."""
    import torch.distributions.normal as tdn
    torch.manual_seed(seed)
    eps32 = torch.empty([x.shape[0], model.latent_dim]).normal_()
    orig = tdn._standard_normal
    tdn._standard_normal = lambda shape, dtype, device: eps32.to(dtype)
    try:
        with torch.no_grad(), torch.autocast('cpu', dtype=torch.bfloat16):
            total, recon, kl = model.elbo(x, t)
            y = model(x, t, training=True)
    finally:
        tdn._standard_normal = orig
    return dict(ac_total=float(total), ac_recon=float(recon), ac_kl=float(kl), ac_post_output=y.float().numpy())


def probunet_case(tag, B, H, L, grads=True):
    sys.path.insert(0, REF)
    from prob_unet import ProbabilisticUNet
    torch.manual_seed(0)
    model = ProbabilisticUNet(3, 3, latent_dim=L, num_filters=[64, 128, 256, 512])
    schema = dump_schema(model, f'schema_probunet_L{L}.json')
    model.load_state_dict(synth.make_weights(schema, seed=0))
    disable_dropout(model)
    model.train()
    x, t = synth.make_inputs(B, H, H, seed=1)
    # rsample draws torch.empty([B,L]).normal_() from the global generator (first draw: dropout is off)
    torch.manual_seed(1234)
    eps = torch.empty([B, L]).normal_()
    torch.manual_seed(1234)
    total, recon, kl = model.elbo(x, t)
    out = {
        'eps': eps.numpy(), 'total': total.item(), 'recon': recon.item(), 'kl': kl.item(),
        'mu_p': model.prior_latent_space.base_dist.loc.detach().numpy(),
        'sigma_p': model.prior_latent_space.base_dist.scale.detach().numpy(),
        'mu_q': model.posterior_latent_space.base_dist.loc.detach().numpy(),
        'sigma_q': model.posterior_latent_space.base_dist.scale.detach().numpy(),
    }
    if grads:
        total.backward()
        names, rows = grad_digest((n, p.grad) for n, p in model.named_parameters())
        out['grad_digest'] = rows
        out['grad_names'] = np.asarray(names)
        out['none_grad_names'] = np.asarray([n for n, p in model.named_parameters() if p.grad is None])
    # sampling path: forward(training=False) -> prior branch (prob_unet.py:190-193)
    model.eval()
    with torch.no_grad():
        torch.manual_seed(77)
        eps_s = torch.empty([B, L]).normal_()
        torch.manual_seed(77)
        y = model(x, training=False)
        feat = model.unet(x)
    out['sample_eps'] = eps_s.numpy()
    out['sample_output'] = y.numpy()
    out['unet_features_digest'] = np.asarray([feat.double().norm().item(), feat.double().sum().item()])
    out['unet_features_corner'] = feat[:, :, :4, :4].numpy()
    # posterior-branch forward (training=True with a target)
    with torch.no_grad():
        torch.manual_seed(78)
        y2 = model(x, t, training=True)
    torch.manual_seed(78)
    out['post_eps'] = torch.empty([B, L]).normal_().numpy()
    out['post_output'] = y2.numpy()
    # noise floor of bf16 arithmetic, measured on the reference itself (posterior-branch logits for post_eps, losses)
    out.update(_autocast_bf16(model, x, t, 78))
    d = out['ac_post_output'].astype(np.float64) - out['post_output'].astype(np.float64)
    out['ac_post_output_relerr'] = float(np.linalg.norm(d) / np.linalg.norm(out['post_output'].astype(np.float64)))
    del out['ac_post_output']          # only its error is needed; keeps the fixture small
    with torch.no_grad():
        torch.manual_seed(78)
        tot78, rec78, kl78 = model.elbo(x, t)
    out['ac_total_relerr'] = abs(out['ac_total'] - float(tot78)) / abs(float(tot78))
    out['ac_kl_relerr'] = abs(out['ac_kl'] - float(kl78)) / abs(float(kl78))
    print(tag, 'reference under CPU autocast(bf16): logits rel err', out['ac_post_output_relerr'], 'ELBO rel err',
          out['ac_total_relerr'], 'KL rel err', out['ac_kl_relerr'])
    np.savez_compressed(os.path.join(HERE, f'{tag}.npz'), **out)
    print(tag, 'total', out['total'], 'recon', out['recon'], 'kl', out['kl'])


def detunet_case(tag, B, H):
    sys.path.insert(0, os.path.join(REF, 'baseline'))
    import deterministic_unet as du
    torch.manual_seed(0)
    model = du.UNet(img_resolution=(64, 64), in_channels=3, out_channels=3, label_dim=0, use_diffuse=False)
    schema = dump_schema(model, 'schema_detunet.json')
    model.load_state_dict(synth.make_weights(schema, seed=3))
    disable_dropout(model)
    model.train()
    x, t = synth.make_inputs(B, H, H, seed=5)
    y = model(x, class_labels=None)
    loss = torch.nn.MSELoss()(y, t)          # baseline/main.py:69 (mean reduction)
    loss.backward()
    names, rows = grad_digest((n, p.grad) for n, p in model.named_parameters())
    np.savez_compressed(os.path.join(HERE, f'{tag}.npz'), output=y.detach().numpy(), loss=loss.item(),
                        grad_digest=rows, grad_names=np.asarray(names))
    print(tag, 'loss', loss.item())


if __name__ == '__main__':
    torch.set_num_threads(8)
    probunet_case('probunet_32_L6_B2', 2, 32, 6)
    probunet_case('probunet_64_L16_B1', 1, 64, 16)
    probunet_case('probunet_128_L16_B1', 1, 128, 16)
    detunet_case('detunet_64_B1', 1, 64)
