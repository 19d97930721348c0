"""Repeat the fp32 gradient comparison of tests/test_model_gpu.py several times and print per-run statistics."""
import os, sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests/golden'); sys.path.insert(0, 'tests')
import numpy as np
import synth
from oracle import probunet_oracle as O
from prob_unet_mds_b200 import ProbabilisticUNet

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
L, B, H = 6, 2, 32
fx = np.load('tests/golden/probunet_32_L6_B2.npz')
sd = synth.make_weights(synth.load_schema('schema_probunet_L6.json'), seed=0)
x, t = synth.make_inputs(B, H, H, seed=1)
eps = torch.from_numpy(fx['eps'])
dt = torch.float64
leaf = {k: (v.to(dt).clone().requires_grad_(True) if 'resample_filter' not in k else v.to(dt)) for k, v in sd.items()}
r = O.elbo(leaf, x.to(dt), t.to(dt), eps.to(dt)); r['total'].backward()
ref = {k: v.grad for k, v in leaf.items() if getattr(v, 'grad', None) is not None}

def rel(a, b): return (a.double() - b.double()).norm().item() / (b.double().norm().item() + 1e-30)

m = ProbabilisticUNet(3, 3, latent_dim=L); m.load_state_dict(sd); m.set_precision('fp32')
for b in m.unet.modules():
    if hasattr(b, 'dropout'): b.dropout = 0
m.train()
named = dict(m.named_parameters())
prev = None
for it in range(int(os.environ.get('N', 6))):
    for p in m.parameters(): p.grad = None
    m.eps_override = eps
    total, recon, kl = m.elbo(x.cuda(), t.cuda()); total.backward(); torch.cuda.synchronize()
    errs = sorted(((rel(named[k].grad.cpu(), g), k) for k, g in ref.items() if g.abs().max() > 0), reverse=True)
    med = errs[len(errs) // 2][0]
    cur = {k: named[k].grad.clone() for k in ref}
    selfd = max(rel(cur[k], prev[k]) for k in ref if ref[k].abs().max() > 0) if prev else 0.0
    prev = cur
    print(f'run {it}: total {total.item():.4f} median {med:.3e} max {errs[0][0]:.3e} ({errs[0][1]}) min {errs[-1][0]:.3e} '
          f'({errs[-1][1]})  max diff vs previous run {selfd:.3e}', flush=True)
print('logits err', rel(m.last_output.cpu(), r['output'].detach()))
small = [(e, k) for e, k in errs if e < 1e-4]
print(len(small), 'of', len(errs), 'tensors below 1e-4:', [k for _, k in small][:20])
