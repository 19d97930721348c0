import os, sys, torch, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests/golden')
import synth
from prob_unet_mds_b200 import ProbabilisticUNet
torch.backends.cudnn.allow_tf32 = False
L, B, H = 6, 2, 32
sd = synth.make_weights(synth.load_schema('schema_probunet_L6.json'), seed=0)
x, t = synth.make_inputs(B, H, H, seed=1)
eps = synth.make_eps(B, L, seed=2)
poison = os.environ.get('POISON')
if poison:
    junk = torch.full((1 << 28,), float(poison), device='cuda'); del junk   # 1 GiB of junk returned to the allocator
m = ProbabilisticUNet(3, 3, latent_dim=L); m.load_state_dict(sd); m.set_precision('fp32')
for b in m.unet.modules():
    if hasattr(b, 'dropout'): b.dropout = 0
m.train()
ref = None
for it in range(4):
    for p in m.parameters(): p.grad = None
    m.eps_override = eps
    total, _, _ = m.elbo(x.cuda(), t.cuda()); total.backward(); torch.cuda.synchronize()
    g = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    if ref is None: ref = g
    worst = max(((g[k] - ref[k]).norm() / (ref[k].norm() + 1e-30)).item() for k in g)
    wk = max(g, key=lambda k: ((g[k] - ref[k]).norm() / (ref[k].norm() + 1e-30)).item())
    print(it, total.item(), 'max rel diff vs first iteration', worst, wk)
