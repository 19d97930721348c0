import os, sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests/golden')
import synth
from prob_unet_mds_b200 import ProbabilisticUNet, ops
L, B, H = 6, 2, 32
sd = synth.make_weights(synth.load_schema('schema_probunet_L6.json'), seed=0)
x, t = synth.make_inputs(B, H, H, seed=1)
eps = synth.make_eps(B, L, seed=2)
m = ProbabilisticUNet(3, 3, latent_dim=L); m.load_state_dict(sd); m.set_precision('fp32')
for b in m.unet.modules():
    if hasattr(b, 'dropout'): b.dropout = 0
m.train()
rec = []
names = ['conv2d', 'conv2d_wgrad', 'bias_grad', 'gn_stats', 'gn_apply', 'gn_bwd', 'attention_fwd', 'attention_bwd',
         'avgpool2', 'upsample2', 'relu_pool_bwd', 'global_mean', 'relu_mean_bwd', 'relu_mask', 'heads_fwd', 'heads_bwd',
         'fcomb_fwd', 'fcomb_z_bwd', 'mse_fwd_bwd', 'unpack_wgrad']
orig = {n: getattr(ops, n) for n in names}
def wrap(n):
    def f(*a, **k):
        out = orig[n](*a, **k)
        outs = out if isinstance(out, (tuple, list)) else (out,)
        extra = []
        if n == 'gn_bwd':
            extra = [a[5], a[6]] + ([k['dada']] if k.get('dada') is not None else [])
        rec[-1].append((n, [o.detach().clone().double() for o in list(outs) + extra if isinstance(o, torch.Tensor)],
                        [tuple(v.shape) for v in a if isinstance(v, torch.Tensor)][:2]))
        return out
    return f
for n in names: setattr(ops, n, wrap(n))
for it in range(5):
    rec.append([])
    for p in m.parameters(): p.grad = None
    m.eps_override = eps
    total, _, _ = m.elbo(x.cuda(), t.cuda()); total.backward(); torch.cuda.synchronize()
    if it > 0:
        first = None
        for idx, ((n0, o0, s0), (n1, o1, s1)) in enumerate(zip(rec[0], rec[it])):
            d = max(((a - b).norm() / (a.norm() + 1e-30)).item() for a, b in zip(o0, o1)) if o0 else 0.0
            if d > 0:
                first = (idx, n0, s0, d); break
        print(it, 'first op whose output differs from iteration 0:', first, 'of', len(rec[0]))
print('--- detail: iteration 0 vs last')
a_, b_ = rec[0], rec[-1]
for idx in range(len(a_)):
    n0, o0, s0 = a_[idx]; n1, o1, s1 = b_[idx]
    if n0 in ('fcomb_fwd',) or 296 <= idx <= 301:
        for j, (u, v) in enumerate(zip(o0, o1)):
            d = (u - v).abs()
            print(idx, n0, s0, 'out', j, 'max abs diff', d.max().item(), 'n>1e-6', int((d > 1e-6).sum()), 'of', d.numel(),
                  'zeros', int((u == 0).sum()), int((v == 0).sum()))
