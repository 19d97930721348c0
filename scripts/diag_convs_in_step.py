"""Every conv / wgrad / GroupNorm call of one real ELBO training step (batch 64) with its shape, flags, duration and
TFLOP/s (or GB/s), aggregated by shape -- to set beside scripts/bench_layers.py (the same shapes in isolation)."""
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prob_unet_mds_b200 import ProbabilisticUNet, AdamW, _lib as L  # noqa: E402

_orig = L.CallProfiler._conv_info


def _info(name, args):
    kind, flops = _orig(name, args)
    if name in ('pu_conv2d', 'pu_conv2d_wgrad'):
        a = args[0]._obj
        tag = f'{kind} {a.H}x{a.W} {a.C0}+{a.C1}->{a.Cout} k{a.ksize}'
        if name == 'pu_conv2d':
            tag += (' qstats' if a.qstats else '') + (' gnb' if a.gn_bwd else '') + (' res' if a.residual else '')
        return tag, flops
    if name in ('pu_gn_apply', 'pu_gn_bwd'):
        a = args[0]._obj
        f = a.f if name == 'pu_gn_bwd' else a
        tag = f'{name} {f.H}x{f.W} {f.C0}+{f.C1} rs{f.resample} p{f.dropout_p:.1f}'
        if name == 'pu_gn_bwd':
            tag += (' du_ready' if a.du_ready else '') + (' dres' if a.dres else '')
        return tag, 0.0
    return kind, flops


L.CallProfiler._conv_info = staticmethod(_info)

torch.manual_seed(0)
dev = torch.device('cuda')
B = int(os.environ.get('B', '64'))
m = ProbabilisticUNet(3, 3, latent_dim=16).to(dev)
m.train()
opt = AdamW(m.parameters(), lr=1e-4)
x = torch.randn(B, 3, 128, 128, device=dev)
t = torch.randn(B, 3, 128, 128, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    total, _, _ = m.elbo(x, t)
    total.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
prof = L.start_profiling()
step()
torch.cuda.synchronize()
L.stop_profiling()
agg = OrderedDict()
for name, kind, flops, s, e in prof.records:
    d = agg.setdefault(kind, [0, 0.0, 0.0])
    d[0] += 1
    d[1] += s.elapsed_time(e)
    d[2] += flops
tot = sum(v[1] for v in agg.values())
print(f'total {tot:.2f} ms in {len(prof.records)} calls')
for k, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if ms < 0.15:
        continue
    print(f'{ms:7.3f} ms  x{n:<3d} {ms / n:7.3f} each  {fl / ms / 1e9 if fl else 0:7.1f} TF/s  {k}')
