"""Per-call duration of pu_attention_bwd / pu_attention_fwd inside a real ELBO training step (batch 64), next to the
isolated micro-benchmark of the same shapes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from prob_unet_mds_b200 import ProbabilisticUNet, AdamW, _lib as L, ops  # noqa: E402
import bench_layers as bl  # noqa: E402

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
for name, kind, flops, s, e in prof.records:
    if name in ('pu_attention_bwd', 'pu_attention_fwd'):
        print(f'{name}: {s.elapsed_time(e):.3f} ms')
for heads, T in ((4, 4096),):
    C = heads * 64
    qkv = torch.randn(B, T, 3 * C, device=dev).bfloat16()
    out, lse = ops.attention_fwd(qkv, heads)
    dout = torch.randn_like(out)
    print('micro bwd (random data)        %.3f ms' % bl.timeit(lambda: ops.attention_bwd(qkv, out, dout, lse, heads)))
    print('micro bwd (random, want_dbias) %.3f ms' % bl.timeit(lambda: ops.attention_bwd(qkv, out, dout, lse, heads, want_dbias=True)))
    z = torch.zeros_like(dout)
    print('micro bwd (dout = 0)           %.3f ms' % bl.timeit(lambda: ops.attention_bwd(qkv, out, z, lse, heads)))
    small = (dout.float() * 1e-6).bfloat16()
    print('micro bwd (dout * 1e-6)        %.3f ms' % bl.timeit(lambda: ops.attention_bwd(qkv, out, small, lse, heads)))
