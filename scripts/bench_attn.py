"""Attention kernel microbenchmark at the three (heads, T) configurations of one ELBO step, batch 64."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from prob_unet_mds_b200 import ops  # noqa: E402
import bench_layers as bl  # noqa: E402

B = int(os.environ.get('B', '64'))
for heads, T in ((4, 4096), (6, 1024), (8, 256)):
    C = heads * 64
    qkv = torch.randn(B, T, 3 * C, device='cuda').bfloat16()
    out, lse = ops.attention_fwd(qkv, heads)
    dout = torch.randn_like(out)
    fl = 4.0 * B * heads * T * T * 64
    tf = bl.timeit(lambda: ops.attention_fwd(qkv, heads))
    tb = bl.timeit(lambda: ops.attention_bwd(qkv, out, dout, lse, heads))
    print(f'attn heads={heads} T={T}: fwd {tf:.3f} ms {fl / tf / 1e9:.1f} TF/s | bwd {tb:.3f} ms {2.5 * fl / tb / 1e9:.1f} TF/s',
          flush=True)
