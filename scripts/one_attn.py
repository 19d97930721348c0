"""Runs the attention forward / backward kernels once each at one (heads, T) configuration (for ncu):
    python scripts/one_attn.py [heads] [T] [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prob_unet_mds_b200 import ops  # noqa: E402

heads = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
C = heads * 64
qkv = torch.randn(B, T, 3 * C, device='cuda').bfloat16()
for _ in range(2):
    out, lse = ops.attention_fwd(qkv, heads)
    dout = torch.randn_like(out)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, heads)
torch.cuda.synchronize()
print('ok')
