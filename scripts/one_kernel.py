"""Runs one kernel a few times (for ncu): python scripts/one_kernel.py wgrad|conv|dgrad C0 Cout HW K [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prob_unet_mds_b200 import _lib as L  # noqa: E402
from prob_unet_mds_b200 import ops  # noqa: E402

kind, c0, co, hw, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
B = int(sys.argv[6]) if len(sys.argv) > 6 else 64
dt = torch.bfloat16
x = torch.randn(B, hw, hw, c0, device='cuda').to(dt)
dy = torch.randn(B, hw, hw, co, device='cuda').to(dt)
w = torch.randn(co, c0, k, k, device='cuda') * 0.02
wf, wd = ops.pack_weight(w, 0, dt), ops.pack_weight(w, 1, dt)
for _ in range(3):
    if kind == 'wgrad':
        ops.conv2d_wgrad(x, dy, k, flags=L.CONV_FORCE_TC)
    elif kind == 'conv':
        ops.conv2d(x, wf, co, k, flags=L.CONV_FORCE_TC)
    else:
        ops.conv2d(dy, wd, c0, k, flags=L.CONV_FORCE_TC)
torch.cuda.synchronize()
print('ok')
