"""Ablation timings of the attention kernels at T = 4096, heads = 4, batch 64: each variant removes ONE class of work
(results are wrong on purpose) to show what the kernel is actually bound by.  One interpreter per variant (the switches
are read once per process).  PU_ATTN_BWD: 20 / 3 / 33 (default) are real kernels; 2xx = ablations of attn_bwd_tc2_kernel
(20), 3xx of attn_bwd_tc3_kernel with per-lane reductions (3); xx is a bit mask: 1 no STS of P / dS (2xx) / no LDS of lse,
delta (3xx), 4 = no dQ reductions, 8 = no softmax work at all (synchronisation chain + MMAs); 334 = the default without
issuing its bulk reductions.  PU_ATTN_FWD_ABL: 1 no per-tile O read-back, 2 no exponentials, 5 no softmax work;
PU_ATTN_FWD_POLY: every n-th pair of exponentials on the FMA pipe (default 4).  Results: profiles/r2_attention_ablation.md."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SNIPPET = r"""
import sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + '/scripts')
from prob_unet_mds_b200 import ops
import bench_layers as bl
B, heads, T = 64, 4, 4096
C = heads * 64
qkv = torch.randn(B, T, 3 * C, device='cuda').bfloat16()
out, lse = ops.attention_fwd(qkv, heads)
dout = torch.randn_like(out)
tf = bl.timeit(lambda: ops.attention_fwd(qkv, heads))
tb = bl.timeit(lambda: ops.attention_bwd(qkv, out, dout, lse, heads))
print('RESULT fwd %.3f ms bwd %.3f ms' % (tf, tb))
"""

if __name__ == '__main__':
    runs = [('default', {})]
    runs += [(f'bwd {v}', {'PU_ATTN_BWD': str(v)}) for v in (20, 201, 204, 208, 212, 3, 301, 304, 305, 308, 312, 334)]
    runs += [('fwd read-back', {'PU_ATTN_FWD': '2'})]
    runs += [(f'fwd abl {v}', {'PU_ATTN_FWD_ABL': str(v)}) for v in (1, 2, 5)]
    runs += [(f'fwd poly {v}', {'PU_ATTN_FWD_POLY': str(v)}) for v in (1, 2, 3)]
    
    for name, env in runs:
        r = subprocess.run([sys.executable, '-c', SNIPPET.format(root=ROOT)], capture_output=True, text=True,
                           env=dict(os.environ, **env), timeout=600)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith('RESULT')]
        print(f'{name:14s} {line[0] if line else "FAILED: " + r.stderr[-400:]}', flush=True)
