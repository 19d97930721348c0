"""Launches every heavy kernel of one ELBO step once (after one warm-up launch each) at the headline shapes
(batch 64, 128x128 input), for `ncu --set full`.  Order = the order of the rows in profiles/r2_ncu_heavy_kernels.tsv."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prob_unet_mds_b200 import _lib as L  # noqa: E402
from prob_unet_mds_b200 import ops  # noqa: E402

B = int(os.environ.get('B', '64'))
dt = torch.bfloat16
dev = 'cuda'


def twice(f):
    f()
    torch.cuda.synchronize()
    f()
    torch.cuda.synchronize()


def conv_case(c0, co, hw, k):
    x = torch.randn(B, hw, hw, c0, device=dev).to(dt)
    dy = torch.randn(B, hw, hw, co, device=dev).to(dt)
    w = torch.randn(co, c0, k, k, device=dev) * 0.02
    wf, wd = ops.pack_weight(w, 0, dt), ops.pack_weight(w, 1, dt)
    bias = torch.randn(co, device=dev)
    twice(lambda: ops.conv2d(x, wf, co, k, bias=bias, flags=L.CONV_FORCE_TC))      # forward
    twice(lambda: ops.conv2d(x, wf, co, k, bias=bias, flags=L.CONV_FORCE_TC, want_qstats=True))   # + GroupNorm statistics
    twice(lambda: ops.conv2d(dy, wd, c0, k, flags=L.CONV_FORCE_TC))                # dgrad
    if k == 3:                                                                     # dgrad + GroupNorm-backward epilogue
        gam, bet, ada = torch.ones(c0, device=dev), torch.zeros(c0, device=dev), torch.zeros(2 * c0, device=dev)
        st = ops.gn_stats(x)
        mask = torch.empty(x.numel() // 8, dtype=torch.uint8, device=dev)
        ops.gn_apply(x, st, gam, bet, ada=ada, silu=True, dropout_p=0.1, seed=1, keep_mask=mask)
        d, sums, keep = ops.gn_bwd_epilogue(x, st, gam, bet, ada=ada, silu=True, dropout_p=0.1, seed=1, keep_mask=mask)
        twice(lambda: ops.conv2d(dy, wd, c0, k, flags=L.CONV_FORCE_TC, gn_bwd=d))
    twice(lambda: ops.conv2d_wgrad(x, dy, k, flags=L.CONV_FORCE_TC))               # wgrad


conv_case(256, 256, 128, 3)     # BN=256 tile, wgrad mode 1
conv_case(128, 128, 128, 3)     # BN=128, two accumulators (16x16-pixel tile), wgrad mode 2
conv_case(256, 768, 64, 1)      # qkv 1x1

for heads, T in ((4, 4096),):
    C = heads * 64
    qkv = torch.randn(B, T, 3 * C, device=dev).to(dt)
    out, lse = ops.attention_fwd(qkv, heads)
    dout = torch.randn_like(out)
    twice(lambda: ops.attention_fwd(qkv, heads))
    twice(lambda: ops.attention_bwd(qkv, out, dout, lse, heads))

Cc, hw = 128, 128
x = torch.randn(B, hw, hw, Cc, device=dev).to(dt)
dy = torch.randn(B, hw, hw, Cc, device=dev).to(dt)
dres = torch.randn(B, hw, hw, Cc, device=dev).to(dt)
gamma, beta, ada = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev), torch.zeros(2 * Cc, device=dev)
dg, db, da = torch.empty(Cc, device=dev), torch.empty(Cc, device=dev), torch.empty(2 * Cc, device=dev)
st = ops.gn_stats(x)
twice(lambda: ops.gn_stats(x))
twice(lambda: ops.gn_apply(x, st, gamma, beta, ada=ada, silu=True, dropout_p=0.1, seed=1))
twice(lambda: ops.gn_bwd(x, st, gamma, beta, dy, dg, db, ada=ada, dada=da, silu=True, dropout_p=0.1, seed=1, dres=dres))

# ensemble decode: 8 inputs x 100 members
N, S, Lz = 8, 100, 16
feat = torch.randn(N, 128, 128, 64, device=dev).to(dt)
z = torch.randn(N, S, Lz, device=dev)
w0 = torch.randn(64, 64 + Lz, 1, 1, device=dev) / 8
w1 = torch.randn(64, 64, 1, 1, device=dev) / 8
w2 = torch.randn(3, 64, 1, 1, device=dev) / 8
b0, b1, b2 = torch.randn(64, device=dev), torch.randn(64, device=dev), torch.randn(3, device=dev)
twice(lambda: ops.fcomb_fwd(feat, z, w0, b0, w1, b1, w2, b2, S=S))
print('ok')
