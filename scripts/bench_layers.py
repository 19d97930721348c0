"""Per-layer microbenchmark of the tcgen05 conv kernels (forward/dgrad and wgrad) and attention at the shapes one
ELBO step issues, batch 64.  CUDA events on the launching stream, L2 flushed between reps by
the size of the working set (every tensor here is >> 126 MB at batch 64 except the 16x16 level)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prob_unet_mds_b200 import _lib as L  # noqa: E402
from prob_unet_mds_b200 import ops  # noqa: E402

B = int(os.environ.get('B', '64'))
REPS = 5
# (Cin0, Cin1, Cout, HW, k, count)
CONV_SHAPES = [
    (128, 0, 128, 128, 3, 7), (256, 0, 128, 128, 3, 2), (256, 0, 256, 128, 3, 2), (384, 0, 128, 128, 3, 1),
    (128, 0, 64, 128, 3, 1),
    (128, 0, 256, 64, 3, 1), (256, 0, 256, 64, 3, 6), (384, 0, 256, 64, 3, 1), (384, 0, 384, 64, 3, 2),
    (512, 0, 256, 64, 3, 1), (640, 0, 256, 64, 3, 1),
    (256, 0, 384, 32, 3, 1), (384, 0, 384, 32, 3, 6), (512, 0, 512, 32, 3, 2), (768, 0, 384, 32, 3, 3),
    (384, 0, 512, 16, 3, 1), (512, 0, 512, 16, 3, 10), (1024, 0, 512, 16, 3, 3),
    (256, 0, 768, 64, 1, 5), (256, 0, 256, 64, 1, 5), (384, 0, 1152, 32, 1, 5), (512, 0, 1536, 16, 1, 6),
    (256, 128, 128, 128, 1, 2),
]


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ms = []
    for _ in range(REPS):
        flush.zero_()
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ms.append(s.elapsed_time(e))
    return min(ms)


def bench_gn():
    dt = torch.bfloat16
    for Cc, hw in ((128, 128), (256, 64), (384, 32), (1024, 16)):
        x = torch.randn(B, hw, hw, Cc, device='cuda').to(dt)
        dy = torch.randn(B, hw, hw, Cc, device='cuda').to(dt)
        gamma = torch.ones(Cc, device='cuda')
        beta = torch.zeros(Cc, device='cuda')
        ada = torch.zeros(2 * Cc, device='cuda')
        dg, db, da = torch.empty(Cc, device='cuda'), torch.empty(Cc, device='cuda'), torch.empty(2 * Cc, device='cuda')
        st = ops.gn_stats(x)
        nbytes = x.numel() * 2
        t_s = timeit(lambda: ops.gn_stats(x))
        t_a = timeit(lambda: ops.gn_apply(x, st, gamma, beta, ada=ada, silu=True, dropout_p=0.1, seed=1))
        t_b = timeit(lambda: ops.gn_bwd(x, st, gamma, beta, dy, dg, db, ada=ada, dada=da, silu=True, dropout_p=0.1,
                                        seed=1, dres=dy))
        t_g = timeit(lambda: ops.bias_grad(dy))
        print(json.dumps(dict(shape=f'gn C={Cc} {hw}x{hw}', stats_ms=round(t_s, 3), stats_gbs=round(nbytes / t_s / 1e6),
                              apply_ms=round(t_a, 3), apply_gbs=round(2 * nbytes / t_a / 1e6),
                              bwd_ms=round(t_b, 3), bwd_gbs=round(6 * nbytes / t_b / 1e6),
                              bias_grad_ms=round(t_g, 3), bias_grad_gbs=round(nbytes / t_g / 1e6))), flush=True)


def main():
    if os.environ.get('ONLY') == 'gn':
        return bench_gn()
    bench_gn()
    dt = torch.bfloat16
    rows = []
    tot = {'fwd': [0.0, 0.0], 'dgrad': [0.0, 0.0], 'wgrad': [0.0, 0.0]}
    for c0, c1, co, hw, k, cnt in CONV_SHAPES:
        x0 = torch.randn(B, hw, hw, c0, device='cuda').to(dt)
        x1 = torch.randn(B, hw, hw, c1, device='cuda').to(dt) if c1 else None
        dy = torch.randn(B, hw, hw, co, device='cuda').to(dt)
        w = torch.randn(co, c0 + c1, k, k, device='cuda') * 0.02
        wf = ops.pack_weight(w, 0, dt)
        wd = ops.pack_weight(w, 1, dt)
        bias = torch.randn(co, device='cuda')
        flops = 2.0 * B * hw * hw * co * k * k * (c0 + c1)
        t_f = timeit(lambda: ops.conv2d(x0, wf, co, k, bias=bias, src1=x1, flags=L.CONV_FORCE_TC))
        t_d = timeit(lambda: ops.conv2d(dy, wd, c0 + c1, k, flags=L.CONV_FORCE_TC))
        # the same data gradient with the GroupNorm-backward epilogue (SiLU + dropout 0.1), and what it replaces: the
        # first pass of pu_gn_bwd over the conv's output
        Cn = c0 + c1
        xn = torch.randn(B, hw, hw, Cn, device='cuda').to(dt)
        gam, bet, ada = torch.ones(Cn, device='cuda'), torch.zeros(Cn, device='cuda'), torch.zeros(2 * Cn, device='cuda')
        st = ops.gn_stats(xn)
        mask = torch.empty(xn.numel() // 8, dtype=torch.uint8, device='cuda')     # keep bits as the forward stores them
        ops.gn_apply(xn, st, gam, bet, ada=ada, silu=True, dropout_p=0.1, seed=1, keep_mask=mask)
        d, sums, keep = ops.gn_bwd_epilogue(xn, st, gam, bet, ada=ada, silu=True, dropout_p=0.1, seed=1, keep_mask=mask)
        t_dg = timeit(lambda: ops.conv2d(dy, wd, Cn, k, flags=L.CONV_FORCE_TC, gn_bwd=d))
        dh = ops.conv2d(dy, wd, Cn, k, flags=L.CONV_FORCE_TC)
        dg_, db_, da_ = torch.empty(Cn, device='cuda'), torch.empty(Cn, device='cuda'), torch.empty(2 * Cn, device='cuda')
        t_b2 = timeit(lambda: ops.gn_bwd(xn, st, gam, bet, dh, dg_, db_, ada=ada, dada=da_, silu=True, dropout_p=0.1, seed=1))
        t_b1 = timeit(lambda: ops.gn_bwd(xn, st, gam, bet, dh, dg_, db_, ada=ada, dada=da_, silu=True, dropout_p=0.1, seed=1,
                                         sums=sums, du_ready=True))
        del xn, dh
        t_w = timeit(lambda: ops.conv2d_wgrad(x0, dy, k, src1=x1, flags=L.CONV_FORCE_TC))
        rows.append(dict(shape=f'{c0}+{c1}->{co} {hw}x{hw} k{k} x{cnt}', fwd_ms=round(t_f, 3), dgrad_ms=round(t_d, 3),
                         wgrad_ms=round(t_w, 3), fwd_tf=round(flops / t_f / 1e9, 1), dgrad_tf=round(flops / t_d / 1e9, 1),
                         wgrad_tf=round(flops / t_w / 1e9, 1), dgrad_gnb_ms=round(t_dg, 3),
                         gn_bwd_two_pass_ms=round(t_b2, 3), gn_bwd_second_pass_ms=round(t_b1, 3),
                         gnb_gain_ms=round((t_d + t_b2) - (t_dg + t_b1), 3)))
        for key, t in (('fwd', t_f), ('dgrad', t_d), ('wgrad', t_w)):
            tot[key][0] += flops * cnt
            tot[key][1] += t * cnt
        print(json.dumps(rows[-1]), flush=True)
        del x0, x1, dy
    for key, (fl, ms) in tot.items():
        print(f'{key}: {ms:.2f} ms per step-equivalent, {fl / ms / 1e9:.1f} TFLOP/s')
    # attention
    for heads, T, cnt in ((4, 4096, 5), (6, 1024, 5), (8, 256, 6)):
        C = heads * 64
        qkv = torch.randn(B, T, 3 * C, device='cuda').to(dt)
        out, lse = ops.attention_fwd(qkv, heads)
        dout = torch.randn_like(out)
        fl = 4.0 * B * heads * T * T * 64
        t_f = timeit(lambda: ops.attention_fwd(qkv, heads))
        t_b = timeit(lambda: ops.attention_bwd(qkv, out, dout, lse, heads))
        print(json.dumps(dict(shape=f'attn heads={heads} T={T} x{cnt}', fwd_ms=round(t_f, 3), bwd_ms=round(t_b, 3),
                              fwd_tf=round(fl / t_f / 1e9, 1), bwd_tf=round(2.5 * fl / t_b / 1e9, 1))), flush=True)


if __name__ == '__main__':
    main()
