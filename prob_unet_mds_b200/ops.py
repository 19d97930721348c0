"""Thin Python wrappers over the C ABI: allocate outputs with torch, pass raw pointers + the current stream.

Activations are NHWC tensors ([N, H, W, C], contiguous) in torch.float32 or torch.bfloat16.
Nothing here computes with torch; every function ends in exactly one (or a fixed few) pu_* kernel launches.
"""
import ctypes as C

import torch

from . import _lib as L
from ._lib import check, dtype_code, lib, ptr, stream_ptr


def _nhwc(t):
    assert t.dim() == 4 and t.is_contiguous() and t.is_cuda, 'expected a contiguous CUDA NHWC tensor'
    return t.shape


def zeros(shape, dtype, device):
    """torch.empty + a stream-ordered cudaMemsetAsync (pu_zero) -- no ATen fill kernel on the path."""
    t = torch.empty(shape, dtype=dtype, device=device)
    check(lib().pu_zero(ptr(t), t.numel() * t.element_size(), stream_ptr()), 'zero')
    return t


def clone(t, out=None):
    assert t.is_contiguous()
    if out is None:
        out = torch.empty_like(t)
    assert out.is_contiguous() and out.numel() == t.numel() and out.dtype == t.dtype
    check(lib().pu_copy(ptr(out), ptr(t), t.numel() * t.element_size(), stream_ptr()), 'copy')
    return out


def zeros_f64(n, device):
    return zeros((n,), torch.float64, device)


# ----------------------------------------------------------------------------- layout / packing
def nchw_to_nhwc(x, dtype, out=None, c_off=0, Cdst=None):
    N, Cc, H, W = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and x.is_cuda
    if out is None:
        Cdst = Cdst or Cc
        out = (zeros if Cdst != Cc else torch.empty)((N, H, W, Cdst), dtype=dtype, device=x.device)
    check(lib().pu_nchw_to_nhwc(ptr(x), ptr(out), N, Cc, H, W, out.shape[3], c_off, dtype_code(out.dtype), stream_ptr()),
          'nchw_to_nhwc')
    return out


def nhwc_to_nchw(x, C=None):
    """C: take only the first C channels of x (default: all)."""
    N, H, W, Csrc = _nhwc(x)
    Cc = C or Csrc
    out = torch.empty((N, Cc, H, W), dtype=torch.float32, device=x.device)
    check(lib().pu_nhwc_to_nchw(ptr(x), ptr(out), N, Cc, H, W, Csrc, dtype_code(x.dtype), stream_ptr()), 'nhwc_to_nchw')
    return out


def pack_weight(w, mode, dtype, Ci_pad=None, perm=None, Ci=None, src_co_stride=0, out=None):
    """w: fp32 OIHW master weight (or a [Co, >=Ci] matrix with row stride src_co_stride for k = 1)."""
    Co = w.shape[0]
    k = w.shape[-1] if w.dim() == 4 else 1
    Ci = Ci if Ci is not None else w.shape[1]
    Ci_pad = Ci_pad or Ci
    shape = (Co, k, k, Ci_pad) if mode == 0 else ((Ci, k, k, Co) if mode == 1 else (Co, k, k, 2 * Ci_pad))
    if out is None:
        out = torch.empty(shape, dtype=dtype, device=w.device)
    check(lib().pu_pack_conv_weight(ptr(w), ptr(out), Co, Ci, k, Ci_pad, mode, ptr(perm), src_co_stride,
                                    dtype_code(dtype), stream_ptr()), 'pack_conv_weight')
    return out


_PACK_DT = None


def pack_item(w, out, mode, perm=None, Ci_pad=None, dtype=None):
    """One row of the pu_pack_conv_weights_multi table (include/probunet_b200.h, PuPackItem) as a tuple."""
    Co = w.shape[0]
    k = w.shape[-1] if w.dim() == 4 else 1
    Ci = w.shape[1]
    Ci_pad = Ci_pad or Ci
    tiles = ((Co + 31) // 32) * ((Ci_pad + 31) // 32)
    return (w.data_ptr(), out.data_ptr(), perm.data_ptr() if perm is not None else 0, Ci * k * k, Co, Ci, k, Ci_pad,
            mode, dtype_code(dtype if dtype is not None else out.dtype), tiles)


def pack_table(items, device):
    """Device table for pu_pack_conv_weights_multi: returns (table tensor, n_items, total_tiles)."""
    import numpy as np
    global _PACK_DT
    if _PACK_DT is None:
        _PACK_DT = np.dtype([('src', '<u8'), ('dst', '<u8'), ('perm', '<u8'), ('stride', '<i8'), ('Co', '<i4'),
                             ('Ci', '<i4'), ('k', '<i4'), ('Ci_pad', '<i4'), ('mode', '<i4'), ('dtype', '<i4'),
                             ('tile_begin', '<i4'), ('pad', '<i4')])
    tab = np.zeros(len(items), dtype=_PACK_DT)
    begin = 0
    for i, it in enumerate(items):
        tab[i] = it[:10] + (begin, 0)
        begin += it[10]
    host = torch.from_numpy(tab.view(np.uint8).copy())
    return host.to(device), len(items), begin


def pack_weights_multi(table, n_items, total_tiles):
    check(lib().pu_pack_conv_weights_multi(ptr(table), n_items, total_tiles, stream_ptr()), 'pack_conv_weights_multi')


def unpack_wgrad(dw_packed, grad, perm=None, Ci=None, dst_co_stride=0, accumulate=False):
    """dw_packed: fp32 [Co', k, k, Ci_pad] with Co' >= Co (rows beyond the destination's Co are padding and ignored);
    grad: fp32 OIHW destination."""
    _, k, _, Ci_pad = dw_packed.shape
    Co = min(dw_packed.shape[0], grad.shape[0])
    Ci = Ci if Ci is not None else grad.shape[1]
    check(lib().pu_unpack_conv_wgrad(ptr(dw_packed), ptr(grad), Co, Ci, k, Ci_pad, ptr(perm), dst_co_stride,
                                     int(accumulate), stream_ptr()), 'unpack_conv_wgrad')
    return grad


def gather(src, perm):
    out = torch.empty_like(src)
    check(lib().pu_gather_f32(ptr(src), ptr(perm), ptr(out), src.numel(), stream_ptr()), 'gather')
    return out


def scatter(src, perm, dst, accumulate=False):
    check(lib().pu_scatter_f32(ptr(src), ptr(perm), ptr(dst), src.numel(), int(accumulate), stream_ptr()), 'scatter')
    return dst


# ----------------------------------------------------------------------------- convolution
def conv2d(src0, weight, Cout, ksize, bias=None, src1=None, residual=None, relu=False, bias_per_sample=False,
           out=None, flags=0, want_qstats=False, gn_bwd=None):
    """want_qstats: also return the [N, Cout/4, 2] fp64 (sum, sumsq) of the stored output per quad of channels, taken in
    the conv epilogue -- the GroupNorm statistics of whatever consumes the output (gn_stats_from_quads)."""
    N, H, W, C0 = _nhwc(src0)
    C1 = src1.shape[3] if src1 is not None else 0
    if out is None:
        out = torch.empty((N, H, W, Cout), dtype=src0.dtype, device=src0.device)
    q = torch.empty((N, Cout // 4, 2), dtype=torch.float64, device=src0.device) if want_qstats else None
    a = L.PuConvArgs(N, H, W, C0, C1, Cout, ksize, dtype_code(src0.dtype), flags | (L.CONV_RELU if relu else 0),
                     int(bias_per_sample), ptr(src0), ptr(src1), ptr(weight), ptr(bias), ptr(residual), ptr(out),
                     ptr(q), 0, C.pointer(gn_bwd) if gn_bwd is not None else None)
    check(lib().pu_conv2d(C.byref(a), stream_ptr()), 'conv2d')
    return (out, q) if want_qstats else out


def conv_tc_applies(x, C1, Cout):
    """Mirror of conv_tc_applicable (csrc/conv_tc.cu): will pu_conv2d run the tcgen05 kernel for this input?"""
    return (x.dtype == torch.bfloat16 and x.shape[3] % 64 == 0 and C1 % 64 == 0 and Cout % 64 == 0
            and x.shape[2] >= 16 and x.shape[1] >= 8)


def gn_bwd_epilogue(src0, stats, gamma, beta, src1=None, ada=None, silu=True, dropout_p=0.0, seed=0, eps=1e-5,
                    keep_mask=None):
    """Descriptor (PuConvGnBwd) that lets the data-gradient conv producing dL/dy of a GroupNorm(+SiLU)(+dropout) do the
    first pass of the GroupNorm backward in its epilogue.  Returns (descriptor, sums, keepalive); pass the descriptor
    to conv2d(gn_bwd=...) and `sums` to gn_bwd(..., sums=sums, du_ready=True)."""
    N, H, W, C0 = _nhwc(src0)
    C1 = src1.shape[3] if src1 is not None else 0
    Cc = C0 + C1
    consts = torch.empty((N, Cc, 4), dtype=torch.float32, device=src0.device)
    f = _gn_args(src0, src1, stats, gamma, beta, ada, silu, L.RS_NONE, dropout_p, seed, None, eps)
    check(lib().pu_gn_bwd_consts(C.byref(f), ptr(consts), stream_ptr()), 'gn_bwd_consts')
    sums = torch.empty((N, Cc, 2), dtype=torch.float64, device=src0.device)
    d = L.PuConvGnBwd(ptr(src0), ptr(src1), C0, C1, ptr(consts), ptr(sums), int(silu), float(dropout_p), int(seed),
                      ptr(keep_mask))
    return d, sums, consts


def conv2d_wgrad(src0, dy, ksize, src1=None, dw=None, accumulate=False, flags=0):
    N, H, W, C0 = _nhwc(src0)
    C1 = src1.shape[3] if src1 is not None else 0
    Cout = dy.shape[3]
    if dw is None:
        dw = torch.empty((Cout, ksize, ksize, C0 + C1), dtype=torch.float32, device=src0.device)
        accumulate = False
    a = L.PuWgradArgs(N, H, W, C0, C1, Cout, ksize, dtype_code(src0.dtype), flags, ptr(src0), ptr(src1), ptr(dy),
                      ptr(dw), int(accumulate))
    check(lib().pu_conv2d_wgrad(C.byref(a), stream_ptr()), 'conv2d_wgrad')
    return dw


def bias_grad(dy, db=None, accumulate=False):
    Cc = dy.shape[-1]
    pixels = dy.numel() // Cc
    if db is None:
        db = torch.empty(Cc, dtype=torch.float32, device=dy.device)
        accumulate = False
    check(lib().pu_bias_grad(ptr(dy), ptr(db), pixels, Cc, dtype_code(dy.dtype), int(accumulate), stream_ptr()),
          'bias_grad')
    return db


# ----------------------------------------------------------------------------- group norm
def gn_groups(Cc):
    return min(32, Cc // 4)


def gn_stats(src0, src1=None, G=None):
    N, H, W, C0 = _nhwc(src0)
    C1 = src1.shape[3] if src1 is not None else 0
    G = G or gn_groups(C0 + C1)
    stats = torch.empty((N, G, 2), dtype=torch.float64, device=src0.device)
    check(lib().pu_gn_stats(ptr(src0), ptr(src1), C0, C1, N, H * W, G, dtype_code(src0.dtype), ptr(stats),
                            stream_ptr()), 'gn_stats')
    return stats


def gn_stats_from_quads(q0, q1=None, G=None):
    """GroupNorm statistics [N, G, 2] of a tensor (or the channel concatenation of two) from the per-quad statistics that
    the producing convolutions emitted (conv2d(..., want_qstats=True))."""
    N, C0 = q0.shape[0], q0.shape[1] * 4
    C1 = q1.shape[1] * 4 if q1 is not None else 0
    G = G or gn_groups(C0 + C1)
    stats = torch.empty((N, G, 2), dtype=torch.float64, device=q0.device)
    check(lib().pu_gn_stats_from_quads(ptr(q0), ptr(q1), C0, C1, N, G, ptr(stats), stream_ptr()), 'gn_stats_from_quads')
    return stats


def _gn_args(src0, src1, stats, gamma, beta, ada, silu, resample, dropout_p, seed, y, eps, keep_mask=None):
    N, H, W, C0 = _nhwc(src0)
    C1 = src1.shape[3] if src1 is not None else 0
    return L.PuGnArgs(N, H, W, C0, C1, stats.shape[1], dtype_code(src0.dtype), int(silu), resample, eps,
                      float(dropout_p), int(seed), ptr(src0), ptr(src1), ptr(stats), ptr(gamma), ptr(beta), ptr(ada),
                      ptr(y), ptr(keep_mask))


def gn_apply(src0, stats, gamma, beta, src1=None, ada=None, silu=True, resample=L.RS_NONE, dropout_p=0.0, seed=0,
             eps=1e-5, keep_mask=None):
    """keep_mask: optional uint8 [N*H*W*C/8] that receives the dropout keep bits (for gn_bwd_epilogue)."""
    N, H, W, C0 = _nhwc(src0)
    Cc = C0 + (src1.shape[3] if src1 is not None else 0)
    OH, OW = (H * 2, W * 2) if resample == L.RS_UP else ((H // 2, W // 2) if resample == L.RS_DOWN else (H, W))
    y = torch.empty((N, OH, OW, Cc), dtype=src0.dtype, device=src0.device)
    a = _gn_args(src0, src1, stats, gamma, beta, ada, silu, resample, dropout_p, seed, y, eps, keep_mask)
    check(lib().pu_gn_apply(C.byref(a), stream_ptr()), 'gn_apply')
    return y


def gn_bwd(src0, stats, gamma, beta, dy, dgamma, dbeta, src1=None, ada=None, dada=None, silu=True,
           resample=L.RS_NONE, dropout_p=0.0, seed=0, eps=1e-5, dres=None, dres_resample=L.RS_NONE,
           dx0=None, dx1=None, acc0=False, acc1=False, acc_params=False, colsum0=None, colsum1=None, sums=None,
           du_ready=False, keep_mask=None):
    """Returns (dx0, dx1).  dgamma/dbeta/dada are written (or accumulated into when acc_params).
    colsum0 / colsum1 (optional fp32 [C0] / [C1]) receive the per-channel sums of the final dx0 / dx1."""
    N, H, W, C0 = _nhwc(src0)
    C1 = src1.shape[3] if src1 is not None else 0
    if dx0 is None:
        dx0 = torch.empty_like(src0)
        acc0 = False
    if src1 is not None and dx1 is None:
        dx1 = torch.empty_like(src1)
        acc1 = False
    if sums is None:
        assert not du_ready
        sums = torch.empty((N, C0 + C1, 2), dtype=torch.float64, device=src0.device)
    f = _gn_args(src0, src1, stats, gamma, beta, ada, silu, resample, dropout_p, seed, None, eps, keep_mask)
    a = L.PuGnBwdArgs(f, ptr(dy), ptr(dres), dres_resample, ptr(sums), ptr(dx0), ptr(dx1), int(acc0), int(acc1),
                      ptr(dgamma), ptr(dbeta), ptr(dada), int(acc_params), ptr(colsum0), ptr(colsum1), int(du_ready))
    check(lib().pu_gn_bwd(C.byref(a), stream_ptr()), 'gn_bwd')
    return dx0, dx1


# ----------------------------------------------------------------------------- attention
def attention_fwd(qkv, heads, flags=0):
    """qkv: [N, H, W, 3C] (or [N, T, 3C]) in the product's (j, head, d) channel order."""
    N = qkv.shape[0]
    C3 = qkv.shape[-1]
    T = qkv.numel() // (N * C3)
    out = torch.empty(qkv.shape[:-1] + (C3 // 3,), dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty((N, heads, T), dtype=torch.float32, device=qkv.device)
    check(lib().pu_attention_fwd(ptr(qkv), ptr(out), ptr(lse), N, T, heads, dtype_code(qkv.dtype), flags,
                                 stream_ptr()), 'attention_fwd')
    return out, lse


def attention_bwd(qkv, out, dout, lse, heads, flags=0, want_dbias=False):
    """want_dbias: also return the fp32 [3C] column sums of dqkv (the qkv conv's bias gradient, in the product's channel
    order), taken in the kernels' epilogues."""
    N = qkv.shape[0]
    C3 = qkv.shape[-1]
    T = qkv.numel() // (N * C3)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty((N, heads, T), dtype=torch.float32, device=qkv.device)
    # dQ accumulation workspace [N][T][C] + (want_dbias) the per-CTA column-sum partials [N*heads][ceil(T/128)][192]
    n_ws = N * T * (C3 // 3) + (N * heads * ((T + 127) // 128) * 192 if want_dbias else 0)
    dq_ws = torch.empty(n_ws, dtype=torch.float32, device=qkv.device)
    dbias = torch.empty(C3, dtype=torch.float32, device=qkv.device) if want_dbias else None
    check(lib().pu_attention_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), ptr(delta), ptr(dq_ws), ptr(dbias),
                                 N, T, heads, dtype_code(qkv.dtype), flags, stream_ptr()), 'attention_bwd')
    return (dqkv, dbias) if want_dbias else dqkv


# ----------------------------------------------------------------------------- encoder glue
def upsample2(x):
    N, H, W, Cc = _nhwc(x)
    y = torch.empty((N, H * 2, W * 2, Cc), dtype=x.dtype, device=x.device)
    check(lib().pu_upsample2(ptr(x), ptr(y), N, H, W, Cc, dtype_code(x.dtype), stream_ptr()), 'upsample2')
    return y


def resample_grad(dy, resample):
    """Gradient wrt the pre-resample activation (see pu_resample_grad): dy is NHWC at the resampled resolution."""
    N, OH, OW, Cc = _nhwc(dy)
    H, W = (OH // 2, OW // 2) if resample == L.RS_UP else (OH * 2, OW * 2)
    g = torch.empty((N, H, W, Cc), dtype=dy.dtype, device=dy.device)
    check(lib().pu_resample_grad(ptr(dy), ptr(g), N, H, W, Cc, resample, dtype_code(dy.dtype), stream_ptr()),
          'resample_grad')
    return g


def avgpool2(x):
    N, H, W, Cc = _nhwc(x)
    y = torch.empty((N, H // 2, W // 2, Cc), dtype=x.dtype, device=x.device)
    check(lib().pu_avgpool2(ptr(x), ptr(y), N, H, W, Cc, dtype_code(x.dtype), stream_ptr()), 'avgpool2')
    return y


def relu_pool_bwd(dp, r):
    N, H, W, Cc = _nhwc(r)
    dr = torch.empty_like(r)
    check(lib().pu_relu_pool_bwd(ptr(dp), ptr(r), ptr(dr), N, H, W, Cc, dtype_code(r.dtype), stream_ptr()),
          'relu_pool_bwd')
    return dr


def global_mean(x):
    N, H, W, Cc = _nhwc(x)
    m = torch.empty((N, Cc), dtype=torch.float32, device=x.device)
    check(lib().pu_global_mean(ptr(x), ptr(m), N, H * W, Cc, dtype_code(x.dtype), stream_ptr()), 'global_mean')
    return m


def relu_mean_bwd(dm, r):
    N, H, W, Cc = _nhwc(r)
    dr = torch.empty_like(r)
    check(lib().pu_relu_mean_bwd(ptr(dm), ptr(r), ptr(dr), N, H * W, Cc, dtype_code(r.dtype), stream_ptr()),
          'relu_mean_bwd')
    return dr


def relu_mask(dy, y, out=None):
    out = out if out is not None else torch.empty_like(dy)
    check(lib().pu_relu_mask(ptr(dy), ptr(y), ptr(out), dy.numel(), dtype_code(dy.dtype), stream_ptr()), 'relu_mask')
    return out


def heads_fwd(m, w, b):
    N, Cc = m.shape
    L2 = w.shape[0]
    out = torch.empty((N, L2), dtype=torch.float32, device=m.device)
    check(lib().pu_heads_fwd(ptr(m), ptr(w), ptr(b), ptr(out), N, Cc, L2, stream_ptr()), 'heads_fwd')
    return out


def heads_bwd(m, w, dout, dw, db, accumulate=False, dm=None):
    N, Cc = m.shape
    L2 = w.shape[0]
    acc_dm = dm is not None
    if dm is None:
        dm = torch.empty_like(m)
    check(lib().pu_heads_bwd(ptr(m), ptr(w), ptr(dout), ptr(dm), ptr(dw), ptr(db), N, Cc, L2, int(accumulate),
                             int(acc_dm), stream_ptr()), 'heads_bwd')
    return dm


# ----------------------------------------------------------------------------- latent
def rsample(mu, log_sigma, eps, flag=None):
    z = torch.empty_like(mu)
    sigma = torch.empty_like(mu)
    check(lib().pu_rsample(ptr(mu), ptr(log_sigma), ptr(eps), ptr(z), ptr(sigma), ptr(flag), mu.numel(),
                           stream_ptr()), 'rsample')
    return z, sigma


def rsample_bwd(dz, eps, sigma, dmu, dls):
    check(lib().pu_rsample_bwd(ptr(dz), ptr(eps), ptr(sigma), ptr(dmu), ptr(dls), dz.numel(), stream_ptr()),
          'rsample_bwd')


def kl_fwd_bwd(mu_q, ls_q, mu_p, ls_p, kl_acc, gscale=None, want_grads=True):
    """gscale: optional device fp32 scalar tensor."""
    g = [torch.empty_like(mu_q) for _ in range(4)] if want_grads else [None] * 4
    check(lib().pu_kl_fwd_bwd(ptr(mu_q), ptr(ls_q), ptr(mu_p), ptr(ls_p), ptr(kl_acc), ptr(g[0]), ptr(g[1]),
                              ptr(g[2]), ptr(g[3]), ptr(gscale), mu_q.numel(), stream_ptr()), 'kl')
    return g


def mse_fwd_bwd(out_nchw, target, recon_acc, dtype=None, gscale=None, Cdst=None):
    """Cdst: channel count of the returned dlogits tensor (>= C, zero padded)."""
    N, Cc, H, W = out_nchw.shape
    Cdst = Cdst or Cc
    dlogits = None
    if dtype is not None:
        dlogits = (zeros if Cdst != Cc else torch.empty)((N, H, W, Cdst), dtype=dtype, device=out_nchw.device)
    check(lib().pu_mse_fwd_bwd(ptr(out_nchw), ptr(target), ptr(recon_acc), ptr(dlogits), ptr(gscale), N, Cc, H * W, Cdst,
                               dtype_code(dtype) if dtype is not None else 0, stream_ptr()), 'mse')
    return dlogits


def loss_finalize(acc, beta):
    t = [torch.empty((), dtype=torch.float32, device=acc.device) for _ in range(3)]
    check(lib().pu_loss_finalize(ptr(acc), float(beta), ptr(t[0]), ptr(t[1]), ptr(t[2]), stream_ptr()),
          'loss_finalize')
    return t


def loss_bwd_scales(g_total, g_recon, g_kl, beta, device):
    out2 = torch.empty(2, dtype=torch.float32, device=device)
    check(lib().pu_loss_bwd_scales(ptr(g_total), ptr(g_recon), ptr(g_kl), float(beta), ptr(out2), stream_ptr()),
          'loss_bwd_scales')
    return out2


# ----------------------------------------------------------------------------- fcomb
def fcomb_fwd(feat, z, w0, b0, w1, b1, w2, b2, S=1, save_hidden=False):
    """feat [N,H,W,64]; z [N,S,L] (or [N,L]); returns (out [N,S,nc,H,W] or [N,nc,H,W] fp32, h1, h2)."""
    N, H, W, Cf = _nhwc(feat)
    assert Cf == 64
    Lz = z.shape[-1]
    nc = w2.shape[0]
    out = torch.empty((N, S, nc, H, W) if z.dim() == 3 else (N, nc, H, W), dtype=torch.float32, device=feat.device)
    h1 = torch.empty_like(feat) if save_hidden else None
    h2 = torch.empty_like(feat) if save_hidden else None
    a = L.PuFcombArgs(N, H * W, Lz, dtype_code(feat.dtype), S, nc, ptr(feat), ptr(z), ptr(w0), ptr(b0), ptr(w1),
                      ptr(b1), ptr(w2), ptr(b2), ptr(out), ptr(h1), ptr(h2))
    check(lib().pu_fcomb_fwd(C.byref(a), stream_ptr()), 'fcomb_fwd')
    return out, h1, h2


def fcomb_z_bwd(rmean, hw, z, w0, dw0, db0, accumulate=False):
    N, Lz = z.shape
    dz = torch.empty_like(z)
    check(lib().pu_fcomb_z_bwd(ptr(rmean), float(hw), ptr(z), ptr(w0), ptr(dz), ptr(dw0), ptr(db0), N, Lz,
                               int(accumulate), stream_ptr()), 'fcomb_z_bwd')
    return dz


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step):
    check(lib().pu_adamw(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr, beta1, beta2, eps, weight_decay, step,
                         stream_ptr()), 'adamw')
