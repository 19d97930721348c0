"""Forward / hand-derived backward execution of the U-Net, the Gaussian encoders and Fcomb over the C ABI.

Nothing here computes with torch: torch allocates buffers (torch.empty), owns the parameters and the RNG, and
wires the result into autograd through two torch.autograd.Function classes (one per public entry point).
Every arithmetic step is a pu_* kernel launch (prob_unet_mds_b200/ops.py).

Data layout in HBM: activations NHWC in the compute dtype (bf16 by default, fp32 in "fp32 mode"); conv weights
are re-packed from the fp32 OIHW master parameters into [Cout][kh][kw][Cin] (forward) and [Cin][kh][kw][Cout]
(flipped, data-gradient) whenever a parameter's version counter changes; parameter gradients are fp32.
"""
import os
import weakref

import torch

from . import _lib as L
from . import ops


def default_compute_dtype():
    v = os.environ.get('PROBUNET_B200_DTYPE', 'bf16').lower()
    if v in ('bf16', 'bfloat16'):
        return torch.bfloat16
    if v in ('fp32', 'f32', 'float32'):
        return torch.float32
    raise ValueError(f'PROBUNET_B200_DTYPE={v!r}: expected bf16 or fp32')


def input_nhwc(x, dtype, extra=None):
    """fp32 NCHW model input (optionally followed channel-wise by `extra`, the posterior's target) -> NHWC in the
    compute dtype.  In bf16 mode the channels are zero-padded to 64 so that the first convolutions (Cin = 3 / 6) and
    their weight gradients run on the tcgen05 kernels instead of the CUDA-core ones (zero channels x zero-padded
    weights contribute nothing)."""
    N, Cx, H, W = x.shape
    Ctot = Cx + (extra.shape[1] if extra is not None else 0)
    pad = 64 if (dtype == torch.bfloat16 and H >= 8 and W >= 16) else Ctot
    out = ops.nchw_to_nhwc(x.contiguous(), dtype, Cdst=max(pad, Ctot))
    if extra is not None:
        ops.nchw_to_nhwc(extra.contiguous(), dtype, out=out, c_off=Cx)
    return out


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f'probunet_b200: {what} must live on a CUDA device (there is no CPU path); got {t.device}')


class _PackCache:
    """Packed copies of parameters, refreshed when the parameter is modified in place (optimizer step) or moved.

    Conv weights are registered with their packing spec: the first use packs them one by one; afterwards
    ``refresh()`` (called at the start of every forward) re-packs ALL stale ones in place with a single
    pu_pack_conv_weights_multi launch -- stable buffers (the TMA descriptor cache keeps hitting) and one launch per
    engine and step instead of one per weight and layout."""

    def __init__(self):
        self._store = {}      # key -> (tag, value)
        self._specs = {}      # key -> (param, dict(mode, perm, Ci_pad)) for conv weights
        self._table = None    # (signature, device table, n_items, total_tiles)

    def clear(self):
        self._store.clear()
        self._specs.clear()
        self._table = None

    @staticmethod
    def _tag(param):
        return (param._version, param.data_ptr())

    def get(self, key, param, make):
        tag = self._tag(param)
        hit = self._store.get(key)
        if hit is not None and hit[0] == tag:
            return hit[1]
        val = make()
        self._store[key] = (tag, val)
        return val

    def get_weight(self, key, param, dtype, mode, perm=None, Ci_pad=None):
        tag = self._tag(param)
        hit = self._store.get(key)
        if hit is not None and hit[0] == tag:
            return hit[1]
        val = ops.pack_weight(param.detach(), mode, dtype, perm=perm, Ci_pad=Ci_pad,
                              out=hit[1] if hit is not None else None)
        self._store[key] = (tag, val)
        self._specs[key] = (param, dict(mode=mode, perm=perm, Ci_pad=Ci_pad, dtype=dtype))
        return val

    def refresh(self):
        """Re-pack every registered conv weight whose parameter changed since it was packed: one launch."""
        stale = [k for k, (p, _) in self._specs.items() if self._store[k][0] != self._tag(p)]
        if not stale:
            return
        sig = tuple((k, self._specs[k][0].data_ptr(), self._store[k][1].data_ptr()) for k in stale)
        if self._table is None or self._table[0] != sig:
            items = []
            for k in stale:
                p, spec = self._specs[k]
                items.append(ops.pack_item(p.detach(), self._store[k][1], **spec))
            self._table = (sig,) + ops.pack_table(items, self._specs[stale[0]][0].device)
        _, table, n_items, total_tiles = self._table
        ops.pack_weights_multi(table, n_items, total_tiles)
        for k in stale:
            self._store[k] = (self._tag(self._specs[k][0]), self._store[k][1])


# ======================================================================================================================
# U-Net
# ======================================================================================================================
class UNetEngine:
    def __init__(self, unet, dtype):
        self.unet = unet
        self.dtype = dtype
        self.cache = _PackCache()
        self._perms = {}
        self._step = 0

    # ---- packed parameters ------------------------------------------------------------------------------------------
    def w_fwd(self, p, perm=None, Ci_pad=None):
        return self.cache.get_weight((id(p), 'f', self.dtype, Ci_pad), p, self.dtype, 0, perm=perm, Ci_pad=Ci_pad)

    def w_dgrad(self, p, perm=None):
        return self.cache.get_weight((id(p), 'd', self.dtype), p, self.dtype, 1, perm=perm)

    def out_pad(self):
        """Channel count the output head is computed with: a head with few output channels (the deterministic U-Net's
        64 -> 3 conv at full resolution, baseline/deterministic_unet.py:296) would run forward, data gradient and weight
        gradient on the CUDA-core kernels (18 of 70 ms per step at 256x256, batch 32); zero-padded to 64 output channels it
        runs on the tcgen05 kernels and the first `out_channels` channels are what leaves the engine."""
        Co = self.unet.out_conv.out_channels
        return 64 if (self.dtype == torch.bfloat16 and Co < 64) else Co

    def _w_out_padded(self, p, mode):
        def make():
            Co, Ci, k, _ = p.shape
            wp = ops.zeros((64, Ci, k, k), torch.float32, p.device)
            ops.clone(p.detach().reshape(-1), out=wp.reshape(-1)[:p.numel()])      # OIHW: the first Co rows are contiguous
            return ops.pack_weight(wp, mode, self.dtype)
        return self.cache.get((id(p), 'opad', mode, self.dtype), p, make)

    def _b_out_padded(self, p):
        def make():
            bp = ops.zeros((64,), torch.float32, p.device)
            ops.clone(p.detach(), out=bp[:p.numel()])
            return bp
        return self.cache.get((id(p), 'obpad'), p, make)

    def qkv_perm(self, C, heads, device):
        """my channel (j, head, d) -> reference channel head*192 + d*3 + j  (networks.py:180 reshape/unbind)."""
        key = (C, heads, str(device))
        if key not in self._perms:
            idx = []
            for j in range(3):
                for h in range(heads):
                    for d in range(64):
                        idx.append(h * 192 + d * 3 + j)
            self._perms[key] = torch.tensor(idx, dtype=torch.int32).to(device)
        return self._perms[key]

    def bias_perm(self, p, perm):
        return self.cache.get((id(p), 'bp'), p, lambda: ops.gather(p.detach(), perm))

    # ---- forward ----------------------------------------------------------------------------------------------------
    def forward(self, x, training, save, seed_base=0):
        """x: NHWC [N,H,W,in_channels] in self.dtype.  Returns (features NHWC, tape or None)."""
        u = self.unet
        tape = [] if save else None
        skips = []
        bi = 0
        # GroupNorm statistics come out of the producing convolutions' epilogues (per quad of channels); self._q maps an
        # activation tensor to them for its consumers (the next block's norm0, also across a skip concatenation)
        self._q = {}
        self._prod = {}       # id(activation) -> bias parameter of the conv that produced it (training tape only)
        # (a consumer can use the quads when its groups are made of whole quads -- _stats decides per tensor: always true for
        # the Probabilistic U-Net's 128-channel backbone; in the 64-channel deterministic baseline the 192-, 320- and
        # 448-channel norms have groups of 6, 10 and 14 channels and take the separate statistics pass)
        self._fused_stats = os.environ.get('PROBUNET_B200_FUSED_GN_STATS', '1') != '0'
        self.cache.refresh()
        for name, mod in u.enc.items():
            if isinstance(mod, torch.nn.Module) and hasattr(mod, 'norm0'):
                x, rec = self._block_fwd(mod, x, None, training, seed_base + bi, save)
                bi += 1
            else:
                xin = x
                # xin may carry zero channels up to a multiple of 64 (input_nhwc) so that the tcgen05 kernel applies
                x = self._conv_q(xin, self.w_fwd(mod.weight, Ci_pad=xin.shape[3]), mod.out_channels, mod.kernel,
                                 bias=mod.bias)
                rec = dict(kind='conv', mod=mod, xin=xin, out=x)
                self._prod[id(x)] = mod.bias
            if save:
                tape.append(rec)
            skips.append(x)
        for name, mod in u.dec.items():
            xb = None
            if x.shape[3] != mod.in_channels:
                xb = skips.pop()
            x, rec = self._block_fwd(mod, x, xb, training, seed_base + bi, save)
            bi += 1
            if save:
                tape.append(rec)
        st = self._stats(x)
        h = ops.gn_apply(x, st, u.out_norm.weight, u.out_norm.bias, silu=True, eps=u.out_norm.eps)
        if self.out_pad() != u.out_conv.out_channels:
            feat = ops.conv2d(h, self._w_out_padded(u.out_conv.weight, 0), self.out_pad(), 3,
                              bias=self._b_out_padded(u.out_conv.bias))
        else:
            feat = ops.conv2d(h, self.w_fwd(u.out_conv.weight), u.out_conv.out_channels, 3, bias=u.out_conv.bias)
        if save:
            tape.append(dict(kind='out', x=x, st=st, h=h, out=feat, bias_x=self._prod.get(id(x))))
        self._q = {}
        self._prod = {}
        return feat, tape

    def _conv_q(self, *args, **kw):
        """conv2d whose output feeds a GroupNorm: its epilogue also emits the per-quad (sum, sumsq)."""
        src, k = args[0], args[3]
        cin = src.shape[3] + (kw['src1'].shape[3] if kw.get('src1') is not None else 0)
        # the statistics epilogue hides behind the tile's MMAs from K = 9 x 128 on; the 64-channel 3x3 convs of the
        # deterministic baseline are epilogue-bound with it (+1.5 ms per step at 256x256, what the separate pass costs)
        hidden = k == 1 or k * k * cin >= 1152 or self.dtype != torch.bfloat16
        if not self._fused_stats or not hidden:
            return ops.conv2d(*args, **kw)
        y, q = ops.conv2d(*args, want_qstats=True, **kw)
        self._q[id(y)] = (weakref.ref(y), q)     # weak: the table must not keep activations alive in eval mode
        return y

    def _quads(self, x):
        ent = self._q.get(id(x)) if x is not None else None
        return ent[1] if ent is not None and ent[0]() is x else None

    def _stats(self, xa, xb=None):
        qa, qb = self._quads(xa), self._quads(xb)
        C = xa.shape[3] + (xb.shape[3] if xb is not None else 0)
        whole_quads = (C // ops.gn_groups(C)) % 4 == 0 and xa.shape[3] % 4 == 0
        if qa is None or (xb is not None and qb is None) or not whole_quads:
            return ops.gn_stats(xa, xb)
        return ops.gn_stats_from_quads(qa, qb)

    def _block_fwd(self, blk, xa, xb, training, seed, save):
        Cout = blk.out_channels
        rs = L.RS_UP if blk.up else (L.RS_DOWN if blk.down else L.RS_NONE)
        st0 = self._stats(xa, xb)
        h0 = ops.gn_apply(xa, st0, blk.norm0.weight, blk.norm0.bias, src1=xb, silu=True, resample=rs, eps=blk.norm0.eps)
        a = self._conv_q(h0, self.w_fwd(blk.conv0.weight), Cout, 3, bias=blk.conv0.bias)
        st1 = self._stats(a)
        p = float(blk.dropout) if training else 0.0
        # training: keep the dropout mask bits (1/16 of the activation's bytes) so that the data-gradient conv's
        # GroupNorm-backward epilogue reads them instead of regenerating the mask (a third of its instructions)
        mask1 = (torch.empty(a.numel() // 8, dtype=torch.uint8, device=a.device)
                 if (save and p > 0.0 and a.shape[3] % 32 == 0) else None)
        h1 = ops.gn_apply(a, st1, blk.norm1.weight, blk.norm1.bias, ada=blk.affine.bias, silu=True, dropout_p=p,
                          seed=seed, eps=blk.norm1.eps, keep_mask=mask1)
        w1 = self.w_fwd(blk.conv1.weight)
        if blk.skip is not None and blk.skip.weight is not None:
            if blk.up or blk.down:
                raise NotImplementedError('resampling 1x1 skip (resample_proj) is not on the path')
            s = ops.conv2d(xa, self.w_fwd(blk.skip.weight), Cout, 1, bias=blk.skip.bias, src1=xb)
            y = self._conv_q(h1, w1, Cout, 3, bias=blk.conv1.bias, residual=s, out=s)
        elif blk.up:
            s = ops.upsample2(xa)
            y = self._conv_q(h1, w1, Cout, 3, bias=blk.conv1.bias, residual=s, out=s)
        elif blk.down:
            s = ops.avgpool2(xa)
            y = self._conv_q(h1, w1, Cout, 3, bias=blk.conv1.bias, residual=s, out=s)
        else:
            y = self._conv_q(h1, w1, Cout, 3, bias=blk.conv1.bias, residual=xa)
        rec = None
        if save:
            rec = dict(kind='block', blk=blk, xa=xa, xb=xb, st0=st0, h0=h0, a=a, st1=st1, h1=h1, y=y, p=p, seed=seed,
                       rs=rs, mask1=mask1, bias_xa=self._prod.get(id(xa)), bias_xb=self._prod.get(id(xb)) if xb is not None else None)
        out = y
        if blk.num_heads:
            heads = blk.num_heads
            perm = self.qkv_perm(Cout, heads, y.device)
            st2 = self._stats(y)
            h2 = ops.gn_apply(y, st2, blk.norm2.weight, blk.norm2.bias, silu=False, eps=blk.norm2.eps)
            qkv = ops.conv2d(h2, self.w_fwd(blk.qkv.weight, perm), 3 * Cout, 1, bias=self.bias_perm(blk.qkv.bias, perm))
            att, lse = ops.attention_fwd(qkv, heads)
            out = self._conv_q(att, self.w_fwd(blk.proj.weight), Cout, 1, bias=blk.proj.bias, residual=y)
            if save:
                rec.update(st2=st2, h2=h2, qkv=qkv, att=att, lse=lse, perm=perm)
        if save:
            rec['out'] = out
            self._prod[id(out)] = blk.proj.bias if blk.num_heads else blk.conv1.bias
        return out, rec

    # ---- backward ---------------------------------------------------------------------------------------------------
    def _wgrad(self, grads, param, src0, dy, k, src1=None, perm=None):
        """Weight gradient, off the critical path: the data-gradient chain (dgrad -> GroupNorm backward -> dgrad ...)
        never needs it, so it is issued on a side stream where its tensor-core work overlaps the HBM-bound
        GroupNorm kernels of the main stream.  backward() joins the streams before returning."""
        side = self._side_stream()
        if side is None:
            g = grads.alloc(param)          # a view into an all-reduce bucket under data parallelism
            cin = src0.shape[3] + (src1.shape[3] if src1 is not None else 0)
            if k == 1 and perm is None and tuple(param.shape[:2]) == (dy.shape[3], cin):
                # a 1x1 weight's packed layout [Cout][1][Cin] IS its OIHW layout: the kernel writes the gradient in place
                ops.conv2d_wgrad(src0, dy, 1, src1=src1, dw=g.view(dy.shape[3], 1, 1, cin))
            else:
                dwp = ops.conv2d_wgrad(src0, dy, k, src1=src1)
                ops.unpack_wgrad(dwp, g, perm=perm)
            grads[id(param)] = g            # "written": may trigger the bucket's all-reduce
            return
        main = torch.cuda.current_stream()
        side.wait_stream(main)              # dy (and src) have been produced on the main stream
        with torch.cuda.stream(side):
            dwp = ops.conv2d_wgrad(src0, dy, k, src1=src1)
            g = grads.alloc(param)
            ops.unpack_wgrad(dwp, g, perm=perm)
            grads[id(param)] = g
        for t in (src0, dy, src1):
            if t is not None:
                t.record_stream(side)       # the caching allocator must not recycle them before the side stream is done
        self._side_used = True

    def _side_stream(self):
        # measured on B200 (batch 64): 154.0 -> 151.9 ms per step only -- a persistent wgrad CTA per SM leaves too
        # little room for the GroupNorm blocks to co-run -- so the side stream is opt-in
        if os.environ.get('PROBUNET_B200_WGRAD_STREAM', '0') != '1':
            return None
        if getattr(self, '_side', None) is None:
            self._side = torch.cuda.Stream()
        return self._side

    def join_side_stream(self):
        if getattr(self, '_side_used', False):
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_used = False

    def _dgrad_gn(self, dy, w, Cin, k, x0, st, norm, x1=None, ada=None, silu=True, p=0.0, seed=0, rs=L.RS_NONE,
                  keep_mask=None):
        """Data gradient of a conv whose input was GroupNorm(+SiLU)(+dropout) of x0 (|| x1).  Where the tcgen05 kernel
        applies (and no resampling sits between the norm and the conv) its epilogue already does the first pass of the
        GroupNorm backward: it returns du = dL/du and the per-(sample, channel) sums; otherwise (dL/dh, None)."""
        # The epilogue costs about as much as the pass it replaces, so it only pays where it hides behind the tile's MMAs:
        # K = taps x channels of dy >= 2304 (measured per layer, scripts/bench_layers.py: 3x3 convs from >= 256 channels)
        deep = k * k * dy.shape[3] >= 2304 or self._fused_gn_bwd_all
        if self._fused_gn_bwd and deep and rs == L.RS_NONE and ops.conv_tc_applies(dy, 0, Cin):
            d, sums, _ = ops.gn_bwd_epilogue(x0, st, norm.weight, norm.bias, src1=x1, ada=ada, silu=silu, dropout_p=p,
                                             seed=seed, eps=norm.eps, keep_mask=keep_mask)
            return ops.conv2d(dy, w, Cin, k, gn_bwd=d), sums
        return ops.conv2d(dy, w, Cin, k), None

    def backward(self, tape, dfeat, grads):
        """dfeat: NHWC gradient wrt the features.  Fills grads[id(param)] for every live parameter."""
        u = self.unet
        self._fused_gn_bwd = os.environ.get('PROBUNET_B200_FUSED_GN_BWD', '1') != '0'       # 0: never, 1: where it pays,
        self._fused_gn_bwd_all = os.environ.get('PROBUNET_B200_FUSED_GN_BWD', '1') == '2'   # 2: every eligible layer
        self._unresample = os.environ.get('PROBUNET_B200_UNRESAMPLE', '1') != '0'           # A/B switch, see _block_bwd
        gbuf = {}   # id(activation tensor) -> gradient tensor accumulated so far
        self._gsum = {}   # id(gradient tensor) -> its per-channel sums (= bias gradient of the producing conv),
                          # emitted by the gn_bwd call that wrote the tensor last

        rec = tape[-1]
        # order everywhere below: data gradient first, then the weight gradient (side stream, starts once the dgrad
        # has finished) so that it runs under the GroupNorm-backward kernels that follow on the main stream
        Co = u.out_conv.out_channels
        if dfeat.shape[3] != Co:            # padded head (out_pad): dfeat carries zero channels beyond Co
            grads[id(u.out_conv.bias)] = ops.clone(ops.bias_grad(dfeat)[:Co].contiguous(), out=grads.alloc(u.out_conv.bias))
            w_d = self._w_out_padded(u.out_conv.weight, 1)
        else:
            grads[id(u.out_conv.bias)] = ops.bias_grad(dfeat, db=grads.alloc(u.out_conv.bias))
            w_d = self.w_dgrad(u.out_conv.weight)
        dh, sums = self._dgrad_gn(dfeat, w_d, rec['h'].shape[3], 3, rec['x'], rec['st'], u.out_norm)
        self._wgrad(grads, u.out_conv.weight, rec['h'], dfeat, 3)
        dg = grads.alloc(u.out_norm.weight)
        db = grads.alloc(u.out_norm.bias)
        cs = self._bias_buf(grads, rec.get('bias_x'), rec['x'].shape[3], dh.device)
        dx, _ = ops.gn_bwd(rec['x'], rec['st'], u.out_norm.weight, u.out_norm.bias, dh, dg, db, silu=True,
                           eps=u.out_norm.eps, colsum0=cs, sums=sums, du_ready=sums is not None)
        grads[id(u.out_norm.weight)] = dg
        grads[id(u.out_norm.bias)] = db
        gbuf[id(rec['x'])] = dx
        self._gsum[id(dx)] = cs

        for rec in reversed(tape[:-1]):
            dout = gbuf.pop(id(rec['out']))
            if rec['kind'] == 'conv':
                mod = rec['mod']
                self._wgrad(grads, mod.weight, rec['xin'], dout, mod.kernel)
                grads[id(mod.bias)] = self._colsum(dout)
                continue
            self._block_bwd(rec, dout, grads, gbuf)
        self.join_side_stream()
        return grads

    def _block_bwd(self, rec, dz, grads, gbuf):
        blk = rec['blk']
        Cout = blk.out_channels
        xa, xb = rec['xa'], rec['xb']
        Cin = xa.shape[3] + (xb.shape[3] if xb is not None else 0)
        if blk.num_heads:
            heads = blk.num_heads
            perm = rec['perm']
            grads[id(blk.proj.bias)] = self._colsum(dz)
            datt = ops.conv2d(dz, self.w_dgrad(blk.proj.weight), Cout, 1)
            self._wgrad(grads, blk.proj.weight, rec['att'], dz, 1)
            dqkv, dbq = ops.attention_bwd(rec['qkv'], rec['att'], datt, rec['lse'], heads, want_dbias=True)
            gq = grads.alloc(blk.qkv.bias)
            ops.scatter(dbq, perm, gq)
            grads[id(blk.qkv.bias)] = gq
            dh2, sums = self._dgrad_gn(dqkv, self.w_dgrad(blk.qkv.weight, perm), Cout, 1, rec['y'], rec['st2'], blk.norm2,
                                       silu=False)
            self._wgrad(grads, blk.qkv.weight, rec['h2'], dqkv, 1, perm=perm)
            dg = grads.alloc(blk.norm2.weight)
            db = grads.alloc(blk.norm2.bias)
            cs = self._bias_buf(grads, blk.conv1.bias, Cout, dz.device)
            dy, _ = ops.gn_bwd(rec['y'], rec['st2'], blk.norm2.weight, blk.norm2.bias, dh2, dg, db, silu=False,
                               eps=blk.norm2.eps, dres=dz, colsum0=cs, sums=sums, du_ready=sums is not None)
            self._gsum[id(dy)] = cs
            grads[id(blk.norm2.weight)] = dg
            grads[id(blk.norm2.bias)] = db
        else:
            dy = dz
        # conv1
        bias_dy = self._colsum(dy)
        grads[id(blk.conv1.bias)] = bias_dy
        dh1, sums = self._dgrad_gn(dy, self.w_dgrad(blk.conv1.weight), Cout, 3, rec['a'], rec['st1'], blk.norm1,
                                   ada=blk.affine.bias, p=rec['p'], seed=rec['seed'], keep_mask=rec.get('mask1'))
        self._wgrad(grads, blk.conv1.weight, rec['h1'], dy, 3)
        dg = grads.alloc(blk.norm1.weight)
        db = grads.alloc(blk.norm1.bias)
        dada = grads.alloc(blk.affine.bias)
        cs = self._bias_buf(grads, blk.conv0.bias, Cout, dy.device)
        da, _ = ops.gn_bwd(rec['a'], rec['st1'], blk.norm1.weight, blk.norm1.bias, dh1, dg, db, ada=blk.affine.bias,
                           dada=dada, silu=True, dropout_p=rec['p'], seed=rec['seed'], eps=blk.norm1.eps, colsum0=cs,
                           sums=sums, du_ready=sums is not None, keep_mask=rec.get('mask1'))
        grads[id(blk.norm1.weight)] = dg
        grads[id(blk.norm1.bias)] = db
        grads[id(blk.affine.bias)] = dada
        # conv0
        grads[id(blk.conv0.bias)] = cs
        rs = rec['rs']
        dh0, sums0 = self._dgrad_gn(da, self.w_dgrad(blk.conv0.weight), Cin, 3, xa, rec['st0'], blk.norm0, x1=xb, rs=rs)
        # skip branch
        if blk.skip is not None and blk.skip.weight is not None:
            grads[id(blk.skip.bias)] = ops.clone(bias_dy, out=grads.alloc(blk.skip.bias))   # same values as conv1.bias' gradient
            dres = ops.conv2d(dy, self.w_dgrad(blk.skip.weight), Cin, 1)
            dres_rs = L.RS_NONE
            self._wgrad(grads, blk.skip.weight, xa, dy, 1, src1=xb)
        else:
            dres = dy
            dres_rs = rs
        self._wgrad(grads, blk.conv0.weight, rec['h0'], da, 3)
        dg = grads.alloc(blk.norm0.weight)
        db = grads.alloc(blk.norm0.bias)
        ga = gbuf.get(id(xa))
        gb = gbuf.get(id(xb)) if xb is not None else None
        csa = self._bias_buf(grads, rec.get('bias_xa'), xa.shape[3], dy.device)
        csb = self._bias_buf(grads, rec.get('bias_xb'), xb.shape[3], dy.device) if xb is not None else None
        if rs != L.RS_NONE and self._unresample:
            # up / down blocks: undo the 2x resampling of both incoming gradients in a small pass of their own, so that
            # the GroupNorm backward runs its streaming kernels (the gathering form is 2.2x slower per element)
            dh0 = ops.resample_grad(dh0, rs)
            if dres_rs != L.RS_NONE:
                dres = ops.resample_grad(dres, dres_rs)
            rs = dres_rs = L.RS_NONE
        dxa, dxb = ops.gn_bwd(xa, rec['st0'], blk.norm0.weight, blk.norm0.bias, dh0, dg, db, src1=xb, silu=True,
                              resample=rs, eps=blk.norm0.eps, dres=dres, dres_resample=dres_rs,
                              dx0=ga, dx1=gb, acc0=ga is not None, acc1=gb is not None, colsum0=csa, colsum1=csb,
                              sums=sums0, du_ready=sums0 is not None)
        grads[id(blk.norm0.weight)] = dg
        grads[id(blk.norm0.bias)] = db
        gbuf[id(xa)] = dxa
        self._gsum[id(dxa)] = csa
        if xb is not None:
            gbuf[id(xb)] = dxb
            self._gsum[id(dxb)] = csb

    @staticmethod
    def _bias_buf(grads, bias_param, n, device):
        """fp32 [n] buffer for the per-channel sums of a data gradient == the bias gradient of the conv that produced the
        tensor: taken from the gradient sink (under data parallelism a view into an all-reduce bucket, no staging copy)
        when that conv's bias parameter is known."""
        if bias_param is not None and bias_param.numel() == n:
            return grads.alloc(bias_param)
        return torch.empty(n, dtype=torch.float32, device=device)

    def _colsum(self, g):
        """Per-channel sum of a gradient tensor: taken from the gn_bwd call that wrote it, else computed."""
        cs = self._gsum.get(id(g))
        return cs if cs is not None else ops.bias_grad(g)


def new_grad_sink(model):
    """Where backward puts parameter gradients: plain tensors, or all-reduce buckets when parallel.DataParallel is
    attached to the model."""
    factory = getattr(model, '_grad_sink_factory', None)
    if factory is not None:
        return factory()
    from .parallel import GradSink
    return GradSink()


def _collect_grads(params, grads, zero_cache):
    """Gradient tuple for autograd.Function.backward: live grads, zeros for affine.weight (the reference gives
    them an all-zero gradient because emb == 0), None for the never-used map_layer*."""
    out = []
    for name, p in params:
        g = grads.get(id(p))
        if g is None and name.endswith('affine.weight'):
            z = zero_cache.get(id(p))
            if z is None or z.shape != p.shape or z.device != p.device:
                z = torch.zeros_like(p)
                zero_cache[id(p)] = z
            g = z
        out.append(g)
    return out


class _UNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, unet, x, *params):
        eng = unet.engine()
        ctx.unet = unet
        xs = input_nhwc(x, eng.dtype)
        eng._step += 1
        feat, tape = eng.forward(xs, unet.training, True, seed_base=_seed_base(eng._step))
        ctx.tape = tape
        return ops.nhwc_to_nchw(feat, C=unet.out_conv.out_channels)

    @staticmethod
    def backward(ctx, dout):
        unet = ctx.unet
        eng = unet.engine()
        dfeat = ops.nchw_to_nhwc(dout.contiguous().float(), eng.dtype, Cdst=eng.out_pad())
        grads = new_grad_sink(unet)
        eng.backward(ctx.tape, dfeat, grads)
        grads.finish()
        ctx.tape = None
        named = list(unet.named_parameters())
        if not hasattr(unet, '_zero_cache'):
            unet._zero_cache = {}
        return (None, None) + tuple(_collect_grads(named, grads, unet._zero_cache))


def _seed_base(step):
    # per-process stream: torch's seed, the rank and the step counter; 64 blocks per step at most
    rank = int(os.environ.get('RANK', '0'))
    return (torch.initial_seed() * 1000003 + rank * 7919 + step) * 64 % (1 << 62)


def unet_apply(unet, x):
    _require_cuda(x, 'input')
    params = [p for _, p in unet.named_parameters()]
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return _UNetFunction.apply(unet, x, *params)
    eng = unet.engine()
    xs = input_nhwc(x, eng.dtype)
    eng._step += 1
    feat, _ = eng.forward(xs, unet.training, False, seed_base=_seed_base(eng._step))
    return ops.nhwc_to_nchw(feat, C=unet.out_conv.out_channels)


# ======================================================================================================================
# prior / posterior encoders (prob_unet.py:8-78)
# ======================================================================================================================
class GaussianEngine:
    def __init__(self, net, dtype):
        self.net = net
        self.dtype = dtype
        self.cache = _PackCache()

    def w_fwd(self, p, Ci_pad=None):
        # bf16 mode: hi/lo split packing (pu_pack_conv_weight mode 2).  The KL term is a function of mu_q - mu_p, and
        # the bf16 rounding of the encoder WEIGHTS alone moves it by 1e-3..2e-3 (activations: 5e-5); the split keeps the
        # tcgen05 kernel and doubles K of these forward convs (<1 % of the step's FLOPs).
        mode = 2 if self.dtype == torch.bfloat16 else 0
        return self.cache.get_weight((id(p), 'f', self.dtype, Ci_pad, mode), p, self.dtype, mode, Ci_pad=Ci_pad)

    def w_dgrad(self, p):
        return self.cache.get_weight((id(p), 'd', self.dtype), p, self.dtype, 1)

    def convs(self):
        return [m for m in self.net.encoder if isinstance(m, torch.nn.Conv2d)]

    def forward(self, xin, save):
        """xin: NHWC [N,H,W,Cin] (x, or x||target for the posterior).  Returns (mu, log_sigma, tape)."""
        self.cache.refresh()
        convs = self.convs()
        x = xin
        acts = []
        for i, c in enumerate(convs):
            split = self.dtype == torch.bfloat16
            r = ops.conv2d(x, self.w_fwd(c.weight, Ci_pad=x.shape[3]), c.out_channels, 3, bias=c.bias, relu=True,
                           src1=x if split else None)
            last = i == len(convs) - 1
            acts.append((x, r))
            if not last:
                x = ops.avgpool2(r)
        # AvgPool2d(2) followed by the global mean == global mean of r when H, W are even; for odd sizes (inputs that are
        # multiples of 8 but not of 16) AvgPool2d floors, i.e. the last row / column never reaches the mean
        r_last = acts[-1][1]
        h2, w2 = r_last.shape[1] // 2 * 2, r_last.shape[2] // 2 * 2
        if h2 == 0 or w2 == 0:
            raise ValueError('prior/posterior encoder: input too small for len(num_filters) 2x pools')
        crop = (h2, w2) if (h2, w2) != tuple(r_last.shape[1:3]) else None
        m = ops.global_mean(r_last[:, :h2, :w2, :].contiguous() if crop else r_last)
        net = self.net
        Lz = net.latent_dim
        mu = ops.heads_fwd(m, net.conv_mu.weight, net.conv_mu.bias)
        ls = ops.heads_fwd(m, net.conv_log_sigma.weight, net.conv_log_sigma.bias)
        tape = dict(acts=acts, m=m, crop=crop) if save else None
        return mu, ls, tape

    def backward(self, tape, dmu, dls, grads):
        net = self.net
        convs = self.convs()
        m = tape['m']
        gw = grads.alloc(net.conv_mu.weight)
        gb = grads.alloc(net.conv_mu.bias)
        dm = ops.heads_bwd(m, net.conv_mu.weight, dmu, gw, gb)
        grads[id(net.conv_mu.weight)] = gw
        grads[id(net.conv_mu.bias)] = gb
        gw = grads.alloc(net.conv_log_sigma.weight)
        gb = grads.alloc(net.conv_log_sigma.bias)
        ops.heads_bwd(m, net.conv_log_sigma.weight, dls, gw, gb, dm=dm)
        grads[id(net.conv_log_sigma.weight)] = gw
        grads[id(net.conv_log_sigma.bias)] = gb
        dp = None
        for i in reversed(range(len(convs))):
            c = convs[i]
            x, r = tape['acts'][i]
            if i == len(convs) - 1:
                if tape.get('crop'):
                    h2, w2 = tape['crop']
                    dr = torch.zeros_like(r)
                    dr[:, :h2, :w2, :] = ops.relu_mean_bwd(dm, r[:, :h2, :w2, :].contiguous())
                else:
                    dr = ops.relu_mean_bwd(dm, r)
            else:
                dr = ops.relu_pool_bwd(dp, r)
            dwp = ops.conv2d_wgrad(x, dr, 3)
            g = grads.alloc(c.weight)
            ops.unpack_wgrad(dwp, g)
            grads[id(c.weight)] = g
            grads[id(c.bias)] = ops.bias_grad(dr, db=grads.alloc(c.bias))
            if i > 0:
                dp = ops.conv2d(dr, self.w_dgrad(c.weight), c.in_channels, 3)
        return grads
