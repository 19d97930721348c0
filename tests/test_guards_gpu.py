"""Out-of-bounds write check for the kernels of the hot path (GPU only).

compute-sanitizer is not available on the GPU pool, so this is the in-house bounds check: every tensor that the op
wrappers allocate while a kernel family runs is carved out of a larger buffer whose head and tail are filled with a
sentinel; after the kernels finish the sentinels must be intact.  Shapes are chosen ragged (partial TMA tiles, partial
128-row tiles, two-source inputs) because that is where an epilogue or a reduction would overrun.
"""
import pytest
import torch

from prob_unet_mds_b200 import _lib as L
from prob_unet_mds_b200 import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'
GUARD = 4096          # bytes on either side; a multiple of 1024 keeps the payload aligned for TMA
SENTINEL = 0xA5


class GuardedAlloc:
    def __init__(self):
        self.bufs = []
        self._empty = torch.empty

    def empty(self, *size, dtype=None, device=None, **kw):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            size = tuple(size[0])
        dtype = dtype or torch.float32
        if device is None or torch.device(device).type != 'cuda':
            return self._empty(size, dtype=dtype, device=device, **kw)
        n = 1
        for s in size:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        pad = (-nbytes) % 1024
        raw = self._empty(GUARD + nbytes + pad + GUARD, dtype=torch.uint8, device=device)
        raw.fill_(SENTINEL)
        self.bufs.append((raw, nbytes))
        return raw[GUARD:GUARD + nbytes].view(dtype).view(size)

    def empty_like(self, t, **kw):
        return self.empty(tuple(t.shape), dtype=kw.get('dtype', t.dtype), device=t.device)

    def check(self):
        torch.cuda.synchronize()
        assert self.bufs, 'no guarded allocation was made'
        for raw, nbytes in self.bufs:
            head = raw[:GUARD]
            tail = raw[GUARD + nbytes + ((-nbytes) % 1024):]
            assert bool((head == SENTINEL).all()), f'write before a {nbytes}-byte output'
            assert bool((tail == SENTINEL).all()), f'write past a {nbytes}-byte output'


@pytest.fixture
def guarded(monkeypatch):
    g = GuardedAlloc()
    monkeypatch.setattr(torch, 'empty', g.empty)
    monkeypatch.setattr(torch, 'empty_like', g.empty_like)
    yield g
    monkeypatch.undo()


def _rnd(*shape, dtype=torch.bfloat16, seed=0):
    gen = torch.Generator(device='cpu').manual_seed(seed)
    return torch.randn(*shape, generator=gen).to(DEV).to(dtype)


@pytest.mark.parametrize('N,H,W,C0,C1,Cout,k', [(2, 20, 24, 64, 0, 128, 3), (1, 40, 48, 128, 64, 256, 3),
                                                 (3, 16, 16, 256, 0, 64, 1), (2, 24, 40, 128, 0, 128, 3)])
def test_conv_family_stays_in_bounds(guarded, N, H, W, C0, C1, Cout, k):
    x0 = _rnd(N, H, W, C0, seed=1)
    x1 = _rnd(N, H, W, C1, seed=2) if C1 else None
    dy = _rnd(N, H, W, Cout, seed=3)
    w = torch.randn(Cout, C0 + C1, k, k, device=DEV) * 0.05
    bias = torch.randn(Cout, device=DEV)
    res = _rnd(N, H, W, Cout, seed=4)
    wf = ops.pack_weight(w, 0, torch.bfloat16)
    wd = ops.pack_weight(w, 1, torch.bfloat16)
    ops.conv2d(x0, wf, Cout, k, bias=bias, src1=x1, residual=res, flags=L.CONV_FORCE_TC)
    ops.conv2d(dy, wd, C0 + C1, k, flags=L.CONV_FORCE_TC)
    dwp = ops.conv2d_wgrad(x0, dy, k, src1=x1, flags=L.CONV_FORCE_TC)
    g = torch.empty_like(w)
    ops.unpack_wgrad(dwp, g)
    ops.bias_grad(dy)
    guarded.check()


@pytest.mark.parametrize('N,T,heads', [(2, 256, 2), (1, 384, 4), (1, 1024, 1)])
def test_attention_stays_in_bounds(guarded, N, T, heads):
    C = heads * 64
    qkv = _rnd(N, T, 3 * C, seed=5)
    out, lse = ops.attention_fwd(qkv, heads, flags=L.CONV_FORCE_TC)
    dout = _rnd(N, T, C, seed=6)
    ops.attention_bwd(qkv, out, dout, lse, heads)
    guarded.check()


@pytest.mark.parametrize('N,H,W,C0,C1', [(2, 20, 24, 128, 0), (3, 9, 7, 64, 64), (1, 64, 64, 256, 128)])
def test_groupnorm_stays_in_bounds(guarded, N, H, W, C0, C1):
    C = C0 + C1
    x0 = _rnd(N, H, W, C0, seed=7)
    x1 = _rnd(N, H, W, C1, seed=8) if C1 else None
    gamma, beta = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    ada = torch.randn(2 * C, device=DEV) * 0.1
    st = ops.gn_stats(x0, x1)
    ops.gn_apply(x0, st, gamma, beta, src1=x1, ada=ada, silu=True, dropout_p=0.1, seed=3)
    dy = _rnd(N, H, W, C, seed=9)
    dres = _rnd(N, H, W, C, seed=10)
    dg, db, da = torch.empty(C, device=DEV), torch.empty(C, device=DEV), torch.empty(2 * C, device=DEV)
    cs0 = torch.zeros(C0, device=DEV)
    cs1 = torch.zeros(C1, device=DEV) if C1 else None
    ops.gn_bwd(x0, st, gamma, beta, dy, dg, db, src1=x1, ada=ada, dada=da, silu=True, dropout_p=0.1, seed=3, dres=dres,
               colsum0=cs0, colsum1=cs1)
    guarded.check()


@pytest.mark.parametrize('N,H,W,S', [(2, 16, 16, 7), (1, 32, 24, 130)])
def test_fcomb_members_stay_in_bounds(guarded, N, H, W, S):
    Lz = 16
    feat = _rnd(N, H, W, 64, seed=11)
    z = torch.randn(N, S, Lz, device=DEV)
    w0 = torch.randn(64, 64 + Lz, 1, 1, device=DEV) / 8
    w1 = torch.randn(64, 64, 1, 1, device=DEV) / 8
    w2 = torch.randn(3, 64, 1, 1, device=DEV) / 8
    b0, b1, b2 = torch.randn(64, device=DEV), torch.randn(64, device=DEV), torch.randn(3, device=DEV)
    ops.fcomb_fwd(feat, z, w0, b0, w1, b1, w2, b2, S=S)
    guarded.check()
