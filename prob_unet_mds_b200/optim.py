"""Fused multi-tensor AdamW.

Drop-in for ``torch.optim.AdamW(params, lr, betas, eps, weight_decay)`` as the reference constructs it (``main.py:95``):
same defaults, same update (decoupled weight decay, bias-corrected moments), same ``state_dict`` layout
(``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter), but ``step()`` is ONE kernel launch for the whole model
(``pu_adamw_multi``) instead of a loop over 446 tensors.  fp32 CUDA parameters only; no amsgrad / maximize.
"""
import numpy as np
import torch

from . import _lib as L
from .ops import check, stream_ptr

CHUNK = 65536                      # elements per table entry (include/probunet_b200.h, PuAdamWChunk)
_CHUNK_DT = np.dtype([('p', '<u8'), ('g', '<u8'), ('m', '<u8'), ('v', '<u8'), ('n', '<i4'), ('pad', '<i4')])


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, maximize=False):
        if amsgrad or maximize:
            raise ValueError('prob_unet_mds_b200.optim.AdamW: amsgrad / maximize are not supported')
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError('prob_unet_mds_b200.optim.AdamW: invalid hyper-parameter')
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._plans = {}

    def _plan(self, key, tensors):
        """Chunk layout of a list of parameters: (tensor index, element offset, length) per table entry; cached."""
        plan = self._plans.get(key)
        if plan is None:
            idx, off, cnt = [], [], []
            for i, p in enumerate(tensors):
                n = p.numel()
                for o in range(0, n, CHUNK):
                    idx.append(i)
                    off.append(o)
                    cnt.append(min(CHUNK, n - o))
            plan = (np.asarray(idx, dtype=np.int64), np.asarray(off, dtype=np.uint64) * np.uint64(4),
                    np.asarray(cnt, dtype=np.int32))
            self._plans[key] = plan
        return plan

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            by_step = {}
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.dtype != torch.float32 or not p.is_cuda or p.grad.dtype != torch.float32:
                    raise RuntimeError('prob_unet_mds_b200.optim.AdamW needs fp32 CUDA parameters and gradients')
                if p.grad.is_sparse:
                    raise RuntimeError('prob_unet_mds_b200.optim.AdamW does not support sparse gradients')
                st = self.state[p]
                if not st:
                    st['step'] = torch.tensor(0.0)
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st['step'] += 1
                by_step.setdefault(int(st['step']), []).append(p)
            for step, ps in by_step.items():          # normally a single bucket: every parameter has the same age
                if not all(p.is_contiguous() for p in ps):
                    raise RuntimeError('prob_unet_mds_b200.optim.AdamW needs contiguous parameters')
                grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
                key = (gi, tuple(id(p) for p in ps))
                idx, off, cnt = self._plan(key, ps)
                table = np.empty(len(idx), dtype=_CHUNK_DT)
                for name, ts in (('p', ps), ('g', grads), ('m', [self.state[p]['exp_avg'] for p in ps]),
                                 ('v', [self.state[p]['exp_avg_sq'] for p in ps])):
                    base = np.fromiter((t.data_ptr() for t in ts), dtype=np.uint64, count=len(ts))
                    table[name] = base[idx] + off
                table['n'] = cnt
                table['pad'] = 0
                dev = ps[0].device
                host = torch.from_numpy(table.view(np.uint8)).pin_memory()
                tab = host.to(dev, non_blocking=True)
                b1, b2 = group['betas']
                with torch.cuda.device(dev):
                    check(L.lib().pu_adamw_multi(tab.data_ptr(), len(idx), float(group['lr']), float(b1), float(b2),
                                                 float(group['eps']), float(group['weight_decay']), step, stream_ptr()),
                          'adamw_multi')
                # the kernel wrote the parameters through raw pointers: tell autograd (and every cache keyed on
                # param._version, e.g. engine._PackCache's packed conv weights) that they changed
                torch.autograd.graph.increment_version(ps)
                # keep the table and the (possibly re-laid-out) gradients alive until the kernel has been enqueued
                self._keep = (host, tab, grads)
        return loss
