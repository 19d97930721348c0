"""Data-parallel training and ensemble sharding (one process per GPU, torch.distributed / NCCL over NVLink-NVSwitch).

The reference has no multi-GPU code.  The path shards naturally over the batch: every op is per-sample
(GroupNorm, attention) and the losses are *sums* over samples (prob_unet.py:227,230), so the only exchange step is
an all-reduce(SUM) of the parameter gradients -- no division by the world size: the result equals the single-process
gradient of the global batch.  Parameters without a gradient (unet.map_layer*) never enter a bucket.

Overlap with backward: the hand-derived backward (engine.py) asks a *gradient sink* for the memory of every
parameter gradient and tells it when the gradient has been written.  `BucketedSink` hands out views into a few flat
fp32 buckets laid out in backward order; as soon as the last gradient of a bucket is enqueued, the bucket is
all-reduced on a side stream while the main stream keeps running the backward kernels.  There is no copy in or out
of the buckets: the views are what autograd stores in `param.grad`.
"""
import contextlib

import torch
import torch.distributed as dist


class GradSink(dict):
    """Default sink: independent tensors, no communication.  Maps id(param) -> gradient tensor."""

    def alloc(self, p):
        return torch.empty_like(p)

    def finish(self):
        pass


class BucketedSink(GradSink):
    def __init__(self, owner):
        super().__init__()
        self.owner = owner
        self._pending = {}

    def alloc(self, p):
        slot = self.owner.slot_of(p)
        if slot is None:
            return torch.empty_like(p)
        b, off = slot
        view = self.owner.flat[b][off:off + p.numel()].view_as(p)
        if p.grad is not None and p.grad.data_ptr() == view.data_ptr():
            # gradient accumulation (a second backward without zero_grad(set_to_none=True)): param.grad still IS the
            # bucket view of the previous backward.  Move the old gradient out of the bucket so that this backward
            # can overwrite the bucket, all-reduce it, and let autograd add it to the preserved old value.
            p.grad = p.grad.clone()
        return view

    def __setitem__(self, pid, g):
        super().__setitem__(pid, g)
        self.owner.ready(pid, g, self)

    def finish(self):
        self.owner.finish(self)


class DataParallel:
    """Attach to a model: ``dp = DataParallel(model)``; afterwards ``model.elbo(...)[0].backward()`` leaves globally
    summed gradients in ``param.grad`` on every rank.  The first backward learns the order in which gradients are
    produced (and reduces them unbucketed at the end); later steps use the overlapped buckets."""

    def __init__(self, model, bucket_bytes=48 << 20, group=None):
        self.model = model
        self.bucket_bytes = bucket_bytes
        self.group = group
        self.order = []          # params in the order their gradients are first allocated
        self.slots = None        # id(param) -> (bucket, offset)
        self.flat = None
        self.bucket_members = None
        self._count = None
        self._works = []
        self._stream = None
        self.by_id = {id(p): p for p in model.parameters()}
        self._sync = True        # False inside no_sync(): gradients stay local (gradient accumulation micro-steps)
        self._local_acc = False  # param.grad currently holds un-reduced gradients of no_sync micro-steps
        self.bucket_params = None
        model._grad_sink_factory = self.make_sink
        self.broadcast_parameters()

    # -- setup --------------------------------------------------------------------------------------------------------
    def world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def broadcast_parameters(self):
        if self.world() == 1:
            return
        with torch.no_grad():
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t, src=0, group=self.group)     # on the tensor itself: bumps its version counter

    def make_sink(self):
        if not self._sync:
            self._local_acc = True
            return GradSink()        # plain tensors, no communication; autograd accumulates them into param.grad
        return BucketedSink(self)

    @contextlib.contextmanager
    def no_sync(self):
        """Gradient accumulation (a global batch larger than world x per-GPU micro-batch, BASELINE.json configs[2] on
        fewer than 8 GPUs): backward passes inside the context leave their gradients local in ``param.grad``; the
        first backward outside it folds them into the all-reduce buckets, so the exchanged gradient is the SUM over all
        micro-batches of all ranks.  Start every accumulation window with ``zero_grad(set_to_none=True)``."""
        prev, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = prev

    def _fold_local(self, b):
        """Adds the locally accumulated (un-reduced) gradients of bucket b's parameters into the bucket and clears
        them, so that autograd stores the reduced bucket view as param.grad afterwards."""
        views, olds = [], []
        for p, off in self.bucket_params[b]:
            if p.grad is not None:
                views.append(self.flat[b][off:off + p.numel()].view_as(p))
                olds.append(p.grad)
                p.grad = None
        if views:
            torch._foreach_add_(views, olds)

    def slot_of(self, p):
        return None if self.slots is None else self.slots.get(id(p))

    def record(self, p):
        if all(q is not p for q in self.order):
            self.order.append(p)

    def _build_buckets(self):
        self.slots, self.flat, self.bucket_members, self.bucket_params = {}, [], [], []
        cur, size = [], 0
        groups = []
        for p in self.order:
            cur.append(p)
            size += p.numel() * 4
            if size >= self.bucket_bytes:
                groups.append(cur)
                cur, size = [], 0
        if cur:
            groups.append(cur)
        for b, ps in enumerate(groups):
            n = sum(((p.numel() + 3) // 4) * 4 for p in ps)   # keep every view 16-byte aligned
            self.flat.append(torch.zeros(n, dtype=torch.float32, device=ps[0].device))
            off = 0
            members = []
            for p in ps:
                self.slots[id(p)] = (b, off)
                members.append((p, off))
                off += ((p.numel() + 3) // 4) * 4
            self.bucket_members.append(len(ps))
            self.bucket_params.append(members)

    # -- per-step protocol --------------------------------------------------------------------------------------------
    def ready(self, pid, g, sink):
        if self.slots is None:
            p = self.by_id.get(pid)
            if p is not None:
                self.record(p)      # first step: learn the order in which gradients complete
            return
        if self.world() == 1:
            return
        slot = self.slots.get(pid)
        if slot is None:
            return
        b, off = slot
        view = self.flat[b][off:off + g.numel()]
        if g.data_ptr() != view.data_ptr():      # produced elsewhere (e.g. a cached zero gradient): stage it
            p = self.by_id.get(pid)
            if p is not None and p.grad is not None and p.grad.data_ptr() == view.data_ptr():
                p.grad = p.grad.clone()          # gradient accumulation: see BucketedSink.alloc
            view.copy_(g.reshape(-1))
            dict.__setitem__(sink, pid, view.view_as(g))
        cnt = sink._pending.get(b, 0) + 1
        sink._pending[b] = cnt
        if cnt == self.bucket_members[b]:
            if self._local_acc:
                self._fold_local(b)
            self._launch(self.flat[b])

    def _launch(self, flat):
        if flat.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream()
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self, sink):
        if self.world() > 1:
            if self.slots is None:
                # first step: order just learned; reduce everything now, bucketed from the next step on
                for pid, g in list(sink.items()):
                    p = self.by_id.get(pid)
                    if self._local_acc and p is not None and p.grad is not None:
                        g.add_(p.grad)
                        p.grad = None
                    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            else:
                for b, members in enumerate(self.bucket_members):
                    if sink._pending.get(b, 0) != members and sink._pending.get(b, 0) > 0:
                        if self._local_acc:
                            self._fold_local(b)
                        self._launch(self.flat[b])     # partially filled (some grads legitimately absent)
                for w in self._works:
                    w.wait()
                self._works = []
                if self._stream is not None:
                    torch.cuda.current_stream().wait_stream(self._stream)
        self._local_acc = False
        if self.slots is None and self.order:
            self._build_buckets()


def allreduce_losses(*scalars, group=None):
    """SUM the logged scalars over ranks (they are sums over the local batch)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return scalars
    t = torch.stack([s.detach() for s in scalars])
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tuple(t.unbind())


def shard_range(n, rank=None, world=None):
    """Contiguous slice [lo, hi) of n work items (inputs to encode, ensemble members to decode) owned by a rank."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def ensemble_sharded(model, x, num_samples, eps=None, group=None):
    """Ensemble generation over all ranks: the *inputs* are sharded for the U-Net / prior encode, the
    features and (mu, log_sigma) are all-gathered (the one real exchange step), then every rank decodes its slice of
    the *members* for all inputs with the fused Fcomb kernel.  Returns this rank's [B, S_local, C, H, W] block and
    its member range."""
    from . import ops
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = x.shape[0]
    dt = model.compute_dtype
    model.unet.compute_dtype = dt
    lo, hi = shard_range(B, rank, world)
    with torch.no_grad():
        xs = x[lo:hi].contiguous()
        from . import engine
        feat, _ = model.unet.engine().forward(engine.input_nhwc(xs, dt), False, False)
        mu, ls, _ = model.prior.engine(dt).forward(model.prior._input(xs, None, dt), save=False)
        if world > 1:
            per = (B + world - 1) // world
            assert B % world == 0, 'ensemble_sharded: the input batch must divide evenly over the ranks'
            feat_all = torch.empty((B,) + feat.shape[1:], dtype=feat.dtype, device=feat.device)
            mu_all = torch.empty((B, mu.shape[1]), dtype=mu.dtype, device=mu.device)
            ls_all = torch.empty_like(mu_all)
            dist.all_gather_into_tensor(feat_all, feat.contiguous(), group=group)
            dist.all_gather_into_tensor(mu_all, mu.contiguous(), group=group)
            dist.all_gather_into_tensor(ls_all, ls.contiguous(), group=group)
            assert per * world == B
        else:
            feat_all, mu_all, ls_all = feat, mu, ls
        s_lo, s_hi = shard_range(num_samples, rank, world)
        eps_local = None if eps is None else eps[:, s_lo:s_hi].contiguous()
        out = model.decode_ensemble(feat_all, mu_all, ls_all, s_hi - s_lo, eps_local)
    return out, (s_lo, s_hi)
