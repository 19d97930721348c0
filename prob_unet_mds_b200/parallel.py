"""Data-parallel training and ensemble sharding over NCCL (one process per GPU, torch.distributed plumbing).

The reference has no multi-GPU code; the path shards naturally over the batch because every op is
per-sample (GroupNorm, attention) and the losses are *sums* over samples (prob_unet.py:227,230).  Hence the
gradient exchange is an all-reduce with SUM and **no** division by the world size: the result equals the
single-process gradient of the global batch.  Parameters without a gradient (unet.map_layer*) are skipped.
"""
import torch
import torch.distributed as dist


class GradAllReduce:
    """Bucketed SUM all-reduce of the live gradients over NVLink / NVSwitch."""

    def __init__(self, model, bucket_bytes=64 << 20):
        self.model = model
        self.bucket_bytes = bucket_bytes
        self._buckets = None

    def _build(self, params):
        buckets, cur, size = [], [], 0
        for p in params:
            cur.append(p)
            size += p.numel() * 4
            if size >= self.bucket_bytes:
                buckets.append(cur)
                cur, size = [], 0
        if cur:
            buckets.append(cur)
        self._buckets = buckets

    def allreduce(self):
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        params = [p for p in self.model.parameters() if p.grad is not None]
        if self._buckets is None or sum(len(b) for b in self._buckets) != len(params):
            self._build(list(reversed(params)))
        works = []
        for bucket in self._buckets:
            grads = [p.grad for p in bucket]
            flat = torch.cat([g.reshape(-1) for g in grads])
            works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, grads))
        for work, flat, grads in works:
            work.wait()
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n


def allreduce_losses(*scalars):
    """SUM the three logged scalars over ranks (they are sums over the local batch)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return scalars
    t = torch.stack([s.detach() for s in scalars])
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return tuple(t.unbind())


def shard_members(num_samples, rank=None, world=None):
    """Contiguous slice of ensemble members owned by this rank (members shard, inputs are encoded once)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    per = (num_samples + world - 1) // world
    lo = min(num_samples, rank * per)
    return lo, min(num_samples, lo + per)
