"""Data-parallel pieces of the training step (BASELINE config 5).

The reference wraps the models in `nn.DataParallel` (train/adversarial_wasserstein_gp.py:192-196): one process
scatters the batch over the GPUs, gathers the outputs and reduces the gradients on GPU 0 every step.  Here the
layout is one process per GPU (torchrun) with the batch sharded by rank and ONE collective per step: the
gradients of the step are packed into a few flat buckets and all-reduced (sum) over NCCL -- 24,073,584 elements
for a generator step, 119,233 for a discriminator step -- then divided by the world size, which reproduces the
gradient of the mean loss over the global batch.  Text2Mel and the discriminator use LayerNorm only, so there
are no batch statistics to synchronise.

    for step ...:
        loss = generator_loss(m1(mel, ids, spk), ...)      # spoofsv_b200 kernels, autograd graph
        loss.backward()
        allreduce_gradients(m1.parameters())                # NCCL over NVLink, bucketed
        optimizer.step()
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def plan_buckets(numels: Sequence[int], bucket_elems: int) -> List[List[int]]:
    """Greedy, order-preserving packing of tensors (by index) into buckets of at most bucket_elems elements;
    a tensor larger than the limit gets a bucket of its own."""
    if bucket_elems < 1:
        raise ValueError("bucket_elems must be positive")
    buckets, cur, used = [], [], 0
    for i, n in enumerate(numels):
        if cur and used + n > bucket_elems:
            buckets.append(cur)
            cur, used = [], 0
        cur.append(i)
        used += n
    if cur:
        buckets.append(cur)
    return buckets


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_mb: float = 32.0, average: bool = True,
                        group=None) -> int:
    """Sum the `.grad` of every parameter over the ranks of the process group (in place), bucketed into flat
    buffers of about bucket_mb megabytes; divides by the world size when `average`.  Parameters without a
    gradient on this rank contribute zeros (every rank must pass the same parameter list).  Returns the number of
    elements reduced.  A single-process run (no process group) is a no-op."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    plist = [p for p in params if p.requires_grad]
    if world == 1 or not plist:
        return 0
    for p in plist:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    total = 0
    by_kind = {}
    for p in plist:
        by_kind.setdefault((p.grad.device, p.grad.dtype), []).append(p)
    for (device, dtype), ps in by_kind.items():
        elem = torch.empty((), dtype=dtype).element_size()
        limit = max(1, int(bucket_mb * (1 << 20) / elem))
        handles = []
        for idxs in plan_buckets([p.grad.numel() for p in ps], limit):
            grads = [ps[i].grad for i in idxs]
            flat = torch.cat([g.reshape(-1) for g in grads])
            handles.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True), flat, grads))
            total += flat.numel()
        for work, flat, grads in handles:           # buckets are in flight together; unpack in order
            work.wait()
            if average:
                flat.div_(world)
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
    return total


def shard_batch(n: int, world: int, rank: int) -> slice:
    """Contiguous slice of a global batch of n items for this rank (sizes differ by at most one)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return slice(lo, lo + base + (1 if rank < extra else 0))
