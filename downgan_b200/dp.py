"""Data-parallel host logic (one process per GPU).

Every loss term of the WGAN-GP iteration is a mean over independent samples
(the gradient penalty is per-sample: ones seed, per-sample norm,
GAN/wasserstein.py:100-117 of the reference), so sharding the batch over
ranks and summing the flat gradient buckets, scaled by 1/world, reproduces
the global-batch gradient exactly.  The only exchange step is that sum-
all-reduce (NCCL on GPUs; gloo in the CPU tests); there is no other
collective on the data path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard(t: torch.Tensor, r: int | None = None, w: int | None = None) -> torch.Tensor:
    """Contiguous equal shard of the batch axis for rank r (the global batch must divide evenly)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    if t.shape[0] % w:
        raise ValueError(f"global batch {t.shape[0]} is not divisible by world size {w}")
    n = t.shape[0] // w
    return t[r * n:(r + 1) * n]


def allreduce_sum_(flat: torch.Tensor) -> float:
    """In-place sum-all-reduce of a flat gradient bucket; returns the scale (1/world) the fused
    Adam applies so that equal local batches give the global-batch mean."""
    w = world_size()
    if w > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return 1.0 / w
    return 1.0


def allreduce_sum_async_(flat: torch.Tensor):
    """Asynchronous in-place sum-all-reduce (world > 1): the collective is ordered after the work already enqueued on the
    current stream and runs on the backend's own stream; ``.wait()`` on the returned handle orders the current stream
    after it.  Used to overlap the classifier-gradient bucket with the conv weight gradients still in flight."""
    return dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
