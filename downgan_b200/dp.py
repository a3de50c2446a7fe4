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


# ---------------------------------------------------------------------------------------------------------------------
# Fused exchange + optimizer step: ONE kernel of this repo (csrc/dg_dp.cu, dg_dp_allreduce_adam) sums the flat gradient
# bucket over NVLink peer memory (NVSwitch multimem reduce / broadcast when the bucket has a multicast mapping) and applies
# Adam, instead of NCCL all-reduce + dg_adam_step.  The bucket and a small flag block live in symmetric memory.
# ---------------------------------------------------------------------------------------------------------------------
import os
import sys


class FusedBucket:
    """Symmetric-memory gradient bucket of one network (becomes the module's flat gradient buffer) + its flag block."""

    def __init__(self, n: int, device: torch.device, use_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        group = dist.group.WORLD
        self.n = n
        self.grads = symm.empty(n, dtype=torch.float32, device=device)
        self.flags = symm.empty(32, dtype=torch.int32, device=device)
        self.grads.zero_()
        self.flags.zero_()
        gh = symm.rendezvous(self.grads, group)
        fh = symm.rendezvous(self.flags, group)
        self._handles = (gh, fh)  # keep the mappings alive
        self.rank, self.world = int(gh.rank), int(gh.world_size)
        if self.world > 8:
            raise RuntimeError("dg_dp_allreduce_adam handles at most 8 ranks (one node)")
        goff, foff = int(getattr(gh, "offset", 0) or 0), int(getattr(fh, "offset", 0) or 0)
        gp = [int(p) + goff for p in gh.buffer_ptrs]
        fp = [int(p) + foff for p in fh.buffer_ptrs]
        if gp[self.rank] != self.grads.data_ptr() or fp[self.rank] != self.flags.data_ptr():
            raise RuntimeError("symmetric memory: local mapping does not match the tensor address")
        self.peers = _lib.DpPeers()
        self.peers.rank, self.peers.world = self.rank, self.world
        for r in range(self.world):
            self.peers.grad_ptrs[r] = gp[r]
            self.peers.flag_ptrs[r] = fp[r]
        mc = int(getattr(gh, "multicast_ptr", 0) or 0)
        self.multicast = bool(mc) and use_multicast
        self.peers.grad_multicast = (mc + goff) if self.multicast else None
        self.epoch = 0
        torch.cuda.synchronize(device)
        dist.barrier()  # every rank's zeroed flag block is in place before the first kernel signals into it

    def step(self, params: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, lr: float, b1: float, b2: float,
             eps: float, step: int, grad_scale: float) -> None:
        from . import _lib
        self.epoch += 1
        _lib.check(_lib.load().dg_dp_allreduce_adam(self.peers, params.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                                    self.n, float(lr), float(b1), float(b2), float(eps), int(step),
                                                    float(grad_scale), self.epoch, _lib.stream_ptr()))

    def self_test(self) -> None:
        """One exchange on known values (lr = 0 on scratch parameters): every rank must end up with the exact sum."""
        dev = self.grads.device
        n = self.n
        pattern = (torch.arange(n, device=dev, dtype=torch.float32) % 257.0) - 128.0  # small integers: exact in fp32
        self.grads.copy_(pattern * float(self.rank + 1))
        p = torch.zeros(n, device=dev)
        m = torch.zeros(n, device=dev)
        v = torch.zeros(n, device=dev)
        self.step(p, m, v, 0.0, 0.9, 0.99, 1e-8, 1, 1.0)
        torch.cuda.synchronize(dev)
        want = pattern * float(self.world * (self.world + 1) // 2)
        if not torch.equal(self.grads, want):
            bad = int((self.grads != want).sum())
            raise RuntimeError(f"fused exchange self-test failed on rank {self.rank}: {bad} of {n} elements differ")
        if not torch.allclose(m, want * 0.1, rtol=1e-5, atol=0.0):  # exp_avg after one step = (1 - beta1) * sum
            raise RuntimeError(f"fused exchange self-test: Adam moment differs on rank {self.rank}")
        self.grads.zero_()
        torch.cuda.synchronize(dev)
        dist.barrier()


_fused_warned = False


def fused_bucket(module) -> "FusedBucket | None":
    """The module's symmetric gradient bucket (created, self-tested and installed as its flat gradient buffer on first use), or
    None when the fused exchange is switched off (DG_DP_FUSED=0), the world is one rank, or symmetric memory is not
    available - the callers then use NCCL all-reduce + dg_adam_step."""
    global _fused_warned
    if world_size() <= 1 or os.environ.get("DG_DP_FUSED", "1") != "1":
        return None
    b = getattr(module, "_dp_bucket", None)
    if b is False:
        return None
    flat = module.flat_params()
    if b is None or b.n != flat.numel() or b.grads.device != flat.device:
        ok = 1
        try:
            b = FusedBucket(flat.numel(), flat.device, use_multicast=os.environ.get("DG_DP_MULTICAST", "1") == "1")
            b.self_test()
        except Exception as e:  # no symmetric memory on this system / driver: NCCL path
            ok = 0
            if not _fused_warned:
                sys.stderr.write(f"[downgan_b200.dp] fused exchange unavailable on rank {rank()} ({type(e).__name__}: {e}); using NCCL all-reduce\n")
                _fused_warned = True
        # all ranks must take the same path
        flag = torch.tensor([ok], device=flat.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            module._dp_bucket = False
            return None
        module._dp_bucket = b
    if module.flat_grads().data_ptr() != b.grads.data_ptr():
        module.use_grad_bucket(b.grads)
    return b
