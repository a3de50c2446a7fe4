"""B200 WGAN-GP trainer: drop-in for ``DoWnGAN/GAN/wasserstein.py:16-189``.

Same class name, constructor and method names as the reference trainer;
``_critic_train_iteration`` / ``_generator_train_iteration`` return ``None``
and mutate parameters and optimizer state.  Underneath, each iteration is ONE
call into ``libdowngan_b200.so`` (``dg_critic_step`` / ``dg_generator_step``)
followed by one optimizer launch: ``dg_adam_step`` on one GPU; under data
parallelism (one process per GPU) ``dg_dp_allreduce_adam``, this repo's kernel
that sums the symmetric-memory gradient bucket over NVLink and applies Adam
(``dp.FusedBucket``; NCCL all-reduce + ``dg_adam_step`` when it is unavailable
or switched off with ``DG_DP_FUSED=0``).

Differences from the reference that do not change results:
  * the generator backward of the critic iteration (wasserstein.py:52) is
    skipped — the reference discards those gradients at :65;
  * the gradient penalty uses the closed-form double backward instead of
    ``autograd.grad(create_graph=True)`` (wasserstein.py:100-106);
  * ``alpha`` can be injected (parity tests); the default draws it on the
    device with ``torch.rand`` exactly like wasserstein.py:91;
  * the flattened gradient uses the actual batch size, not ``hp.batch_size``
    (the reference mis-shapes ragged batches, wasserstein.py:110);
  * mlflow logging / plotting (wasserstein.py:140-179) is out of scope; loss
    scalars stay on the device in ``last_critic`` / ``last_generator``.  The
    per-batch metric pass itself (``gen_batch_and_log_metrics``,
    mlflow_epoch.py:53-63: MAE / MSE / Wass with the updated weights) and the
    test-set pass (wasserstein.py:159-172) are here: ``track_metrics`` /
    ``testdataloader`` of ``_train_epoch``, epoch means in ``last_epoch_metrics``.
    MS-SSIM needs the third-party pytorch_msssim package and is not provided.
"""
from __future__ import annotations

import os
import time
from typing import Optional

import torch

from .. import _lib, dp
from ..config import hyperparams as hp
from ..networks.critic import Critic
from ..networks.generator import Generator

CRITIC_SCALARS = ("critic_loss", "c_real_mean", "c_fake_mean", "gp", "penalty")
GENERATOR_SCALARS = ("g_loss", "c_fake_mean", "l1")
METRIC_SCALARS = ("MAE", "MSE", "Wass", "c_real_mean", "c_fake_mean")  # dg_metrics out8; hp.metrics_to_calculate = the first three


class _FlatAdam:
    """torch.optim.Adam semantics (stage.py:63-64) as one fused launch over the flat buffer.
    Hyper-parameters are read from the caller's optimizer every step."""

    def __init__(self, module, optimizer):
        self.module = module
        self.optimizer = optimizer
        self.exp_avg: Optional[torch.Tensor] = None
        self.exp_avg_sq: Optional[torch.Tensor] = None
        self.step_count = 0

    def _group(self):
        if self.optimizer is None:
            return {"lr": hp.lr, "betas": (0.9, 0.99), "eps": 1e-8}
        if len(self.optimizer.param_groups) != 1:
            raise _lib.DgError("the fused Adam step supports a single param group")
        g = self.optimizer.param_groups[0]
        if g.get("weight_decay", 0) or g.get("amsgrad", False) or g.get("maximize", False):
            raise _lib.DgError("fused Adam: weight_decay / amsgrad / maximize are not supported")
        return g

    def step(self, grads: torch.Tensor, grad_scale: float = 1.0) -> None:
        flat = self.module.flat_params()
        if self.exp_avg is None or self.exp_avg.data_ptr() == 0 or self.exp_avg.numel() != flat.numel():
            self.exp_avg = torch.zeros_like(flat)
            self.exp_avg_sq = torch.zeros_like(flat)
        g = self._group()
        self.step_count += 1
        b1, b2 = g["betas"]
        _lib.check(_lib.load().dg_adam_step(flat.data_ptr(), grads.data_ptr(), self.exp_avg.data_ptr(),
                                            self.exp_avg_sq.data_ptr(), flat.numel(), float(g["lr"]), float(b1),
                                            float(b2), float(g["eps"]), self.step_count, float(grad_scale),
                                            _lib.stream_ptr()))
        self.module.mark_params_changed()

    def step_fused(self, bucket, grad_scale: float) -> None:
        """Data parallel: sum-all-reduce of the symmetric gradient bucket + this Adam step in one kernel (dp.FusedBucket)."""
        flat = self.module.flat_params()
        if self.exp_avg is None or self.exp_avg.data_ptr() == 0 or self.exp_avg.numel() != flat.numel():
            self.exp_avg = torch.zeros_like(flat)
            self.exp_avg_sq = torch.zeros_like(flat)
        g = self._group()
        self.step_count += 1
        b1, b2 = g["betas"]
        bucket.step(flat, self.exp_avg, self.exp_avg_sq, g["lr"], b1, b2, g["eps"], self.step_count, grad_scale)
        self.module.mark_params_changed()

    def sync_to_optimizer(self) -> None:
        """Mirror the fused state into ``optimizer.state`` so ``optimizer.state_dict()`` is meaningful."""
        if self.optimizer is None or self.exp_avg is None:
            return
        offs = self.module.param_offsets()
        for p, o in zip(self.module._param_list(), offs):
            self.optimizer.state[p] = {
                "step": torch.tensor(float(self.step_count)),
                "exp_avg": self.exp_avg[o:o + p.numel()].view(p.shape),
                "exp_avg_sq": self.exp_avg_sq[o:o + p.numel()].view(p.shape),
            }


class WassersteinGAN:
    """Implements Wasserstein GAN with gradient penalty (B200-native iteration)."""

    def __init__(self, G: Generator, C: Critic, G_optimizer, C_optimizer) -> None:
        if not isinstance(G, Generator) or not isinstance(C, Critic):
            raise TypeError("WassersteinGAN needs downgan_b200.networks.Generator / Critic modules")
        self.G = G
        self.C = C
        self.G_optimizer = G_optimizer
        self.C_optimizer = C_optimizer
        self.num_steps = 0
        self._g_adam = _FlatAdam(G, G_optimizer)
        self._c_adam = _FlatAdam(C, C_optimizer)
        self.last_critic: Optional[torch.Tensor] = None     # 8 floats on device, see CRITIC_SCALARS
        self.last_generator: Optional[torch.Tensor] = None  # 8 floats on device, see GENERATOR_SCALARS
        self._c_scal = None
        self._g_scal = None
        self.lookahead = os.environ.get("DG_NO_LOOKAHEAD", "0") != "1"  # _train_epoch computes the fakes of the critic steps between two generator updates in one pass
        # data parallel: classifier-gradient all-reduce started while the conv weight gradients still run (DG_OVERLAP_AR=1).
        # Off by default: measured 0.4 % slower than one all-reduce per iteration on 2 GPUs (two collectives + one more call
        # cost more than hiding 3.3 MB over NVLink saves), results identical (tools/dp_overlap_check.py).
        self.overlap_allreduce = os.environ.get("DG_OVERLAP_AR", "0") == "1"
        self.freq_sep = False        # True: the iterations of GAN/wasserstein_fs.py (see WassersteinGANFS)
        self.track_metrics = False   # True: _train_epoch runs the reference's per-batch metric pass on the training batches
        self.last_epoch_metrics = None  # {"train": {"MAE", "MSE", "Wass"}, "test": {...}}: means over batches (mlflow_epoch.py:38-49)
        self._m_scal = None

    # ---- helpers -------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return self.G.conv1.weight.device

    def _hyper(self) -> _lib.Hyper:
        return _lib.Hyper(float(hp.gp_lambda), float(hp.gamma), float(hp.content_lambda),
                          1 if getattr(self, "freq_sep", False) else 0, int(getattr(hp, "filter_size", 5)))

    def _prep(self, t: torch.Tensor) -> torch.Tensor:
        return t.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()

    def _handles(self, coarse: torch.Tensor, lazy_critic: bool = False):
        b, _, h, _ = coarse.shape
        g = self.G.native(h, b)
        c = self.C.native(b)
        self.G.ensure_packed(g)
        self.C.ensure_packed(c, lazy=lazy_critic)
        return g, c

    def _allreduce(self, grads: torch.Tensor) -> float:
        """Sum-all-reduce the flat gradient bucket over NCCL; returns the 1/world scale for Adam."""
        return dp.allreduce_sum_(grads)

    # ---- iterations ------------------------------------------------------------
    def _critic_train_iteration(self, coarse, fine, alpha: Optional[torch.Tensor] = None, _fake_offset: Optional[int] = None):
        """One critic update (wasserstein.py:27-55).  `_fake_offset` (set by `_train_epoch`'s look-ahead) takes
        fake = G(coarse) from the generator's look-ahead buffer instead of recomputing it."""
        coarse, fine = self._prep(coarse), self._prep(fine)
        b = coarse.shape[0]
        with torch.cuda.device(self.device):
            if alpha is None:
                alpha = torch.rand(b, 1, 1, 1, device=self.device)  # wasserstein.py:91
            alpha = self._prep(alpha).reshape(b)
            g, c = self._handles(coarse, lazy_critic=True)  # the critic's re-pack runs beside the batch assembly
            if self._c_scal is None or self._c_scal.device != self.device:
                self._c_scal = torch.zeros(8, device=self.device)
            # data parallel: the classifier gradients (74 % of the bucket) are final before the conv weight gradients,
            # which still run on the handle's side stream - their all-reduce starts while those finish
            overlap = dp.world_size() > 1 and self.overlap_allreduce
            # default exchange: the flat gradients are written into a symmetric-memory bucket and ONE kernel does the
            # sum over NVLink peer memory + Adam (dp.FusedBucket); None = one rank, switched off, or unavailable -> NCCL
            bucket = None if overlap else dp.fused_bucket(self.C)
            grads = self.C.flat_grads()
            hyp = self._hyper()
            lib = _lib.load()
            _lib.check(lib.dg_critic_defer_conv_grads(c, 1 if overlap else 0))
            if _fake_offset is None:
                _lib.check(lib.dg_critic_step(g, c, hyp, coarse.data_ptr(), fine.data_ptr(), alpha.data_ptr(), b,
                                              grads.data_ptr(), self._c_scal.data_ptr(), _lib.stream_ptr()))
            else:
                _lib.check(lib.dg_critic_step_fake(g, c, hyp, int(_fake_offset), fine.data_ptr(), alpha.data_ptr(),
                                                   b, grads.data_ptr(), self._c_scal.data_ptr(), _lib.stream_ptr()))
            if overlap:
                off = self.C.param_offsets()[9]  # classifier.0.weight: everything before it is conv weights + features.0.bias
                work = dp.allreduce_sum_async_(grads[off:])
                _lib.check(lib.dg_critic_step_finish(c, grads.data_ptr(), _lib.stream_ptr()))
                scale = self._allreduce(grads[:off])
                work.wait()
            elif bucket is None:
                scale = self._allreduce(grads)
            if bucket is not None:
                self._c_adam.step_fused(bucket, 1.0 / dp.world_size())
            else:
                self._c_adam.step(grads, scale)
        self.last_critic = self._c_scal

    def _generator_lookahead(self, coarse_all: torch.Tensor, save_first: int = 0, first: Optional[tuple] = None) -> None:
        """fake = G(coarse) for several upcoming critic batches in one pass (same generator weights: the
        generator is only updated every `critic_iterations` steps, wasserstein.py:136-137).  The first
        `save_first` samples keep their activations for the generator iteration on that batch."""
        coarse_all = self._prep(coarse_all)
        total, _, h, _ = coarse_all.shape
        with torch.cuda.device(self.device):
            g = self.G.native(h, total)
            self.G.ensure_packed(g)
            if first is not None and save_first <= first[0]:
                # only the next critic iteration's batch in stream order, the rest overlaps that iteration
                _lib.check(_lib.load().dg_generator_lookahead_first(g, coarse_all.data_ptr(), total, int(save_first),
                                                                    int(first[0]), int(first[1]), _lib.stream_ptr()))
            else:
                _lib.check(_lib.load().dg_generator_lookahead(g, coarse_all.data_ptr(), total, int(save_first),
                                                              _lib.stream_ptr()))

    def _generator_train_iteration(self, coarse, fine, _saved_forward: bool = False):
        """One generator update (wasserstein.py:58-83).  `_saved_forward` (set by `_train_epoch`'s look-ahead):
        G(coarse) and its activations are still resident from the look-ahead pass and are not recomputed."""
        coarse, fine = self._prep(coarse), self._prep(fine)
        b = coarse.shape[0]
        with torch.cuda.device(self.device):
            g, c = self._handles(coarse)
            if self._g_scal is None or self._g_scal.device != self.device:
                self._g_scal = torch.zeros(8, device=self.device)
            bucket = dp.fused_bucket(self.G)
            grads = self.G.flat_grads()
            hyp = self._hyper()
            if _saved_forward:
                _lib.check(_lib.load().dg_generator_step_saved(g, c, hyp, fine.data_ptr(), b, grads.data_ptr(),
                                                               self._g_scal.data_ptr(), _lib.stream_ptr()))
            else:
                _lib.check(_lib.load().dg_generator_step(g, c, hyp, coarse.data_ptr(), fine.data_ptr(), b,
                                                         grads.data_ptr(), self._g_scal.data_ptr(), _lib.stream_ptr()))
            if bucket is not None:
                self._g_adam.step_fused(bucket, 1.0 / dp.world_size())
            else:
                scale = self._allreduce(grads)
                self._g_adam.step(grads, scale)
        self.last_generator = self._g_scal

    def _metrics_batch(self, coarse, fine, _fake_offset: Optional[int] = None) -> torch.Tensor:
        """`gen_batch_and_log_metrics` (mlflow_epoch.py:53-63) for one batch: 8 floats on the device in METRIC_SCALARS
        order, computed with the CURRENT weights of both networks.  `_fake_offset` (set by `_train_epoch`): G(coarse) is
        still resident from the look-ahead pass and the generator has not been updated since, so it is not recomputed."""
        coarse, fine = self._prep(coarse), self._prep(fine)
        b = coarse.shape[0]
        with torch.cuda.device(self.device):
            g, c = self._handles(coarse)
            out = torch.empty(8, device=self.device)
            _lib.check(_lib.load().dg_metrics(g, c, None if _fake_offset is not None else coarse.data_ptr(),
                                              int(_fake_offset or 0), fine.data_ptr(), b, out.data_ptr(), _lib.stream_ptr()))
        return out

    def _gp(self, real, fake, critic, alpha: Optional[torch.Tensor] = None, want_grads: bool = False):
        """Gradient penalty value ``hp.gp_lambda * mean((||grad||-1)^2)`` (wasserstein.py:87-117).
        With ``want_grads`` also returns d(hp.gp_lambda * value)/d(critic params) as a flat tensor."""
        real, fake = self._prep(real), self._prep(fake.detach())
        b = real.shape[0]
        with torch.cuda.device(self.device):
            if alpha is None:
                alpha = torch.rand(b, 1, 1, 1, device=self.device)
            alpha = self._prep(alpha).reshape(b)
            c = critic.native(b)
            critic.ensure_packed(c)
            out = torch.zeros(1, device=self.device)
            norms = torch.empty(b, device=self.device)
            grads = torch.empty_like(critic.flat_params()) if want_grads else None
            hyp = self._hyper()
            _lib.check(_lib.load().dg_gp(c, hyp, real.data_ptr(), fake.data_ptr(), alpha.data_ptr(), b, out.data_ptr(),
                                         norms.data_ptr(), grads.data_ptr() if grads is not None else None,
                                         _lib.stream_ptr()))
        self.last_gp_norms = norms
        return (out[0], grads) if want_grads else out[0]

    def prepare(self, coarse_shape, fine_shape, with_alpha: bool = True) -> None:
        """Allocate everything `_train_epoch` needs for host batches of these shapes (staging slots, the look-ahead
        buffer, the pinned scalar ring, native handles sized for the look-ahead pass) so that the first epoch neither
        allocates nor synchronises inside its loop.  Optional: `_train_epoch` allocates lazily without it."""
        dev = self.device
        n_critic = int(hp.critic_iterations)
        b = int(coarse_shape[0])
        with torch.cuda.device(dev):
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
            shapes = [tuple(coarse_shape), tuple(fine_shape)] + ([(b, 1, 1, 1)] if with_alpha else [])
            self._slots = [{"bufs": [torch.empty(sh, device=dev, dtype=torch.float32) for sh in shapes], "done": None}
                           for _ in range(n_critic + 3)]
            self._slot_i = 0
            self._la_buf = torch.empty(((n_critic + 1) * b,) + tuple(coarse_shape[1:]), device=dev, dtype=torch.float32)
            if getattr(self, "_log_host", None) is None:
                self._log_host = torch.empty(1024, 8, dtype=torch.float32, pin_memory=True)
            la = (n_critic + 1) * b if getattr(self, "lookahead", True) and n_critic > 1 else b
            g = self.G.native(int(coarse_shape[2]), la)
            c = self.C.native(b)
            self.G.ensure_packed(g)
            self.C.ensure_packed(c)
            if self._c_scal is None or self._c_scal.device != dev:
                self._c_scal = torch.zeros(8, device=dev)
            if self._g_scal is None or self._g_scal.device != dev:
                self._g_scal = torch.zeros(8, device=dev)
            torch.cuda.synchronize(dev)

    # ---- epoch loop (wasserstein.py:120-189, logging/plotting out of scope) ----
    def _train_epoch(self, dataloader, testdataloader=None, epoch: int = 0):
        """One pass over `dataloader` with the reference schedule (wasserstein.py:131-147): a critic
        iteration per batch, a generator iteration on the same batch when num_steps % n_critic == 0.
        Batches may live on the host: batch i+1 is copied to the device on a side stream while batch
        i trains, and the loss scalars of every step are copied back asynchronously into pinned
        memory (no per-step device synchronisation).  Returns a (steps, 8) CPU tensor of the critic
        scalars (CRITIC_SCALARS order) — the reference logs the same quantities through mlflow.
        With `self.track_metrics` every training batch is followed by the metric pass (wasserstein.py:140-146); with a
        `testdataloader` the test-set pass runs after the loop (:159-172).  Their means over batches land in
        `self.last_epoch_metrics`."""
        dev = self.device
        with torch.cuda.device(dev):
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
            main = torch.cuda.current_stream()

            # fixed staging slots reused round-robin: no allocation (hence no implicit device synchronisation)
            # in the steady state; a slot is rewritten only after the step that read it
            n_critic = int(hp.critic_iterations)
            if getattr(self, "_slots", None) is None or len(self._slots) != n_critic + 3:
                self._slots = [{"bufs": None, "done": None} for _ in range(n_critic + 3)]
            self._slot_i = getattr(self, "_slot_i", 0)

            def stage(data):
                src = list(data[:3])
                if all(isinstance(t, torch.Tensor) and t.device == dev and t.dtype == torch.float32 and t.is_contiguous()
                       for t in src):
                    return {"bufs": src, "done": None}, None  # already resident: used in place, no staging copy
                slot = self._slots[self._slot_i % len(self._slots)]
                self._slot_i += 1
                shapes = [tuple(t.shape) for t in src]
                with torch.cuda.stream(self._copy_stream):
                    if slot["bufs"] is None or [tuple(b.shape) for b in slot["bufs"]] != shapes:
                        # (re)allocated ON the copy stream: the caching allocator may hand back a block whose last user is
                        # still queued on the main stream (a freed look-ahead or alpha temporary), so the first write into a
                        # fresh block is also ordered after everything the main stream has enqueued so far
                        slot["bufs"] = [torch.empty(sh, device=dev, dtype=torch.float32) for sh in shapes]
                        fresh = torch.cuda.Event()
                        fresh.record(main)
                        self._copy_stream.wait_event(fresh)
                    if slot["done"] is not None:
                        self._copy_stream.wait_event(slot["done"])
                    for b, t in zip(slot["bufs"], src):
                        b.copy_(t, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
                return slot, ev

            it = iter(dataloader)
            pending = []  # staged batches, oldest first
            t_enqueue0 = time.perf_counter()

            def fill(n):
                while len(pending) < n:
                    data = next(it, None)
                    if data is None:
                        return
                    pending.append(stage(data))

            depth = n_critic + 1 if getattr(self, "lookahead", True) else 2  # host->device copies run this far ahead
            fill(depth)
            offsets = []  # look-ahead: sample offsets of the fakes of the next critic steps
            saved_steps = set()  # steps whose generator iteration reuses the look-ahead forward
            logs = []
            track = bool(getattr(self, "track_metrics", False))
            mlogs = []
            if (track or testdataloader is not None) and getattr(self, "_metric_host", None) is None:
                self._metric_host = torch.empty(1024, 8, dtype=torch.float32, pin_memory=True)
            while pending:
                s = self.num_steps
                if getattr(self, "lookahead", True) and not offsets and n_critic > 1 and s % n_critic != 0:
                    # steps s .. next multiple of n_critic all see the current generator weights
                    fill(n_critic - (s % n_critic) + 1)
                    want = n_critic - (s % n_critic) + 1
                    group = pending[:want]
                    if len(group) > 1:
                        for _, ev in group:
                            if ev is not None:
                                main.wait_event(ev)
                        # the group's LAST batch is the one the generator iteration runs on (step % n_critic == 0):
                        # it goes first in the pass and keeps its activations, so that iteration needs no forward
                        full = len(group) == want
                        order = [len(group) - 1] + list(range(len(group) - 1)) if full else list(range(len(group)))
                        save_first = group[-1][0]["bufs"][0].shape[0] if full else 0
                        offs = [0] * len(group)
                        off = 0
                        for i in order:
                            offs[i] = off
                            off += group[i][0]["bufs"][0].shape[0]
                        # the group's coarse fields side by side in a fixed buffer (no allocation in the steady state)
                        c0 = group[0][0]["bufs"][0]
                        la = getattr(self, "_la_buf", None)
                        if la is None or la.shape[0] < off or la.shape[1:] != c0.shape[1:] or la.device != c0.device:
                            la = torch.empty((max(off, (n_critic + 1) * c0.shape[0]),) + tuple(c0.shape[1:]), device=dev,
                                             dtype=torch.float32)
                            self._la_buf = la
                        for i in order:
                            t = group[i][0]["bufs"][0]
                            la[offs[i]:offs[i] + t.shape[0]].copy_(t)
                        coarse_all = la[:off]
                        self._generator_lookahead(coarse_all, save_first,
                                                  first=(offs[0], group[0][0]["bufs"][0].shape[0]))
                        offsets.extend(offs)
                        saved_steps = {s + len(group) - 1} if full else set()
                slot, ev = pending.pop(0)
                fill(depth)
                if ev is not None:
                    main.wait_event(ev)
                ts = slot["bufs"]
                coarse, fine = ts[0], ts[1]
                alpha = ts[2] if len(ts) > 2 else None
                off = offsets.pop(0) if offsets else None
                self._critic_train_iteration(coarse, fine, alpha, _fake_offset=off)
                g_updated = self.num_steps % n_critic == 0
                if g_updated:
                    self._generator_train_iteration(coarse, fine, _saved_forward=self.num_steps in saved_steps)
                    saved_steps.discard(self.num_steps)
                if track:
                    # the look-ahead fake of this batch is still G(coarse) unless the generator was just updated
                    m = self._metrics_batch(coarse, fine, _fake_offset=None if g_updated else off)
                    k = len(mlogs)
                    if k >= self._metric_host.shape[0]:
                        main.synchronize()
                        grown = torch.empty(2 * (k + 1), 8, dtype=torch.float32, pin_memory=True)
                        grown[:k].copy_(self._metric_host[:k])
                        self._metric_host = grown
                    self._metric_host[k].copy_(m, non_blocking=True)
                    mlogs.append(k)
                self.num_steps += 1
                slot["done"] = torch.cuda.Event()
                slot["done"].record(main)
                k = len(logs)
                if getattr(self, "_log_host", None) is None or k >= self._log_host.shape[0]:
                    main.synchronize()  # growing the pinned ring: rare (every 1024 steps)
                    grown = torch.empty(max(1024, 2 * (k + 1)), 8, dtype=torch.float32, pin_memory=True)
                    if getattr(self, "_log_host", None) is not None:
                        grown[:k].copy_(self._log_host[:k])
                    self._log_host = grown
                self._log_host[k].copy_(self.last_critic, non_blocking=True)
                logs.append(k)
            self.last_enqueue_seconds = time.perf_counter() - t_enqueue0  # host time to enqueue the epoch (diagnostic)
            main.synchronize()
            result = {}
            if track:
                result["train"] = self._metric_means(self._metric_host[:len(mlogs)])
                if hasattr(dataloader, "skip_first_batch"):
                    dataloader.skip_first_batch()  # RNG side effect of the reference's plot batch (wasserstein.py:155)
            if testdataloader is not None:
                # test-set pass (wasserstein.py:159-172): metrics only, no update; every batch needs its own generator forward
                k = 0
                for data in testdataloader:
                    m = self._metrics_batch(data[0], data[1])
                    if k >= self._metric_host.shape[0]:
                        main.synchronize()
                        grown = torch.empty(2 * (k + 1), 8, dtype=torch.float32, pin_memory=True)
                        grown[:k].copy_(self._metric_host[:k])
                        self._metric_host = grown
                    self._metric_host[k].copy_(m, non_blocking=True)
                    k += 1
                main.synchronize()
                result["test"] = self._metric_means(self._metric_host[:k])
                if hasattr(testdataloader, "skip_first_batch"):
                    testdataloader.skip_first_batch()  # wasserstein.py:175
            if result:
                self.last_epoch_metrics = result
        return self._log_host[:len(logs)].clone() if logs else torch.zeros(0, 8)

    @staticmethod
    def _metric_means(rows: torch.Tensor) -> dict:
        """`post_epoch_metric_mean` (mlflow_epoch.py:38-49): the mean over batches of every tracked metric."""
        if rows.shape[0] == 0:
            return {k: float("nan") for k in hp.metrics_to_calculate}
        m = rows.double().mean(0)
        return {k: float(m[METRIC_SCALARS.index(k)]) for k in hp.metrics_to_calculate}

    def train(self, dataloader, testdataloader=None):
        self.num_steps = 0
        for epoch in range(hp.epochs):
            self._train_epoch(dataloader, testdataloader, epoch)
        self.sync_optimizer_state()

    def sync_optimizer_state(self) -> None:
        self._g_adam.sync_to_optimizer()
        self._c_adam.sync_to_optimizer()


class WassersteinGANFS(WassersteinGAN):
    """Wasserstein GAN with gradient penalty and frequency separation: drop-in for ``GAN/wasserstein_fs.py:15-91``.
    The critic and the penalty see the high-pass parts ``x - low(x)`` of the real and generated fields, the content
    loss compares their low-pass parts; ``low`` = ``hp.low(hp.rf(.))`` (hyperparams.py:31-35), one stencil kernel here
    (``dg_lowpass``) with its adjoint in the generator's backward.  Upstream this class is not reachable from
    ``train.py`` (``hp.freq_sep = False`` and its imports are broken); it is mirrored from its source statements."""

    def __init__(self, G: Generator, C: Critic, G_optimizer, C_optimizer) -> None:
        super().__init__(G, C, G_optimizer, C_optimizer)
        self.freq_sep = True
