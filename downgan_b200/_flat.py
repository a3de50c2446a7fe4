"""Flat parameter storage shared by the Generator and Critic hosts.

The C ABI takes ONE contiguous fp32 buffer per network, tensors laid out in
the reference's ``state_dict()`` order (include/downgan_b200.h).  The modules
keep ordinary ``nn.Parameter`` objects — so ``state_dict`` / ``load_state_dict``
/ ``parameters()`` / ``.to(device)`` behave exactly like the reference
modules — and re-point their ``.data`` at slices of the flat
buffer the first time they run on a CUDA device.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn


class FlatParamsMixin:
    _flat: Optional[torch.Tensor] = None
    _flat_grad: Optional[torch.Tensor] = None
    _offsets: Optional[List[int]] = None
    _dirty: bool = True
    _seen_version: int = -1

    def _param_list(self) -> List[nn.Parameter]:
        return list(self.parameters())  # registration order == state_dict order

    def flat_params(self) -> torch.Tensor:
        """Ensure every parameter is a view into one flat CUDA fp32 buffer and return it."""
        params = self._param_cache if getattr(self, "_param_cache", None) is not None else self._param_list()
        self._param_cache = params
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("downgan_b200 modules run on CUDA only (no CPU fallback): call .to('cuda') first")
        flat = self._flat
        if flat is not None and self._offsets is not None and flat.device == dev:
            # fast path: `.to()` / `load_state_dict(assign=True)` re-point every parameter, so checking the
            # first and the last one is enough to know the views are still in place
            base = flat.data_ptr()
            p0, p1 = params[0], params[-1]
            if p0.data_ptr() == base and p1.data_ptr() == base + 4 * self._offsets[-1] and p1.dtype == torch.float32:
                return flat
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += p.numel()
        flat = self._flat
        ok = flat is not None and flat.device == dev and flat.numel() == n
        if ok:
            base = flat.data_ptr()
            for p, o in zip(params, offs):
                if p.data_ptr() != base + 4 * o or p.dtype != torch.float32:
                    ok = False
                    break
        if not ok:
            flat = torch.empty(n, device=dev, dtype=torch.float32)
            with torch.no_grad():
                for p, o in zip(params, offs):
                    flat[o:o + p.numel()].copy_(p.detach().reshape(-1).to(torch.float32))
                    p.data = flat[o:o + p.numel()].view(p.shape)
            self._flat = flat
            self._flat_grad = torch.zeros_like(flat)
            self._offsets = offs
            self._dirty = True
        return self._flat

    def flat_grads(self) -> torch.Tensor:
        self.flat_params()
        return self._flat_grad

    def use_grad_bucket(self, bucket: torch.Tensor) -> None:
        """Make `bucket` (same size / device, e.g. a symmetric-memory tensor for the fused data-parallel exchange,
        downgan_b200/dp.py) the flat gradient buffer; `p.grad` views are re-pointed if they were bound."""
        flat = self.flat_params()
        if bucket.numel() != flat.numel() or bucket.device != flat.device or bucket.dtype != torch.float32:
            raise ValueError("gradient bucket does not match the flat parameter buffer")
        bound = self._param_list()[0].grad is not None
        bucket.copy_(self._flat_grad)
        self._flat_grad = bucket
        if bound:
            self.bind_grads()

    def param_offsets(self) -> List[int]:
        self.flat_params()
        return list(self._offsets)

    def bind_grads(self) -> None:
        """Expose the flat gradient buffer through ``p.grad`` (views, no copies)."""
        g = self.flat_grads()
        for p, o in zip(self._param_list(), self._offsets):
            p.grad = g[o:o + p.numel()].view(p.shape)

    def mark_params_changed(self) -> None:
        self._dirty = True

    def _params_version(self) -> int:
        ps = self._param_cache if getattr(self, "_param_cache", None) is not None else self._param_list()
        return sum(p._version for p in ps)

    def _needs_pack(self) -> bool:
        v = self._params_version()
        if self._dirty or v != self._seen_version:
            self._seen_version = v
            return True
        return False
