"""B200 host for the strided-conv critic.

Drop-in for ``DoWnGAN/networks/critic.py:9-106`` of the reference: same
constructor ``Critic(coarse_dim, fine_dim, nc)``, same ``state_dict`` keys
(``features.{0,2,...,14}.weight``, ``features.0.bias``,
``classifier.{0,2}.{weight,bias}``), same initialisation stream,
``forward(x: (B,nc,fine,fine) fp32) -> (B,1)``.  Runs in
``libdowngan_b200.so`` (``dg_critic_fwd`` / ``dg_critic_bwd``).  Double
backward through this module is not an autograd graph: the gradient penalty
uses the closed-form path ``dg_gp`` (see ``GAN/wasserstein.py``).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib
from .._flat import FlatParamsMixin
from .generator import _default_precision


class _CriticFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module: "Critic", x: torch.Tensor, *params):
        ctx.module = module
        ctx.needs_dx = x.requires_grad
        out = module._run_forward(x)
        ctx.save_for_backward(x)
        ctx.fwd_id = module._fwd_id
        return out

    @staticmethod
    def backward(ctx, d_out: torch.Tensor):
        m: Critic = ctx.module
        if m._fwd_id != ctx.fwd_id:  # another forward overwrote the saved activations: recompute
            m._run_forward(ctx.saved_tensors[0])
            ctx.fwd_id = m._fwd_id
        grads, dx = m._run_backward(d_out, ctx.needs_dx)
        outs = [None, dx]
        for p, o in zip(m._param_list(), m._offsets):
            outs.append(grads[o:o + p.numel()].view(p.shape))
        return tuple(outs)


class Critic(FlatParamsMixin, nn.Module):
    def __init__(self, coarse_dim, fine_dim, nc, *, precision: Optional[str] = None):
        super().__init__()
        self.coarse_dim = int(coarse_dim)
        self.fine_dim = int(fine_dim)
        self.nc = int(nc)
        self.precision = precision or _default_precision()
        if self.precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        if self.fine_dim % 16:
            raise ValueError("fine_dim must be a multiple of 16 (four stride-2 stages)")
        w = self.coarse_dim
        plan = [(self.nc, w, 1, True), (w, w, 2, False), (w, 2 * w, 1, False), (2 * w, 2 * w, 2, False),
                (2 * w, 4 * w, 1, False), (4 * w, 4 * w, 2, False), (4 * w, 8 * w, 1, False), (8 * w, 8 * w, 2, False)]
        feats = []
        for ci, co, s, bias in plan:  # parameter holders; bias only on the first conv (critic.py:21-23)
            feats += [nn.Conv2d(ci, co, kernel_size=3, stride=s, padding=1, bias=bias),
                      nn.LeakyReLU(negative_slope=0.2, inplace=True)]
        self.features = nn.Sequential(*feats)
        fc_in = int((w * 2 ** 3) * (self.fine_dim / 2 ** 4) ** 2)
        self.classifier = nn.Sequential(nn.Linear(fc_in, 100), nn.LeakyReLU(negative_slope=0.2, inplace=True),
                                        nn.Linear(100, 1))
        self._handle: Optional[int] = None
        self._handle_key = None
        self._fwd_id = 0

    # ---- native handle -----------------------------------------------------
    def _config(self, max_batch: int) -> _lib.CriticConfig:
        return _lib.CriticConfig(self.coarse_dim, self.fine_dim, self.nc, max_batch,
                                 _lib.DG_BF16 if self.precision == "bf16" else _lib.DG_FP32)

    def native(self, batch: int) -> int:
        """dg_critic* sized for `batch` samples per call (3x internally for the fused critic step)."""
        lib = _lib.load()
        dev = self.classifier[0].weight.device
        key = (self.precision, dev.index)
        if self._handle is not None and self._handle_key is not None:
            k, mb = self._handle_key
            if k == key and batch <= mb:
                return self._handle
            self._free()
        cfg = self._config(batch)
        flat = self.flat_params()
        n = lib.dg_critic_param_count(C.byref(cfg))
        if n != flat.numel():
            raise _lib.DgError(f"parameter layout mismatch: library expects {n} floats, module has {flat.numel()}")
        for i, o in enumerate(self._offsets):
            if lib.dg_critic_param_offset(C.byref(cfg), i) != o:
                raise _lib.DgError(f"parameter offset mismatch at tensor {i}")
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.dg_critic_create(C.byref(cfg), C.byref(h)))
        self._handle, self._handle_key = h.value, (key, batch)
        self._dirty = True
        return self._handle

    def _free(self):
        if self._handle is not None:
            try:
                _lib.load().dg_critic_destroy(self._handle)
            except Exception:
                pass
            self._handle, self._handle_key = None, None

    def __del__(self):
        self._free()

    def __getstate__(self):
        d = self.__dict__.copy()
        d["_handle"], d["_handle_key"], d["_flat"], d["_flat_grad"], d["_offsets"] = None, None, None, None, None
        d["_dirty"] = True
        return d

    def ensure_packed(self, handle: int, lazy: bool = False) -> None:
        """Re-pack the weights if they changed.  `lazy` (the trainer's critic iteration): only record the request; the native
        iteration runs the pack launch beside its batch assembly (dg_critic_pack_lazy)."""
        flat = self.flat_params()
        if self._needs_pack():
            if lazy:
                _lib.check(_lib.load().dg_critic_pack_lazy(handle, flat.data_ptr()))
            else:
                _lib.check(_lib.load().dg_critic_pack(handle, flat.data_ptr(), _lib.stream_ptr()))
            self._dirty = False

    # ---- forward / backward ------------------------------------------------
    def _run_forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4 or x.shape[1] != self.nc or x.shape[2] != self.fine_dim or x.shape[3] != self.fine_dim:
            raise RuntimeError(f"Critic expects (B,{self.nc},{self.fine_dim},{self.fine_dim}), got {tuple(x.shape)}")
        x = x.detach().to(device=self.classifier[0].weight.device, dtype=torch.float32).contiguous()
        b = x.shape[0]
        hd = self.native(b)
        with torch.cuda.device(x.device):
            self.ensure_packed(hd)
            out = torch.empty(b, 1, device=x.device, dtype=torch.float32)
            _lib.check(_lib.load().dg_critic_fwd(hd, x.data_ptr(), b, out.data_ptr(), _lib.stream_ptr()))
        self._fwd_id += 1
        return out

    def _run_backward(self, d_out: torch.Tensor, needs_dx: bool):
        d_out = d_out.detach().to(torch.float32).contiguous()
        b = d_out.shape[0]
        grads = torch.empty_like(self.flat_params())
        dx = torch.empty(b, self.nc, self.fine_dim, self.fine_dim, device=d_out.device, dtype=torch.float32) \
            if needs_dx else None
        with torch.cuda.device(d_out.device):
            _lib.check(_lib.load().dg_critic_bwd(self._handle, d_out.data_ptr(), grads.data_ptr(),
                                                 dx.data_ptr() if dx is not None else None, _lib.stream_ptr()))
        return grads, dx

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        self.flat_params()
        if torch.is_grad_enabled() and (input.requires_grad or any(p.requires_grad for p in self.parameters())):
            return _CriticFn.apply(self, input, *self._param_list())
        return self._run_forward(input)
