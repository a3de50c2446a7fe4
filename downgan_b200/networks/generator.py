"""B200 host for the ESRGAN-style RRDB generator.

Drop-in for ``DoWnGAN/networks/generator.py:56-90`` of the reference: same
constructor signature, same ``state_dict`` keys / shapes / order
(``conv1``, ``res_blocks.{r}.dense_blocks.{d}.b{k}.0``, ``conv2``,
``upsampling.{0,3,6}``, ``conv3.{0,2}``), same default initialisation stream,
``forward(x: NCHW fp32) -> NCHW fp32``.  The arithmetic runs in
``libdowngan_b200.so`` (``dg_generator_fwd`` / ``dg_generator_bwd``); there is
no PyTorch or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib
from .._flat import FlatParamsMixin


def _default_precision() -> str:
    return os.environ.get("DOWNGAN_PRECISION", "bf16")


def _conv(ci: int, co: int) -> nn.Conv2d:
    # parameter holder only (never called); default init == reference init
    return nn.Conv2d(ci, co, kernel_size=3, stride=1, padding=1)


class _DenseHolder(nn.Module):
    """Parameters of one dense residual block (generator.py:14-41): b1..b5, b_k maps k*F -> F."""

    def __init__(self, filters: int):
        super().__init__()
        for k in range(1, 6):
            layers = [_conv(k * filters, filters)]
            if k < 5:
                layers.append(nn.LeakyReLU())
            setattr(self, f"b{k}", nn.Sequential(*layers))


class _RRDBHolder(nn.Module):
    """Parameters of one residual-in-residual dense block (generator.py:44-53)."""

    def __init__(self, filters: int):
        super().__init__()
        self.dense_blocks = nn.Sequential(*[_DenseHolder(filters) for _ in range(3)])


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module: "Generator", x: torch.Tensor, *params):
        ctx.module = module
        ctx.needs_dx = x.requires_grad
        out = module._run_forward(x)
        ctx.save_for_backward(x)
        ctx.fwd_id = module._fwd_id
        return out

    @staticmethod
    def backward(ctx, d_out: torch.Tensor):
        m: Generator = ctx.module
        if m._fwd_id != ctx.fwd_id:  # another forward overwrote the saved activations: recompute
            m._run_forward(ctx.saved_tensors[0])
            ctx.fwd_id = m._fwd_id
        grads, dx = m._run_backward(d_out, ctx.needs_dx)
        outs = [None, dx]
        for p, o in zip(m._param_list(), m._offsets):
            outs.append(grads[o:o + p.numel()].view(p.shape))
        return tuple(outs)


class Generator(FlatParamsMixin, nn.Module):
    # coarse_dim_n, fine_dim_n, n_covariates, n_predictands  (GAN/stage.py:60)
    def __init__(self, filters, fine_dims, channels, n_predictands=2, num_res_blocks=16, num_upsample=3,
                 *, precision: Optional[str] = None):
        super().__init__()
        self.filters = int(filters)
        self.fine_dims = fine_dims  # unused by the reference as well (generator.py:58)
        self.channels = int(channels)
        self.n_predictands = int(n_predictands)
        self.num_res_blocks = int(num_res_blocks)
        self.num_upsample = int(num_upsample)
        self.precision = precision or _default_precision()
        if self.precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        f = self.filters
        self.conv1 = _conv(self.channels, f)
        self.res_blocks = nn.Sequential(*[_RRDBHolder(f) for _ in range(self.num_res_blocks)])
        self.conv2 = _conv(f, f)
        ups = []
        for _ in range(self.num_upsample):
            ups += [_conv(f, 4 * f), nn.LeakyReLU(), nn.PixelShuffle(upscale_factor=2)]
        self.upsampling = nn.Sequential(*ups)
        self.conv3 = nn.Sequential(_conv(f, f), nn.LeakyReLU(), _conv(f, self.n_predictands))
        self._handle: Optional[int] = None
        self._handle_key = None
        self._fwd_id = 0

    # ---- native handle -----------------------------------------------------
    def _config(self, coarse_dim: int, max_batch: int) -> _lib.GeneratorConfig:
        return _lib.GeneratorConfig(self.filters, self.channels, self.n_predictands, self.num_res_blocks,
                                    self.num_upsample, coarse_dim, max_batch,
                                    _lib.DG_BF16 if self.precision == "bf16" else _lib.DG_FP32)

    def native(self, coarse_dim: int, batch: int) -> int:
        """dg_generator* for this shape (re-created when the grid or the batch grows)."""
        lib = _lib.load()
        dev = self.conv1.weight.device
        key = (coarse_dim, self.precision, dev.index)
        if self._handle is not None and self._handle_key is not None:
            k, mb = self._handle_key
            if k == key and batch <= mb:
                return self._handle
            self._free()
        cfg = self._config(coarse_dim, batch)
        flat = self.flat_params()
        n = lib.dg_generator_param_count(C.byref(cfg))
        if n != flat.numel():
            raise _lib.DgError(f"parameter layout mismatch: library expects {n} floats, module has {flat.numel()}")
        for i, o in enumerate(self._offsets):
            if lib.dg_generator_param_offset(C.byref(cfg), i) != o:
                raise _lib.DgError(f"parameter offset mismatch at tensor {i}")
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.dg_generator_create(C.byref(cfg), C.byref(h)))
        self._handle, self._handle_key = h.value, (key, batch)
        self._dirty = True
        return self._handle

    def _free(self):
        if self._handle is not None:
            try:
                _lib.load().dg_generator_destroy(self._handle)
            except Exception:
                pass
            self._handle, self._handle_key = None, None

    def __del__(self):
        self._free()

    def __getstate__(self):  # handles are process-local
        d = self.__dict__.copy()
        d["_handle"], d["_handle_key"], d["_flat"], d["_flat_grad"], d["_offsets"] = None, None, None, None, None
        d["_dirty"] = True
        return d

    def ensure_packed(self, handle: int) -> None:
        flat = self.flat_params()
        if self._needs_pack():
            _lib.check(_lib.load().dg_generator_pack(handle, flat.data_ptr(), _lib.stream_ptr()))
            self._dirty = False

    # ---- forward / backward ------------------------------------------------
    def _run_forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4 or x.shape[1] != self.channels or x.shape[2] != x.shape[3]:
            raise RuntimeError(f"Generator expects (B,{self.channels},H,H), got {tuple(x.shape)}")
        x = x.detach().to(device=self.conv1.weight.device, dtype=torch.float32).contiguous()
        b, _, h, _ = x.shape
        hd = self.native(h, b)
        with torch.cuda.device(x.device):
            self.ensure_packed(hd)
            up = h << self.num_upsample
            out = torch.empty(b, self.n_predictands, up, up, device=x.device, dtype=torch.float32)
            _lib.check(_lib.load().dg_generator_fwd(hd, x.data_ptr(), b, out.data_ptr(), 1, _lib.stream_ptr()))
        self._fwd_id += 1
        return out

    def _run_backward(self, d_out: torch.Tensor, needs_dx: bool):
        d_out = d_out.detach().to(torch.float32).contiguous()
        b = d_out.shape[0]
        h = d_out.shape[-1] >> self.num_upsample
        grads = torch.empty_like(self.flat_params())
        dx = torch.empty(b, self.channels, h, h, device=d_out.device, dtype=torch.float32) if needs_dx else None
        with torch.cuda.device(d_out.device):
            _lib.check(_lib.load().dg_generator_bwd(self._handle, d_out.data_ptr(), grads.data_ptr(),
                                                    dx.data_ptr() if dx is not None else None, _lib.stream_ptr()))
        return grads, dx

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.flat_params()
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return _GeneratorFn.apply(self, x, *self._param_list())
        return self._run_forward(x)
