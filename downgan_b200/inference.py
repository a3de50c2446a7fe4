"""Checkpoint loading and chunked full-domain inference: the callers on the far side of ``Generator.forward``.

Reference (paths under /root/reference/DoWnGAN):
  * mlflow_tools/mlflow_epoch.py:65-69  ``log_network_models``: every epoch ``mlflow.pytorch.log_state_dict(G.state_dict(), ...)``
    — i.e. ``torch.save(state_dict, <artifact dir>/state_dict.pth)``;
  * helpers/gen_fake_ds.py:148-158      ``gen_chunks``: ``G.load_state_dict(state_dict)``, then
    ``for chunk in torch.chunk(coarse, chunks=100): ds[i1:i2] = G(chunk.to(device).float()).detach().cpu()``.

Here the same loop runs as a three-stage pipeline on two pinned staging buffers per direction: the host->device copy of
chunk k+1 and the device->host copy of chunk k-1 overlap the forward of chunk k (the reference's loop is synchronous:
``.cpu()`` blocks every iteration).  Chunk boundaries are ``torch.chunk``'s, so the rows written are the reference's.
"""
from __future__ import annotations

import os
from typing import Mapping, Optional, Union

import torch

from . import _lib


def load_reference_checkpoint(module: torch.nn.Module, source: Union[str, os.PathLike, Mapping[str, torch.Tensor]]):
    """Loads a reference ``state_dict`` (a mapping, a ``state_dict.pth`` file, or the mlflow artifact directory holding one)
    into a downgan_b200 ``Generator`` / ``Critic`` with ``strict=True``: keys, shapes and dtypes are the reference's
    (tests/test_abi.py), so checkpoints move between the two code bases unchanged."""
    if not isinstance(source, Mapping):
        path = os.fspath(source)
        if os.path.isdir(path):
            path = os.path.join(path, "state_dict.pth")
        source = torch.load(path, map_location="cpu", weights_only=True)
    module.load_state_dict(source, strict=True)
    if hasattr(module, "mark_params_changed"):
        module.mark_params_changed()
    return module


def generate_chunks(G, coarse: torch.Tensor, chunks: int = 100, out: Optional[torch.Tensor] = None,
                    device: Optional[torch.device] = None) -> torch.Tensor:
    """``gen_chunks`` of helpers/gen_fake_ds.py:148-158: the generator over every row of `coarse` (N, C, h, w; any float
    dtype, on the host), `torch.chunk(coarse, chunks)` rows at a time; returns / fills `out` (N, n_predictands, 8h, 8w) fp32
    on the host."""
    if coarse.dim() != 4:
        raise RuntimeError(f"expected (N, C, h, w), got {tuple(coarse.shape)}")
    dev = device or next(G.parameters()).device
    if dev.type != "cuda":
        raise _lib.DgError("generate_chunks runs on a CUDA device only (no CPU fallback)")
    n, _c, h, w = coarse.shape
    scale = 1 << G.num_upsample
    shape = (n, G.n_predictands, h * scale, w * scale)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32)
    elif tuple(out.shape) != shape or out.dtype != torch.float32:
        raise RuntimeError(f"out must be float32 {shape}, got {out.dtype} {tuple(out.shape)}")
    parts = torch.chunk(coarse, chunks=chunks) if n > 0 else ()
    if not parts:
        return out
    rows = parts[0].shape[0]
    with torch.cuda.device(dev), torch.no_grad():
        main = torch.cuda.current_stream()
        h2d, d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        pin_in = [torch.empty((rows,) + tuple(coarse.shape[1:]), dtype=torch.float32, pin_memory=True) for _ in range(2)]
        dev_in = [torch.empty((rows,) + tuple(coarse.shape[1:]), dtype=torch.float32, device=dev) for _ in range(2)]
        pin_out = [torch.empty((rows,) + shape[1:], dtype=torch.float32, pin_memory=True) for _ in range(2)]
        in_ready = [None, None]    # H2D of the chunk in slot s has finished
        in_free = [None, None]     # the forward that read dev_in[s] has finished
        out_ready = [None, None]   # D2H into pin_out[s] has finished
        pending = [None, None]     # (i1, i2) of the rows sitting in pin_out[s]

        def stage(k):
            s = k & 1
            p = parts[k]
            b = p.shape[0]
            pin_in[s][:b].copy_(p)  # dtype conversion (`.float()`) + pinning on the host
            with torch.cuda.stream(h2d):
                if in_free[s] is not None:
                    h2d.wait_event(in_free[s])
                dev_in[s][:b].copy_(pin_in[s][:b], non_blocking=True)
                in_ready[s] = torch.cuda.Event()
                in_ready[s].record(h2d)

        def drain(s):
            if pending[s] is not None:
                out_ready[s].synchronize()
                i1, i2 = pending[s]
                out[i1:i2].copy_(pin_out[s][:i2 - i1])
                pending[s] = None

        stage(0)
        i1 = 0
        for k, p in enumerate(parts):
            s = k & 1
            b = p.shape[0]
            if k + 1 < len(parts):
                # pin_in[(k+1)&1] was last read by the H2D of chunk k-1, which finished before chunk k-1's forward ran
                if in_ready[(k + 1) & 1] is not None:
                    in_ready[(k + 1) & 1].synchronize()
                stage(k + 1)
            main.wait_event(in_ready[s])
            y = G(dev_in[s][:b])
            in_free[s] = torch.cuda.Event()
            in_free[s].record(main)
            done = torch.cuda.Event()
            done.record(main)
            drain(s)  # pin_out[s] still holds chunk k-2
            with torch.cuda.stream(d2h):
                d2h.wait_event(done)
                pin_out[s][:b].copy_(y, non_blocking=True)
                y.record_stream(d2h)
                out_ready[s] = torch.cuda.Event()
                out_ready[s].record(d2h)
            pending[s] = (i1, i1 + b)
            i1 += b
        drain(0)
        drain(1)
    return out
