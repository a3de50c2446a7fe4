"""Functional CPU restatement of the reference generator and critic.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Both networks are expressed as pure functions of a ``state_dict``-shaped
mapping (same key names, OIHW fp32 shapes as the reference modules) so the
same code runs in fp32 or fp64 and against weights taken from the reference
modules, the product modules, or ``init_*_state`` below.

Reference being restated (all paths under /root/reference):
  * DoWnGAN/networks/generator.py:14-41   dense residual block
  * DoWnGAN/networks/generator.py:44-53   residual-in-residual block
  * DoWnGAN/networks/generator.py:56-90   generator ctor + forward
  * DoWnGAN/networks/critic.py:12-106     critic ctor + forward
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Mapping, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

# LeakyReLU slopes: generator uses the nn.LeakyReLU() default
# (generator.py:26,72,79); critic uses 0.2 (critic.py:24..87,97).
G_SLOPE = 0.01
C_SLOPE = 0.2
RES_SCALE = 0.2  # generator.py:19,45


@dataclass(frozen=True)
class GeneratorSpec:
    """Constructor arguments of the reference Generator (generator.py:58)."""

    filters: int
    channels: int
    n_predictands: int = 2
    num_res_blocks: int = 16
    num_upsample: int = 3
    fine_dims: int = 0  # accepted and ignored by the reference (generator.py:58)


@dataclass(frozen=True)
class CriticSpec:
    """Constructor arguments of the reference Critic (critic.py:12)."""

    coarse_dim: int
    fine_dim: int
    nc: int

    @property
    def widths(self) -> List[Tuple[int, int, int]]:
        """(in, out, stride) of the 8 feature convs (critic.py:20-92)."""
        w = self.coarse_dim
        return [
            (self.nc, w, 1), (w, w, 2),
            (w, 2 * w, 1), (2 * w, 2 * w, 2),
            (2 * w, 4 * w, 1), (4 * w, 4 * w, 2),
            (4 * w, 8 * w, 1), (8 * w, 8 * w, 2),
        ]

    @property
    def fc_in(self) -> int:
        # critic.py:95
        return int((self.coarse_dim * 2 ** 3) * (self.fine_dim / 2 ** 4) ** 2)


# --------------------------------------------------------------------------
# key enumeration (state_dict order of the reference modules)
# --------------------------------------------------------------------------
def generator_keys(spec: GeneratorSpec) -> List[Tuple[str, Tuple[int, ...]]]:
    """(key, shape) in the reference's ``state_dict()`` order."""
    f = spec.filters
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(name, ci, co):
        out.append((f"{name}.weight", (co, ci, 3, 3)))
        out.append((f"{name}.bias", (co,)))

    conv("conv1", spec.channels, f)
    for r in range(spec.num_res_blocks):
        for d in range(3):
            for k in range(1, 6):
                conv(f"res_blocks.{r}.dense_blocks.{d}.b{k}.0", k * f, f)
    conv("conv2", f, f)
    for u in range(spec.num_upsample):
        conv(f"upsampling.{3 * u}", f, 4 * f)
    conv("conv3.0", f, f)
    conv("conv3.2", f, spec.n_predictands)
    return out


def critic_keys(spec: CriticSpec) -> List[Tuple[str, Tuple[int, ...]]]:
    out: List[Tuple[str, Tuple[int, ...]]] = []
    for i, (ci, co, _s) in enumerate(spec.widths):
        out.append((f"features.{2 * i}.weight", (co, ci, 3, 3)))
        if i == 0:  # only the first conv carries a bias (critic.py:21-23)
            out.append((f"features.{2 * i}.bias", (co,)))
    out.append(("classifier.0.weight", (100, spec.fc_in)))
    out.append(("classifier.0.bias", (100,)))
    out.append(("classifier.2.weight", (1, 100)))
    out.append(("classifier.2.bias", (1,)))
    return out


# --------------------------------------------------------------------------
# initialisation: consume the torch RNG exactly as the reference ctor does
# (one nn.Conv2d / nn.Linear per layer, created in the reference's order)
# --------------------------------------------------------------------------
def init_generator_state(spec: GeneratorSpec, seed: int | None = None,
                         dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    if seed is not None:
        torch.manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    keys = generator_keys(spec)
    for i in range(0, len(keys), 2):
        (wk, wshape), (bk, _b) = keys[i], keys[i + 1]
        layer = nn.Conv2d(wshape[1], wshape[0], 3, 1, 1)
        sd[wk] = layer.weight.detach().to(dtype).clone()
        sd[bk] = layer.bias.detach().to(dtype).clone()
    return sd


def init_critic_state(spec: CriticSpec, seed: int | None = None,
                      dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    if seed is not None:
        torch.manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for i, (ci, co, s) in enumerate(spec.widths):
        layer = nn.Conv2d(ci, co, 3, s, 1, bias=(i == 0))
        sd[f"features.{2 * i}.weight"] = layer.weight.detach().to(dtype).clone()
        if i == 0:
            sd[f"features.{2 * i}.bias"] = layer.bias.detach().to(dtype).clone()
    for name, (fi, fo) in (("classifier.0", (spec.fc_in, 100)), ("classifier.2", (100, 1))):
        layer = nn.Linear(fi, fo)
        sd[f"{name}.weight"] = layer.weight.detach().to(dtype).clone()
        sd[f"{name}.bias"] = layer.bias.detach().to(dtype).clone()
    return sd


# --------------------------------------------------------------------------
# LeakyReLU mask tape (parity diagnostics, tests/test_gpu_masks.py)
# --------------------------------------------------------------------------
class MaskTape:
    """Records, or replays, the LeakyReLU sign masks of a forward pass in call order.

    ``MaskTape()`` records the pre-activations ``z`` of every LeakyReLU the pass executes
    (``tape.z``; the mask is ``z > 0``).  ``MaskTape(masks)`` replays: the k-th LeakyReLU becomes
    ``z * where(masks[k], 1, slope)`` — the same piecewise-linear function with the branch chosen
    by the caller instead of by sign(z).  With the masks of the CUDA path replayed, the oracle
    differentiates exactly the linear map the CUDA path differentiated, so the remaining
    disagreement is arithmetic rounding only; without, bf16 storage flips the branch wherever
    |z| is below its own rounding error (the reference's nn.LeakyReLU: generator.py:26,72,79,
    critic.py:24..97 — the derivative at the flipped positions jumps between 1 and the slope).
    ``bf16=True`` additionally rounds every conv / linear input and weight to bf16 (straight-through
    gradient): the emulated storage precision of the CUDA path's bf16 mode."""

    def __init__(self, masks=None, bf16: bool = False):
        self.replay = masks is not None
        self.masks = list(masks) if masks is not None else []
        self.z: List[torch.Tensor] = []
        self.i = 0
        self.bf16 = bf16

    def q(self, t: torch.Tensor) -> torch.Tensor:
        if not self.bf16:
            return t
        return t + (t.detach().to(torch.bfloat16).to(t.dtype) - t.detach())


def _lrelu(z: torch.Tensor, slope: float, tape: "MaskTape | None") -> torch.Tensor:
    if tape is None:
        return F.leaky_relu(z, slope)
    if tape.replay:
        m = tape.masks[tape.i]
        tape.i += 1
        if m.shape != z.shape:
            raise ValueError(f"mask {tape.i - 1}: shape {tuple(m.shape)} != activation {tuple(z.shape)}")
        return z * torch.where(m, torch.ones((), dtype=z.dtype), torch.full((), slope, dtype=z.dtype))
    tape.z.append(z.detach())
    return F.leaky_relu(z, slope)


# --------------------------------------------------------------------------
# forward passes
# --------------------------------------------------------------------------
def _conv(sd: Mapping[str, torch.Tensor], name: str, x: torch.Tensor, stride: int = 1, tape=None) -> torch.Tensor:
    w = sd[f"{name}.weight"]
    if tape is not None and tape.bf16:
        x, w = tape.q(x), tape.q(w)
    return F.conv2d(x, w, sd.get(f"{name}.bias"), stride=stride, padding=1)


def dense_block_forward(sd, prefix: str, x: torch.Tensor, tape=None) -> torch.Tensor:
    """generator.py:36-41 — concat order [x, o1, o2, o3, o4]; b5 has no activation."""
    feats = x
    o = x
    for k in range(1, 6):
        o = _conv(sd, f"{prefix}.b{k}.0", feats, tape=tape)
        if k < 5:
            o = _lrelu(o, G_SLOPE, tape)
            feats = torch.cat((feats, o), dim=1)
    return o * RES_SCALE + x


def rrdb_forward(sd, prefix: str, x: torch.Tensor, tape=None) -> torch.Tensor:
    """generator.py:52-53."""
    y = x
    for d in range(3):
        y = dense_block_forward(sd, f"{prefix}.dense_blocks.{d}", y, tape)
    return y * RES_SCALE + x


def trunk_forward(sd, spec: GeneratorSpec, x: torch.Tensor, tape=None) -> torch.Tensor:
    """``self.res_blocks(out1)`` of generator.py:85: the RRDB Sequential alone."""
    y = x
    for r in range(spec.num_res_blocks):
        y = rrdb_forward(sd, f"res_blocks.{r}", y, tape)
    return y


def generator_forward(sd: Mapping[str, torch.Tensor], spec: GeneratorSpec, x: torch.Tensor, tape=None) -> torch.Tensor:
    """generator.py:83-90.  x: (B, channels, H, W) -> (B, n_predictands, H*2^u, W*2^u)."""
    first = _conv(sd, "conv1", x, tape=tape)
    y = trunk_forward(sd, spec, first, tape)
    y = first + _conv(sd, "conv2", y, tape=tape)
    for u in range(spec.num_upsample):
        # conv -> LeakyReLU -> PixelShuffle(2)  (generator.py:70-74)
        y = F.pixel_shuffle(_lrelu(_conv(sd, f"upsampling.{3 * u}", y, tape=tape), G_SLOPE, tape), 2)
    y = _lrelu(_conv(sd, "conv3.0", y, tape=tape), G_SLOPE, tape)
    return _conv(sd, "conv3.2", y, tape=tape)


def critic_forward(sd: Mapping[str, torch.Tensor], spec: CriticSpec, x: torch.Tensor, tape=None) -> torch.Tensor:
    """critic.py:101-106.  x: (B, nc, fine, fine) -> (B, 1)."""
    y = x
    for i, (_ci, _co, s) in enumerate(spec.widths):
        y = _lrelu(_conv(sd, f"features.{2 * i}", y, stride=s, tape=tape), C_SLOPE, tape)
    y = torch.flatten(y, 1)  # NCHW order: c*H*W + h*W + w
    w1 = sd["classifier.0.weight"]
    if tape is not None and tape.bf16:
        y = tape.q(y)  # the CUDA path stores the features in bf16; the classifier itself runs in fp32
    y = _lrelu(F.linear(y, w1, sd["classifier.0.bias"]), C_SLOPE, tape)
    return F.linear(y, sd["classifier.2.weight"], sd["classifier.2.bias"])


def as_leaf_params(sd: Mapping[str, torch.Tensor], dtype=None) -> Dict[str, torch.Tensor]:
    """Detached copies with requires_grad=True (optionally cast)."""
    out = OrderedDict()
    for k, v in sd.items():
        t = v.detach().clone()
        if dtype is not None:
            t = t.to(dtype)
        out[k] = t.requires_grad_(True)
    return out
