"""Shared helpers for the GPU parity tests (CUDA path vs the CPU oracle)."""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import torch

from downgan_b200 import _lib
from downgan_b200.GAN.wasserstein import WassersteinGAN
from downgan_b200.networks import Critic, Generator
from oracle import networks as onet
from oracle import trainer as otr

PREC = {"fp32": _lib.DG_FP32, "bf16": _lib.DG_BF16}


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| (L2, fp64)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    d = float(b.norm())
    return float((a - b).norm()) / d if d > 0 else float(a.norm())


def build_pair(gspec: onet.GeneratorSpec, cspec: onet.CriticSpec, precision: str, seed=0, g_sd=None, c_sd=None,
               critic_scale: float = 1.0):
    torch.manual_seed(seed)
    C_ = Critic(cspec.coarse_dim, cspec.fine_dim, cspec.nc, precision=precision)
    G_ = Generator(gspec.filters, cspec.fine_dim, gspec.channels, gspec.n_predictands, gspec.num_res_blocks,
                   gspec.num_upsample, precision=precision)
    if g_sd is not None:
        G_.load_state_dict(g_sd)
    if c_sd is not None:
        C_.load_state_dict(c_sd)
    if critic_scale != 1.0:
        with torch.no_grad():
            for p in C_.features.parameters():
                if p.dim() == 4:
                    p.mul_(critic_scale)
    g_sd = OrderedDict((k, v.detach().clone()) for k, v in G_.state_dict().items())
    c_sd = OrderedDict((k, v.detach().clone()) for k, v in C_.state_dict().items())
    G_.cuda(), C_.cuda()
    return G_, C_, g_sd, c_sd


def flat_to_dict(module, flat: torch.Tensor):
    out = OrderedDict()
    for (k, p), o in zip(module.named_parameters(), module.param_offsets()):
        out[k] = flat[o:o + p.numel()].view(p.shape).detach().cpu()
    return out


def grad_report(got: dict, ref: dict):
    """per-tensor worst and flat relative error."""
    worst, worst_k = 0.0, None
    num = den = 0.0
    for k, r in ref.items():
        g = got[k].double()
        r = r.double()
        e = float((g - r).norm())
        n = float(r.norm())
        num += e * e
        den += n * n
        rk = e / n if n > 0 else float(g.norm())
        if rk > worst:
            worst, worst_k = rk, k
    return worst, worst_k, (num ** 0.5) / (den ** 0.5 if den > 0 else 1.0)


def conv_fwd(x, w, bias, stride, slope, precision):
    lib = _lib.load()
    b, ci, h, wd = x.shape
    co = w.shape[0]
    ho, wo = (h - 1) // stride + 1, (wd - 1) // stride + 1
    xd, wd_, = x.cuda().contiguous(), w.cuda().contiguous()
    bd = bias.cuda().contiguous() if bias is not None else None
    y = torch.empty(b, co, ho, wo, device="cuda")
    _lib.check(lib.dg_conv3x3_fwd(xd.data_ptr(), wd_.data_ptr(), bd.data_ptr() if bd is not None else None,
                                  y.data_ptr(), b, ci, co, h, x.shape[3], stride, float(slope), PREC[precision],
                                  _lib.stream_ptr()))
    return y.cpu()


def conv_dgrad(dy, w, hin, win, stride, precision):
    lib = _lib.load()
    b, co = dy.shape[:2]
    ci = w.shape[1]
    dyd, wd_ = dy.cuda().contiguous(), w.cuda().contiguous()
    dx = torch.empty(b, ci, hin, win, device="cuda")
    _lib.check(lib.dg_conv3x3_dgrad(dyd.data_ptr(), wd_.data_ptr(), dx.data_ptr(), b, ci, co, hin, win, stride,
                                    PREC[precision], _lib.stream_ptr()))
    return dx.cpu()


def conv_wgrad(x, dy, stride, precision):
    lib = _lib.load()
    b, ci, h, wd = x.shape
    co = dy.shape[1]
    xd, dyd = x.cuda().contiguous(), dy.cuda().contiguous()
    dw = torch.empty(co, ci, 3, 3, device="cuda")
    db = torch.empty(co, device="cuda")
    _lib.check(lib.dg_conv3x3_wgrad(xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), db.data_ptr(), b, ci, co, h, wd,
                                    stride, PREC[precision], _lib.stream_ptr()))
    return dw.cpu(), db.cpu()
