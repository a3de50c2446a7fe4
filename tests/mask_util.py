"""Mask-pinned parity helpers (GPU): run one fused iteration of the CUDA path, read back the LeakyReLU
sign masks it actually used (``dg_generator_activation`` / ``dg_critic_activation``: sign of the stored
post-activation) and hand them to the oracle as ``oracle.networks.MaskTape`` replays.

Why: the reference's networks are piecewise linear (LeakyReLU after every conv, generator.py:26,72,79,
critic.py:24..97).  bf16 storage moves a pre-activation z by ~|z|*2^-9, so where |z| is below its own
rounding error the CUDA path and the fp32 oracle take different branches and the derivative of that
element jumps between 1 and the slope — a flipped fraction p of a layer's elements changes gradients
flowing through it by ~sqrt(p) relative, whatever the kernel's arithmetic quality.  Pinning the masks
separates that effect from kernel errors: with the CUDA path's masks replayed, the oracle differentiates
exactly the linear map the kernels differentiated, and every gradient tensor must meet north_star's 2e-2.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

from downgan_b200 import _lib
from oracle import networks as onet
from oracle import trainer as otr

import parity_util as pu


def _gen_act(g, which, shape):
    out = torch.empty(shape, device="cuda", dtype=torch.float32)
    _lib.check(_lib.load().dg_generator_activation(g, which, shape[0], out.data_ptr(), _lib.stream_ptr()))
    return out


def _crit_act(c, which, s0, shape):
    out = torch.empty(shape, device="cuda", dtype=torch.float32)
    _lib.check(_lib.load().dg_critic_activation(c, which, s0, shape[0], out.data_ptr(), _lib.stream_ptr()))
    return out


def generator_masks(g, gspec: onet.GeneratorSpec, b: int, hc: int):
    """Masks of the generator's LeakyReLUs in the oracle's call order (networks.generator_forward)."""
    f = gspec.filters
    masks = []
    for i in range(gspec.num_res_blocks * 3):
        buf = _gen_act(g, i, (b, 5 * f, hc, hc))
        for k in range(1, 5):
            masks.append((buf[:, k * f:(k + 1) * f] > 0).cpu())
    h = hc
    for u in range(gspec.num_upsample):
        h *= 2
        up = _gen_act(g, 1100 + u, (b, f, h, h))
        masks.append((F.pixel_unshuffle(up, 2) > 0).cpu())  # LeakyReLU precedes PixelShuffle (generator.py:70-74)
    masks.append((_gen_act(g, 1200, (b, f, h, h)) > 0).cpu())
    return masks


def critic_masks(c, cspec: onet.CriticSpec, s0: int, b: int, kept: bool = False):
    """Masks of one critic forward (8 convs + classifier.0) for samples [s0, s0+b) of the critic's batch;
    ``kept``: the interpolates' conv activations the fused iteration put aside (dg_set_tuning(15, 1))."""
    masks = []
    h = cspec.fine_dim
    for i, (_ci, co, s) in enumerate(cspec.widths):
        h //= s
        which = (101 if kept else 1) + i
        masks.append((_crit_act(c, which, s0, (b, co, h, h)) > 0).cpu())
    return masks


def _fc_mask(c, s0, b):
    return (_crit_act(c, 9, s0, (b, 100)) > 0).cpu()


def critic_step_with_masks(G, C, gspec, cspec, coarse, fine, alpha):
    """dg_critic_step on the CUDA path; returns (scalars, grads dict, fake NCHW cpu, mask lists {real,fake,interp})."""
    lib = _lib.load()
    tr = pu.WassersteinGAN(G, C, None, None)
    b = coarse.shape[0]
    cd, fd, ad = coarse.cuda(), fine.cuda(), alpha.reshape(b).cuda().contiguous()
    g, c = tr._handles(cd)
    sc = torch.zeros(8, device="cuda")
    cg = torch.zeros_like(C.flat_params())
    prev = lib.dg_set_tuning(15, 1)
    try:
        _lib.check(lib.dg_critic_step(g, c, tr._hyper(), cd.data_ptr(), fd.data_ptr(), ad.data_ptr(), b, cg.data_ptr(),
                                      sc.data_ptr(), _lib.stream_ptr()))
        torch.cuda.synchronize()
        masks = {
            "real": critic_masks(c, cspec, 0, b) + [_fc_mask(c, 0, b)],
            "fake": critic_masks(c, cspec, b, b) + [_fc_mask(c, b, b)],
            "interp": critic_masks(c, cspec, 0, b, kept=True) + [_fc_mask(c, 2 * b, b)],
        }
    finally:
        lib.dg_set_tuning(15, prev)
    with torch.no_grad():
        fake = G(cd).cpu()  # the forward kernels are deterministic: the same fake the iteration used
    return sc.cpu(), pu.flat_to_dict(C, cg), fake, masks


def generator_step_with_masks(G, C, gspec, cspec, coarse, fine):
    """dg_generator_step on the CUDA path; returns (scalars, grads dict, mask lists {gen, fake})."""
    lib = _lib.load()
    tr = pu.WassersteinGAN(G, C, None, None)
    b, _, hc, _ = coarse.shape
    cd, fd = coarse.cuda(), fine.cuda()
    g, c = tr._handles(cd)
    sg = torch.zeros(8, device="cuda")
    gg = torch.zeros_like(G.flat_params())
    _lib.check(lib.dg_generator_step(g, c, tr._hyper(), cd.data_ptr(), fd.data_ptr(), b, gg.data_ptr(), sg.data_ptr(),
                                     _lib.stream_ptr()))
    torch.cuda.synchronize()
    masks = {"gen": generator_masks(g, gspec, b, hc), "fake": critic_masks(c, cspec, 0, b) + [_fc_mask(c, 0, b)]}
    with torch.no_grad():
        fake = G(cd).cpu()  # deterministic forward kernels: the fake the iteration differentiated
    # the L1 term's kink: sign(fake - fine) as the CUDA path saw it (losses.py:51-53 via wasserstein.py:78)
    masks["l1_sign"] = torch.sign(fake - fine)
    return sg.cpu(), pu.flat_to_dict(G, gg), masks


def flip_stats(masks, tape: onet.MaskTape):
    """Per LeakyReLU: fraction of elements whose branch differs between the CUDA path (masks) and the oracle's free
    run (tape.z), and the largest |z| / rms(z) among them (how close to zero the oracle's pre-activation was)."""
    rows = []
    for m, z in zip(masks, tape.z):
        flipped = m != (z > 0)
        n = int(flipped.sum())
        rms = float(z.double().pow(2).mean().sqrt())
        zmax = float(z[flipped].abs().max()) / rms if n else 0.0
        rows.append((n / flipped.numel(), zmax))
    return rows


def tensor_errors(got: dict, ref: dict):
    """{name: (relative L2 error, |ref|)}; tensors whose reference gradient is exactly zero report the absolute norm."""
    out = OrderedDict()
    for k, r in ref.items():
        g = got[k].double()
        r = r.double()
        n = float(r.norm())
        e = float((g - r).norm())
        out[k] = (e / n if n > 0 else e, n)
    return out


def critic_parity(G, C, gspec, cspec, g_sd, c_sd, coarse, fine, alpha, hp=None, emulate=True):
    """One critic iteration three ways.  Returns a dict with scalars, per-tensor errors of the CUDA gradients against
    (a) the free-running fp32 oracle, (b) the oracle with the CUDA path's masks replayed, plus (c) the emulated-bf16
    oracle's own error against (a), and the flip statistics."""
    hp = hp or otr.Hyper()
    sc, cg, fake, masks = critic_step_with_masks(G, C, gspec, cspec, coarse, fine, alpha)
    rec = {k: onet.MaskTape() for k in ("real", "fake", "interp")}
    free = otr.critic_loss_and_grads(g_sd, gspec, c_sd, cspec, coarse, fine, alpha, hp, tapes=rec, fake=fake)
    own = otr.critic_loss_and_grads(g_sd, gspec, c_sd, cspec, coarse, fine, alpha, hp)  # oracle's own generator forward
    pin = otr.critic_loss_and_grads(g_sd, gspec, c_sd, cspec, coarse, fine, alpha, hp, fake=fake,
                                    tapes={k: onet.MaskTape(v) for k, v in masks.items()})
    out = {"scalars": sc, "free": own, "pinned": pin,
           "err_free": tensor_errors(cg, own["grads"]), "err_pinned": tensor_errors(cg, pin["grads"]),
           "flips": {k: flip_stats(masks[k], rec[k]) for k in masks}, "fake_rel": pu.rel(fake, own["fake"]), "grads": cg}
    if emulate:
        emu = otr.critic_loss_and_grads(g_sd, gspec, c_sd, cspec, coarse, fine, alpha, hp, fake=fake,
                                        tapes={k: onet.MaskTape(bf16=True) for k in ("real", "fake", "interp")})
        out["err_emulated"] = tensor_errors(emu["grads"], free["grads"])
    return out


def generator_parity(G, C, gspec, cspec, g_sd, c_sd, coarse, fine, hp=None, emulate=True):
    hp = hp or otr.Hyper()
    sg, gg, masks = generator_step_with_masks(G, C, gspec, cspec, coarse, fine)
    rec = {k: onet.MaskTape() for k in ("gen", "fake")}
    free = otr.generator_loss_and_grads(g_sd, gspec, c_sd, cspec, coarse, fine, hp, tapes=rec)
    l1_sign = masks.pop("l1_sign")
    pin_tapes = {k: onet.MaskTape(v) for k, v in masks.items()}
    pin_tapes["l1_sign"] = l1_sign
    pin = otr.generator_loss_and_grads(g_sd, gspec, c_sd, cspec, coarse, fine, hp, tapes=pin_tapes)
    out = {"scalars": sg, "free": free, "pinned": pin,
           "err_free": tensor_errors(gg, free["grads"]), "err_pinned": tensor_errors(gg, pin["grads"]),
           "flips": {k: flip_stats(masks[k], rec[k]) for k in masks}, "grads": gg,
           "l1_sign_flips": float((l1_sign != torch.sign(free["fake"] - fine)).float().mean())}
    if emulate:
        emu = otr.generator_loss_and_grads(g_sd, gspec, c_sd, cspec, coarse, fine, hp,
                                           tapes={k: onet.MaskTape(bf16=True) for k in ("gen", "fake")})
        out["err_emulated"] = tensor_errors(emu["grads"], free["grads"])
    return out


def flat_err(errs: dict) -> float:
    """Relative L2 error of the concatenated gradient from per-tensor (rel, |ref|) pairs."""
    num = sum((e * n) ** 2 if n > 0 else e ** 2 for e, n in errs.values())
    den = sum(n ** 2 for _e, n in errs.values())
    return (num ** 0.5) / (den ** 0.5 if den > 0 else 1.0)
