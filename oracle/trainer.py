"""CPU restatement of the reference WGAN-GP trainer.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference being restated (paths under /root/reference/DoWnGAN):
  * GAN/wasserstein.py:27-55    _critic_train_iteration
  * GAN/wasserstein.py:58-83    _generator_train_iteration
  * GAN/wasserstein.py:87-117   _gp  (alpha made injectable)
  * GAN/wasserstein.py:131-147  n_critic schedule
  * GAN/losses.py:40-55         content_loss == nn.L1Loss (mean |a-b|)
  * config/hyperparams.py:16-22 gp_lambda 10, critic_iterations 5, gamma 0.01, content_lambda 5, lr 2.5e-4
  * GAN/stage.py:63-64          Adam(lr, betas=(0.9, 0.99)), eps 1e-8, no weight decay

``gp_param_grads_closed_form`` is the *specification* of the hand-written
double-backward; ``tests/test_oracle.py`` checks it
against autograd ``create_graph=True`` in fp64.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, Mapping, Optional, Tuple

import torch
import torch.nn.functional as F

from .networks import (
    C_SLOPE,
    CriticSpec,
    GeneratorSpec,
    as_leaf_params,
    critic_forward,
    generator_forward,
)


@dataclass
class Hyper:
    """config/hyperparams.py:16-22."""

    gp_lambda: float = 10.0
    critic_iterations: int = 5
    gamma: float = 0.01
    content_lambda: float = 5.0
    lr: float = 0.00025
    betas: Tuple[float, float] = (0.9, 0.99)
    eps: float = 1e-8


# --------------------------------------------------------------------------
# gradient penalty, as written (autograd double backward)
# --------------------------------------------------------------------------
def gradient_penalty(c_params, cspec: CriticSpec, real, fake, alpha, hp: Hyper, tape=None):
    """wasserstein.py:87-117.  Returns ``gp_lambda * mean((||g||-1)^2)`` (the
    caller multiplies by gp_lambda a second time, wasserstein.py:40) together
    with the per-sample norms.  ``alpha``: (B,1,1,1)."""
    b = real.shape[0]
    a = alpha.to(real.dtype).expand_as(real)
    interp = (a * real.detach() + (1 - a) * fake.detach()).requires_grad_(True)
    score = critic_forward(c_params, cspec, interp, tape)
    (g,) = torch.autograd.grad(score, interp, grad_outputs=torch.ones_like(score),
                               create_graph=True, retain_graph=True)
    norms = torch.sqrt(torch.sum(g.reshape(b, -1) ** 2, dim=1) + 1e-12)
    return hp.gp_lambda * ((norms - 1) ** 2).mean(), norms, g


def critic_loss_and_grads(g_sd, gspec: GeneratorSpec, c_sd, cspec: CriticSpec,
                          coarse, fine, alpha, hp: Hyper, dtype=None, tapes=None, fake=None):
    """One critic objective + parameter grads (wasserstein.py:35-52).
    ``tapes``: optional {"gen", "real", "fake", "interp"} -> networks.MaskTape (parity diagnostics);
    ``fake``: use this G(coarse) instead of running the generator (mask-pinned comparisons feed the
    CUDA path's own fake so that the replayed masks belong to the critic input they were taken on)."""
    tapes = tapes or {}
    gp_ = {k: (v.to(dtype) if dtype is not None else v) for k, v in g_sd.items()}
    cp = as_leaf_params(c_sd, dtype)
    if dtype is not None:
        coarse, fine, alpha = coarse.to(dtype), fine.to(dtype), alpha.to(dtype)
    if fake is None:
        with torch.no_grad():  # the generator backward of wasserstein.py:52 is discarded (:65)
            fake = generator_forward(gp_, gspec, coarse, tapes.get("gen"))
    elif dtype is not None:
        fake = fake.to(dtype)
    c_real = critic_forward(cp, cspec, fine, tapes.get("real"))
    c_fake = critic_forward(cp, cspec, fake, tapes.get("fake"))
    gp, norms, g = gradient_penalty(cp, cspec, fine, fake, alpha, hp, tapes.get("interp"))
    penalty = hp.gp_lambda * gp
    loss = c_fake.mean() - c_real.mean() + penalty
    grads = torch.autograd.grad(loss, list(cp.values()), allow_unused=True)
    gd = OrderedDict()
    for (k, p), gr in zip(cp.items(), grads):
        gd[k] = torch.zeros_like(p) if gr is None else gr.detach()
    return {
        "fake": fake.detach(), "c_real": c_real.detach(), "c_fake": c_fake.detach(),
        "c_real_mean": c_real.mean().detach(), "c_fake_mean": c_fake.mean().detach(),
        "gp": gp.detach(), "penalty": penalty.detach(), "loss": loss.detach(),
        "norms": norms.detach(), "gp_input_grad": g.detach(), "grads": gd,
    }


def generator_loss_and_grads(g_sd, gspec: GeneratorSpec, c_sd, cspec: CriticSpec,
                             coarse, fine, hp: Hyper, dtype=None, tapes=None):
    """One generator objective + parameter grads (wasserstein.py:65-80).
    ``tapes``: optional {"gen", "fake"} -> networks.MaskTape (parity diagnostics); ``tapes["l1_sign"]``: a tensor of
    the signs sign(fake - fine) to use for the L1 term's derivative (|.| is the generator objective's other kink: where
    |fake - fine| is below the CUDA path's rounding error of `fake` its sign — the whole gradient seed of that element —
    differs; pinning it makes the L1 term the linear function ((fake - fine) * sign).mean())."""
    tapes = tapes or {}
    gp_ = as_leaf_params(g_sd, dtype)
    cp = {k: (v.to(dtype) if dtype is not None else v) for k, v in c_sd.items()}
    if dtype is not None:
        coarse, fine = coarse.to(dtype), fine.to(dtype)
    fake = generator_forward(gp_, gspec, coarse, tapes.get("gen"))
    c_fake = critic_forward(cp, cspec, fake, tapes.get("fake"))
    adv = -c_fake.mean() * hp.gamma
    if tapes.get("l1_sign") is not None:
        l1 = ((fake - fine) * tapes["l1_sign"].to(fake.dtype)).mean()
    else:
        l1 = (fake - fine).abs().mean()  # nn.L1Loss default 'mean' over all elements
    loss = adv + hp.content_lambda * l1
    grads = torch.autograd.grad(loss, list(gp_.values()))
    gd = OrderedDict((k, g.detach()) for k, g in zip(gp_.keys(), grads))
    return {"fake": fake.detach(), "c_fake": c_fake.detach(), "c_fake_mean": c_fake.mean().detach(),
            "l1": l1.detach(), "adv": adv.detach(), "loss": loss.detach(), "grads": gd}


# --------------------------------------------------------------------------
# closed-form gradient-penalty parameter gradients (spec for the CUDA path)
# --------------------------------------------------------------------------
def gp_param_grads_closed_form(c_sd: Mapping[str, torch.Tensor], cspec: CriticSpec,
                               real, fake, alpha, hp: Hyper):
    """d(gp_lambda^2 * mean((||g||-1)^2))/dW for every critic tensor without
    autograd double-backward.  LeakyReLU is piecewise linear, so with masks
    m_l in {1, 0.2} held fixed:

        dz_l : the ones-seeded input-gradient chain (dz_fc2 -> ... -> g)
        u    = lam_eff * (2/B) * (n-1)/n * g                     (dGP/dg)
        v_0  = u ;  v_l = m_l * conv_l(v_{l-1})    (bias-free JVP forward)
        dW_l = wgrad(x := v_{l-1}, dy := dz_l)
        dW_fc1 = dz_fc1^T flat(v_8);  v_fc = m_fc * (flat(v_8) W_fc1^T);  dW_fc2 = sum_i v_fc,i
        all bias gradients are exactly zero.
    Returns (penalty_value_with_both_lambdas, norms, grads dict)."""
    b = real.shape[0]
    dt = real.dtype
    a = alpha.to(dt).expand_as(real)
    x = a * real + (1 - a) * fake
    # forward, recording masks
    masks = []
    acts = x
    for i, (_ci, _co, s) in enumerate(cspec.widths):
        z = F.conv2d(acts, c_sd[f"features.{2 * i}.weight"], c_sd.get(f"features.{2 * i}.bias"), stride=s, padding=1)
        m = torch.where(z > 0, torch.ones_like(z), torch.full_like(z, C_SLOPE))
        masks.append(m)
        acts = z * m
    feat_shape = acts.shape
    zf = F.linear(acts.flatten(1), c_sd["classifier.0.weight"], c_sd["classifier.0.bias"])
    mf = torch.where(zf > 0, torch.ones_like(zf), torch.full_like(zf, C_SLOPE))
    # ones-seeded backward chain
    dz_fc = mf * c_sd["classifier.2.weight"].expand(b, -1)           # (B,100)
    da = (dz_fc @ c_sd["classifier.0.weight"]).reshape(feat_shape)   # (B,C,H,W)
    dzs = [None] * 8
    for i in reversed(range(8)):
        ci, co, s = cspec.widths[i]
        dz = da * masks[i]
        dzs[i] = dz
        hin = dz.shape[-1] * s
        da = torch.nn.grad.conv2d_input((b, ci, hin, hin), c_sd[f"features.{2 * i}.weight"], dz, stride=s, padding=1)
    g = da
    norms = torch.sqrt((g.reshape(b, -1) ** 2).sum(1) + 1e-12)
    lam_eff = hp.gp_lambda * hp.gp_lambda
    value = lam_eff * ((norms - 1) ** 2).mean()
    u = (lam_eff * (2.0 / b) * (norms - 1) / norms).reshape(b, 1, 1, 1) * g
    grads = OrderedDict((k, torch.zeros_like(v)) for k, v in c_sd.items())
    v = u
    for i, (ci, co, s) in enumerate(cspec.widths):
        w = c_sd[f"features.{2 * i}.weight"]
        grads[f"features.{2 * i}.weight"] = torch.nn.grad.conv2d_weight(v, w.shape, dzs[i], stride=s, padding=1)
        v = masks[i] * F.conv2d(v, w, None, stride=s, padding=1)
    vf = v.flatten(1)
    grads["classifier.0.weight"] = dz_fc.t() @ vf
    v_fc = mf * (vf @ c_sd["classifier.0.weight"].t())
    grads["classifier.2.weight"] = v_fc.sum(0, keepdim=True)
    return value, norms, grads


# --------------------------------------------------------------------------
# Adam, restated (torch.optim.Adam defaults: eps 1e-8, no amsgrad, no decay)
# --------------------------------------------------------------------------
@dataclass
class AdamState:
    step: int = 0
    m: Dict[str, torch.Tensor] = field(default_factory=dict)
    v: Dict[str, torch.Tensor] = field(default_factory=dict)


def adam_update(params: Dict[str, torch.Tensor], grads: Mapping[str, torch.Tensor],
                st: AdamState, hp: Hyper) -> None:
    st.step += 1
    b1, b2 = hp.betas
    bc1 = 1 - b1 ** st.step
    bc2 = 1 - b2 ** st.step
    for k, p in params.items():
        g = grads[k].to(p.dtype)
        if k not in st.m:
            st.m[k] = torch.zeros_like(p)
            st.v[k] = torch.zeros_like(p)
        st.m[k].mul_(b1).add_(g, alpha=1 - b1)
        st.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (st.v[k].sqrt() / (bc2 ** 0.5)).add_(hp.eps)
        p.addcdiv_(st.m[k], denom, value=-hp.lr / bc1)


class OracleTrainer:
    """State + schedule of the reference trainer (wasserstein.py:16-55,131-147)
    with injectable alpha.  Metrics logging and mlflow are out of scope."""

    def __init__(self, g_sd, gspec: GeneratorSpec, c_sd, cspec: CriticSpec, hp: Optional[Hyper] = None, dtype=None,
                 emulate_bf16: bool = False):
        # emulate_bf16: every conv / linear input and weight rounded to bf16 (networks.MaskTape(bf16=True)), fp32 master weights
        # and fp32 Adam - what bf16 STORAGE alone does to the reference's trajectory (yardstick of the bf16 loss-curve test)
        self.emulate_bf16 = emulate_bf16
        self.hp = hp or Hyper()
        self.gspec, self.cspec = gspec, cspec
        self.dtype = dtype
        cast = (lambda t: t.detach().clone().to(dtype)) if dtype is not None else (lambda t: t.detach().clone())
        self.g = OrderedDict((k, cast(v)) for k, v in g_sd.items())
        self.c = OrderedDict((k, cast(v)) for k, v in c_sd.items())
        self.g_adam, self.c_adam = AdamState(), AdamState()
        self.num_steps = 0

    def _tapes(self, names):
        from .networks import MaskTape
        return {k: MaskTape(bf16=True) for k in names} if self.emulate_bf16 else None

    def critic_iteration(self, coarse, fine, alpha):
        out = critic_loss_and_grads(self.g, self.gspec, self.c, self.cspec, coarse, fine, alpha, self.hp, self.dtype,
                                    tapes=self._tapes(("gen", "real", "fake", "interp")))
        adam_update(self.c, out["grads"], self.c_adam, self.hp)
        return out

    def generator_iteration(self, coarse, fine):
        out = generator_loss_and_grads(self.g, self.gspec, self.c, self.cspec, coarse, fine, self.hp, self.dtype,
                                       tapes=self._tapes(("gen", "fake")))
        adam_update(self.g, out["grads"], self.g_adam, self.hp)
        return out

    def batch(self, coarse, fine, alpha):
        """One DataLoader batch of _train_epoch (wasserstein.py:131-147)."""
        c_out = self.critic_iteration(coarse, fine, alpha)
        g_out = None
        if self.num_steps % self.hp.critic_iterations == 0:
            g_out = self.generator_iteration(coarse, fine)
        self.num_steps += 1
        return c_out, g_out


# --------------------------------------------------------------------------
# per-batch metric pass (mlflow_tools/mlflow_epoch.py:53-63, table config/hyperparams.py:38-43)
# --------------------------------------------------------------------------
def batch_metrics(g_sd, gspec: GeneratorSpec, c_sd, cspec: CriticSpec, coarse, fine, dtype=None):
    """``gen_batch_and_log_metrics``: fake = G(coarse).detach(); creal = mean C(real); cfake = mean C(fake);
    MAE = content_loss = nn.L1Loss (losses.py:40-55), MSE = content_MSELoss = nn.MSELoss (:58-68),
    Wass = wass_loss(creal, cfake) = creal - cfake (:8-9).  MSSSIM (:12-38) needs pytorch_msssim (absent) and is left out."""
    cast = (lambda t: t.to(dtype)) if dtype is not None else (lambda t: t)
    g = {k: cast(v) for k, v in g_sd.items()}
    c = {k: cast(v) for k, v in c_sd.items()}
    coarse, fine = cast(coarse), cast(fine)
    with torch.no_grad():
        fake = generator_forward(g, gspec, coarse)
        creal = critic_forward(c, cspec, fine).mean()
        cfake = critic_forward(c, cspec, fake).mean()
        return {"MAE": (fine - fake).abs().mean(), "MSE": ((fine - fake) ** 2).mean(), "Wass": creal - cfake,
                "c_real_mean": creal, "c_fake_mean": cfake}


# --------------------------------------------------------------------------
# frequency separation (GAN/wasserstein_fs.py:28-91; filters config/hyperparams.py:31-35)
# --------------------------------------------------------------------------
def low_pass(x: torch.Tensor, filter_size: int = 5) -> torch.Tensor:
    """``hp.low(hp.rf(x))``: ReplicationPad2d(filter_size // 2) then AvgPool2d(filter_size, stride=1, padding=0)."""
    p = filter_size // 2
    return F.avg_pool2d(F.pad(x, (p, p, p, p), mode="replicate"), filter_size, stride=1, padding=0)


def critic_loss_and_grads_fs(g_sd, gspec: GeneratorSpec, c_sd, cspec: CriticSpec, coarse, fine, alpha, hp: Hyper,
                             filter_size: int = 5, dtype=None):
    """wasserstein_fs.py:36-60: the critic and the penalty see the high-pass parts x - low(x)."""
    gp_ = {k: (v.to(dtype) if dtype is not None else v) for k, v in g_sd.items()}
    cp = as_leaf_params(c_sd, dtype)
    if dtype is not None:
        coarse, fine, alpha = coarse.to(dtype), fine.to(dtype), alpha.to(dtype)
    with torch.no_grad():
        fake = generator_forward(gp_, gspec, coarse)
        fake_high = fake - low_pass(fake, filter_size)
        real_high = fine - low_pass(fine, filter_size)
    c_real = critic_forward(cp, cspec, real_high)
    c_fake = critic_forward(cp, cspec, fake_high)
    gp, norms, _g = gradient_penalty(cp, cspec, real_high, fake_high, alpha, hp)
    loss = c_fake.mean() - c_real.mean() + hp.gp_lambda * gp
    grads = torch.autograd.grad(loss, list(cp.values()), allow_unused=True)
    gd = OrderedDict((k, torch.zeros_like(p) if gr is None else gr.detach()) for (k, p), gr in zip(cp.items(), grads))
    return {"loss": loss.detach(), "gp": gp.detach(), "c_real_mean": c_real.mean().detach(),
            "c_fake_mean": c_fake.mean().detach(), "norms": norms.detach(), "grads": gd}


def generator_loss_and_grads_fs(g_sd, gspec: GeneratorSpec, c_sd, cspec: CriticSpec, coarse, fine, hp: Hyper,
                                filter_size: int = 5, dtype=None):
    """wasserstein_fs.py:71-91: adversarial term on the high-pass fake, L1 between the low-pass parts."""
    gp_ = as_leaf_params(g_sd, dtype)
    cp = {k: (v.to(dtype) if dtype is not None else v) for k, v in c_sd.items()}
    if dtype is not None:
        coarse, fine = coarse.to(dtype), fine.to(dtype)
    fake = generator_forward(gp_, gspec, coarse)
    fake_low = low_pass(fake, filter_size)
    real_low = low_pass(fine, filter_size)
    c_fake = critic_forward(cp, cspec, fake - fake_low)
    l1 = (fake_low - real_low).abs().mean()
    loss = -c_fake.mean() * hp.gamma + hp.content_lambda * l1
    grads = torch.autograd.grad(loss, list(gp_.values()))
    return {"loss": loss.detach(), "l1": l1.detach(), "c_fake_mean": c_fake.mean().detach(),
            "grads": OrderedDict((k, g.detach()) for k, g in zip(gp_.keys(), grads))}
