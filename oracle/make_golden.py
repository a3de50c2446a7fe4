"""Pin the oracle against the real reference and write the golden fixtures.

TEST INFRASTRUCTURE ONLY.  Runs in the build container, where
``/root/reference`` exists (it does not exist on the GPU box).  It

  1. imports the reference's own ``DoWnGAN.networks.{generator,critic}``,
  2. asserts that ``oracle.networks`` reproduces their default init
     (same RNG consumption) and forward outputs BIT-EXACTLY in fp32,
  3. executes the reference trainer's statements (wasserstein.py:35-52,
     65-80, 87-117 — autograd ``create_graph`` GP, ``loss.backward()``) on
     the reference modules with an injected alpha and asserts the oracle
     trainer gives bit-identical losses and parameter gradients,
  4. writes ``tests/golden/tiny.npz`` (inputs, weights, outputs, gradients of
     a small configuration) and ``tests/golden/cfg1_summary.json`` (per-tensor
     gradient norms etc. of BASELINE cfg-1) for the tests that run where the
     reference is absent,
  5. (round 2) pins the adjacent rows of SURVEY.md §8f on the same tiny configuration and writes
     ``tests/golden/tiny_extra.json`` + ``tests/golden/ref_tiny_{generator,critic}_state_dict.pth``:
       * the per-batch metrics through the reference's OWN ``content_loss`` / ``content_MSELoss`` / ``wass_loss``
         (GAN/losses.py, imported with a stub for its absent third-party ``pytorch_msssim`` dependency) as
         ``gen_batch_and_log_metrics`` calls them (mlflow_epoch.py:53-63),
       * the frequency-separation iterations by executing the statements of GAN/wasserstein_fs.py:36-91 on the
         reference modules with the filters of config/hyperparams.py:31-35 (that file itself cannot be imported:
         ``import stage as s``), asserted bit-equal to ``oracle.trainer.*_fs``,
       * ``torch.save(module.state_dict())`` of the reference modules = what ``mlflow.pytorch.log_state_dict`` stores
         (mlflow_epoch.py:65-69), for the checkpoint-compatibility tests.

Usage:  python -m oracle.make_golden
"""
from __future__ import annotations

import json
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

from . import networks as onet  # noqa: E402
from . import trainer as otr  # noqa: E402


def _ref_modules():
    sys.path.insert(0, REF)
    from DoWnGAN.networks.generator import Generator  # type: ignore
    from DoWnGAN.networks.critic import Critic  # type: ignore
    return Generator, Critic


from downgan_b200.synthetic import synth_batch  # noqa: E402  (shared synthetic-field generator)


def ref_critic_step_grads(G, C, coarse, fine, alpha, hp: otr.Hyper):
    """The reference's statements (wasserstein.py:35-52, 87-117) on its modules."""
    fake = G(coarse)
    c_real = C(fine)
    c_fake = C(fake)
    a = alpha.expand_as(fine)
    interpolated = (a * fine.data + (1 - a) * fake.data).requires_grad_(True)
    ci = C(interpolated)
    gradients = torch.autograd.grad(outputs=ci, inputs=interpolated, grad_outputs=torch.ones(ci.size()),
                                    create_graph=True, retain_graph=True)[0]
    gradients = gradients.view(fine.size(0), -1)
    gn = torch.sqrt(torch.sum(gradients ** 2, dim=1) + 1e-12)
    gp = hp.gp_lambda * ((gn - 1) ** 2).mean()
    gradient_penalty = hp.gp_lambda * gp
    C.zero_grad()
    loss = torch.mean(c_fake) - torch.mean(c_real) + gradient_penalty
    loss.backward(retain_graph=True)
    grads = OrderedDict((k, (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)))
                        for k, p in C.named_parameters())
    return fake.detach(), c_real.detach(), c_fake.detach(), gp.detach(), loss.detach(), grads


def ref_generator_step_grads(G, C, coarse, fine, hp: otr.Hyper):
    """wasserstein.py:65-80 on the reference modules."""
    G.zero_grad()
    fake = G(coarse)
    c_fake = C(fake)
    g_loss = -torch.mean(c_fake) * hp.gamma
    g_loss = g_loss + hp.content_lambda * torch.nn.L1Loss()(fake, fine)
    g_loss.backward()
    grads = OrderedDict((k, p.grad.detach().clone()) for k, p in G.named_parameters())
    return g_loss.detach(), grads


def check_config(name, gspec: onet.GeneratorSpec, cspec: onet.CriticSpec, b, seed=0, warm_steps=0):
    Generator, Critic = _ref_modules()
    hp = otr.Hyper()
    torch.manual_seed(seed)
    C = Critic(cspec.coarse_dim, cspec.fine_dim, cspec.nc)
    G = Generator(gspec.filters, cspec.fine_dim, gspec.channels, gspec.n_predictands,
                  gspec.num_res_blocks, gspec.num_upsample)
    # same seed, same construction order (critic first, as GAN/stage.py:59-60)
    torch.manual_seed(seed)
    c_sd = onet.init_critic_state(cspec)
    g_sd = onet.init_generator_state(gspec)
    for k, v in C.state_dict().items():
        assert torch.equal(v, c_sd[k]), f"critic init mismatch {k}"
    for k, v in G.state_dict().items():
        assert torch.equal(v, g_sd[k]), f"generator init mismatch {k}"
    assert list(C.state_dict().keys()) == [k for k, _ in onet.critic_keys(cspec)]
    assert list(G.state_dict().keys()) == [k for k, _ in onet.generator_keys(gspec)]

    coarse, fine, alpha = synth_batch(b, gspec.channels, cspec.coarse_dim,
                                      up=2 ** gspec.num_upsample, npred=gspec.n_predictands)

    if warm_steps:
        # move away from the degenerate random init (SURVEY §7 hard-part 3)
        tr = otr.OracleTrainer(g_sd, gspec, c_sd, cspec, hp)
        gC = torch.optim.Adam(C.parameters(), hp.lr, betas=hp.betas)
        gG = torch.optim.Adam(G.parameters(), hp.lr, betas=hp.betas)
        for s in range(warm_steps):
            al = torch.rand(b, 1, 1, 1, generator=torch.Generator().manual_seed(100 + s))
            tr.batch(coarse, fine, al)
            # reference statements with torch.optim.Adam
            *_x, gr = ref_critic_step_grads(G, C, coarse, fine, al, hp)
            gC.step()
            if s % hp.critic_iterations == 0:
                ref_generator_step_grads(G, C, coarse, fine, hp)
                gG.step()
        g_sd, c_sd = tr.g, tr.c
        worst = max(float((C.state_dict()[k] - c_sd[k]).abs().max()) for k in c_sd)
        worstg = max(float((G.state_dict()[k] - g_sd[k]).abs().max()) for k in g_sd)
        print(f"[{name}] after {warm_steps} Adam steps: max |ref-oracle| critic {worst:.3e} generator {worstg:.3e}")
        assert worst < 1e-5 and worstg < 1e-5
        C.load_state_dict(c_sd)
        G.load_state_dict(g_sd)

    with torch.no_grad():
        ref_fake = G(coarse)
        ref_score = C(fine)
    assert torch.equal(ref_fake, onet.generator_forward(g_sd, gspec, coarse)), "generator forward not bit-exact"
    assert torch.equal(ref_score, onet.critic_forward(c_sd, cspec, fine)), "critic forward not bit-exact"

    fake, c_real, c_fake, gp, loss, cgr = ref_critic_step_grads(G, C, coarse, fine, alpha, hp)
    oc = otr.critic_loss_and_grads(g_sd, gspec, c_sd, cspec, coarse, fine, alpha, hp)
    assert torch.equal(oc["loss"], loss), (oc["loss"], loss)
    assert torch.equal(oc["gp"], gp)
    for k in cgr:
        assert torch.equal(cgr[k], oc["grads"][k]), f"critic grad mismatch {k}"
    gl, ggr = ref_generator_step_grads(G, C, coarse, fine, hp)
    og = otr.generator_loss_and_grads(g_sd, gspec, c_sd, cspec, coarse, fine, hp)
    assert torch.equal(og["loss"], gl)
    for k in ggr:
        assert torch.equal(ggr[k], og["grads"][k]), f"generator grad mismatch {k}"

    # closed form vs autograd (fp64)
    c64 = {k: v.double() for k, v in c_sd.items()}
    f64 = onet.generator_forward({k: v.double() for k, v in g_sd.items()}, gspec, coarse.double())
    val, norms, gcf = otr.gp_param_grads_closed_form(c64, cspec, fine.double(), f64, alpha.double(), hp)
    cp = onet.as_leaf_params(c64)
    gp64, n64, _ = otr.gradient_penalty(cp, cspec, fine.double(), f64, alpha.double(), hp)
    ag = torch.autograd.grad(hp.gp_lambda * gp64, list(cp.values()), allow_unused=True)
    worst = 0.0
    for (k, _p), a in zip(cp.items(), ag):
        a = torch.zeros_like(gcf[k]) if a is None else a
        denom = float(a.norm()) + 1e-300
        worst = max(worst, float((a - gcf[k]).norm()) / denom if float(a.norm()) > 0 else float(gcf[k].norm()))
    print(f"[{name}] reference==oracle bit-exact (init, G fwd, C fwd, critic/generator losses+grads); "
          f"closed-form GP vs autograd fp64 worst rel {worst:.2e}; gp={float(gp):.6f} norms~{float(norms.mean()):.3e}")
    assert worst < 1e-8
    return g_sd, c_sd, coarse, fine, alpha, oc, og


def _ref_losses():
    """GAN/losses.py with a stub for `pytorch_msssim` (only SSIM_Loss, which is out of scope, needs it)."""
    import types
    if "pytorch_msssim" not in sys.modules:
        stub = types.ModuleType("pytorch_msssim")
        stub.MS_SSIM = object
        sys.modules["pytorch_msssim"] = stub
    sys.path.insert(0, REF)
    from DoWnGAN.GAN import losses  # type: ignore
    return losses


def ref_fs_critic_step_grads(G, C, coarse, fine, alpha, hp: otr.Hyper, filter_size=5):
    """wasserstein_fs.py:36-60 + its _gp (:94-124) on the reference modules; filters as hyperparams.py:31-35."""
    low = torch.nn.AvgPool2d(filter_size, stride=1, padding=0)
    rf = torch.nn.ReplicationPad2d(filter_size // 2)
    fake = G(coarse)
    fake_low = low(rf(fake))
    real_low = low(rf(fine))
    fake_high = fake - fake_low
    real_high = fine - real_low
    c_real = C(real_high)
    c_fake = C(fake_high)
    a = alpha.expand_as(real_high)
    interpolated = (a * real_high.data + (1 - a) * fake_high.data).requires_grad_(True)
    ci = C(interpolated)
    gradients = torch.autograd.grad(outputs=ci, inputs=interpolated, grad_outputs=torch.ones(ci.size()),
                                    create_graph=True, retain_graph=True)[0]
    gn = torch.sqrt(torch.sum(gradients.view(fine.size(0), -1) ** 2, dim=1) + 1e-12)
    gp = hp.gp_lambda * torch.mean((gn - 1) ** 2)
    C.zero_grad()
    loss = torch.mean(c_fake) - torch.mean(c_real) + hp.gp_lambda * gp
    loss.backward(retain_graph=True)
    grads = OrderedDict((k, (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)))
                        for k, p in C.named_parameters())
    return loss.detach(), gp.detach(), grads


def ref_fs_generator_step_grads(G, C, coarse, fine, hp: otr.Hyper, filter_size=5):
    """wasserstein_fs.py:71-91 on the reference modules."""
    low = torch.nn.AvgPool2d(filter_size, stride=1, padding=0)
    rf = torch.nn.ReplicationPad2d(filter_size // 2)
    G.zero_grad()
    fake = G(coarse)
    fake_low = low(rf(fake))
    real_low = low(rf(fine))
    fake_high = fake - fake_low
    c_fake = C(fake_high)
    g_loss = -torch.mean(c_fake) * hp.gamma
    g_loss = g_loss + hp.content_lambda * torch.nn.L1Loss()(fake_low, real_low)
    g_loss.backward()
    return g_loss.detach(), OrderedDict((k, p.grad.detach().clone()) for k, p in G.named_parameters())


def extra_rows(gspec, cspec, g_sd, c_sd, coarse, fine, alpha):
    """Step 5 of the module docstring; returns the dict written to tiny_extra.json."""
    Generator, Critic = _ref_modules()
    hp = otr.Hyper()
    C = Critic(cspec.coarse_dim, cspec.fine_dim, cspec.nc)
    G = Generator(gspec.filters, cspec.fine_dim, gspec.channels, gspec.n_predictands, gspec.num_res_blocks, gspec.num_upsample)
    C.load_state_dict(c_sd)
    G.load_state_dict(g_sd)
    torch.save(G.state_dict(), os.path.join(GOLD, "ref_tiny_generator_state_dict.pth"))
    torch.save(C.state_dict(), os.path.join(GOLD, "ref_tiny_critic_state_dict.pth"))
    # metrics, as gen_batch_and_log_metrics computes them
    losses = _ref_losses()
    with torch.no_grad():
        fake = G(coarse).detach()
        creal = torch.mean(C(fine)).detach()
        cfake = torch.mean(C(fake)).detach()
    ref_m = {"MAE": losses.content_loss(fine, fake, "cpu"), "MSE": losses.content_MSELoss(fine, fake, "cpu"),
             "Wass": losses.wass_loss(creal, cfake, "cpu")}
    om = otr.batch_metrics(g_sd, gspec, c_sd, cspec, coarse, fine)
    for k, v in ref_m.items():
        assert torch.equal(v, om[k]), f"metric {k}: reference {float(v)} oracle {float(om[k])}"
    # frequency separation
    loss, gp, cgr = ref_fs_critic_step_grads(G, C, coarse, fine, alpha, hp)
    oc = otr.critic_loss_and_grads_fs(g_sd, gspec, c_sd, cspec, coarse, fine, alpha, hp)
    assert torch.equal(oc["loss"], loss) and torch.equal(oc["gp"], gp)
    for k in cgr:
        assert torch.equal(cgr[k], oc["grads"][k]), f"FS critic grad mismatch {k}"
    gl, ggr = ref_fs_generator_step_grads(G, C, coarse, fine, hp)
    og = otr.generator_loss_and_grads_fs(g_sd, gspec, c_sd, cspec, coarse, fine, hp)
    assert torch.equal(og["loss"], gl)
    for k in ggr:
        assert torch.equal(ggr[k], og["grads"][k]), f"FS generator grad mismatch {k}"
    print("[tiny] reference==oracle bit-exact (metrics MAE/MSE/Wass via GAN/losses.py; frequency-separation critic and "
          "generator losses + all gradients)")
    return {"metrics": {k: float(v) for k, v in ref_m.items()}, "c_real_mean": float(creal), "c_fake_mean": float(cfake),
            "fs_critic_loss": float(loss), "fs_gp": float(gp), "fs_gen_loss": float(gl), "fs_l1": float(og["l1"]),
            "fs_dC_norm": {k: float(v.norm()) for k, v in cgr.items()},
            "fs_dG_total_norm": float(torch.sqrt(sum((v.double() ** 2).sum() for v in ggr.values()))),
            "fs_dG_norm_tail": {k: float(v.norm()) for k, v in list(ggr.items())[-10:]}}


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    # ---- tiny: small enough to commit in full --------------------------------
    gspec = onet.GeneratorSpec(filters=8, channels=3, n_predictands=2, num_res_blocks=2, num_upsample=3)
    cspec = onet.CriticSpec(coarse_dim=8, fine_dim=64, nc=2)
    g_sd, c_sd, coarse, fine, alpha, oc, og = check_config("tiny", gspec, cspec, b=4, seed=0, warm_steps=6)
    blob = {"coarse": coarse.numpy(), "fine": fine.numpy(), "alpha": alpha.numpy(),
            "fake": oc["fake"].numpy(), "c_real": oc["c_real"].numpy(), "c_fake": oc["c_fake"].numpy(),
            "gp": oc["gp"].numpy(), "critic_loss": oc["loss"].numpy(), "norms": oc["norms"].numpy(),
            "gp_input_grad": oc["gp_input_grad"].numpy(),
            "gen_loss": og["loss"].numpy(), "l1": og["l1"].numpy()}
    for k, v in g_sd.items():
        blob["G/" + k] = v.numpy()
    for k, v in c_sd.items():
        blob["C/" + k] = v.numpy()
    for k, v in oc["grads"].items():
        blob["dC/" + k] = v.numpy()
    for k, v in og["grads"].items():
        blob["dG/" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, "tiny.npz"), **blob)
    print("wrote tiny.npz", os.path.getsize(os.path.join(GOLD, "tiny.npz")) // 1024, "KiB")
    extra = extra_rows(gspec, cspec, g_sd, c_sd, coarse, fine, alpha)
    with open(os.path.join(GOLD, "tiny_extra.json"), "w") as f:
        json.dump(extra, f, indent=1)
    print("wrote tiny_extra.json and the reference state_dict checkpoints")

    # ---- cfg-1 (BASELINE.json configs[0]): summary only ----------------------
    gspec = onet.GeneratorSpec(filters=16, channels=2)
    cspec = onet.CriticSpec(coarse_dim=16, fine_dim=128, nc=2)
    g_sd, c_sd, coarse, fine, alpha, oc, og = check_config("cfg1", gspec, cspec, b=16, seed=0)
    summ = {
        "config": "cfg-1: B=16, 2ch, 16->128, F=16, 16 RRDB, seed 0, data seed 1234, alpha seed 4321",
        "fake_mean": float(oc["fake"].mean()), "fake_std": float(oc["fake"].std()),
        "fake_sha_first8": [float(x) for x in oc["fake"].flatten()[:8]],
        "c_real_mean": float(oc["c_real_mean"]), "c_fake_mean": float(oc["c_fake_mean"]),
        "gp": float(oc["gp"]), "critic_loss": float(oc["loss"]), "gen_loss": float(og["loss"]), "l1": float(og["l1"]),
        "norms": [float(x) for x in oc["norms"]],
        "dC_norm": {k: float(v.norm()) for k, v in oc["grads"].items()},
        "dG_norm_first": {k: float(v.norm()) for k, v in list(og["grads"].items())[:6]},
        "dG_total_norm": float(torch.sqrt(sum((v.double() ** 2).sum() for v in og["grads"].values()))),
    }
    with open(os.path.join(GOLD, "cfg1_summary.json"), "w") as f:
        json.dump(summ, f, indent=1)
    print("wrote cfg1_summary.json")


if __name__ == "__main__":
    main()
