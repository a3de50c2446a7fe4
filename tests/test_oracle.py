"""CPU tests of the oracle itself: golden fixtures (generated from the real
reference modules by oracle/make_golden.py) and the closed-form GP spec."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import networks as onet
from oracle import trainer as otr

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def tiny():
    z = np.load(os.path.join(GOLD, "tiny.npz"))
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    gspec = onet.GeneratorSpec(filters=8, channels=3, n_predictands=2, num_res_blocks=2, num_upsample=3)
    cspec = onet.CriticSpec(coarse_dim=8, fine_dim=64, nc=2)
    g_sd = {k[2:]: v for k, v in t.items() if k.startswith("G/")}
    c_sd = {k[2:]: v for k, v in t.items() if k.startswith("C/")}
    return t, gspec, cspec, g_sd, c_sd


def test_state_dict_keys_follow_reference(tiny):
    t, gspec, cspec, g_sd, c_sd = tiny
    assert [k for k, _ in onet.generator_keys(gspec)] == list(g_sd.keys())
    assert [k for k, _ in onet.critic_keys(cspec)] == list(c_sd.keys())
    for k, shp in onet.generator_keys(gspec):
        assert tuple(g_sd[k].shape) == shp
    for k, shp in onet.critic_keys(cspec):
        assert tuple(c_sd[k].shape) == shp


def test_forward_matches_golden(tiny):
    t, gspec, cspec, g_sd, c_sd = tiny
    fake = onet.generator_forward(g_sd, gspec, t["coarse"])
    assert torch.allclose(fake, t["fake"], rtol=0, atol=1e-6)
    assert torch.allclose(onet.critic_forward(c_sd, cspec, t["fine"]), t["c_real"], rtol=0, atol=1e-6)
    assert torch.allclose(onet.critic_forward(c_sd, cspec, t["fake"]), t["c_fake"], rtol=0, atol=1e-6)


def test_steps_match_golden(tiny):
    t, gspec, cspec, g_sd, c_sd = tiny
    hp = otr.Hyper()
    oc = otr.critic_loss_and_grads(g_sd, gspec, c_sd, cspec, t["coarse"], t["fine"], t["alpha"], hp)
    assert torch.allclose(oc["loss"], t["critic_loss"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(oc["gp"], t["gp"], rtol=1e-5)
    for k, v in oc["grads"].items():
        ref = t["dC/" + k]
        assert (v - ref).norm() <= 1e-4 * ref.norm() + 1e-7, k
    og = otr.generator_loss_and_grads(g_sd, gspec, c_sd, cspec, t["coarse"], t["fine"], hp)
    assert torch.allclose(og["loss"], t["gen_loss"], rtol=1e-5)
    for k, v in og["grads"].items():
        ref = t["dG/" + k]
        assert (v - ref).norm() <= 1e-4 * ref.norm() + 1e-7, k


def test_gp_closed_form_equals_autograd_fp64(tiny):
    """The hand-written double backward (spec of the CUDA path) against create_graph autograd."""
    t, gspec, cspec, g_sd, c_sd = tiny
    hp = otr.Hyper()
    torch.manual_seed(3)
    # scale the critic so that ||grad|| straddles 1 (random init gives ~1e-4 and hides sign errors)
    c64 = {k: v.double() * (1.9 if k.endswith("weight") and v.dim() == 4 else 1.0) for k, v in c_sd.items()}
    real, fake, alpha = t["fine"].double(), t["fake"].double(), t["alpha"].double()
    val, norms, gcf = otr.gp_param_grads_closed_form(c64, cspec, real, fake, alpha, hp)
    cp = onet.as_leaf_params(c64)
    gp, n2, _ = otr.gradient_penalty(cp, cspec, real, fake, alpha, hp)
    ag = torch.autograd.grad(hp.gp_lambda * gp, list(cp.values()), allow_unused=True)
    assert torch.allclose(val, hp.gp_lambda * gp, rtol=1e-12)
    assert float(norms.min()) < 1.0 < float(norms.max()) or float(norms.mean()) > 1e-2
    for (k, _p), a in zip(cp.items(), ag):
        if a is None:
            assert float(gcf[k].abs().max()) == 0.0, k
        else:
            assert (a - gcf[k]).norm() <= 1e-9 * (a.norm() + 1e-30), k


def test_cfg1_summary_reproduces():
    """BASELINE cfg-1 scalars recorded from the reference modules (make_golden.py)."""
    with open(os.path.join(GOLD, "cfg1_summary.json")) as f:
        s = json.load(f)
    from downgan_b200.synthetic import synth_batch
    gspec = onet.GeneratorSpec(filters=16, channels=2)
    cspec = onet.CriticSpec(coarse_dim=16, fine_dim=128, nc=2)
    torch.manual_seed(0)
    c_sd = onet.init_critic_state(cspec)
    g_sd = onet.init_generator_state(gspec)
    coarse, fine, alpha = synth_batch(16, 2, 16)
    with torch.no_grad():
        fake = onet.generator_forward(g_sd, gspec, coarse[:2])
    assert np.allclose(fake.flatten()[:8].numpy(), np.array(s["fake_sha_first8"]), atol=1e-6)


def test_adam_matches_torch():
    torch.manual_seed(0)
    p = {"w": torch.randn(37)}
    q = torch.nn.Parameter(p["w"].clone())
    opt = torch.optim.Adam([q], 2.5e-4, betas=(0.9, 0.99))
    st = otr.AdamState()
    hp = otr.Hyper()
    for i in range(5):
        g = torch.randn(37)
        q.grad = g.clone()
        opt.step()
        otr.adam_update(p, {"w": g}, st, hp)
    assert torch.allclose(p["w"], q.detach(), atol=1e-7)


def test_mask_tape_record_replay(tiny):
    """MaskTape: replaying the masks a free run recorded reproduces that run (losses and every gradient,
    double backward of the gradient penalty included); a flipped mask changes them; the bf16-storage
    emulation stays close to fp32 on a tiny model."""
    t, gspec, cspec, g_sd, c_sd = tiny
    hp = otr.Hyper()
    c_sd = {k: v * (1.9 if k.endswith("weight") and v.dim() == 4 else 1.0) for k, v in c_sd.items()}
    rec = {k: onet.MaskTape() for k in ("gen", "real", "fake", "interp")}
    a = otr.critic_loss_and_grads(g_sd, gspec, c_sd, cspec, t["coarse"], t["fine"], t["alpha"], hp, tapes=rec)
    assert len(rec["real"].z) == 9 and len(rec["gen"].z) == 2 * 3 * 4 + 3 + 1
    rep = {k: onet.MaskTape([z > 0 for z in v.z]) for k, v in rec.items()}
    b = otr.critic_loss_and_grads(g_sd, gspec, c_sd, cspec, t["coarse"], t["fine"], t["alpha"], hp, tapes=rep)
    assert all(tp.i == len(tp.masks) for tp in rep.values())
    assert torch.allclose(a["loss"], b["loss"], rtol=1e-6)
    for k in a["grads"]:
        assert (a["grads"][k] - b["grads"][k]).norm() <= 1e-5 * a["grads"][k].norm() + 1e-9, k
    # flipping 2 % of one layer's masks must move the gradients (the replay really drives the branch)
    bad = [z > 0 for z in rec["interp"].z]
    bad[3] = bad[3] ^ (torch.rand(bad[3].shape, generator=torch.Generator().manual_seed(1)) < 0.02)
    rep2 = {k: onet.MaskTape([z > 0 for z in v.z]) for k, v in rec.items()}
    rep2["interp"] = onet.MaskTape(bad)
    c = otr.critic_loss_and_grads(g_sd, gspec, c_sd, cspec, t["coarse"], t["fine"], t["alpha"], hp, tapes=rep2)
    k0 = "features.0.weight"
    assert (c["grads"][k0] - a["grads"][k0]).norm() > 1e-2 * a["grads"][k0].norm()
    # generator objective
    rg = {k: onet.MaskTape() for k in ("gen", "fake")}
    ga = otr.generator_loss_and_grads(g_sd, gspec, c_sd, cspec, t["coarse"], t["fine"], hp, tapes=rg)
    gb = otr.generator_loss_and_grads(g_sd, gspec, c_sd, cspec, t["coarse"], t["fine"], hp,
                                      tapes={k: onet.MaskTape([z > 0 for z in v.z]) for k, v in rg.items()})
    for k in ga["grads"]:
        assert (ga["grads"][k] - gb["grads"][k]).norm() <= 1e-5 * ga["grads"][k].norm() + 1e-9, k
    ge = otr.generator_loss_and_grads(g_sd, gspec, c_sd, cspec, t["coarse"], t["fine"], hp,
                                      tapes={k: onet.MaskTape(bf16=True) for k in ("gen", "fake")})
    assert abs(float(ge["loss"]) - float(ga["loss"])) < 2e-2 * abs(float(ga["loss"]))
    num = sum(float((ge["grads"][k] - ga["grads"][k]).norm()) ** 2 for k in ga["grads"])
    den = sum(float(ga["grads"][k].norm()) ** 2 for k in ga["grads"])
    assert 0 < (num / den) ** 0.5 < 0.3


def test_adjacent_rows_match_reference_values(tiny):
    """tiny_extra.json holds what the REFERENCE's own code gave (oracle/make_golden.py step 5): GAN/losses.py metrics as
    gen_batch_and_log_metrics calls them, and the frequency-separation statements of GAN/wasserstein_fs.py on the reference modules."""
    t, gspec, cspec, g_sd, c_sd = tiny
    with open(os.path.join(GOLD, "tiny_extra.json")) as f:
        x = json.load(f)
    hp = otr.Hyper()
    m = otr.batch_metrics(g_sd, gspec, c_sd, cspec, t["coarse"], t["fine"])
    for k in ("MAE", "MSE", "Wass"):
        assert abs(float(m[k]) - x["metrics"][k]) <= 1e-6 * max(abs(x["metrics"][k]), 1e-3), k
    oc = otr.critic_loss_and_grads_fs(g_sd, gspec, c_sd, cspec, t["coarse"], t["fine"], t["alpha"], hp)
    og = otr.generator_loss_and_grads_fs(g_sd, gspec, c_sd, cspec, t["coarse"], t["fine"], hp)
    assert abs(float(oc["loss"]) - x["fs_critic_loss"]) <= 1e-5 * abs(x["fs_critic_loss"])
    assert abs(float(og["loss"]) - x["fs_gen_loss"]) <= 1e-5 * abs(x["fs_gen_loss"])
    for k, n in x["fs_dC_norm"].items():
        assert abs(float(oc["grads"][k].norm()) - n) <= 1e-4 * n + 1e-9, k
    # the low-pass filter restated = the module pair of config/hyperparams.py:31-35
    xin = torch.randn(2, 2, 9, 11, generator=torch.Generator().manual_seed(0))
    ref = torch.nn.AvgPool2d(5, stride=1, padding=0)(torch.nn.ReplicationPad2d(2)(xin))
    assert torch.equal(otr.low_pass(xin, 5), ref)


def test_reference_checkpoints_load_on_cpu():
    """The torch.save'd reference state_dicts have exactly the keys / shapes the product modules register (strict load)."""
    from downgan_b200.networks import Critic, Generator
    g = Generator(8, 64, 3, 2, num_res_blocks=2)
    c = Critic(8, 64, 2)
    g.load_state_dict(torch.load(os.path.join(GOLD, "ref_tiny_generator_state_dict.pth"), weights_only=True), strict=True)
    c.load_state_dict(torch.load(os.path.join(GOLD, "ref_tiny_critic_state_dict.pth"), weights_only=True), strict=True)
