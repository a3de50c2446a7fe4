"""Size-independent properties of the CUDA path at BASELINE.json's full single-GPU size (cfg-2:
B=64, 2-ch 16x16 -> 128x128, F=16, 16 RRDB) and on the other configs' shapes, where the CPU oracle
would take minutes: batch-split linearity of the gradients, a directional-derivative check of the
critic loss against its analytic gradient, ragged / single-sample batches, cfg-3 and cfg-4 shapes."""
import pytest
import torch

from downgan_b200 import _lib
from downgan_b200.synthetic import synth_batch
from oracle import networks as onet
from oracle import trainer as otr

import parity_util as pu
from test_gpu_parity import _run_steps

pytestmark = pytest.mark.gpu

CFG2_G = onet.GeneratorSpec(filters=16, channels=2)
CFG2_C = onet.CriticSpec(coarse_dim=16, fine_dim=128, nc=2)


def _flat(d):
    return torch.cat([v.reshape(-1).double() for v in d.values()])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_batch_split_linearity_cfg2(precision):
    """Every loss term is a mean over independent samples (GP included): the B=64 gradients must equal
    the average of the gradients of the two B=32 halves (same kernels, different tiling / batching)."""
    G, C, _, _ = pu.build_pair(CFG2_G, CFG2_C, precision, seed=0, critic_scale=1.9)
    coarse, fine, alpha = synth_batch(64, 2, 16)
    sc, cg, sg, gg = _run_steps(G, C, coarse, fine, alpha)
    halves = [_run_steps(G, C, coarse[i:i + 32], fine[i:i + 32], alpha[i:i + 32]) for i in (0, 32)]
    cg2 = 0.5 * (_flat(halves[0][1]) + _flat(halves[1][1]))
    gg2 = 0.5 * (_flat(halves[0][3]) + _flat(halves[1][3]))
    tol = 1e-4 if precision == "fp32" else 2e-3
    assert float((_flat(cg) - cg2).norm() / cg2.norm()) < tol
    assert float((_flat(gg) - gg2).norm() / gg2.norm()) < tol
    # scalars: losses are means too
    assert abs(float(sc[0]) - 0.5 * (float(halves[0][0][0]) + float(halves[1][0][0]))) < 1e-3 * abs(float(sc[0]))
    assert abs(float(sg[0]) - 0.5 * (float(halves[0][2][0]) + float(halves[1][2][0]))) < 1e-3 * abs(float(sg[0]))


def test_directional_derivative_of_critic_loss_cfg2(monkeypatch):
    """d/de L_C(theta + e*v) at e=0 by central differences of the fp32-mode loss against <grad, v>
    from dg_critic_step, at full cfg-2 size.  The Wasserstein part E[C(fake)] - E[C(real)] is continuous
    and piecewise smooth in the weights, so finite differences are meaningful; the gradient penalty is
    not (||dC/dx|| jumps whenever a LeakyReLU mask flips), it is checked against autograd instead
    (test_gp_standalone, test_steps_*), so gp_lambda is 0 here."""
    from downgan_b200.config import hyperparams as hpm
    monkeypatch.setattr(hpm, "gp_lambda", 0.0)
    G, C, _, _ = pu.build_pair(CFG2_G, CFG2_C, "fp32", seed=0, critic_scale=1.9)
    coarse, fine, alpha = synth_batch(64, 2, 16)
    sc, cg, _sg, _gg = _run_steps(G, C, coarse, fine, alpha)
    grad = _flat(cg)
    flat = C.flat_params()
    # direction = the gradient itself: the directional derivative is then ||grad||, far above the
    # resolution of an fp32 loss value of ~90 (a random direction gives a derivative lost in rounding)
    v = grad / grad.norm()
    base = flat.detach().clone()
    eps = 1e-2 / float(grad.norm())
    losses = []
    for sgn in (+1.0, -1.0):
        with torch.no_grad():
            flat.copy_(base + (sgn * eps * v).float().cuda())
        C.mark_params_changed()
        s, *_ = _run_steps(G, C, coarse, fine, alpha)
        losses.append(float(s[0]))
    with torch.no_grad():
        flat.copy_(base)
    C.mark_params_changed()
    fd = (losses[0] - losses[1]) / (2 * eps)
    an = float(torch.dot(grad, v))
    assert an > 0 and abs(fd - an) <= 3e-2 * an, (fd, an)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("batch", [1, 5])
def test_ragged_and_single_sample_batches(precision, batch):
    """The reference mis-shapes ragged batches (wasserstein.py:110 uses hp.batch_size); here any batch works."""
    G, C, g_sd, c_sd = pu.build_pair(CFG2_G, CFG2_C, precision, seed=1)
    coarse, fine, alpha = synth_batch(batch, 2, 16, seed=77)
    hp = otr.Hyper()
    oc = otr.critic_loss_and_grads(g_sd, CFG2_G, c_sd, CFG2_C, coarse, fine, alpha, hp)
    sc, cg, sg, gg = _run_steps(G, C, coarse, fine, alpha)
    tol = 1e-3 if precision == "fp32" else 2e-2
    assert abs(float(sc[0]) - float(oc["loss"])) <= tol * abs(float(oc["loss"]))
    _w, _k, flat = pu.grad_report(cg, oc["grads"])
    # a single sample has no averaging over the batch: the bf16 mask-flip floor is a little higher
    assert flat < (1e-3 if precision == "fp32" else 1.5e-1)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cfg3_seven_covariates_step(precision):
    """cfg-3 shapes: 7 input channels (land-sea mask channel is 0/1)."""
    gspec = onet.GeneratorSpec(filters=16, channels=7)
    G, C, g_sd, c_sd = pu.build_pair(gspec, CFG2_C, precision, seed=2)
    coarse, fine, alpha = synth_batch(4, 7, 16)
    hp = otr.Hyper()
    og = otr.generator_loss_and_grads(g_sd, gspec, c_sd, CFG2_C, coarse, fine, hp)
    _sc, _cg, sg, gg = _run_steps(G, C, coarse, fine, alpha)
    tol = 1e-3 if precision == "fp32" else 2e-2
    assert abs(float(sg[0]) - float(og["loss"])) <= tol * abs(float(og["loss"]))
    _w, _k, flat = pu.grad_report(gg, og["grads"])
    assert flat < (1e-3 if precision == "fp32" else 1e-1)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cfg4_like_shapes_step(precision):
    """cfg-4 family (F = 32, coarse 32x32 -> 256x256) at reduced depth (2 RRDB) and batch 2: exercises the
    per-layer tcgen05 kernels with 32..256-channel layers and the non-fused trunk path."""
    gspec = onet.GeneratorSpec(filters=32, channels=7, num_res_blocks=2)
    cspec = onet.CriticSpec(coarse_dim=32, fine_dim=256, nc=2)
    G, C, g_sd, c_sd = pu.build_pair(gspec, cspec, precision, seed=3)
    coarse, fine, alpha = synth_batch(2, 7, 32)
    hp = otr.Hyper()
    oc = otr.critic_loss_and_grads(g_sd, gspec, c_sd, cspec, coarse, fine, alpha, hp)
    og = otr.generator_loss_and_grads(g_sd, gspec, c_sd, cspec, coarse, fine, hp)
    sc, cg, sg, gg = _run_steps(G, C, coarse, fine, alpha)
    tol = 1e-3 if precision == "fp32" else 2e-2
    assert abs(float(sc[0]) - float(oc["loss"])) <= tol * abs(float(oc["loss"]))
    assert abs(float(sg[0]) - float(og["loss"])) <= tol * abs(float(og["loss"]))
    for got, ref in ((cg, oc["grads"]), (gg, og["grads"])):
        _w, _k, flat = pu.grad_report(got, ref)
        assert flat < (2e-3 if precision == "fp32" else 1.5e-1), flat


def test_inference_tiles_cfg5_like():
    """cfg-5: generator-only inference under no_grad on a larger tile (fully convolutional forward)."""
    gspec = onet.GeneratorSpec(filters=16, channels=7, num_res_blocks=2)
    G, _C, g_sd, _ = pu.build_pair(gspec, CFG2_C, "bf16", seed=4)
    coarse, _f, _a = synth_batch(2, 7, 64)
    with torch.no_grad():
        out = G(coarse.cuda())
        ref = onet.generator_forward(g_sd, gspec, coarse)
    assert out.shape == (2, 2, 512, 512)
    assert pu.rel(out, ref) < 2e-2


def test_lookahead_epoch_equals_plain_iterations():
    """`_train_epoch` with the look-ahead generator forward (fakes of the critic steps between two generator
    updates computed in one pass) must reproduce the step-by-step schedule exactly: same kernels per sample."""
    from downgan_b200.synthetic import synth_batch
    batches = [synth_batch(4, 2, 16, seed=50 + i, aseed=90 + i) for i in range(12)]
    logs = {}
    for la in (False, True):
        G, C, _, _ = pu.build_pair(onet.GeneratorSpec(filters=16, channels=2), onet.CriticSpec(coarse_dim=16, fine_dim=128, nc=2),
                                   "bf16", seed=3)
        from downgan_b200.GAN.wasserstein import WassersteinGAN
        tr = WassersteinGAN(G, C, torch.optim.Adam(G.parameters(), 2.5e-4, betas=(0.9, 0.99)),
                            torch.optim.Adam(C.parameters(), 2.5e-4, betas=(0.9, 0.99)))
        tr.lookahead = la
        logs[la] = tr._train_epoch(batches).clone()
        params = torch.cat([p.detach().flatten().cpu() for p in list(G.parameters()) + list(C.parameters())])
        logs[(la, "p")] = params
    assert torch.isfinite(logs[True]).all()
    # fp32 reductions across CTAs use atomics (summation order varies run to run), so this is a tolerance, not
    # bit equality.  Parameters: Adam's first steps move every weight by about +-lr whatever the gradient's
    # magnitude, so order noise on near-zero gradients appears at ~steps*lr/|w|; a wrong fake (offset bug)
    # would instead show up in the critic scalars at O(1).
    assert pu.rel(logs[True], logs[False]) < 1e-3
    assert pu.rel(logs[(True, "p")], logs[(False, "p")]) < 2e-2


def test_deferred_conv_gradients_equal_plain_step():
    """Data-parallel overlap hook: with dg_critic_defer_conv_grads the critic iteration returns with the classifier gradients
    final and the conv gradients pending; after dg_critic_step_finish the bucket must equal the plain iteration's (same kernels;
    fp32 atomics -> tolerance).  Until the finish call every other critic entry point refuses to run."""
    G, C, _g_sd, _c_sd = pu.build_pair(CFG2_G, CFG2_C, "bf16", seed=1, critic_scale=1.9)
    coarse, fine, alpha = synth_batch(8, 2, 16, seed=5, aseed=6)
    tr = pu.WassersteinGAN(G, C, None, None)
    lib = _lib.load()
    b = coarse.shape[0]
    cd, fd, ad = coarse.cuda(), fine.cuda(), alpha.reshape(b).cuda().contiguous()
    g, c = tr._handles(cd)
    off = C.param_offsets()[9]

    def step(defer):
        sc = torch.zeros(8, device="cuda")
        cg = torch.full_like(C.flat_params(), float("nan"))
        _lib.check(lib.dg_critic_defer_conv_grads(c, 1 if defer else 0))
        _lib.check(lib.dg_critic_step(g, c, tr._hyper(), cd.data_ptr(), fd.data_ptr(), ad.data_ptr(), b, cg.data_ptr(),
                                      sc.data_ptr(), _lib.stream_ptr()))
        if defer:
            torch.cuda.synchronize()
            assert torch.isfinite(cg[off:]).all(), "classifier gradients must be final before the finish call"
            assert lib.dg_critic_pack(c, C.flat_params().data_ptr(), _lib.stream_ptr()) != 0, "pending finish must be refused"
            _lib.check(lib.dg_critic_step_finish(c, cg.data_ptr(), _lib.stream_ptr()))
        torch.cuda.synchronize()
        return sc.cpu(), cg.cpu()

    # The conv / classifier weight-gradient reductions use fp32 atomics (red.global.add), so two runs of the SAME code differ by
    # their summation order.  That run-to-run floor is measured here (plain vs plain, three pairs) and the deferred path must
    # stay within 4x of it (+2e-4) on each slice; a real ordering bug (conv gradients unpacked before the side stream has finished,
    # a stale classifier slice) is O(1), orders of magnitude above any floor.
    plain = [step(False) for _ in range(4)]
    sc0, g0 = plain[0]
    floor_fc = max(pu.rel(g[off:], g0[off:]) for _s, g in plain[1:])
    floor_conv = max(pu.rel(g[:off], g0[:off]) for _s, g in plain[1:])
    sc1, g1 = step(True)
    _lib.check(lib.dg_critic_defer_conv_grads(c, 0))
    assert torch.isfinite(g1).all()
    assert pu.rel(sc1, sc0) < 1e-5
    e_fc, e_conv = pu.rel(g1[off:], g0[off:]), pu.rel(g1[:off], g0[:off])
    print(f"deferred vs plain: classifier {e_fc:.2e} (floor {floor_fc:.2e}), conv {e_conv:.2e} (floor {floor_conv:.2e})")
    assert e_fc <= 4 * floor_fc + 2e-4 and e_conv <= 4 * floor_conv + 2e-4, (e_fc, floor_fc, e_conv, floor_conv)
    assert max(e_fc, e_conv) < 5e-3  # sanity ceiling, far below an ordering bug
    # both also agree with the fp32 oracle as well as the plain bf16 iteration does
    oc = otr.critic_loss_and_grads(_g_sd, CFG2_G, _c_sd, CFG2_C, coarse, fine, alpha, otr.Hyper())
    ref = torch.cat([v.reshape(-1) for v in oc["grads"].values()])
    assert abs(pu.rel(g1, ref) - pu.rel(g0, ref)) < 1e-3


@pytest.mark.parametrize("key,value", [(19, 1), (20, 1), (21, 0), (21, 2), (22, 0), (18, 1), (18, 4), (23, 0), (24, 1), (25, 0)])
def test_kernel_selection_switches_keep_the_results(key, value):
    """Every alternative kernel path behind dg_set_tuning (second critic side stream, planar first-layer kernel, bf16 masks /
    whole-tile sign words instead of per-piece sign bits, general conv epilogue, trunk backward in 1 / 4 RRDB ranges) computes the
    same critic and generator iteration as the default path: scalars and flat gradients to 1e-3 (measured: 6e-8 .. 3e-7 = the
    atomics' run-to-run floor for every switch, i.e. identical LeakyReLU branches with sign bits and with bf16 masks; 1.8e-4 for
    the planar first-layer kernel, whose operand rounding differs)."""
    from downgan_b200 import _lib
    from test_gpu_parity import _run_steps
    lib = _lib.load()
    gspec, cspec = onet.GeneratorSpec(filters=16, channels=2), onet.CriticSpec(coarse_dim=16, fine_dim=128, nc=2)
    G, C, _, _ = pu.build_pair(gspec, cspec, "bf16", seed=3, critic_scale=1.9)
    coarse, fine, alpha = synth_batch(8, 2, 16, seed=21, aseed=22)
    sc0, cg0, sg0, gg0 = _run_steps(G, C, coarse, fine, alpha)
    prev = lib.dg_set_tuning(key, value)
    try:
        sc1, cg1, sg1, gg1 = _run_steps(G, C, coarse, fine, alpha)
    finally:
        lib.dg_set_tuning(key, prev)
    assert abs(float(sc1[0]) - float(sc0[0])) <= 1e-3 * abs(float(sc0[0])) and abs(float(sg1[0]) - float(sg0[0])) <= 1e-3 * abs(float(sg0[0]))
    flat = lambda d: torch.cat([v.reshape(-1) for v in d.values()])
    ec, eg = pu.rel(flat(cg1), flat(cg0)), pu.rel(flat(gg1), flat(gg0))
    print(f"tuning {key}={value}: critic grads {ec:.2e}, generator grads {eg:.2e}")
    assert ec < 1e-3 and eg < 1e-3
