"""SURVEY.md §8f rows on the GPU, each against the oracle / the committed golden values:
f-1 per-batch metric pass, f-2 HBM-resident loader + epoch loop with train and test metric means, f-3 reference checkpoints +
chunked full-domain inference, f-4 frequency-separation iterations and their filter kernel."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from downgan_b200 import _lib, inference
from downgan_b200.GAN.dataloader import DeviceLoader, NetCDFSR
from downgan_b200.GAN.wasserstein import WassersteinGAN
from downgan_b200.GAN.wasserstein_fs import WassersteinGANFS
from downgan_b200.networks import Critic, Generator
from downgan_b200.synthetic import synth_batch
from oracle import dataloader as odl
from oracle import networks as onet
from oracle import trainer as otr

import parity_util as pu
from test_gpu_parity import TINY_C, TINY_G

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = {"fp32": 1e-3, "bf16": 2e-2}


def _tiny():
    z = np.load(os.path.join(GOLD, "tiny.npz"))
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    g_sd = {k[2:]: v for k, v in t.items() if k.startswith("G/")}
    c_sd = {k[2:]: v for k, v in t.items() if k.startswith("C/")}
    with open(os.path.join(GOLD, "tiny_extra.json")) as f:
        extra = json.load(f)
    return t, g_sd, c_sd, extra


def _close(a, b, tol, floor=1e-6):
    return abs(float(a) - float(b)) <= tol * max(abs(float(b)), floor)


# ---------------------------------------------------------------- f-1: metric pass
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_metric_pass_against_reference_values(precision):
    """dg_metrics vs the values the reference's own content_loss / content_MSELoss / wass_loss gave (tiny_extra.json)."""
    t, g_sd, c_sd, extra = _tiny()
    G, C, _, _ = pu.build_pair(TINY_G, TINY_C, precision, g_sd=g_sd, c_sd=c_sd)
    tr = WassersteinGAN(G, C, None, None)
    m = tr._metrics_batch(t["coarse"], t["fine"]).cpu()
    tol = TOL[precision]
    assert _close(m[0], extra["metrics"]["MAE"], tol) and _close(m[1], extra["metrics"]["MSE"], tol)
    # Wass is a difference of two nearly equal means: judged on the scale of the means
    scale = max(abs(extra["c_real_mean"]), abs(extra["c_fake_mean"]))
    assert abs(float(m[2]) - extra["metrics"]["Wass"]) <= tol * scale
    assert _close(m[3], extra["c_real_mean"], tol) and _close(m[4], extra["c_fake_mean"], tol)
    assert abs(float(m[2]) - (float(m[3]) - float(m[4]))) < 1e-7


def test_epoch_metrics_and_test_pass_follow_the_reference_loop():
    """`_train_epoch(dataloader, testdataloader)` with track_metrics: per-batch metrics are taken AFTER the batch's updates
    (wasserstein.py:131-147 then mlflow_epoch.py:53-63), means over batches for the train and the test set
    (mlflow_epoch.py:38-49) - against the oracle stepping the same schedule, fp32 mode, look-ahead on (the fake of a
    non-generator step is reused from the look-ahead pass, a generator step forces a fresh forward)."""
    G, C, g_sd, c_sd = pu.build_pair(TINY_G, TINY_C, "fp32", seed=7)
    tr = WassersteinGAN(G, C, torch.optim.Adam(G.parameters(), 2.5e-4, betas=(0.9, 0.99)),
                        torch.optim.Adam(C.parameters(), 2.5e-4, betas=(0.9, 0.99)))
    tr.track_metrics = True
    train = [synth_batch(4, 3, 8, seed=300 + i, aseed=400 + i) for i in range(11)]
    test = [synth_batch(4, 3, 8, seed=500 + i)[:2] for i in range(3)]
    ref = otr.OracleTrainer(g_sd, TINY_G, c_sd, TINY_C)
    want = {"train": [], "test": []}
    for c_, f_, a_ in train:
        ref.batch(c_, f_, a_)
        want["train"].append(otr.batch_metrics(ref.g, TINY_G, ref.c, TINY_C, c_, f_))
    for c_, f_ in test:
        want["test"].append(otr.batch_metrics(ref.g, TINY_G, ref.c, TINY_C, c_, f_))
    tr._train_epoch(train, test)
    got = tr.last_epoch_metrics
    for split in ("train", "test"):
        for k in ("MAE", "MSE", "Wass"):
            w = float(torch.stack([m[k] for m in want[split]]).double().mean())
            scale = 1.0 if k != "Wass" else 0.0
            assert abs(got[split][k] - w) <= 2e-3 * max(abs(w), 1e-2 if k == "Wass" else 1e-6), (split, k, got[split][k], w)


# ---------------------------------------------------------------- f-2: loader
def test_device_loader_equals_torch_dataloader_batches():
    """Same seed -> the HBM gather delivers bit-identical batches, in the same order, as DataLoader(NetCDFSR, shuffle=True)."""
    n = 53
    gen = torch.Generator().manual_seed(9)
    coarse = torch.randn(n, 7, 16, 16, generator=gen)
    fine = torch.randn(n, 2, 128, 128, generator=gen)
    torch.manual_seed(321)
    want = odl.epoch_batches(odl.reference_loader(coarse, fine, 8, shuffle=True))
    torch.manual_seed(321)
    dl = DeviceLoader(NetCDFSR(coarse.cuda(), fine.cuda(), device=torch.device("cuda")), 8, shuffle=True)
    got = [(c.cpu(), f.cpu()) for c, f in dl]
    assert len(got) == len(want) == 7 and got[-1][0].shape[0] == 5
    for (gc, gf), (wc, wf) in zip(got, want):
        assert torch.equal(gc, wc) and torch.equal(gf, wf)


def test_training_from_device_loader_runs_the_schedule():
    G, C, _, _ = pu.build_pair(onet.GeneratorSpec(filters=16, channels=2, num_res_blocks=2), onet.CriticSpec(16, 128, 2), "bf16", seed=1)
    tr = WassersteinGAN(G, C, None, None)
    coarse, fine, _ = synth_batch(26, 2, 16)
    torch.manual_seed(5)
    dl = DeviceLoader(NetCDFSR(coarse.cuda(), fine.cuda()), 4, shuffle=True)
    logs = tr._train_epoch(dl)
    assert logs.shape == (7, 8) and bool(torch.isfinite(logs).all()) and tr.num_steps == 7
    assert tr._g_adam.step_count == 2 and tr._c_adam.step_count == 7  # generator at steps 0 and 5


# ---------------------------------------------------------------- f-3: checkpoints + chunked inference
def test_reference_checkpoint_round_trip():
    """A state_dict torch.save'd from the REFERENCE modules (what mlflow.pytorch.log_state_dict stores) loads strictly and
    reproduces the reference's outputs; saving from here gives a file the reference layout accepts (same keys / shapes)."""
    t, _g, _c, _e = _tiny()
    G = Generator(8, 64, 3, 2, num_res_blocks=2, precision="fp32")
    C = Critic(8, 64, 2, precision="fp32")
    inference.load_reference_checkpoint(G, os.path.join(GOLD, "ref_tiny_generator_state_dict.pth"))
    inference.load_reference_checkpoint(C, os.path.join(GOLD, "ref_tiny_critic_state_dict.pth"))
    G.cuda(), C.cuda()
    with torch.no_grad():
        assert pu.rel(G(t["coarse"].cuda()), t["fake"]) < 1e-5
        assert pu.rel(C(t["fine"].cuda()), t["c_real"]) < 1e-5
    ref_sd = torch.load(os.path.join(GOLD, "ref_tiny_generator_state_dict.pth"), weights_only=True)
    ours = {k: v.cpu() for k, v in G.state_dict().items()}
    assert list(ours.keys()) == list(ref_sd.keys())
    for k in ref_sd:
        assert ours[k].dtype == ref_sd[k].dtype and torch.equal(ours[k], ref_sd[k])


@pytest.mark.parametrize("n,chunks", [(23, 5), (4, 100), (17, 1)])
def test_generate_chunks_equals_direct_forward(n, chunks):
    """helpers/gen_fake_ds.py:152-158: the pipelined chunk loop writes exactly what G(chunk) gives, chunk boundaries as torch.chunk."""
    gspec = onet.GeneratorSpec(filters=16, channels=7, num_res_blocks=2)
    G, _C, g_sd, _ = pu.build_pair(gspec, onet.CriticSpec(16, 128, 2), "bf16", seed=4)
    coarse = synth_batch(n, 7, 16)[0].double()  # the reference converts with .float() per chunk
    out = inference.generate_chunks(G, coarse, chunks=chunks)
    assert out.shape == (n, 2, 128, 128) and out.dtype == torch.float32
    with torch.no_grad():
        parts = [G(p.float().cuda()).cpu() for p in torch.chunk(coarse, chunks)]
    assert torch.equal(out, torch.cat(parts))
    ref = onet.generator_forward(g_sd, gspec, coarse[:3].float())
    assert pu.rel(out[:3], ref) < 2e-2


# ---------------------------------------------------------------- f-4: frequency separation
@pytest.mark.parametrize("shape", [(3, 2, 16, 24), (2, 1, 5, 5), (1, 2, 128, 128)])
def test_lowpass_kernel_modes(shape):
    lib = _lib.load()
    x = torch.randn(shape, generator=torch.Generator().manual_seed(3))
    xd = x.cuda()
    n, c, h, w = shape

    def run(src, mode):
        y = torch.empty_like(src)
        _lib.check(lib.dg_lowpass(src.data_ptr(), y.data_ptr(), n * c, h, w, 5, mode, _lib.stream_ptr()))
        return y.cpu()

    low = otr.low_pass(x, 5)
    assert torch.allclose(run(xd, 0), low, atol=1e-6)
    assert torch.allclose(run(xd, 1), x - low, atol=1e-6)
    # adjoint: <low(x), v> == <x, low^T(v)>, and against autograd
    v = torch.randn(shape, generator=torch.Generator().manual_seed(4))
    xr = x.clone().requires_grad_(True)
    (otr.low_pass(xr, 5) * v).sum().backward()
    assert torch.allclose(run(v.cuda(), 2), xr.grad, atol=1e-5)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_frequency_separation_iterations(precision):
    """WassersteinGANFS (GAN/wasserstein_fs.py:28-91) against the oracle restatement and the values recorded from the reference
    statements on the reference modules (tiny_extra.json)."""
    t, g_sd, c_sd, extra = _tiny()
    G, C, _, _ = pu.build_pair(TINY_G, TINY_C, precision, g_sd=g_sd, c_sd=c_sd)
    tr = WassersteinGANFS(G, C, None, None)
    hp = otr.Hyper()
    oc = otr.critic_loss_and_grads_fs(g_sd, TINY_G, c_sd, TINY_C, t["coarse"], t["fine"], t["alpha"], hp)
    og = otr.generator_loss_and_grads_fs(g_sd, TINY_G, c_sd, TINY_C, t["coarse"], t["fine"], hp)
    lib = _lib.load()
    b = 4
    cd, fd, ad = t["coarse"].cuda(), t["fine"].cuda(), t["alpha"].reshape(b).cuda().contiguous()
    g, c = tr._handles(cd)
    sc, sg = torch.zeros(8, device="cuda"), torch.zeros(8, device="cuda")
    cg, gg = torch.zeros_like(C.flat_params()), torch.zeros_like(G.flat_params())
    _lib.check(lib.dg_critic_step(g, c, tr._hyper(), cd.data_ptr(), fd.data_ptr(), ad.data_ptr(), b, cg.data_ptr(), sc.data_ptr(),
                                  _lib.stream_ptr()))
    _lib.check(lib.dg_generator_step(g, c, tr._hyper(), cd.data_ptr(), fd.data_ptr(), b, gg.data_ptr(), sg.data_ptr(),
                                     _lib.stream_ptr()))
    torch.cuda.synchronize()
    tol = TOL[precision]
    assert _close(sc[0], extra["fs_critic_loss"], tol) and _close(sc[3], extra["fs_gp"], tol)
    assert _close(sg[0], extra["fs_gen_loss"], tol) and _close(sg[2], extra["fs_l1"], tol)
    assert _close(sc[0], oc["loss"], tol) and _close(sg[0], og["loss"], tol)
    gtol = 1e-3 if precision == "fp32" else 1e-1  # free-running bf16: LeakyReLU mask-flip floor (profiles/parity_r02.md)
    _w, _k, flat = pu.grad_report(pu.flat_to_dict(C, cg), oc["grads"])
    assert flat < gtol, flat
    _w, _k, flat = pu.grad_report(pu.flat_to_dict(G, gg), og["grads"])
    assert flat < gtol, flat
    if precision == "fp32":
        got = pu.flat_to_dict(G, gg)
        for k, n_ref in extra["fs_dG_norm_tail"].items():
            assert abs(float(got[k].norm()) - n_ref) <= 1e-3 * n_ref, k
