"""Diagnostic (not a test): print every parity error of the CUDA path vs the oracle."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.nn.functional as F

import parity_util as pu
from downgan_b200.synthetic import synth_batch
from oracle import networks as onet
from oracle import trainer as otr
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as T


def main():
    torch.set_num_threads(os.cpu_count())
    print("device", torch.cuda.get_device_name(0))
    for precision in ("fp32", "bf16"):
        worst = 0
        for case in T.CONV_CASES:
            b, ci, co, h, w, s = case
            g = torch.Generator().manual_seed(1)
            x = torch.randn(b, ci, h, w, generator=g); wt = torch.randn(co, ci, 3, 3, generator=g) / (3 * ci ** .5); bias = torch.randn(co, generator=g)
            y_ref = F.leaky_relu(F.conv2d(x, wt, bias, stride=s, padding=1), 0.2)
            e1 = pu.rel(pu.conv_fwd(x, wt, bias, s, 0.2, precision), y_ref)
            dy = torch.randn_like(y_ref)
            e2 = pu.rel(pu.conv_dgrad(dy, wt, h, w, s, precision), torch.nn.grad.conv2d_input(x.shape, wt, dy, stride=s, padding=1))
            dw, db = pu.conv_wgrad(x, dy, s, precision)
            e3 = pu.rel(dw, torch.nn.grad.conv2d_weight(x, wt.shape, dy, stride=s, padding=1))
            e4 = pu.rel(db, dy.sum((0, 2, 3)))
            print(f"[{precision}] conv {case}: fwd {e1:.2e} dgrad {e2:.2e} wgrad {e3:.2e} dbias {e4:.2e}")
    z = np.load(os.path.join(ROOT, "tests/golden/tiny.npz"))
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    g_sd = {k[2:]: v for k, v in t.items() if k.startswith("G/")}
    c_sd = {k[2:]: v for k, v in t.items() if k.startswith("C/")}
    for precision in ("fp32", "bf16"):
        G, C, _, _ = pu.build_pair(T.TINY_G, T.TINY_C, precision, g_sd=g_sd, c_sd=c_sd)
        with torch.no_grad():
            print(f"[{precision}] tiny fwd: fake {pu.rel(G(t['coarse'].cuda()), t['fake']):.2e} c_real {pu.rel(C(t['fine'].cuda()), t['c_real']):.2e}")
        sc, cg, sg, gg = T._run_steps(G, C, t["coarse"], t["fine"], t["alpha"])
        print(f"[{precision}] tiny critic scalars {sc[:5].tolist()} ref loss {float(t['critic_loss']):.6f} gp {float(t['gp']):.6f}")
        print(f"[{precision}] tiny gen scalars {sg[:3].tolist()} ref loss {float(t['gen_loss']):.6f} l1 {float(t['l1']):.6f}")
        ref = {k[3:]: v for k, v in t.items() if k.startswith("dC/")}
        for k in ref:
            print(f"    dC {k:28s} rel {pu.rel(cg[k], ref[k]):.3e}  |ref| {float(ref[k].norm()):.3e}")
        w, wk, fl = pu.grad_report(gg, {k[3:]: v for k, v in t.items() if k.startswith("dG/")})
        print(f"[{precision}] tiny dG worst {w:.3e} ({wk}) flat {fl:.3e}")
        refg = {k[3:]: v for k, v in t.items() if k.startswith("dG/")}
        for k in list(refg)[:6] + list(refg)[-12:]:
            print(f"    dG {k:44s} rel {pu.rel(gg[k], refg[k]):.3e}  |ref| {float(refg[k].norm()):.3e}")
    for precision in ("fp32", "bf16"):
        for scale in (1.0, 1.9):
            G, C, g_sd1, c_sd1 = pu.build_pair(T.CFG1_G, T.CFG1_C, precision, seed=0, critic_scale=scale)
            coarse, fine, alpha = synth_batch(16, 2, 16)
            hp = otr.Hyper()
            t0 = time.time()
            oc = otr.critic_loss_and_grads(g_sd1, T.CFG1_G, c_sd1, T.CFG1_C, coarse, fine, alpha, hp)
            og = otr.generator_loss_and_grads(g_sd1, T.CFG1_G, c_sd1, T.CFG1_C, coarse, fine, hp)
            t1 = time.time()
            sc, cg, sg, gg = T._run_steps(G, C, coarse, fine, alpha)
            t2 = time.time()
            print(f"[{precision}] cfg1 scale {scale}: oracle {t1-t0:.2f}s cuda(first) {t2-t1:.2f}s")
            print(f"    critic: loss {float(sc[0]):.6f}/{float(oc['loss']):.6f} real {float(sc[1]):.6f}/{float(oc['c_real_mean']):.6f} fake {float(sc[2]):.6f}/{float(oc['c_fake_mean']):.6f} gp {float(sc[3]):.6f}/{float(oc['gp']):.6f} norms~{float(oc['norms'].mean()):.3e}")
            print(f"    gen: loss {float(sg[0]):.6f}/{float(og['loss']):.6f} l1 {float(sg[2]):.6f}/{float(og['l1']):.6f}")
            w, wk, fl = pu.grad_report(cg, oc["grads"]); print(f"    dC worst {w:.3e} ({wk}) flat {fl:.3e}")
            w, wk, fl = pu.grad_report(gg, og["grads"]); print(f"    dG worst {w:.3e} ({wk}) flat {fl:.3e}")


if __name__ == "__main__":
    main()
