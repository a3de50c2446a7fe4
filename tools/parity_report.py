#!/usr/bin/env python
"""Writes profiles/parity_r02.md: gradient parity of the benchmarked bf16/tcgen05 path against the fp32 oracle, free-running
and with the LeakyReLU masks pinned to the CUDA path's (tests/mask_util.py), next to the emulated-bf16 oracle's own error.
Run on the B200 box:  python tools/parity_report.py > gpurun_out/parity_r02.md"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import torch  # noqa: E402

from downgan_b200.synthetic import synth_batch  # noqa: E402
from oracle import networks as onet  # noqa: E402

import mask_util as mu  # noqa: E402
import parity_util as pu  # noqa: E402

G_SPEC = onet.GeneratorSpec(filters=16, channels=2)
C_SPEC = onet.CriticSpec(coarse_dim=16, fine_dim=128, nc=2)


def table(title, r, top=None):
    print(f"\n### {title}\n")
    print(f"flat gradient error: free-running **{mu.flat_err(r['err_free']):.3e}**, masks pinned **{mu.flat_err(r['err_pinned']):.3e}**, "
          f"emulated-bf16 oracle vs fp32 oracle {mu.flat_err(r['err_emulated']):.3e}\n")
    print("| tensor | \\|ref\\| | free-running | masks pinned | emulated bf16 (oracle vs oracle) |")
    print("|---|---|---|---|---|")
    keys = list(r["err_pinned"].keys())
    if top is not None and len(keys) > top:  # the generator has 494 tensors: worst `top` by pinned error + the non-trunk ones
        worst = sorted(keys, key=lambda k: -r["err_pinned"][k][0])[:top]
        keys = [k for k in keys if k in worst or not k.startswith("res_blocks.")]
    for k in keys:
        ef, n = r["err_free"][k]
        ep, _ = r["err_pinned"][k]
        ee, _ = r["err_emulated"][k]
        print(f"| `{k}` | {n:.3e} | {ef:.3e} | {ep:.3e} | {ee:.3e} |")
    if top is not None:
        tr = [k for k in r["err_pinned"] if k.startswith("res_blocks.")]
        for name, sel in (("weights", ".weight"), ("biases", ".bias")):
            ks = [k for k in tr if k.endswith(sel)]
            if ks:
                print(f"| res_blocks.* {name} ({len(ks)} tensors): max / median | | "
                      f"{max(r['err_free'][k][0] for k in ks):.3e} / {sorted(r['err_free'][k][0] for k in ks)[len(ks)//2]:.3e} | "
                      f"{max(r['err_pinned'][k][0] for k in ks):.3e} / {sorted(r['err_pinned'][k][0] for k in ks)[len(ks)//2]:.3e} | "
                      f"{max(r['err_emulated'][k][0] for k in ks):.3e} / {sorted(r['err_emulated'][k][0] for k in ks)[len(ks)//2]:.3e} |")
    print("\nLeakyReLU branches that differ from the free-running fp32 oracle (fraction of the layer's elements; largest "
          "oracle |z| / rms(z) among them):\n")
    for name, rows in r["flips"].items():
        fr = ", ".join(f"{f:.2e}" for f, _ in rows[:12]) + (" ..." if len(rows) > 12 else "")
        print(f"* `{name}`: max fraction {max(f for f, _ in rows):.3e}, max |z|/rms {max(z for _, z in rows):.3e}; per layer: {fr}")


def main():
    print("# Gradient parity of the bf16 / tcgen05 path (round 2)\n")
    print(f"Device {torch.cuda.get_device_name(0)}; CUDA path = `dg_critic_step` / `dg_generator_step` through the C ABI, bf16 storage, "
          "fp32 accumulate.  Oracle = `oracle/trainer.py` in fp32 on the CPU.  \"masks pinned\" = the oracle replays the LeakyReLU "
          "branches the CUDA path took (`oracle.networks.MaskTape`), so both differentiate the same linear map; north_star's "
          "tolerance for bf16 is 2e-2.  \"emulated bf16\" = the fp32 oracle with every conv input / weight rounded to bf16 "
          "(free-running masks) against the plain fp32 oracle: the error bf16 STORAGE alone causes, whatever the kernels.\n")
    for batch, scale in ((16, 1.0), (16, 1.9), (64, 1.9)):
        G, C, g_sd, c_sd = pu.build_pair(G_SPEC, C_SPEC, "bf16", seed=0, critic_scale=scale)
        coarse, fine, alpha = synth_batch(batch, 2, 16)
        r = mu.critic_parity(G, C, G_SPEC, C_SPEC, g_sd, c_sd, coarse, fine, alpha)
        sc = r["scalars"]
        print(f"\n## cfg-{'1' if batch == 16 else '2'} size, B = {batch}, critic conv weights x{scale}\n")
        print(f"critic loss {float(sc[0]):.6f} (oracle {float(r['free']['loss']):.6f}), gp {float(sc[3]):.6f} "
              f"(oracle {float(r['free']['gp']):.6f}), fake rel. error {r['fake_rel']:.3e}")
        table("critic iteration (`wasserstein.py:27-52`)", r)
        if scale != 1.0:
            rg = mu.generator_parity(G, C, G_SPEC, C_SPEC, g_sd, c_sd, coarse, fine)
            sg = rg["scalars"]
            print(f"\ngenerator loss {float(sg[0]):.6f} (oracle {float(rg['free']['loss']):.6f}); sign(fake - fine) of the L1 term differs "
                  f"from the oracle's on {rg['l1_sign_flips']:.3e} of the elements (pinned as well in the \"masks pinned\" column)")
            table("generator iteration (`wasserstein.py:58-80`)", rg, top=12)
        del G, C


if __name__ == "__main__":
    main()
