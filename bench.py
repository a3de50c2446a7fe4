#!/usr/bin/env python
"""bench.py — WGAN-GP train samples/s of the DoWnGAN iteration on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one batch of the reference's epoch loop (GAN/wasserstein.py:131-147):
a critic iteration on a fresh batch, plus a generator iteration on the same batch
when step % 5 == 0.  Workload = BASELINE.json configs[1]: 2-ch 16x16 -> 128x128,
F=16, 16 RRDB, batch 64 per GPU, n_critic=5, synthetic ERA-shaped fields,
random-init weights.  Data parallel is weak scaling: every rank trains on its own
64-sample shard and the flat gradient buckets are all-reduced over NCCL.

Prints ONE JSON line (rank 0).  `value` is timed with the batches resident in
HBM; `e2e` is the same loop through the public trainer API with pinned HOST
batches (H2D of every batch and D2H of the loss scalars inside the timed
region).  `--impl reference` times the CPU oracle (the reference's PyTorch
step restated in oracle/, pinned bit-exact against the reference modules) on
the host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "wgan_gp_train_samples_per_sec"
CFG = dict(filters=16, channels=2, n_pred=2, rrdb=16, up=3, coarse=16, fine=128, batch=64, n_critic=5)
# algorithmic conv+linear FLOPs per sample (necessary work only)
F_G, F_C = 1.0347e9, 0.1998e9
FLOPS_CRITIC_STEP = F_G + 10 * F_C
FLOPS_GEN_STEP = 3 * F_G + 2 * F_C


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []   # (host time the line arrived, line)
        self.windows = []  # (t0, t1) host times of the timed regions

    def mark(self, t0: float, t1: float):
        self.windows.append((t0, t1))

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # the sampler is started before the warm-up (its start-up cost stays out of the timed regions); only samples
        # that arrived inside a timed region (+100 ms: one sampling period) count, all of them if none did
        inside = [ln for t, ln in self.lines if any(a <= t <= b + 0.1 for a, b in self.windows)]
        scope = "timed regions" if inside else "whole run (timed regions shorter than the sampling period)"
        for ln in (inside or [ln for _, ln in self.lines]):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


# ------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------
def cpu_oracle_steps(n_steps: int, warmup: int, batch: int, seed: int = 0):
    """Times `n_steps` steps of the schedule on the host with the oracle. Returns (seconds, samples)."""
    from downgan_b200.synthetic import synth_batch
    from oracle import networks as onet
    from oracle import trainer as otr
    torch.set_num_threads(os.cpu_count() or 1)
    gspec = onet.GeneratorSpec(CFG["filters"], CFG["channels"], CFG["n_pred"], CFG["rrdb"], CFG["up"])
    cspec = onet.CriticSpec(CFG["coarse"], CFG["fine"], CFG["n_pred"])
    torch.manual_seed(seed)
    c_sd = onet.init_critic_state(cspec)
    g_sd = onet.init_generator_state(gspec)
    tr = otr.OracleTrainer(g_sd, gspec, c_sd, cspec)
    coarse, fine, alpha = synth_batch(batch, CFG["channels"], CFG["coarse"])
    for _ in range(warmup):
        tr.batch(coarse, fine, alpha)
    tr.num_steps = 0
    t0 = time.perf_counter()
    for _ in range(n_steps):
        tr.batch(coarse, fine, alpha)
    return time.perf_counter() - t0, n_steps * batch


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    sample_b = 16
    secs, samples = cpu_oracle_steps(args.steps, args.warmup, sample_b)
    val = samples / secs
    cores = os.cpu_count() or 1
    sample = (f"{args.steps} steps of the schedule (critic every step, generator every 5th) at batch {sample_b} "
              f"(a quarter of each 64-sample batch), fp32, torch {torch.__version__} CPU, {cores} threads")
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, 1),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def config_dict(args, world):
    return {"workload": "cfg-2: DoWnGAN WGAN-GP, 2-ch (u10,v10) 16x16->128x128 (8x), F=16, 16 RRDB, batch 64 per GPU, "
                        "n_critic=5 (critic every step, generator every 5th step)",
            "global_batch": CFG["batch"] * world, "per_gpu_batch": CFG["batch"], "parallelism": f"dp{world}",
            "precision": args.precision,
            "l2_policy": "per-step working set (activations of a 192-sample critic batch, >1 GB) exceeds the 126 MB L2; "
                         "8 distinct input batches are cycled"}


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DOWNGAN_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist
    from downgan_b200 import _lib
    from downgan_b200.GAN.wasserstein import WassersteinGAN
    from downgan_b200.networks import Critic, Generator
    from downgan_b200.synthetic import synth_batch

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created; keep stdout for the
        # single JSON line by pointing fd 1 at stderr during initialisation and the first collective.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    lib = _lib.load()

    B = CFG["batch"]
    torch.manual_seed(0)  # identical replicas on every rank
    Cn = Critic(CFG["coarse"], CFG["fine"], CFG["n_pred"], precision=args.precision).to(dev)
    Gn = Generator(CFG["filters"], CFG["fine"], CFG["channels"], CFG["n_pred"], CFG["rrdb"], CFG["up"],
                   precision=args.precision).to(dev)
    gopt = torch.optim.Adam(Gn.parameters(), 2.5e-4, betas=(0.9, 0.99))
    copt = torch.optim.Adam(Cn.parameters(), 2.5e-4, betas=(0.9, 0.99))
    tr = WassersteinGAN(Gn, Cn, gopt, copt)

    NBATCH = 8
    host = [synth_batch(B, CFG["channels"], CFG["coarse"], seed=1234 + 100 * rank + i, aseed=4321 + 100 * rank + i)
            for i in range(NBATCH)]
    host = [(c.pin_memory(), f.pin_memory(), a.pin_memory()) for c, f, a in host]
    devb = [(c.to(dev), f.to(dev), a.to(dev)) for c, f, a in host]
    h2d = sum(t.numel() * 4 for t in host[0])
    d2h = 8 * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(steps, batches, read_scalars, step0=0):
        for s in range(steps):
            c, f, a = batches[(step0 + s) % NBATCH]
            tr._critic_train_iteration(c, f, a)
            if (step0 + s) % CFG["n_critic"] == 0:
                tr._generator_train_iteration(c, f)
            if read_scalars:
                tr.last_critic.cpu()  # D2H of the step's loss scalars (synchronises, as a logging caller would)

    def timed(steps, batches, read_scalars):
        # the public epoch loop on DEVICE-resident batches (schedule + look-ahead generator forward included)
        tr.num_steps = 0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.dg_launch_count()
        e0.record()
        tr._train_epoch([batches[s % NBATCH] for s in range(steps)])
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), lib.dg_launch_count() - l0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # warm-up (also creates the native handles and workspaces)
    run(args.warmup, devb, False)
    tr.num_steps = 0
    tr._train_epoch([devb[s % NBATCH] for s in range(max(args.warmup, 6))])
    barrier()
    t_w0 = time.time()
    ms, launches = timed(args.steps, devb, False)
    sampler.mark(t_w0, time.time())

    # ---- e2e: the public trainer API on HOST batches (wasserstein.py:120-147 `_train_epoch`):
    # every batch is copied host->device inside the timed region (pinned memory, side stream) and the
    # loss scalars of every step are copied device->host.
    def epoch_batches(k, step0):
        return [host[(step0 + s) % NBATCH] for s in range(k)]

    tr.num_steps = 0
    tr._train_epoch(epoch_batches(args.warmup, 0))
    tr.num_steps = 0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_w0 = time.time()
    e0.record()
    logs = tr._train_epoch(epoch_batches(args.steps, 0))
    e1.record()
    barrier()
    sampler.mark(t_w0, time.time())
    assert logs.shape == (args.steps, 8) and bool(torch.isfinite(logs).all())
    t_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = float(t_e2e.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel-class CUDA-event pass (same steps, events around every launch) ---------
    prof = None
    if not args.no_profile:
        # kernels timed one at a time: the side-stream overlap (dg_set_tuning keys 9, 10, 12) is switched off for this pass,
        # otherwise an event pair around a launch also spans whatever runs concurrently on the other stream
        prev9, prev10, prev12 = lib.dg_set_tuning(9, 0), lib.dg_set_tuning(10, 0), lib.dg_set_tuning(12, 0)
        lib.dg_profile(1)
        tr.num_steps = 0
        tr._train_epoch([devb[s % NBATCH] for s in range(args.steps)])
        buf = (C.c_double * (4 * len(_lib.PROFILE_CLASSES)))()
        _lib.check(lib.dg_profile_report(buf, len(_lib.PROFILE_CLASSES)))
        lib.dg_profile(0)
        lib.dg_set_tuning(9, prev9); lib.dg_set_tuning(10, prev10); lib.dg_set_tuning(12, prev12)
        prof = {n: {"launches": int(buf[4 * i]), "ms": buf[4 * i + 1], "flops": buf[4 * i + 2], "bytes": buf[4 * i + 3]}
                for i, n in enumerate(_lib.PROFILE_CLASSES) if buf[4 * i] > 0}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    samples = args.steps * B * world
    value = samples / (ms * 1e-3)
    e2e = samples / (ms_e2e * 1e-3)
    hbm, tf_burst, tf_sus, which = peaks()
    n_gen = (args.steps + CFG["n_critic"] - 1) // CFG["n_critic"]
    alg_flops_per_rank = B * (args.steps * FLOPS_CRITIC_STEP + n_gen * FLOPS_GEN_STEP)

    roofline = None
    if prof:
        conv = [k for k in prof if k in ("conv_direct", "wgrad_direct", "conv_tcgen05", "wgrad_tcgen05", "dense_block_tcgen05")]
        dom = max(conv, key=lambda k: prof[k]["ms"])
        d = prof[dom]
        ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
        if os.path.exists(tpath):  # dram__bytes_read+write per launch from the committed ncu --set full captures
            with open(tpath) as f:
                traffic = json.load(f).get(dom, {}).get("bytes_per_launch")
        # The dominant conv class mixes low-intensity layers (N = 16..32, ~30 FLOP/B, below the 254 FLOP/B
        # ridge) with denser ones: report the bound under which it stands closer to its ceiling, and both.
        gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9
        frac_t, frac_h = ach / tf_sus, gbs / hbm
        by_hbm = frac_h > frac_t
        roofline = {"bound": "hbm" if by_hbm else "tensor", "kernel": dom,
                    "achieved": gbs if by_hbm else ach, "peak": hbm if by_hbm else tf_sus,
                    "unit": "GB/s" if by_hbm else "TFLOP/s", "frac": frac_h if by_hbm else frac_t,
                    "frac_tensor": frac_t, "frac_hbm": frac_h, "achieved_tflops": ach, "achieved_gbs": gbs,
                    "traffic": traffic,
                    "peak_source": f"{which} (copy bandwidth / sustained cuBLAS bf16 from MEASURED_PEAKS.json)",
                    "launches": d["launches"], "avg_launch_us": 1e3 * d["ms"] / d["launches"],
                    "share_of_profiled_time": d["ms"] / sum(v["ms"] for v in prof.values()),
                    "whole_step_algorithmic_tflops": alg_flops_per_rank / (ms * 1e-3) / 1e12,
                    "classes": {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                                    "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 3) if v["flops"] else None,
                                    "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)} for k, v in prof.items()},
                    "hbm_peak_gbs": hbm}

    cpu = None
    if not args.no_cpu_baseline and world == 1:  # reported at N=1 only (the reference arm covers every N)
        secs, smp = cpu_oracle_steps(5, 1, 16)
        cores = os.cpu_count() or 1
        cpu = {"value": smp / secs, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"5 steps of the schedule (5 critic + 1 generator iterations) at batch 16, fp32 oracle, {cores} threads"}

    out = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": config_dict(args, world),
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "tcgen05": bool(lib.dg_has_tcgen05()),
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
