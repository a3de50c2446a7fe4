#!/usr/bin/env python
"""bench.py — WGAN-GP train samples/s of the DoWnGAN iteration on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config cfg2|cfg3|cfg4|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one batch of the reference's epoch loop (GAN/wasserstein.py:131-147):
a critic iteration on a fresh batch, plus a generator iteration on the same batch
when step % 5 == 0.  Default workload = BASELINE.json configs[1] (cfg2): 2-ch 16x16 ->
128x128, F=16, 16 RRDB, batch 64 per GPU, n_critic=5, synthetic ERA-shaped fields,
random-init weights.  --config selects configs[2..4]: cfg3 (7 covariates), cfg4
(32x32 -> 256x256, F=32, 23 RRDB, batch 32 per GPU), cfg5 (generator-only inference
on 64x64 -> 512x512 tiles, F=64, batch 128 per GPU; a "step" is one forward pass).
Data parallel is weak scaling: every rank trains on its own shard and the flat
gradient buckets are all-reduced over NCCL (cfg5: independent replicas).

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
# BASELINE.json configs[1..4].  filters = coarse grid edge as the reference passes it (GAN/stage.py:59-60).
# ref_batch: batch of the CPU reference arm's bounded sample (the full batch where one CPU step takes about a second).
CONFIGS = {
    "cfg2": dict(filters=16, channels=2, n_pred=2, rrdb=16, up=3, coarse=16, fine=128, batch=64, n_critic=5, ref_batch=64,
                 mode="train", label="cfg-2: DoWnGAN WGAN-GP, 2-ch (u10,v10) 16x16->128x128 (8x), F=16, 16 RRDB, batch 64 per GPU, "
                                     "n_critic=5 (critic every step, generator every 5th step)"),
    "cfg3": dict(filters=16, channels=7, n_pred=2, rrdb=16, up=3, coarse=16, fine=128, batch=64, n_critic=5, ref_batch=64,
                 mode="train", label="cfg-3: 7-ch covariate input (u10,v10,land-sea mask,...) 16x16->128x128, F=16, 16 RRDB, "
                                     "batch 64 per GPU, n_critic=5"),
    "cfg4": dict(filters=32, channels=7, n_pred=2, rrdb=23, up=3, coarse=32, fine=256, batch=32, n_critic=5, ref_batch=4,
                 mode="train", label="cfg-4: larger domain 32x32->256x256 (8x), 7-ch input, F=32, 23 RRDB, batch 32 per GPU, n_critic=5"),
    "cfg5": dict(filters=64, channels=7, n_pred=2, rrdb=16, up=3, coarse=64, fine=512, batch=128, n_critic=5, ref_batch=2,
                 mode="infer", label="cfg-5: generator-only inference on full-domain tiles 64x64->512x512, 7-ch input, F=64 "
                                     "(filters = coarse edge, stage.py:60), 16 RRDB, batch 128 per GPU"),
}
CFG = dict(CONFIGS["cfg2"])


def model_flops(cfg):
    """Algorithmic conv+linear FLOPs per sample of one forward pass (2*MAC): (F_G, F_C)."""
    f, hc, hf = cfg["filters"], cfg["coarse"], cfg["fine"]
    conv = lambda ci, co, h: 2.0 * 9 * ci * co * h * h
    fg = conv(cfg["channels"], f, hc) + cfg["rrdb"] * 3 * sum(conv(k * f, f, hc) for k in range(1, 6)) + conv(f, f, hc)
    fg += sum(conv(f, 4 * f, hc << u) for u in range(cfg["up"])) + conv(f, f, hf) + conv(f, cfg["n_pred"], hf)
    w = cfg["coarse"]
    plan = [(cfg["n_pred"], w, 1), (w, w, 2), (w, 2 * w, 1), (2 * w, 2 * w, 2), (2 * w, 4 * w, 1), (4 * w, 4 * w, 2),
            (4 * w, 8 * w, 1), (8 * w, 8 * w, 2)]
    fc, h = 0.0, hf
    for ci, co, st in plan:
        h //= st
        fc += conv(ci, co, h)
    fc += 2.0 * (8 * w * h * h) * 100 + 2.0 * 100
    return fg, fc


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
    """Times `n_steps` steps of the schedule (cfg5: generator forwards) on the host with the oracle. Returns (seconds, samples)."""
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
    if CFG["mode"] == "infer":
        step = lambda: onet.generator_forward(g_sd, gspec, coarse)
    else:
        step = lambda: tr.batch(coarse, fine, alpha)
    with torch.no_grad() if CFG["mode"] == "infer" else torch.enable_grad():
        for _ in range(warmup):
            step()
        tr.num_steps = 0
        t0 = time.perf_counter()
        for _ in range(n_steps):
            step()
    return time.perf_counter() - t0, n_steps * batch


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    sample_b = CFG["ref_batch"]
    secs, samples = cpu_oracle_steps(args.steps, args.warmup, sample_b)
    val = samples / secs
    cores = os.cpu_count() or 1
    what = "generator forwards" if CFG["mode"] == "infer" else "steps of the schedule (critic every step, generator every 5th)"
    part = "the full per-GPU batch" if sample_b == CFG["batch"] else f"{sample_b} of each {CFG['batch']}-sample batch"
    sample = (f"{args.steps} {what} at batch {sample_b} ({part}), fp32, torch {torch.__version__} CPU, {cores} threads, "
              "one process (n_gpus echoes the launch)")
    out = {
        "impl": "reference", "metric": metric_name(), "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, 1),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def metric_name():
    return "generator_inference_samples_per_sec" if CFG["mode"] == "infer" else METRIC


def config_dict(args, world):
    return {"workload": CFG["label"], "name": args.config,
            "global_batch": CFG["batch"] * world, "per_gpu_batch": CFG["batch"],
            "parallelism": f"dp{world}" if CFG["mode"] == "train" else f"replicas{world}",
            "precision": args.precision,
            "l2_policy": "per-step working set (activations of one batch through both networks, >1 GB) exceeds the 126 MB L2; "
                         "8 distinct input batches are cycled"}


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default=os.environ.get("DOWNGAN_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    args = ap.parse_args()
    CFG.clear()
    CFG.update(CONFIGS[args.config])
    big = args.config in ("cfg4", "cfg5")
    if args.steps is None:
        args.steps = 20 if big else 200
    if args.warmup is None:
        args.warmup = 3 if big else 10
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
    infer = CFG["mode"] == "infer"

    B = CFG["batch"]
    torch.manual_seed(0)  # identical replicas on every rank
    Cn = Critic(CFG["coarse"], CFG["fine"], CFG["n_pred"], precision=args.precision).to(dev) if not infer else None
    Gn = Generator(CFG["filters"], CFG["fine"], CFG["channels"], CFG["n_pred"], CFG["rrdb"], CFG["up"],
                   precision=args.precision).to(dev)
    tr = None
    if not infer:
        gopt = torch.optim.Adam(Gn.parameters(), 2.5e-4, betas=(0.9, 0.99))
        copt = torch.optim.Adam(Cn.parameters(), 2.5e-4, betas=(0.9, 0.99))
        tr = WassersteinGAN(Gn, Cn, gopt, copt)

    NBATCH = 8 if not big else 4
    host = [synth_batch(B, CFG["channels"], CFG["coarse"], seed=1234 + 100 * rank + i, aseed=4321 + 100 * rank + i)
            for i in range(NBATCH)]
    if infer:
        host = [(c.pin_memory(),) for c, _f, _a in host]
    else:
        host = [(c.pin_memory(), f.pin_memory(), a.pin_memory()) for c, f, a in host]
    devb = [tuple(t.to(dev) for t in hb) for hb in host]
    h2d = sum(t.numel() * 4 for t in host[0])
    out_elems = B * CFG["n_pred"] * CFG["fine"] * CFG["fine"]
    d2h = out_elems * 4 if infer else 8 * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # host->device bandwidth of this box (pinned memory, one batch, best of 5): reported beside e2e so that a slow host
    # link can be told from a regression
    probe = torch.empty_like(devb[0][-1] if infer else devb[0][1])
    src = host[0][-1] if infer else host[0][1]
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        probe.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    h2d_gbs = src.numel() * 4 / best / 1e9
    del probe

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    if infer:
        # ---- cfg5: generator-only inference; a step = one forward pass over a resident 128-tile batch -------------
        out_dev = None

        def fwd(x):
            with torch.no_grad():
                return Gn(x)

        for i in range(args.warmup):
            out_dev = fwd(devb[i % NBATCH][0])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.dg_launch_count()
        t_w0 = time.time()
        e0.record()
        for s_ in range(args.steps):
            out_dev = fwd(devb[s_ % NBATCH][0])
        e1.record()
        barrier()
        sampler.mark(t_w0, time.time())
        ms, launches = max_over_ranks(e0.elapsed_time(e1)), lib.dg_launch_count() - l0
        # e2e: Generator.forward on pinned HOST tiles, output copied back to pinned host memory (what the reference's
        # tiled inference writer does with every chunk, helpers/gen_fake_ds.py:152-158)
        copy_stream = torch.cuda.Stream(device=dev)
        stage = [torch.empty_like(devb[0][0]) for _ in range(2)]
        out_host = [torch.empty(out_dev.shape, dtype=torch.float32, pin_memory=True) for _ in range(2)]
        main_stream = torch.cuda.current_stream()

        def e2e_loop(k):
            evs = [None, None]
            done = [None, None]
            with torch.cuda.stream(copy_stream):
                stage[0].copy_(host[0][0], non_blocking=True)
                evs[0] = torch.cuda.Event(); evs[0].record(copy_stream)
            for s_ in range(k):
                cur, nxt = s_ & 1, (s_ + 1) & 1
                if s_ + 1 < k:
                    with torch.cuda.stream(copy_stream):
                        if done[nxt] is not None:
                            copy_stream.wait_event(done[nxt])
                        stage[nxt].copy_(host[(s_ + 1) % NBATCH][0], non_blocking=True)
                        evs[nxt] = torch.cuda.Event(); evs[nxt].record(copy_stream)
                main_stream.wait_event(evs[cur])
                o = fwd(stage[cur])
                done[cur] = torch.cuda.Event(); done[cur].record(main_stream)
                out_host[cur].copy_(o, non_blocking=True)
            main_stream.synchronize()

        e2e_loop(2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_w0 = time.time()
        e0.record()
        e2e_loop(args.steps)
        e1.record()
        barrier()
        sampler.mark(t_w0, time.time())
        assert bool(torch.isfinite(out_host[0]).all())
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    else:
        def run(steps, batches, step0=0):
            for s_ in range(steps):
                c, f, a = batches[(step0 + s_) % NBATCH]
                tr._critic_train_iteration(c, f, a)
                if (step0 + s_) % CFG["n_critic"] == 0:
                    tr._generator_train_iteration(c, f)

        def timed(steps, batches):
            # the public epoch loop on DEVICE-resident batches (schedule + look-ahead generator forward included)
            tr.num_steps = 0
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = lib.dg_launch_count()
            e0.record()
            tr._train_epoch([batches[s_ % NBATCH] for s_ in range(steps)])
            e1.record()
            barrier()
            return max_over_ranks(e0.elapsed_time(e1)), lib.dg_launch_count() - l0

        # every staging slot, the look-ahead buffer, the pinned scalar ring and the native handles exist before any timed
        # region: the epochs below neither allocate nor synchronise inside their loops
        tr.prepare(host[0][0].shape, host[0][1].shape)
        run(args.warmup, devb)
        tr.num_steps = 0
        tr._train_epoch([devb[s_ % NBATCH] for s_ in range(max(args.warmup, 6))])
        barrier()
        t_w0 = time.time()
        ms, launches = timed(args.steps, devb)
        sampler.mark(t_w0, time.time())

        # ---- e2e: the public trainer API on HOST batches (wasserstein.py:120-147 `_train_epoch`):
        # every batch is copied host->device inside the timed region (pinned memory, side stream) and the
        # loss scalars of every step are copied device->host.
        def epoch_batches(k, step0):
            return [host[(step0 + s_) % NBATCH] for s_ in range(k)]

        tr.num_steps = 0
        tr._train_epoch(epoch_batches(max(args.warmup, 2 * CFG["n_critic"] + 2), 0))  # cycles through every staging slot
        tr.num_steps = 0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_w0 = time.time()
        e0.record()
        logs = tr._train_epoch(epoch_batches(args.steps, 0))
        e1.record()
        barrier()
        sampler.mark(t_w0, time.time())
        assert logs.shape == (args.steps, 8) and (bool(torch.isfinite(logs).all()) or os.environ.get("DG_ABLATE"))  # (ablation runs compute garbage)
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        enqueue_s = getattr(tr, "last_enqueue_seconds", None)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel-class CUDA-event pass (same steps, events around every launch) ---------
    prof = None
    if not args.no_profile:
        # kernels timed one at a time: the side-stream overlap (dg_set_tuning keys 9, 10, 12) is switched off for this pass,
        # otherwise an event pair around a launch also spans whatever runs concurrently on the other stream
        prev9, prev10, prev12 = lib.dg_set_tuning(9, 0), lib.dg_set_tuning(10, 0), lib.dg_set_tuning(12, 0)
        lib.dg_profile(1)
        if infer:
            for s_ in range(min(args.steps, 5)):
                fwd(devb[s_ % NBATCH][0])
        else:
            tr.num_steps = 0
            tr._train_epoch([devb[s_ % NBATCH] for s_ in range(args.steps)])
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
    f_g, f_c = model_flops(CFG)
    if infer:
        alg_flops_per_rank = B * args.steps * f_g
    else:
        n_gen = (args.steps + CFG["n_critic"] - 1) // CFG["n_critic"]
        # necessary work only: critic step F_G + 10 F_C, generator step 3 F_G + 2 F_C
        alg_flops_per_rank = B * (args.steps * (f_g + 10 * f_c) + n_gen * (3 * f_g + 2 * f_c))

    roofline = None
    if prof:
        conv = [k for k in prof if k in ("conv_direct", "wgrad_direct", "conv_tcgen05", "wgrad_tcgen05", "dense_block_tcgen05")]
        dom = max(conv, key=lambda k: prof[k]["ms"])
        d = prof[dom]
        ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
        traffic = None
        for name in ("traffic_r02.json", "traffic_r01.json"):
            tpath = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tpath) and args.config == "cfg2":  # dram__bytes_read+write per launch from the committed ncu --set full captures
                with open(tpath) as f:
                    traffic = json.load(f).get(dom, {}).get("bytes_per_launch")
                break
        # BASELINE.json's metric for the convolutions is tensor-pipe utilisation: `frac` is against the SUSTAINED bf16 tensor peak;
        # the class's algorithmic bytes against the HBM peak are kept beside it (the N = 16..32 layers sit below the 254 FLOP/B ridge)
        gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9
        frac_t, frac_h = ach / tf_sus, gbs / hbm
        roofline = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s", "frac": frac_t,
                    "frac_tensor": frac_t, "frac_hbm": frac_h, "achieved_tflops": ach, "achieved_gbs": gbs,
                    "traffic": traffic,
                    "peak_source": f"{which} (sustained cuBLAS bf16 / copy bandwidth from MEASURED_PEAKS.json)",
                    "launches": d["launches"], "avg_launch_us": 1e3 * d["ms"] / d["launches"],
                    "share_of_profiled_time": d["ms"] / sum(v["ms"] for v in prof.values()),
                    "whole_step_algorithmic_tflops": alg_flops_per_rank / (ms * 1e-3) / 1e12,
                    "whole_step_frac_tensor": alg_flops_per_rank / (ms * 1e-3) / 1e12 / tf_sus,
                    "classes": {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                                    "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 3) if v["flops"] else None,
                                    "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)} for k, v in prof.items()},
                    "hbm_peak_gbs": hbm}

    cpu = None
    if not args.no_cpu_baseline and world == 1:  # reported at N=1 only (the reference arm covers every N)
        rb = CFG["ref_batch"]
        n_cpu = 5 if not infer else 3
        secs, smp = cpu_oracle_steps(n_cpu, 2, rb)
        cores = os.cpu_count() or 1
        what = "generator forwards" if infer else "steps of the schedule (5 critic + 1 generator iterations)"
        cpu = {"value": smp / secs, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} {what} at batch {rb} after 2 warm-up steps, fp32 oracle, {cores} threads"}

    e2e_obj = {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": ms_e2e / args.steps, "h2d_gbs_probe": round(h2d_gbs, 2),
               "h2d_gbs_needed": round(h2d / (ms / args.steps * 1e-3) / 1e9, 2)}
    if not infer and enqueue_s is not None:
        e2e_obj["host_enqueue_ms_per_step"] = round(1e3 * enqueue_s / args.steps, 4)
    out = {
        "metric": metric_name(), "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": config_dict(args, world),
        "e2e": e2e_obj,
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "tcgen05": bool(lib.dg_has_tcgen05()),
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
