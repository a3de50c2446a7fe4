"""Where does the e2e gap come from?  Times _train_epoch on device batches vs pinned host batches."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from downgan_b200.GAN.wasserstein import WassersteinGAN
from downgan_b200.networks import Critic, Generator
from downgan_b200.synthetic import synth_batch
dev = torch.device("cuda:0")
torch.manual_seed(0)
C = Critic(16, 128, 2).to(dev); G = Generator(16, 128, 2, 2).to(dev)
tr = WassersteinGAN(G, C, torch.optim.Adam(G.parameters(), 2.5e-4, betas=(0.9, 0.99)), torch.optim.Adam(C.parameters(), 2.5e-4, betas=(0.9, 0.99)))
host = [tuple(t.pin_memory() for t in synth_batch(64, 2, 16, seed=i)) for i in range(8)]
devb = [tuple(t.to(dev) for t in b) for b in host]
def run(batches, k, label):
    tr.num_steps = 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tr._train_epoch([batches[i % 8] for i in range(k)]); e1.record(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"{label}: events {e0.elapsed_time(e1)/k:.3f} ms/step, wall {(t1-t0)*1e3/k:.3f} ms/step, "
          f"host enqueue {tr.last_enqueue_seconds*1e3/k:.3f} ms/step", flush=True)
def run_plain(k, label):
    tr.num_steps = 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for s in range(k):
        c, f, a = devb[s % 8]
        tr._critic_train_iteration(c, f, a)
        if s % 5 == 0: tr._generator_train_iteration(c, f)
    t_cpu = time.perf_counter()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{label}: wall {(t1-t0)*1e3/k:.3f} ms/step, cpu-enqueue {(t_cpu-t0)*1e3/k:.3f} ms/step", flush=True)
# mixed: `fine` (98 % of the bytes) already on the device, coarse / alpha on the host -> the whole staging machinery (copy stream,
# events, slot ring) runs, but almost nothing crosses PCIe: separates DMA interference from the machinery's own cost
mixed = [(h[0], d[1], h[2]) for h, d in zip(host, devb)]
for _ in range(2):
    run(devb, 200, "epoch(device)"); run(host, 200, "epoch(host)"); run(mixed, 200, "epoch(mixed: fine resident)")
    run(host, 12, "epoch(host, 12 steps: host enqueue time is not throttled by a full launch queue)")
    run_plain(20, "plain(device)")
os._exit(0)
