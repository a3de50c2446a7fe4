"""Is a CUDA graph worth building?  One critic iteration of the cfg-2 schedule (fake taken from a resident look-ahead buffer,
fused step + Adam + re-pack) timed eagerly and as a captured graph replayed on the same inputs.  The Adam step count is baked
into the captured launch, so this is a timing probe only."""
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
B = 64
batches = [tuple(t.to(dev) for t in synth_batch(B, 2, 16, seed=i)) for i in range(8)]
tr.prepare(batches[0][0].shape, batches[0][1].shape)
tr._train_epoch([batches[i % 8] for i in range(11)])           # warm-up incl. look-ahead passes: handles sized, kernels initialised
coarse_all = torch.cat([batches[i][0] for i in range(5)])
tr._generator_lookahead(coarse_all, 0)                           # fakes of five batches resident at offsets 0, 64, ...
torch.cuda.synchronize()
c, f, a = batches[0]
a = a.reshape(B).contiguous()

def step():
    tr._critic_train_iteration(c, f, a, _fake_offset=0)

def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, 1e3 * t_host / n

for _ in range(10):
    step()
print("eager   : %.4f ms per critic iteration (host enqueue %.3f ms)" % timed(step, 200), flush=True)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    step()
torch.cuda.synchronize()
print("captured one iteration", flush=True)
for _ in range(10):
    g.replay()
print("graph   : %.4f ms per critic iteration (host enqueue %.3f ms)" % timed(g.replay, 200), flush=True)
print("eager   : %.4f ms per critic iteration (host enqueue %.3f ms)" % timed(step, 200), flush=True)
print("graph   : %.4f ms per critic iteration (host enqueue %.3f ms)" % timed(g.replay, 200), flush=True)
os._exit(0)
