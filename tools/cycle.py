"""One steady-state 5-step cycle of the cfg-2 schedule for ncu (`--profile-from-start off`): six warm-up steps,
then steps 6..10 (five critic iterations, look-ahead generator forward, one generator iteration) between
cudaProfilerStart/Stop."""
import os, sys
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
batches = [tuple(t.to(dev) for t in synth_batch(64, 2, 16, seed=i)) for i in range(8)]
tr._train_epoch([batches[i % 8] for i in range(6)])
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr._train_epoch([batches[(6 + i) % 8] for i in range(5)])   # num_steps continues at 6: steps 6..10
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", tr.num_steps)
os._exit(0)
