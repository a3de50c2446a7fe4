"""2-GPU check of the all-reduce overlap (torchrun --nproc-per-node 2): the critic parameters after a few iterations with the
classifier bucket reduced early (dg_critic_defer_conv_grads / dg_critic_step_finish) against one all-reduce per iteration,
and against a second run of the latter (the floor set by the fp32 atomics of the weight-gradient kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from downgan_b200.GAN.wasserstein import WassersteinGAN
from downgan_b200.networks import Critic, Generator
from downgan_b200.synthetic import synth_batch

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
batches = [tuple(t.to(dev) for t in synth_batch(64, 2, 16, seed=100 * rank + i, aseed=7 + 100 * rank + i)) for i in range(6)]


def run(overlap):
    torch.manual_seed(0)
    C = Critic(16, 128, 2).to(dev); G = Generator(16, 128, 2, 2).to(dev)
    tr = WassersteinGAN(G, C, torch.optim.Adam(G.parameters(), 2.5e-4, betas=(0.9, 0.99)),
                        torch.optim.Adam(C.parameters(), 2.5e-4, betas=(0.9, 0.99)))
    tr.overlap_allreduce = overlap
    p0 = C.flat_params().clone()
    tr._train_epoch(batches)
    torch.cuda.synchronize()
    return p0, C.flat_params().clone(), G.flat_params().clone()


p0, c_off, g_off = run(False)
_, c_off2, g_off2 = run(False)
_, c_on, g_on = run(True)
den = float((c_off - p0).norm())
msg = (f"rank {rank}: critic |on-off|/|update| = {float((c_on - c_off).norm()) / den:.3e}, "
       f"|off-off'|/|update| = {float((c_off2 - c_off).norm()) / den:.3e}, generator |on-off| rel = "
       f"{float((g_on - g_off).norm() / g_off.norm()):.3e}")
# identical replicas: every rank must hold the same parameters
t = c_on.clone(); dist.all_reduce(t, op=dist.ReduceOp.MAX); same = bool((t == c_on).all())
print(msg, "| replicas identical:", same, flush=True)
dist.destroy_process_group()
