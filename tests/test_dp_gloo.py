"""world_size-2 gloo test (CPU) of the data-parallel host logic: sharding the batch over ranks
and sum-all-reducing the flat gradient buckets (scaled 1/world) reproduces the global-batch
gradients of both iterations, including the per-sample gradient penalty."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from downgan_b200 import dp
from downgan_b200.synthetic import synth_batch
from oracle import networks as onet
from oracle import trainer as otr

GS = onet.GeneratorSpec(filters=4, channels=2, n_predictands=2, num_res_blocks=1, num_upsample=2)
CS = onet.CriticSpec(coarse_dim=4, fine_dim=16, nc=2)


def _flat(d):
    return torch.cat([v.reshape(-1) for v in d.values()])


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    torch.manual_seed(0)
    c_sd = onet.init_critic_state(CS, dtype=torch.float64)
    g_sd = onet.init_generator_state(GS, dtype=torch.float64)
    coarse, fine, alpha = synth_batch(8, 2, 4, up=4)
    coarse, fine, alpha = coarse.double(), fine.double(), alpha.double()
    hp = otr.Hyper()
    oc = otr.critic_loss_and_grads(g_sd, GS, c_sd, CS, dp.shard(coarse), dp.shard(fine), dp.shard(alpha), hp)
    og = otr.generator_loss_and_grads(g_sd, GS, c_sd, CS, dp.shard(coarse), dp.shard(fine), hp)
    cg, gg = _flat(oc["grads"]).clone(), _flat(og["grads"]).clone()
    s1 = dp.allreduce_sum_(cg)
    s2 = dp.allreduce_sum_(gg)
    assert s1 == s2 == 1.0 / world
    if rank == 0:
        full_c = otr.critic_loss_and_grads(g_sd, GS, c_sd, CS, coarse, fine, alpha, hp)
        full_g = otr.generator_loss_and_grads(g_sd, GS, c_sd, CS, coarse, fine, hp)
        ec = float((cg * s1 - _flat(full_c["grads"])).norm() / _flat(full_c["grads"]).norm())
        eg = float((gg * s2 - _flat(full_g["grads"])).norm() / _flat(full_g["grads"]).norm())
        torch.save({"ec": ec, "eg": eg, "world": dp.world_size()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradients_equal_global_batch(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["world"] == 2
    assert r["ec"] < 1e-12 and r["eg"] < 1e-12, r


def test_shard_rejects_ragged_batch():
    import pytest
    with pytest.raises(ValueError):
        dp.shard(torch.zeros(5, 1), 0, 2)
    assert dp.shard(torch.arange(8).reshape(8, 1), 1, 2).flatten().tolist() == [4, 5, 6, 7]
    assert dp.allreduce_sum_(torch.ones(3)) == 1.0
