"""world_size-2 CPU (gloo) test of the multi-rank host logic: shard arithmetic, the flat-arena gradient exchange
(sum + 1/world in the optimiser = DDP's mean, template.py:243-244), per-epoch statistic reduction, rank-identical
model construction, the rank-0 BatchNorm-statistics broadcast of the task boundary.  The kernels themselves are covered
by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import wacil_b200 as w
    import wa_oracle as wo
    import argparse
    args = argparse.Namespace(dist_url="env://", seed=0)
    w.init_distributed_mode(args)                       # gloo on a CPU host
    assert args.distributed and args.world_size == world and args.dist_backend == "gloo"
    w.init_seed(args)                                   # same seed on every rank (template.py:199)
    model = w.CilModel("resnet20")
    model.prev_model_adaption(10)
    ref = model.arena.detach().clone()
    dist.broadcast(ref, src=0)
    same_init = torch.equal(ref, model.arena.detach())
    # shards of one epoch partition the dataset
    n = 7000
    idx = w.train.sampler_indices(n, epoch=3, rank=rank, world_size=world)
    gathered = [torch.zeros_like(idx) for _ in range(world)]
    dist.all_gather(gathered, idx)
    covered = torch.cat(gathered).unique().numel()
    # gradient exchange on the flat arena: per-rank oracle gradients of different shards -> mean
    sd = wo.make_state(20, [10], 3)
    gen = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(4, 3, 32, 32, generator=gen)
    y = torch.randint(0, 10, (4,), generator=gen)
    r = wo.train_step(dict(sd), None, x, y, 20, 0, 0.1, {}, apply_update=False)
    flat = torch.cat([r["grads"][k].reshape(-1) for k in sorted(r["grads"])])
    mine = flat.clone()
    scale = w.train.allreduce_grads_(flat, world)
    others = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(others, mine)
    mean_ok = torch.allclose(flat * scale, torch.stack(others).mean(0), rtol=1e-6, atol=1e-9)
    # per-epoch statistics
    acc = torch.tensor([1.0 + rank, 0.5, 1.5 + rank, 30.0 + rank])
    stats = w.train.reduce_epoch_stats(acc, steps=2, n_samples=200, world_size=world)
    # evaluation: DistributedSampler(shuffle=False) shards + the packed float64 meter reduced once (template.py:185)
    ev_idx = w.train.sampler_indices(101, 0, rank, world, shuffle=False)
    ev_all = [torch.zeros_like(ev_idx) for _ in range(world)]
    dist.all_gather(ev_all, ev_idx)
    ev_cover = (torch.cat(ev_all).unique().numel(), torch.cat(ev_all).numel())  # 101 samples, 102 with the padding duplicate
    meter = torch.tensor([51.0, 2.0, 1.0 + rank, 40.0 + rank, 50.0, 0, 0, 0], dtype=torch.float64)
    acc1, acc5, loss = w.train._finish_meter(meter, log=None)
    # DDP's buffer broadcast at the task boundary (template.py:243-244): ranks that trained on different shards hold
    # different running statistics; afterwards every rank holds rank 0's, and plans built on the old ones are stale
    with torch.no_grad():
        model.bn_running.add_(0.25 * (rank + 1))
    mine_bn = model.bn_running.clone()
    version = model._param_version
    w.train.broadcast_buffers(model, world)
    rank0_bn = mine_bn.clone()
    dist.broadcast(rank0_bn, src=0)
    bn_ok = torch.equal(model.bn_running, rank0_bn) and (rank == 0 or not torch.equal(mine_bn, rank0_bn)) \
        and model._param_version != version
    q.put((rank, same_init, covered, mean_ok, abs(scale - 1.0 / world) < 1e-12, stats, ev_cover, (acc1, acc5, loss), bn_ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_host_logic():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same_init, covered, mean_ok, scale_ok, stats, ev_cover, ev, bn_ok in results:
        assert same_init and covered == 7000 and mean_ok and scale_ok and bn_ok
        assert abs(stats["ce"] - (1.0 + 2.0) / 4) < 1e-9 and abs(stats["acc1"] - 100 * 61 / 200) < 1e-9
        assert ev_cover == (101, 102)
        # every rank ends with the same global averages: 81 top-1 hits of 102 samples, loss 3 / 4 batches
        assert abs(ev[0] - 100 * 81 / 102) < 1e-9 and abs(ev[1] - 100 * 100 / 102) < 1e-9 and abs(ev[2] - 0.75) < 1e-12
