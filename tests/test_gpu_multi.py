"""Two-GPU tests: the fused NVLink all-reduce + SGD step (wacil_allreduce_sgd) against NCCL all-reduce + SGD, and
template.main under torch.distributed — every rank must finish each task with the same BatchNorm statistics, the same
teacher and the same rehearsal memory (what DDP's buffer broadcast gives the reference, template.py:243-244).
Needs >= 2 CUDA devices (skipped on a 1-GPU box; run with `gpurun --gpus 2`)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import wa_oracle as wo
    import wacil_b200 as w
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    sd, tsd = wo.make_state(20, [10, 10], 5), wo.make_state(20, [10], 6)
    gen = torch.Generator().manual_seed(100 + rank)  # every rank trains on its own shard
    xs = [torch.randn(16, 3, 32, 32, generator=gen).cuda() for _ in range(4)]
    ys = [torch.randint(0, 20, (16,), generator=gen).cuda() for _ in range(4)]
    out = {}
    for fused in (True, False):
        m = w.CilModel("resnet20"); m.load_state_dict(sd); m = m.cuda().train()
        t = w.CilModel("resnet20"); t.load_state_dict(tsd); t = t.cuda().copy().freeze()
        tr = w.TaskTrainer(m, t, 10, lr=0.05, world_size=world, fused_allreduce=fused, use_graph=True)
        assert (tr.peer is not None) == fused, "symmetric memory must be available on an NVLink box"
        first = None
        for i, (x, y) in enumerate(zip(xs, ys)):
            tr.step(x, y)
            if i == 0:
                torch.cuda.synchronize()
                first = m.arena.detach().clone().cpu()
        torch.cuda.synchronize()
        p = m.arena.detach().clone()
        others = [torch.zeros_like(p) for _ in range(world)]
        dist.all_gather(others, p)
        out[fused] = (p.cpu(), all(torch.equal(o, others[0]) for o in others), first)
        dist.barrier()
    rel = lambda a, b: ((a - b).double().norm() / b.double().norm()).item()
    # the kernel alone, on identical inputs: fused peer-memory mean + SGD vs NCCL all-reduce + wacil_sgd_step
    import ctypes
    import torch.distributed._symmetric_memory as symm_mem
    n = 100000
    g = symm_mem.empty(n, dtype=torch.float32, device="cuda")
    pad = symm_mem.empty(64, dtype=torch.int32, device="cuda")
    pad.zero_()
    hg, hp = symm_mem.rendezvous(g, dist.group.WORLD), symm_mem.rendezvous(pad, dist.group.WORLD)
    gp = (ctypes.c_void_p * world)(*[int(v) for v in hg.buffer_ptrs])
    pp = (ctypes.c_void_p * world)(*[int(v) for v in hp.buffer_ptrs])
    epoch = torch.zeros(2, dtype=torch.int32, device="cuda")
    gen2 = torch.Generator().manual_seed(7)
    p0 = torch.randn(n, generator=gen2).cuda()
    pa, va, pb, vb = p0.clone(), torch.zeros(n, device="cuda"), p0.clone(), torch.zeros(n, device="cuda")
    lr = torch.full((1,), 0.1, device="cuda")
    kerr = 0.0
    for it in range(3):
        g.copy_(torch.randn(n, generator=torch.Generator().manual_seed(1000 * rank + it)).cuda())
        ref = g.clone()
        torch.cuda.synchronize(); dist.barrier()
        w.ops.allreduce_sgd_(pa, va, gp, pp, world, rank, lr, epoch)
        dist.all_reduce(ref)
        w.ops.sgd_step_(pb, ref, vb, lr, 0.9, 5e-4, 1.0 / world)
        torch.cuda.synchronize(); dist.barrier()
        kerr = max(kerr, (pa - pb).abs().max().item())
    q.put((rank, out[True][1], out[False][1], rel(out[True][2], out[False][2]), rel(out[True][0], out[False][0]), kerr))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_fused_p2p_allreduce_sgd_matches_nccl():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, fused_identical, nccl_identical, rel1, rel4, kerr in results:
        assert fused_identical, "fused path: replicas must be bit-identical (fixed rank-order sums)"
        assert kerr < 2e-6, kerr  # the kernel on identical inputs: fp32 summation order only
        # whole training runs: NCCL's ring order vs the kernel's rank order differ in the last bits of the gradient mean
        # and the bf16 forward amplifies that step by step; everything else is bit-reproducible
        assert rel1 < 1e-4, rel1
        assert rel4 < 1e-3, rel4


def _main_worker(rank, world, port, q, tmp):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import argparse
    import torch.distributed as dist
    import wacil_b200 as w
    p = argparse.ArgumentParser(parents=[w.get_args_parser()])
    w.train.add_runner_args(p)
    args = p.parse_args(["--data_set", "synthetic", "--nb_classes", "6", "--num_bases", "2", "--increment", "2",
                         "--backbone", "resnet20", "--batch_size", "16", "--num_epochs", "10", "--lr", "0.05",
                         "--memory_size", "40", "--synthetic_per_class", "96", "--synthetic_val_per_class", "32",
                         "--eval_every_epoch", "5", "--max_tasks", "2", "--checkpoint_dir", os.path.join(tmp, "ck")])
    captured = {}
    orig_add = w.RehearsalMemory.add

    def add(self, x, y, t, z):  # what each rank herds from, and what it keeps
        captured.setdefault("feats", []).append(z.detach().clone())
        orig_add(self, x, y, t, z)
        captured["mem"] = tuple(a.clone() for a in self.get())

    orig_copy = w.CilModel.copy

    def copy(self):  # the teacher snapshot of template.py:290
        new = orig_copy(self)
        captured["teacher"] = (new.arena.detach().clone(), new.bn_running.clone())
        return new

    w.RehearsalMemory.add, w.CilModel.copy = add, copy
    acc = w.template.main(args, log=lambda *a, **k: None)

    def same(tensor):
        parts = [torch.zeros_like(tensor) for _ in range(world)]
        dist.all_gather(parts, tensor.contiguous())
        return all(torch.equal(parts[0], o) for o in parts)

    ok = {"teacher_params": same(captured["teacher"][0]), "teacher_bn": same(captured["teacher"][1]),
          "features": all(same(f) for f in captured["feats"]),
          "memory_x": same(captured["mem"][0]), "memory_y": same(captured["mem"][1])}
    q.put((rank, ok, acc, int(captured["mem"][1].numel())))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_template_main_keeps_ranks_identical(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_main_worker, args=(r, world, port, q, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    accs = [r[2] for r in results]
    for rank, ok, acc, nmem in results:
        assert all(ok.values()), (rank, ok)   # teacher, herding features and exemplar memory identical on every rank
        assert len(acc) == 2 and nmem > 0
    assert accs[0] == accs[1]  # the meter is all-reduced: same numbers everywhere
    # two separable classes, global batch 32 (the single-GPU run of this configuration reaches 100 %; at global batch 64
    # the 30 steps are too few: 53 %, tools/synthetic_run.py): far above the 50 % of chance
    assert accs[0][0] > 60.0, accs
