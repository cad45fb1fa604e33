#!/usr/bin/env python
"""bench.py — images/sec of the WA class-incremental training step (BASELINE.json metric) on N B200s.

  python bench.py --gpus 1 --steps 50 --warmup 10
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      (the reference's CPU path: the oracle port on the host cores)

Workload (BASELINE.json configs[1]): ResNet-32, CIFAR-100 10x10 protocol at task 1 (20 seen / 10 known classes),
per-GPU batch 128 of synthetic 32x32x3 images, one step = student forward+backward + frozen-teacher forward +
CE + 0.5*KD(T=2) + gradient all-reduce (N>1) + SGD(momentum 0.9, wd 5e-4).  bf16 tensor-core convolutions, fp32
master weights / BatchNorm statistics / loss.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

DEPTH, BATCH, HEADS, KNOWN = 32, 128, [10, 10], 10
F_FWD = 137.72e6                 # conv FLOPs (2*MAC) of one ResNet-32 forward on one image
F_STEP = 4 * F_FWD               # student fwd + dgrad + wgrad + teacher fwd
POOL = 64                        # distinct synthetic batches the inputs rotate over (100 MB on the device)
WORKLOAD = "resnet32-cifar100-10x10-task1-b128"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while a timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self._stop, self._t = gpu_index, [], threading.Event(), None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.idx)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(float(r[0])) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(float(self.rows[0][1])),
                "reasons": reasons, "samples": len(self.rows)}


def make_states(heads=None):
    import wa_oracle as wo
    heads = HEADS if heads is None else heads
    return wo.make_state(DEPTH, heads, 0, perturb_bn=False), wo.make_state(DEPTH, heads[:-1], 1, perturb_bn=False)


def synthetic_pool(n_batches, seed, n_classes=None):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_batches, BATCH, 3, 32, 32, generator=g)
    y = torch.randint(0, n_classes or sum(HEADS), (n_batches, BATCH), generator=g)
    return x, y


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_steps(n_steps, warmup, budget_s=None):
    """The reference's step on the host cores: the oracle port of template.py:255-271 (fp32, torch CPU, all threads),
    one 128-image batch per step.  Returns (img/s, steps run, threads)."""
    import wa_oracle as wo
    cores = os.cpu_count() or 1
    sd, tsd = make_states()
    x, y = synthetic_pool(4, 1234)
    bufs = {}
    # "all the host threads it can use": torch's intra-op pool stops scaling (and collapses under oversubscription on a
    # shared 100+-core host) long before it runs out of cores on these tiny layers, so time one step at a few pool
    # sizes, stop as soon as a larger pool is slower, and report the size actually used
    best, threads = None, min(cores, 8)
    for cand in sorted({min(cores, c) for c in (8, 16, 32, 64)}):
        torch.set_num_threads(cand)
        t0 = time.perf_counter()
        wo.train_step(sd, tsd, x[1], y[1], DEPTH, KNOWN, 0.1, bufs)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, threads = dt, cand
        elif dt > 1.2 * best:
            break
    torch.set_num_threads(threads)
    for i in range(warmup):
        wo.train_step(sd, tsd, x[i % 4], y[i % 4], DEPTH, KNOWN, 0.1, bufs)
    t0 = time.perf_counter()
    done = 0
    for i in range(n_steps):
        wo.train_step(sd, tsd, x[i % 4], y[i % 4], DEPTH, KNOWN, 0.1, bufs)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done * BATCH / dt, done, threads, dt


def _gloo_worker(rank, k, port, threads, n_steps, q):
    """One gloo rank of the reference's DDP step on the CPU (template.py:258-272: forward, losses, backward, gradient
    all-reduce, SGD, per-step barrier), through the oracle port."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(k))
    import wa_oracle as wo
    torch.set_num_threads(threads)
    dist.init_process_group("gloo", rank=rank, world_size=k)
    sd, tsd = make_states()
    x, y = synthetic_pool(2, 1234 + rank)
    names = wo.trainable_names(sd)
    bufs = {}

    def step(i):
        r = wo.train_step(sd, tsd, x[i % 2], y[i % 2], DEPTH, KNOWN, 0.1, bufs, apply_update=False)
        flat = torch.cat([r["grads"][n].reshape(-1) for n in names])
        dist.all_reduce(flat)  # DDP's gradient mean (template.py:243-244)
        flat /= k
        o = 0
        grads = {}
        for n in names:
            m = r["grads"][n].numel()
            grads[n] = flat[o: o + m].view_as(r["grads"][n])
            o += m
        wo.sgd_step({n: sd[n] for n in names}, grads, bufs, 0.1, 0.9, 5e-4)
        dist.barrier()  # template.py:272

    step(0)
    t0 = time.perf_counter()
    for i in range(n_steps):
        step(i)
    dt = time.perf_counter() - t0
    if rank == 0:
        q.put(dt)
    dist.destroy_process_group()


def cpu_gloo_baseline(ks=(1, 2, 4), n_steps=5):
    """BASELINE.md B1 / SURVEY 8(d): the reference step under torch.distributed gloo with k CPU ranks on this box's cores."""
    import torch.multiprocessing as mp
    cores = os.cpu_count() or 1
    out = {"os_cpu_count": cores, "steps": n_steps, "runs": []}
    ctx = mp.get_context("spawn")
    for k in ks:
        threads = max(1, min(16, cores // k))
        q = ctx.Queue()
        port = 29800 + (os.getpid() % 500) + k
        procs = [ctx.Process(target=_gloo_worker, args=(r, k, port, threads, n_steps, q)) for r in range(k)]
        for p_ in procs:
            p_.start()
        try:
            dt = q.get(timeout=300)
            out["runs"].append({"ranks": k, "threads_per_rank": threads, "img_s": k * n_steps * BATCH / dt,
                                "ms_per_step": 1e3 * dt / n_steps})
        except Exception as ex:  # a baseline must never take the bench line down
            out["runs"].append({"ranks": k, "error": str(ex)[:120]})
        for p_ in procs:
            p_.join(timeout=30)
    return out


def gpu_eager_steps(n_steps, autocast):
    """the reference's step sequence as PyTorch eager on this GPU (cuDNN/cuBLAS/ATen), through the oracle
    port of template.py:255-271 with CUDA tensors — the incumbent GPU path next to which the native kernels are read.
    fp32 like the reference, or under bf16 autocast.  Returns img/s."""
    import wa_oracle as wo
    sd, tsd = make_states()
    sd = {k: v.cuda() for k, v in sd.items()}
    tsd = {k: v.cuda() for k, v in tsd.items()}
    x, y = synthetic_pool(4, 1234)
    x, y = x.cuda(), y.cuda()
    bufs = {}

    def step(i):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            wo.train_step(sd, tsd, x[i % 4], y[i % 4], DEPTH, KNOWN, 0.05, bufs)

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n_steps):
        step(i)
    torch.cuda.synchronize()
    return n_steps * BATCH / (time.perf_counter() - t0)


def run_reference(args, rank, world):
    if rank != 0:
        return
    ips, done, threads, dt = cpu_steps(args.steps, max(args.warmup, 1))
    line = {"impl": "reference", "metric": "images/sec per CIL training step", "value": ips, "unit": "img/s",
            "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / done,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH,
                       "note": "the CPU arm is ONE process on rank 0 stepping one rank's 128-image batch whatever N is: at "
                               "N > 1 the driver's ratio compares N GPUs with one CPU process (the gloo k-rank CPU runs are in the "
                               "N = 1 line of the GPU arm: cpu_gloo)"},
            "cpu_baseline": {"value": ips, "unit": "img/s", "cores": threads, "kind": "port",
                             "sample": f"{done} steps of one 128-image batch, oracle/wa_oracle.train_step"},
            "e2e": {"value": ips, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def timed(fn, steps, world):
    """barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    return ms.item()


def kernel_rooflines(w, pk):
    """Isolated timings of the dominant conv kernel and of the KD kernel, CUDA events over rotating buffers whose
    total footprint exceeds L2 (so every launch streams from HBM)."""
    out = {}
    dev = "cuda"
    # The dominant kernel of the step: the fused stage-1 layer (3x3 conv 16->16 over 128x32x32, batch statistics, grid
    # barrier, BatchNorm + identity shortcut + ReLU) — 22 such launches per step (student forward + data gradients).
    nbuf = 12
    xs = [torch.randn(BATCH, 32, 32, 16, device=dev).bfloat16() for _ in range(nbuf)]
    rs = [torch.randn(BATCH, 32, 32, 16, device=dev).bfloat16() for _ in range(nbuf)]
    wt = (torch.randn(9, 16, 16, device=dev) * 0.1).bfloat16()
    gamma, beta = torch.ones(16, device=dev), torch.zeros(16, device=dev)
    rm, rv = torch.zeros(16, device=dev), torch.ones(16, device=dev)
    ys = torch.empty((nbuf, BATCH, 32, 32, 16), dtype=torch.bfloat16, device=dev)
    acts = torch.empty((nbuf, BATCH, 32, 32, 16), dtype=torch.bfloat16, device=dev)
    saved = torch.empty(32, device=dev)
    lib, st = w._lib.lib(), torch.cuda.current_stream().cuda_stream
    wsb = torch.zeros(lib.wacil_conv3x3_bn_relu_workspace_bytes(16), dtype=torch.uint8, device=dev)

    def fused(i):
        k = i % nbuf
        rc = lib.wacil_conv3x3_bn_relu(xs[k].data_ptr(), wt.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rs[k].data_ptr(), 1,
                                       ys[k].data_ptr(), acts[k].data_ptr(), saved.data_ptr(), rm.data_ptr(), rv.data_ptr(),
                                       BATCH, 32, 32, 16, 16, 1, wsb.data_ptr(), st)
        assert rc == 0, lib.wacil_last_error()

    for i in range(5):
        fused(i)
    torch.cuda.synchronize()
    reps = 120  # 12 x 16.8 MB rotating working set (201 MB) exceeds the 126 MB L2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fused(i)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    bytes_alg = 4 * BATCH * 32 * 32 * 16 * 2 + 9 * 16 * 16 * 2 + 4 * 16 * 4  # x + shortcut in, y + a out (bf16), weights, stats
    flops = 2 * BATCH * 32 * 32 * 16 * 16 * 9
    out["conv"] = {"kernel": "conv_gemm_kernel_bn_fwd (fused stage-1 layer 128x32x32x16->16: conv + batch stats + "
                             "exchange of the per-CTA sums + BN + shortcut + ReLU, one launch)", "us": us,
                   "gbs": bytes_alg / us / 1e3, "tflops": flops / us / 1e6, "bytes": bytes_alg}
    del rs, acts
    del xs, ys
    # KD kernel (BASELINE metric 2): B=1024, C=8192 fp32, 3*B*C*4 bytes
    B, C = 1024, 8192
    nb = 6
    ss = [torch.randn(B, C, device=dev) for _ in range(nb)]
    ts = [torch.randn(B, C, device=dev) for _ in range(nb)]
    gs = [torch.empty(B, C, device=dev) for _ in range(nb)]
    for i in range(3):
        w.ops.kd_fwd_bwd(ss[i], ts[i], 2.0, grad_out=gs[i])
    torch.cuda.synchronize()
    loss = torch.empty((), device=dev)
    ws = torch.zeros(lib.wacil_kd_workspace_bytes(B), dtype=torch.uint8, device=dev)
    reps = 60
    e0.record()
    for i in range(reps):
        lib.wacil_kd_fwd_bwd(ss[i % nb].data_ptr(), C, ts[i % nb].data_ptr(), C, B, C, 2.0, 1.0, loss.data_ptr(),
                             gs[i % nb].data_ptr(), C, 0, 0, ws.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    bytes_alg = 3 * B * C * 4 + 4 * B
    out["kd"] = {"kernel": "kd_stream_kernel fp32 B=1024 C=8192 T=2 (TMA-streamed, 60 back-to-back launches over 6 "
                           "rotating buffer sets, PDL)", "us": us, "gbs": bytes_alg / us / 1e3,
                 "frac": bytes_alg / us / 1e3 / pk["hbm_gbs"], "bytes": bytes_alg}
    del ss, ts, gs
    # weight-align kernel (north star): [8192, 2048] fp32, 1024 new rows: reads every row once, writes the new rows
    Cw, Dw, nn = 8192, 2048, 1024
    nbw = 5  # 5 x 67 MB > L2
    Ws = [torch.randn(Cw, Dw, device=dev) * 0.1 for _ in range(nbw)]
    gam_out = torch.empty((), device=dev)
    wws = torch.zeros(lib.wacil_weight_align_workspace_bytes(Cw), dtype=torch.uint8, device=dev)
    for i in range(3):
        lib.wacil_weight_align(Ws[i].data_ptr(), Cw, Dw, nn, gam_out.data_ptr(), 0, wws.data_ptr(), st)
    torch.cuda.synchronize()
    reps = 40
    e0.record()
    for i in range(reps):
        lib.wacil_weight_align(Ws[i % nbw].data_ptr(), Cw, Dw, nn, gam_out.data_ptr(), 0, wws.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    bytes_alg = (Cw + nn) * Dw * 4
    out["wa"] = {"kernel": "weight_align_kernel fp32 [8192,2048] n_new=1024 (persistent, new rows parked in shared memory, "
                           "40 launches over 5 rotating matrices)", "us": us, "gbs": bytes_alg / us / 1e3,
                 "frac": bytes_alg / us / 1e3 / pk["hbm_gbs"], "bytes": bytes_alg}
    return out


def ncu_traffic(csv_name, kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum of the first matching kernel in a committed `ncu --page raw --csv`
    export under profiles/ (the capture is taken with the same entry point, see profiles/README.md); None if absent."""
    import csv
    path = os.path.join(ROOT, "profiles", csv_name)
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return None
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        if kernel_substr in r[ix["Kernel Name"]]:
            tot = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[ix[key]]) * mult.get(units[ix[key]], 1)
            return int(tot)
    return None


def per_task_block(w, model, teacher):
    """The per-task path (not per step): feature extraction for herding (template.py:292-299), herding itself
    (template.py:300-302), eval() (template.py:169-188) and weight align in situ (template.py:156-166)."""
    out = {}
    dev = "cuda"
    n = 7000  # a task set of the 10x10 protocol: 5000 new + 2000 exemplars
    x = torch.randn(n, 3, 32, 32, device=dev)
    for _ in range(2):
        f = w.extract_features(model, x, BATCH)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    f = w.extract_features(model, x, BATCH)
    torch.cuda.synchronize()
    out["feature_extraction_img_s"] = n / (time.perf_counter() - t0)
    model.train()
    import numpy as np
    for name, (k, per, quota) in {"herding_20x500x64_ms": (20, 500, 100), "herding_100x70x64_ms": (100, 70, 20)}.items():
        feats = torch.relu(torch.randn(k * per, 64, device=dev) * 0.3 + 0.5)
        yy = torch.arange(k, device=dev).repeat_interleave(per)
        w.rehearsal.herd_indices(feats, yy, quota)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        w.rehearsal.herd_indices(feats, yy, quota)
        torch.cuda.synchronize()
        out[name] = 1e3 * (time.perf_counter() - t0)
    xv = torch.randn(2000, 3, 32, 32, device=dev)
    yv = torch.randint(0, sum(HEADS), (2000,), device=dev)
    loader = [(xv[i: i + BATCH], yv[i: i + BATCH], None) for i in range(0, 2000, BATCH)]
    w.eval(model, loader)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    w.eval(model, loader)
    torch.cuda.synchronize()
    out["eval_img_s"] = 2000 / (time.perf_counter() - t0)
    model.train()
    m2 = model.copy()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m2.weight_align(HEADS[-1], verbose=False)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        m2.weight_align(HEADS[-1], verbose=False)
    e1.record()
    torch.cuda.synchronize()
    out["weight_align_in_situ_us"] = 1e3 * e0.elapsed_time(e1) / 20
    out["note"] = ("ResNet-32, 20 seen classes; features / eval through the folded-BN inference plan at batch 128, host timer "
                   "around whole calls (one sync each); weight align on the [20,64] classifier (launch-latency bound)")
    return out


def task9_point(w, world, rank, steps):
    """The same step at task 9 of the protocol (100 seen / 90 known classes)."""
    heads = [10] * 10
    sd, tsd = make_states(heads)
    model = w.CilModel(f"resnet{DEPTH}")
    model.load_state_dict(sd)
    model = model.cuda().train()
    teacher = w.CilModel(f"resnet{DEPTH}")
    teacher.load_state_dict(tsd)
    teacher = teacher.cuda().copy().freeze()
    tr = w.TaskTrainer(model, teacher, 90, lr=0.05, lambda_kd=0.5, T=2.0, world_size=world, use_graph=True)
    x, y = synthetic_pool(16, 4321 + rank, n_classes=100)
    x, y = x.cuda(), y.cuda()
    for i in range(30):
        tr.step(x[i % 16], y[i % 16])
    ms = timed(lambda i: tr.step(x[i % 16], y[i % 16]), steps, world)
    return {"workload": "resnet32-cifar100-10x10-task9-b128", "classes_seen": 100, "classes_known": 90, "steps": steps,
            "ms_per_step": ms / steps, "img_s": steps * BATCH * world / (ms / 1e3)}



def _dbg(msg):
    if os.environ.get("BENCH_DEBUG"):
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def run_ours(args, rank, local_rank, world):
    import wacil_b200 as w
    torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pk = peaks()
    sd, tsd = make_states()
    model = w.CilModel(f"resnet{DEPTH}")
    model.load_state_dict(sd)
    model = model.cuda().train()
    teacher = w.CilModel(f"resnet{DEPTH}")
    teacher.load_state_dict(tsd)
    teacher = teacher.cuda().copy().freeze()
    tr = w.TaskTrainer(model, teacher, KNOWN, lr=0.05, lambda_kd=0.5, T=2.0, world_size=world, use_graph=True)
    x_host, y_host = synthetic_pool(POOL, 1234 + rank)
    x_host, y_host = x_host.pin_memory(), y_host.pin_memory()
    x_dev, y_dev = x_host.cuda(), y_host.cuda()
    loss_host = torch.zeros((max(args.steps, 1), 4), dtype=torch.float32).pin_memory()

    def step_dev(i):
        tr.step(x_dev[i % POOL], y_dev[i % POOL])

    def step_e2e(i):
        # the public call a user makes, host buffers in, the step's losses out (async D2H into pinned memory)
        losses = tr.step(x_host[i % POOL], y_host[i % POOL])
        loss_host[i % loss_host.shape[0]].copy_(losses, non_blocking=True)

    warm = max(args.warmup, 3)
    _dbg('model ready; first eager steps + graph capture')
    for i in range(3):  # two eager steps and the graph capture happen before anything is warmed up or timed
        step_dev(i)
    torch.cuda.synchronize()
    _dbg('graph captured; soak')
    with ClockSampler(local_rank) as clk:  # clocks under load: warm-up, both timed regions and the soak in between
        for i in range(max(warm, 400)):  # >= W warm-up steps; a fixed count (same on every rank) long enough
            step_dev(i)                  # (~0.5 s) for the 200 ms sampler to see the clocks under load
        torch.cuda.synchronize()
        _dbg('soak done; timed region')
        ms = timed(step_dev, args.steps, world)
        _dbg(f'timed done: {ms / args.steps:.3f} ms/step; e2e')
        for i in range(3):
            step_e2e(i)
        ms_e2e = timed(step_e2e, args.steps, world)
    _dbg('e2e done')
    launches_per_step = tr.launches_per_step
    t9 = task9_point(w, world, rank, min(args.steps, 50))
    # data-parallel replicas must stay in lock step (DDP semantics, template.py:243-244): compare the whole parameter
    # arena, the momentum buffer and (after rank 0's broadcast, what the reference's DDP does) nothing else — the
    # BatchNorm running statistics are per-rank by design until a task boundary
    replicas_identical = None
    if world > 1:
        torch.cuda.synchronize()
        mine = torch.cat([model.arena.detach().reshape(-1), tr.mom.reshape(-1)])
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        replicas_identical = bool(all(torch.equal(parts[0], p_) for p_ in parts[1:]))
    imgs = args.steps * BATCH * world
    value = imgs / (ms / 1e3)
    e2e = imgs / (ms_e2e / 1e3)
    final_loss = loss_host[(args.steps - 1) % loss_host.shape[0]].tolist()
    if rank != 0:
        return
    roof = kernel_rooflines(w, pk) if world == 1 else None
    per_task = per_task_block(w, model, teacher) if world == 1 else None
    _dbg('rooflines done; cpu baseline')
    cpu = None
    if world == 1:
        ips, done, threads, dt = cpu_steps(60, 1, budget_s=15.0)
        cpu = {"value": ips, "unit": "img/s", "cores": threads, "kind": "port",
               "sample": f"{done} steps of one 128-image batch ({dt:.1f} s), oracle/wa_oracle.train_step, torch CPU fp32"}
    line = {
        "metric": "images/sec per CIL training step", "value": value, "unit": "img/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "backbone": "resnet32", "global_batch": BATCH * world, "per_gpu_batch": BATCH,
                   "classes_seen": sum(HEADS), "classes_known": KNOWN, "parallelism": f"dp{world}",
                   "l2": f"inputs rotate over {POOL} distinct batches; per-step working set (~190 MB of saved "
                         "activations + gradients) exceeds the 126 MB L2",
                   "cuda_graph": True, "peaks": pk["src"]},
        "e2e": {"value": e2e, "unit": "img/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": BATCH * 3 * 32 * 32 * 4 + BATCH * 8, "d2h_bytes_per_step": 16,
                "api": "TaskTrainer.step(x_pinned_host, y_pinned_host) + async D2H of the losses"},
        "gpu_launches": launches_per_step * args.steps,
        "launches_per_step": launches_per_step,
        "clocks": clk.summary(),
        "step_tensor_frac": value / world * F_STEP / (pk["tflops"] * 1e12),
        "final_losses": final_loss,
        "task9": t9,
    }
    if per_task is not None:
        line["per_task"] = per_task
    if replicas_identical is not None:
        line["replicas_identical"] = replicas_identical
    if roof is not None:
        c = roof["conv"]
        line["roofline"] = {"bound": "hbm", "achieved": c["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                            "frac": c["gbs"] / pk["hbm_gbs"],
                            "traffic": ncu_traffic("r02_ncu_full_bn_fwd_stage1.csv", "conv_gemm_kernel_bn_fwd"),
                            "kernel": c["kernel"], "us": c["us"],
                            "traffic_note": "dram__bytes_read+write of one launch, parsed from the committed ncu --set full export "
                                            "profiles/r02_ncu_full_bn_fwd_stage1.csv (same entry point): outputs stay in the 126 MB L2 "
                                            "(write-back), inputs are read once from DRAM",
                            "algorithmic_bytes": c["bytes"], "tflops": c["tflops"], "peak_src": pk["src"]}
        line["kd_kernel"] = roof["kd"]
        line["weight_align_kernel"] = roof["wa"]
    if cpu is not None:
        line["cpu_baseline"] = cpu
        try:
            line["cpu_gloo"] = cpu_gloo_baseline()
        except Exception as ex:  # noqa: BLE001
            line["cpu_gloo"] = {"error": str(ex)[:200]}
    if world == 1:
        try:
            line["gpu_eager_baseline"] = {"fp32_img_s": gpu_eager_steps(20, False), "bf16_autocast_img_s": gpu_eager_steps(20, True),
                                          "what": "the reference's step (oracle port of template.py:255-271) as PyTorch eager on "
                                                  "this B200: cuDNN/cuBLAS/ATen kernels, no CUDA graph; 20 steps, wall clock"}
        except Exception as ex:  # a baseline must never take the bench line down
            line["gpu_eager_baseline"] = {"error": str(ex)[:200]}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    run_ours(args, rank, local_rank, world)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
