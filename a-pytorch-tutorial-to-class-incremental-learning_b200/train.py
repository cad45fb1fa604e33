"""The WA task loop of the reference's script body (template.py:191-303) as functions.

The reference has no train()/update_model() functions — its loop is inline under `if __name__ == "__main__"`.
This module provides them as thin wrappers with the same call order and hyper-parameter names:

    update_model(model, args, nb_new, "before")   template.py:241          (add a head)
    train_task(model, teacher, (x, y), args)      template.py:246-283      (epochs of CE + lambda*KD, SGD, cosine LR)
    update_model(model, args, nb_new, "after")    template.py:285-290      (weight align, teacher snapshot)
    TaskTrainer.step(x, y)                        template.py:255-271      (one optimisation step, CUDA-graph replay)

Per-step host synchronisations of the reference (4x .item(), barrier; template.py:272-276) are gone: losses are
accumulated on the device and read once per epoch."""
from __future__ import annotations

import argparse
import math
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, ops
from .model import CilModel


def get_args_parser():
    """The reference's flags with the same names and defaults (template.py:13-49)."""
    p = argparse.ArgumentParser("Class-Incremental Learning training and evaluation script", add_help=False)
    p.add_argument("--seed", default=0, type=int)
    p.add_argument("--num_bases", default=50, type=int)
    p.add_argument("--increment", default=10, type=int)
    p.add_argument("--backbone", default="resnet32", type=str)
    p.add_argument("--batch_size", default=128, type=int)
    p.add_argument("--input_size", default=32, type=int)
    p.add_argument("--color_jitter", default=0.4, type=float)
    p.add_argument("--aa", type=str, default="rand-m9-mstd0.5-inc1", metavar="NAME")
    p.add_argument("--reprob", type=float, default=0.0, metavar="PCT")
    p.add_argument("--remode", type=str, default="pixel")
    p.add_argument("--recount", type=int, default=1)
    p.add_argument("--resplit", action="store_true", default=False)
    p.add_argument("--herding_method", default="barycenter", type=str)
    p.add_argument("--memory_size", default=2000, type=int)
    p.add_argument("--fixed_memory", default=False, action="store_true")
    p.add_argument("--lr", default=0.1, type=float)
    p.add_argument("--momentum", default=0.9, type=float)
    p.add_argument("--weight_decay", default=5e-4, type=float)
    p.add_argument("--num_epochs", default=140, type=int)
    p.add_argument("--smooth", default=0.0, type=float)
    p.add_argument("--eval_every_epoch", default=5, type=float)
    p.add_argument("--dist_url", default="env://")
    p.add_argument("--data_set", default="cifar")
    p.add_argument("--data_path", default="/data/data/data/cifar100")
    p.add_argument("--lambda_kd", default=0.5, type=float)
    p.add_argument("--dynamic_lambda_kd", action="store_true")
    return p


def add_runner_args(p):
    """Flags of this implementation's runner (template.main) that the reference does not have."""
    p.add_argument("--synthetic_per_class", default=500, type=int, help="--data_set synthetic: train images per class")
    p.add_argument("--synthetic_val_per_class", default=100, type=int)
    p.add_argument("--nb_classes", default=100, type=int, help="--data_set synthetic: number of classes")
    p.add_argument("--checkpoint_dir", default="", type=str, help="write a checkpoint after every task")
    p.add_argument("--resume", default="", type=str, help="checkpoint file to resume from")
    p.add_argument("--dedup_old", action="store_true", help="keep each class once in the rehearsal memory")
    p.add_argument("--max_tasks", default=0, type=int, help="stop after this many tasks (0 = all)")
    return p


def init_seed(args):
    """template.py:52-58."""
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(args.seed)


def init_distributed_mode(args):
    """utils.py:135-153: one process per GPU from torchrun's env; NCCL on GPUs (gloo only for CPU-side tests).
    Unlike the reference, a plain single-process launch is accepted (world_size 1)."""
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        args.rank = int(os.environ["RANK"])
        args.world_size = int(os.environ["WORLD_SIZE"])
        args.gpu = int(os.environ.get("LOCAL_RANK", 0))
        args.distributed = True
        cuda = torch.cuda.is_available()
        if cuda:
            torch.cuda.set_device(args.gpu)
        args.dist_backend = "nccl" if cuda else "gloo"
        if not dist.is_initialized():
            kw = {"device_id": torch.device("cuda", args.gpu)} if cuda else {}
            dist.init_process_group(backend=args.dist_backend, init_method=getattr(args, "dist_url", "env://"),
                                    world_size=args.world_size, rank=args.rank, **kw)
        dist.barrier()
    else:
        args.rank, args.world_size, args.gpu, args.distributed = 0, 1, 0, False
        if torch.cuda.is_available():
            torch.cuda.set_device(0)


def accuracy(output, target, topk=(1,)):
    """timm.utils.accuracy as used at template.py:179,267."""
    maxk = min(max(topk), output.size(1))
    _, pred = output.topk(maxk, 1, True, True)
    pred = pred.t()
    correct = pred.eq(target.reshape(1, -1).expand_as(pred))
    return [correct[: min(k, maxk)].reshape(-1).float().sum(0) * 100.0 / target.size(0) for k in topk]


def cosine_lr(base_lr, epoch, t_max):
    """CosineAnnealingLR(T_max=num_epochs) stepped once per epoch (template.py:248-249,278)."""
    return 0.5 * base_lr * (1.0 + math.cos(math.pi * epoch / t_max))


def sampler_indices(n, epoch, rank, world_size, seed=0, shuffle=True):
    """torch.utils.data.DistributedSampler(shuffle=True, drop_last=False) index arithmetic (template.py:232-233,253)."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        idx = torch.randperm(n, generator=g)
    else:
        idx = torch.arange(n)
    total = int(math.ceil(n / world_size)) * world_size
    if total > n:
        pad = total - n
        idx = torch.cat([idx, idx[:pad]]) if pad <= n else torch.cat([idx, idx.repeat(math.ceil(pad / n))[:pad]])
    return idx[rank:total:world_size]


def allreduce_grads_(flat_grads, world_size, group=None):
    """The DDP exchange of template.py:243-244,270 on the ONE flat gradient arena: sum over ranks (NCCL on GPUs);
    returns the factor (1/world) the SGD kernel folds into its update so that the result is DDP's mean."""
    if world_size > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world_size


def broadcast_buffers(model, world_size=None, group=None, src=0):
    """What DistributedDataParallel(broadcast_buffers=True) gives the reference (template.py:243-244): every rank
    evaluates, snapshots the teacher and extracts herding features with RANK 0's BatchNorm running statistics.  The
    reference broadcasts them before every forward; the training forward itself only overwrites them (train-mode BN
    normalises with batch statistics), so one broadcast of the 2,303-element arena at the points where they are READ
    (before eval / teacher snapshot / feature extraction / checkpoint) leaves the same observable state."""
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world_size > 1:
        dist.broadcast(model.bn_running, src=src, group=group)
        model._touch()
    return model


def reduce_epoch_stats(acc, steps, n_samples, world_size, group=None):
    """acc = per-rank running sums [ce, kd, loss, top1-count] over `steps` steps -> global means and top-1 % over
    `n_samples` samples (the reference's MetricLogger.synchronize_between_processes, utils.py:36-43,117, once per epoch)."""
    acc = acc.clone()
    if world_size > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    ce, kd, loss, top1 = acc.tolist()
    denom = max(steps, 1) * world_size
    return {"ce": ce / denom, "kd": kd / denom, "loss": loss / denom, "acc1": 100.0 * top1 / max(n_samples, 1)}


# ------------------------------------------------------------------------------------------------------------
class _StepState:
    """Static buffers + captured graph for one batch size."""

    def __init__(self, tr, batch):
        dev = tr.device
        self.batch = batch
        self.x = torch.zeros((batch, 3, 32, 32), dtype=torch.float32, device=dev)
        self.y = torch.zeros((batch,), dtype=torch.int64, device=dev)
        self.losses = torch.zeros(4, dtype=torch.float32, device=dev)
        self.dlogits = torch.zeros((batch, tr.c_seen), dtype=torch.float32, device=dev)
        self.ws = torch.zeros(_lib.lib().wacil_kd_workspace_bytes(batch), dtype=torch.uint8, device=dev)
        self.plan_s = tr.model._plan(_lib.NET_TRAIN, batch)
        if tr.overlap:
            self.plan_s.set_side_stream(tr.side_w)
        if tr.fused_opt:  # gradient mean + SGD inside backward() (see Plan.set_optimizer)
            pr = tr.peer
            self.plan_s.set_optimizer(tr.mom, tr.lr_dev, tr.momentum, tr.weight_decay, tr.world_size,
                                      pr["rank"] if pr else 0, pr["grad_ptrs"] if pr else None,
                                      pr["pad_ptrs"] if pr else None, pr["epoch"] if pr else None)
        else:
            self.plan_s.set_optimizer(None, None, 0.0, 0.0)
        self.plan_t = None
        if tr.teacher is not None:
            self.plan_t = tr.teacher._plan(_lib.NET_INFER, batch)
            self.plan_t.fold()
            self.plan_t._folded_version = tr.teacher._fold_version()
        self.graph_fb = None   # forward + loss + backward (+ SGD when world_size == 1)
        self.graph_opt = None  # SGD after the all-reduce (world_size > 1)
        self.warm = 0
        self.stage = None      # host-fed steps: two device staging buffers filled by the copy stream (see stage_in)
        self.flip = 0

    def stage_in(self, tr, x, y):
        """Host batch -> the step's input buffers without stalling the compute stream on PCIe: the host-to-device copy
        goes to a staging buffer on the trainer's copy stream (it runs under the previous step, which the host has
        already queued), the compute stream only waits for its event and moves the batch device-to-device."""
        if self.stage is None:
            self.stage = [(torch.empty_like(self.x), torch.empty_like(self.y), torch.cuda.Event(), torch.cuda.Event())
                          for _ in range(2)]
        sx, sy, ready, free = self.stage[self.flip]
        self.flip ^= 1
        cur = torch.cuda.current_stream()
        cs = tr.copy_stream
        cs.wait_event(free)  # the device-to-device read of this buffer two steps ago (recorded below; no-op at first)
        with torch.cuda.stream(cs):
            sx.copy_(x, non_blocking=True)
            sy.copy_(y, non_blocking=True)
            ready.record(cs)
        cur.wait_event(ready)
        self.x.copy_(sx, non_blocking=True)
        self.y.copy_(sy, non_blocking=True)
        free.record(cur)


class TaskTrainer:
    """One task's optimiser state and step executor (a new one per task, like the reference's per-task SGD and
    schedule, template.py:246-249)."""

    def __init__(self, model: CilModel, teacher, known_classes, lr=0.1, momentum=0.9, weight_decay=5e-4,
                 lambda_kd=0.5, T=2.0, smooth=0.0, world_size=1, use_graph=True, process_group=None, overlap=True,
                 fused_allreduce=True, ce_weight=1.0):
        if not model.arena.is_cuda:
            raise _lib.WacilError("TaskTrainer needs the model on a CUDA device (no CPU path)")
        self.model, self.teacher = model, teacher
        self.device = model.arena.device
        self.c_seen, self.known = model.num_classes, int(known_classes)
        self.momentum, self.weight_decay = momentum, weight_decay
        self.lambda_kd, self.T, self.smooth, self.ce_weight = lambda_kd, T, smooth, ce_weight
        self.world_size, self.pg = world_size, process_group
        self.use_graph = use_graph
        # two extra streams: the frozen teacher's forward runs beside the student's, the weight gradients beside the
        # data-gradient chain (independent work that otherwise leaves these latency-bound kernels' SMs half idle)
        self.overlap = overlap
        self.side_t = torch.cuda.Stream(device=self.device) if overlap else None
        self.side_w = torch.cuda.Stream(device=self.device) if overlap else None
        self.copy_stream = torch.cuda.Stream(device=self.device)  # host-fed steps: H2D under the previous step
        self.peer = None
        if world_size > 1 and fused_allreduce:
            self.peer = self._setup_peer_memory(process_group)
        self.mom = torch.zeros_like(model._arena.params)
        # The optimiser is fused into the backward (C side) when the whole arena trains and the exchange is either local
        # or over NVLink peer memory; the NCCL fallback and partially frozen models (template.py:128-144) use _opt().
        self.train_range = model.trainable_range()
        whole = self.train_range == (0, model._arena.params.numel())
        self.fused_opt = whole and (world_size == 1 or self.peer is not None)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=self.device)
        self.acc = torch.zeros(4, dtype=torch.float32, device=self.device)  # running sums of [ce, kd, loss, top1]
        self.steps_done = 0
        self._states = {}
        self.launches_per_step = 0
        if teacher is not None and teacher.num_classes != self.known:
            raise _lib.WacilError(f"teacher has {teacher.num_classes} classes, known_classes = {self.known}")

    def _setup_peer_memory(self, group):
        """Put the gradient arena and a flag pad into symmetric (peer-mapped) memory so that the fused all-reduce + SGD
        kernel can read every rank's gradients over NVLink.  Returns None (NCCL all-reduce is used instead) when this
        torch build / topology cannot provide peer mappings."""
        import ctypes
        import warnings
        try:
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else dist.group.WORLD
            arena = self.model._arena
            n = arena.grads.numel()
            g = symm_mem.empty(n, dtype=torch.float32, device=self.device)
            g.zero_()
            pad = symm_mem.empty(64, dtype=torch.int32, device=self.device)
            pad.zero_()
            hg = symm_mem.rendezvous(g, grp)
            hp = symm_mem.rendezvous(pad, grp)
            world, rank = hg.world_size, hg.rank
            if world != self.world_size:
                raise RuntimeError("world size mismatch")
            peer = {"grads": g, "pad": pad, "handles": (hg, hp), "rank": rank,
                    "grad_ptrs": (ctypes.c_void_p * world)(*[int(x) for x in hg.buffer_ptrs]),
                    "pad_ptrs": (ctypes.c_void_p * world)(*[int(x) for x in hp.buffer_ptrs]),
                    "epoch": torch.zeros(2, dtype=torch.int32, device=self.device)}
            arena.grads = g          # plans bind the gradient pointer at creation: drop the stale ones
            self.model._plans = {}
            torch.cuda.synchronize()
            dist.barrier(group=group)
            return peer
        except Exception as ex:  # noqa: BLE001
            warnings.warn(f"symmetric memory unavailable ({ex}); gradients go through NCCL all-reduce")
            return None

    def set_lr(self, lr):
        self.lr_dev.fill_(float(lr))

    # -- pieces -------------------------------------------------------------------------------------------------
    def _fwd_bwd(self, st):
        t_logits = None
        n = 0
        main = torch.cuda.current_stream()
        if st.plan_t is not None and self.overlap:
            self.side_t.wait_stream(main)
            with torch.cuda.stream(self.side_t):
                t_logits, _ = st.plan_t.forward(st.x, self.known)
                n += st.plan_t.launches
        logits, _ = st.plan_s.forward(st.x, self.c_seen)
        n += st.plan_s.launches
        if st.plan_t is not None:
            if self.overlap:
                main.wait_stream(self.side_t)
            else:
                t_logits, _ = st.plan_t.forward(st.x, self.known)
                n += st.plan_t.launches
        ops.ce_kd_fwd_bwd(logits, st.y, t_logits, self.known, self.T, self.lambda_kd, self.smooth, st.losses,
                          st.dlogits, st.ws, ce_weight=self.ce_weight)
        st.plan_s.backward(st.dlogits)
        n += st.plan_s.launches + 1
        self.acc.add_(st.losses)
        return n

    def _opt(self):
        a = self.model._arena
        if self.fused_opt:
            return 0  # applied inside plan_s.backward()
        lo, hi = self.train_range
        if (lo, hi) != (0, a.params.numel()):  # freeze(["fc"]) / freeze(["backbone"]): only the trainable range moves
            if hi > lo:
                ops.sgd_step_(a.params[lo:hi], a.grads[lo:hi], self.mom[lo:hi], self.lr_dev, self.momentum,
                              self.weight_decay, 1.0 / self.world_size)
            return 1
        if self.peer is not None:  # gradient mean over NVLink peer memory + SGD in one kernel
            ops.allreduce_sgd_(a.params, self.mom, self.peer["grad_ptrs"], self.peer["pad_ptrs"], self.world_size,
                               self.peer["rank"], self.lr_dev, self.peer["epoch"], self.momentum, self.weight_decay)
        else:
            ops.sgd_step_(a.params, a.grads, self.mom, self.lr_dev, self.momentum, self.weight_decay,
                          1.0 / self.world_size)
        return 1

    def _run(self, st):
        single = self.world_size == 1 or self.fused_opt  # one graph holds the whole step
        if not self.use_graph or st.warm < 2:
            n = self._fwd_bwd(st)
            if not single:
                allreduce_grads_(self.model._arena.grads, self.world_size, self.pg)
            n += self._opt()
            st.warm += 1
            self.launches_per_step = n
            return
        if st.graph_fb is None:
            torch.cuda.synchronize()
            st.graph_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(st.graph_fb):
                self._fwd_bwd(st)
                if single:
                    self._opt()
            if not single:
                st.graph_opt = torch.cuda.CUDAGraph()
                with torch.cuda.graph(st.graph_opt):
                    self._opt()
            return  # capture executes nothing (no forward/backward, no update): the caller replays the step right after
        st.graph_fb.replay()
        if not single:
            allreduce_grads_(self.model._arena.grads, self.world_size, self.pg)
            st.graph_opt.replay()

    def _state(self, b):
        st = self._states.get(b)
        if st is None:
            st = self._states[b] = _StepState(self, b)
        return st

    def _execute(self, st):
        captured_now = self.use_graph and st.warm >= 2 and st.graph_fb is None
        self._run(st)
        if captured_now:
            self._run(st)  # the capture itself did not execute the step
        self.steps_done += 1
        self.model._num_batches_tracked += 1
        self.model._touch()
        return st.losses

    def step(self, x, y):
        """One optimisation step (template.py:255-271) on a [B,3,32,32] fp32 batch and int64 targets; `x`/`y` may live
        in pinned host memory (copied asynchronously) or on the device.  Returns the device tensor [ce, kd, loss, top1
        count] of this step (no host sync)."""
        st = self._state(x.shape[0])
        if x.is_cuda:
            st.x.copy_(x, non_blocking=True)
            st.y.copy_(y, non_blocking=True)
        else:
            st.stage_in(self, x, y)
        return self._execute(st)

    def step_from_store(self, taskset, idx, aug=None):
        """The same step fed by the device data path (utils.py:210-251 without the host loader): gather + RandomCrop +
        flip + ToTensor + Normalize of samples `idx` of the uint8 store straight into the step's input buffer."""
        st = self._state(idx.numel())
        taskset.batch(idx, aug, out=st.x)
        torch.index_select(taskset.y, 0, idx, out=st.y)
        return self._execute(st)

    def epoch_stats(self, n_samples):
        """Means of ce / kd / loss and top-1 % since the last call (one host sync per epoch)."""
        out = reduce_epoch_stats(self.acc, self.steps_done, n_samples, self.world_size, self.pg)
        self.acc.zero_()
        self.steps_done = 0
        return out


# ------------------------------------------------------------------------------------------------------------
def kd_weights(args, known, nb_new, has_teacher):
    """(ce_weight, lambda_kd).  Default: CE + lambda_kd * KD (template.py:262-266).  --dynamic_lambda_kd (declared at
    template.py:48, never read by the reference): the objective of README.md:172-176,
    (1 - lambda) * CE + lambda * KD with lambda = n_old / (n_old + n_new)."""
    if has_teacher and getattr(args, "dynamic_lambda_kd", False) and known > 0:
        lam = known / float(known + nb_new)
        return 1.0 - lam, lam
    return 1.0, args.lambda_kd


def train_task(model, teacher_model, data, args, rank=None, world_size=None, num_epochs=None, use_graph=True,
               on_epoch_end=None, log=print):
    """Train `model` on one task (template.py:246-283).  `data` = (x [N,3,32,32] fp32, y [N] int64) for this task
    (new samples U rehearsal memory), device-resident or pinned host tensors, or a data.DeviceTaskSet (uint8 store in
    HBM, augmentation on the device).  Returns the per-epoch stats."""
    from .data import DeviceTaskSet
    rank = getattr(args, "rank", 0) if rank is None else rank
    world_size = getattr(args, "world_size", 1) if world_size is None else world_size
    epochs = args.num_epochs if num_epochs is None else num_epochs
    bs = args.batch_size
    known = getattr(args, "known_classes", 0)
    if teacher_model is None:
        known = 0
    ce_w, lam = kd_weights(args, known, model.num_classes - known, teacher_model is not None)
    tr = TaskTrainer(model, teacher_model, known, lr=args.lr, momentum=args.momentum, weight_decay=args.weight_decay,
                     lambda_kd=lam, T=2.0, smooth=getattr(args, "smooth", 0.0), world_size=world_size,
                     use_graph=use_graph, ce_weight=ce_w)
    from_store = isinstance(data, DeviceTaskSet)
    if from_store:
        n, dev = len(data), data.x.device
    else:
        x_all, y_all = data
        n, dev = x_all.shape[0], (x_all.device if x_all.is_cuda else None)
    history = []
    for epoch in range(epochs):
        model.train()
        tr.set_lr(cosine_lr(args.lr, epoch, epochs))
        idx = sampler_indices(n, epoch, rank, world_size, seed=0, shuffle=True)
        if dev is not None:
            idx = idx.to(dev)
        aug = data.draw_aug(idx.numel(), seed=getattr(args, "seed", 0), epoch=epoch, rank=rank) if from_store else None
        seen = 0
        for s in range(0, idx.numel(), bs):
            bi = idx[s: s + bs]
            if from_store:
                tr.step_from_store(data, bi, aug[s: s + bs])
            else:
                tr.step(x_all.index_select(0, bi), y_all.index_select(0, bi))
            seen += bi.numel()
        stats = tr.epoch_stats(seen * world_size)
        stats["epoch"], stats["lr"] = epoch + 1, cosine_lr(args.lr, epoch, epochs)
        history.append(stats)
        if log is not None and rank == 0:
            log(f"train states: epoch :[{epoch + 1}/{epochs}] ce: {stats['ce']:.4f}  kd: {stats['kd']:.4f}  "
                f"loss: {stats['loss']:.4f}  acc1: {stats['acc1']:.4f}")
        if on_epoch_end is not None:
            on_epoch_end(epoch, stats)
    return history


def update_model(model, args, nb_new_classes, stage):
    """stage "before": grow the classifier by one head (template.py:241).
    stage "after": weight-align when task_id > 0 and snapshot the frozen teacher (template.py:285-290); returns it."""
    if stage == "before":
        model.prev_model_adaption(nb_new_classes)
        return model
    if stage == "after":
        broadcast_buffers(model, getattr(args, "world_size", 1))  # the teacher is a copy of rank 0's statistics
        model.after_model_adaption(nb_new_classes, args)
        return model.copy().freeze()
    raise ValueError(stage)


@torch.no_grad()
def extract_features(model, x_all, batch_size=128):
    """Features of every sample for herding (template.py:292-299), eval-mode BN, kept on the device.  `x_all`: fp32
    [N,3,32,32] tensor or a data.DeviceTaskSet (centre crop, no flip: the reference extracts through the *train*
    transform, whose random draws differ per rank — SURVEY 8e).  Every rank computes identical features provided the
    caller synchronised the BatchNorm running statistics first (broadcast_buffers; template.main does)."""
    from .data import DeviceTaskSet
    was_training = model.training
    model.eval()
    outs = []
    n = len(x_all) if isinstance(x_all, DeviceTaskSet) else x_all.shape[0]
    for s in range(0, n, batch_size):
        if isinstance(x_all, DeviceTaskSet):
            xb = x_all.batch(torch.arange(s, min(s + batch_size, n), device=x_all.x.device), None)
        else:
            xb = x_all[s: s + batch_size]
            if not xb.is_cuda:
                xb = xb.to(model.arena.device, non_blocking=True)
        outs.append(model.extract_vector(xb.contiguous()))
    model.train(was_training)
    return torch.cat(outs, 0)


def _finish_meter(meter, log=print):
    """MetricLogger.synchronize_between_processes + the global averages eval() prints (template.py:185-188)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(meter)
    n, nb, loss_sum, top1, top5 = meter[:5].tolist()  # the one host sync of the evaluation
    acc1, acc5 = 100.0 * top1 / max(n, 1.0), 100.0 * top5 / max(n, 1.0)
    loss = loss_sum / max(nb, 1.0)
    if log is not None:
        log(" Acc@1 {:.3f}  loss {:.3f}".format(acc1, loss))
    return acc1, acc5, loss


@torch.no_grad()
def eval(model, val_loader):  # noqa: A001  (the reference's name, template.py:169-188)
    """Top-1 accuracy (%) over an iterable of (images, target, task_ids) — this rank's shard, like the reference's
    DistributedSampler'd loader; losses and hit counts are accumulated on the device by one fused kernel per batch
    (wacil_eval_rows) and reduced across ranks once."""
    model.eval()
    dev = model.arena.device
    meter = torch.zeros(8, dtype=torch.float64, device=dev)
    for images, target, *_ in val_loader:
        images, target = images.to(dev, non_blocking=True), target.to(dev, non_blocking=True)
        logits, _ = model(images.contiguous().float())
        ops.eval_rows_(logits, target.long(), meter, topk=5)
    return _finish_meter(meter)[0]


@torch.no_grad()
def evaluate(model, taskset, args=None, batch_size=None, log=print):
    """eval() over a data.DeviceTaskSet (template.py:229,234-239,282-287): each rank takes its
    DistributedSampler(shuffle=False) shard (padding duplicates included, like the reference), eval transform on the
    device, fused meter kernel, one all-reduce.  Returns (acc1, acc5, loss)."""
    model.eval()
    rank, world = getattr(args, "rank", 0), getattr(args, "world_size", 1)
    bs = batch_size or getattr(args, "batch_size", 128)
    dev = taskset.x.device
    idx = sampler_indices(len(taskset), 0, rank, world, shuffle=False).to(dev)
    meter = torch.zeros(8, dtype=torch.float64, device=dev)
    for s in range(0, idx.numel(), bs):
        bi = idx[s: s + bs]
        logits, _ = model(taskset.batch(bi, None))
        ops.eval_rows_(logits, taskset.y.index_select(0, bi), meter, topk=5)
    return _finish_meter(meter, log if rank == 0 else None)
