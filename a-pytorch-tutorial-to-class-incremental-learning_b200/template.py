"""The script body of the reference (template.py:191-303) as a function: the whole WA class-incremental protocol on the
native engine, with the data held in HBM (data.py), one fused kernel per loss / meter / augmentation, and a checkpoint
after every task.

    python -m torch.distributed.run --nproc-per-node N -m wacil_b200.template --data_set synthetic --num_bases 10 ...
    python wacil_b200.py ... (single GPU)

The reference's flags keep their names and defaults (train.get_args_parser); `--data_set synthetic` stands in for
CIFAR-100 where the dataset cannot be downloaded."""
from __future__ import annotations

import argparse
import os

import torch

from . import checkpoint as ckpt
from .data import CIFAR100_CLASS_ORDER, build_dataset
from .model import CilModel
from .rehearsal import RehearsalMemory
from .train import (add_runner_args, broadcast_buffers, evaluate, extract_features, get_args_parser,
                    init_distributed_mode, init_seed, train_task)


def main(args, log=print):
    """Returns the list of per-task top-1 accuracies (acc1s, template.py:225,287-288)."""
    init_distributed_mode(args)
    init_seed(args)
    rank = getattr(args, "rank", 0)
    say = log if rank == 0 else (lambda *a, **k: None)
    if rank == 0 and (str(getattr(args, "aa", "none")).lower() not in ("none", "") or getattr(args, "reprob", 0.0) > 0):
        # utils.py:217-226: the reference's default train transform also applies timm RandAugment (--aa), colour jitter and
        # random erasing on the host (PIL); the device data path covers RandomCrop(32, 4) + flip + ToTensor + Normalize
        log("note: --aa / --color_jitter / --reprob are accepted for command-line compatibility but NOT applied: the device "
            "data path runs RandomCrop(32, padding=4) + RandomHorizontalFlip + Normalize only (pass --aa none to silence)")
    args.class_order = CIFAR100_CLASS_ORDER                                   # template.py:201-202
    dev = torch.device("cuda", getattr(args, "gpu", 0))
    scenario_train, args.nb_classes = build_dataset(True, args, device=dev)   # template.py:203
    scenario_val, _ = build_dataset(False, args, device=dev)                  # template.py:204
    model = CilModel(args, max_classes=args.nb_classes).cuda()               # template.py:206-208
    memory = RehearsalMemory(memory_size=args.memory_size, herding_method=args.herding_method,
                             fixed_memory=args.fixed_memory, nb_total_classes=args.nb_classes,
                             dedup_old=getattr(args, "dedup_old", False))      # template.py:212-216
    teacher_model = None
    args.increment_per_task = list(scenario_train.increments)                 # template.py:222-223
    args.known_classes = 0
    acc1s, start_task = [], 0
    if getattr(args, "resume", ""):
        model, teacher_model, done, acc1s = ckpt.load_checkpoint(args.resume, memory, args, device=dev)
        start_task = done + 1
        say(f"resumed from {args.resume}: task {done} done, known_classes = {args.known_classes}, acc1s = {acc1s}")
    n_tasks = len(scenario_train) if not getattr(args, "max_tasks", 0) else min(len(scenario_train), args.max_tasks)
    for task_id in range(start_task, n_tasks):
        args.task_id = task_id
        dataset_train = scenario_train[task_id]                               # template.py:226
        dataset_val = scenario_val[:task_id + 1]                              # template.py:229
        if task_id > 0 and len(memory) > 0:
            dataset_train.add_samples(*memory.get())                          # template.py:230-231
        inc = args.increment_per_task[task_id]
        model.prev_model_adaption(inc)                                        # template.py:241

        def on_epoch_end(epoch, stats, _val=dataset_val):
            if (epoch + 1) % args.eval_every_epoch == 0:                      # template.py:282-283
                broadcast_buffers(model, args.world_size)                     # DDP: rank 0's BN statistics
                evaluate(model, _val, args, log=say)

        train_task(model, teacher_model, dataset_train, args, on_epoch_end=on_epoch_end, log=say)  # :246-283
        # DDP(broadcast_buffers=True), template.py:243-244: from here on (eval, teacher snapshot, herding features,
        # checkpoint) every rank works with rank 0's BatchNorm running statistics -> identical teacher and memory
        broadcast_buffers(model, args.world_size)
        model.after_model_adaption(inc, args)                                 # template.py:285-286 (weight align)
        acc1 = evaluate(model, dataset_val, args, log=say)[0]                 # template.py:287
        acc1s.append(acc1)
        say(f"task id = {task_id}  @Acc1 = {acc1:.5f}, acc1s = {acc1s}")
        teacher_model = model.copy().freeze()                                 # template.py:290
        features = extract_features(model, dataset_train, args.batch_size)    # template.py:292-299
        memory.add(*dataset_train.get_raw_samples(device=True), features)    # template.py:300-302, all in HBM
        args.known_classes += inc                                             # template.py:303
        if getattr(args, "checkpoint_dir", "") and rank == 0:
            path = os.path.join(args.checkpoint_dir, f"task_{task_id:02d}.pt")
            ckpt.save_checkpoint(path, model, memory, args, task_id, acc1s)
            say(f"checkpoint -> {path}")
    return acc1s


def cli(argv=None):
    parser = argparse.ArgumentParser("Class-Incremental Learning training and evaluation script",
                                     parents=[get_args_parser()])
    add_runner_args(parser)
    return main(parser.parse_args(argv))


if __name__ == "__main__":
    cli()
