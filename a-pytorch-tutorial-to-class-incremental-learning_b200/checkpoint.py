"""Checkpoint / resume of a class-incremental run.  The reference has none (a 10-task run is all-or-nothing); the state
it would need lives at template.py:206 (model), :212-217 (rehearsal memory, teacher), :224-225 (known_classes, acc1s),
:52-58 (RNG).  One file per finished task: everything needed to continue with task_id + 1 and obtain the same run."""
from __future__ import annotations

import os

import numpy as np
import torch

from .model import CilModel

FORMAT = 2


def save_checkpoint(path, model, memory, args, task_id, acc1s, teacher=None, extra=None):
    """State after task `task_id` finished (after weight align, teacher snapshot and memory.add).  The teacher is the
    frozen copy of the model at this point (template.py:290), so it is rebuilt on load unless it differs."""
    mx, my, mt = memory.get_numpy()
    state = {
        "format": FORMAT,
        "backbone": model.backbone_name,
        "max_classes": model._arena.max_classes,
        "head_sizes": [model.fc._bounds[i + 1] - model.fc._bounds[i] for i in range(len(model.fc))],
        "model": {k: v.cpu() for k, v in model.state_dict().items()},
        "teacher": None if teacher is None else {k: v.cpu() for k, v in teacher.state_dict().items()},
        # tensors / plain python only: the file loads with torch.load(weights_only=True) (no arbitrary unpickling)
        "memory": None if mx is None else {"x": torch.from_numpy(np.ascontiguousarray(mx)),
                                           "y": torch.from_numpy(np.ascontiguousarray(my)),
                                           "t": torch.from_numpy(np.ascontiguousarray(mt))},
        "seen_classes": sorted(int(c) for c in memory.seen_classes),
        "task_id": int(task_id),
        "known_classes": int(args.known_classes),
        "acc1s": [float(a) for a in acc1s],
        "rng": {"torch": torch.get_rng_state(), "numpy": _np_rng_to_plain(np.random.get_state()),
                "cuda": torch.cuda.get_rng_state_all() if torch.cuda.is_available() else None},
        "nb_classes": int(getattr(args, "nb_classes", 0) or 0),
        "increments": [int(v) for v in getattr(args, "increment_per_task", [])],
        "extra": extra or {},
    }
    tmp = path + ".tmp"
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(state, tmp)
    os.replace(tmp, path)  # a crash while writing never leaves a truncated checkpoint behind
    return path


def _np_rng_to_plain(st):
    name, keys, pos, has_gauss, cached = st
    return {"name": str(name), "keys": torch.from_numpy(np.asarray(keys, dtype=np.int64)), "pos": int(pos),
            "has_gauss": int(has_gauss), "cached": float(cached)}


def _np_rng_from_plain(d):
    return (d["name"], d["keys"].numpy().astype(np.uint32), d["pos"], d["has_gauss"], d["cached"])


def _rebuild(state, sd, device):
    m = CilModel(state["backbone"], max_classes=state["max_classes"])
    for n in state["head_sizes"]:
        m.prev_model_adaption(n)
    if device is not None:
        m = m.to(device)
    m.load_state_dict(sd)
    return m


def load_checkpoint(path, memory, args, device=None, restore_rng=True):
    """Returns (model, teacher, task_id, acc1s); fills `memory` and args.known_classes in place."""
    state = torch.load(path, map_location="cpu", weights_only=True)
    if state.get("format") != FORMAT:
        raise ValueError(f"unsupported checkpoint format {state.get('format')}")
    # the run being resumed must be the run that was saved: a mismatch would only surface later as shape errors
    want = getattr(args, "backbone", None)
    if want is not None and want != state["backbone"]:
        raise ValueError(f"checkpoint is a {state['backbone']} run, --backbone is {want}")
    incs = [int(v) for v in getattr(args, "increment_per_task", [])]
    if incs and state.get("increments") and incs != state["increments"]:
        raise ValueError(f"checkpoint task increments {state['increments']} differ from this run's {incs}")
    if incs and state["head_sizes"] != incs[: len(state["head_sizes"])]:
        raise ValueError(f"checkpoint heads {state['head_sizes']} do not match the scenario's increments {incs}")
    nbc = int(getattr(args, "nb_classes", 0) or 0)
    if nbc and state.get("nb_classes") and nbc != state["nb_classes"]:
        raise ValueError(f"checkpoint has {state['nb_classes']} classes, this run {nbc}")
    model = _rebuild(state, state["model"], device)
    tsd = state["teacher"] if state["teacher"] is not None else state["model"]
    teacher = _rebuild(state, tsd, device).freeze()
    if state["memory"] is not None:
        mm = state["memory"]
        if device is not None and torch.device(device).type == "cuda":  # the rehearsal path lives in HBM
            memory._x, memory._y, memory._t = (mm[k].to(device) for k in ("x", "y", "t"))
        else:
            memory._x, memory._y, memory._t = (mm[k].numpy() for k in ("x", "y", "t"))
    memory.seen_classes = set(state["seen_classes"])
    args.known_classes = state["known_classes"]
    if restore_rng:
        torch.set_rng_state(state["rng"]["torch"])
        np.random.set_state(_np_rng_from_plain(state["rng"]["numpy"]))
        if state["rng"]["cuda"] is not None and torch.cuda.is_available() and \
                len(state["rng"]["cuda"]) == torch.cuda.device_count():
            torch.cuda.set_rng_state_all(state["rng"]["cuda"])
    return model, teacher, state["task_id"], list(state["acc1s"])
