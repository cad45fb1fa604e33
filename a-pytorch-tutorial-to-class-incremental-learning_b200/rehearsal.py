"""Rehearsal memory + barycenter herding with the selection loop on the GPU.

Mirrors `continuum.rehearsal.RehearsalMemory` / `herd_closest_to_barycenter` as the reference uses them
(template.py:9,212-216,231,300-302; continuum==1.2.2 is a third-party dependency that is not vendored in the
reference, so the semantics follow its published behaviour — see oracle/wa_oracle.py and DESIGN.md)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops


def herd_indices(features, y, nb_per_class, device=None, method="barycenter"):
    """Indices (into the rows of `features`) of the exemplars: classes in ascending id order, herding-rank order
    inside a class.  `features`: [n, D] float32 numpy array or torch tensor; `y`: [n] integer labels.
    method: "barycenter" (iCaRL herding) or "cluster" (closest to the class mean)."""
    if isinstance(features, np.ndarray):
        if features.ndim != 2:
            raise ValueError(f"Expected features to have 2 dimensions, not {features.ndim}d.")
        feats = torch.from_numpy(np.ascontiguousarray(features, dtype=np.float32))
    else:
        if features.dim() != 2:
            raise ValueError(f"Expected features to have 2 dimensions, not {features.dim()}d.")
        feats = features.detach().float()
    if device is None:
        device = feats.device if feats.is_cuda else torch.device("cuda", torch.cuda.current_device())
    if not torch.cuda.is_available():
        raise _lib.WacilError("herding runs on the GPU (no CPU fallback)")
    feats = feats.to(device)
    y_t = torch.as_tensor(np.asarray(y) if not torch.is_tensor(y) else y).to(device).long()
    order = torch.argsort(y_t, stable=True)             # group rows by ascending class id, original order inside
    classes, counts = torch.unique_consecutive(y_t[order], return_counts=True)
    offsets = torch.zeros(classes.numel() + 1, dtype=torch.int32, device=device)
    offsets[1:] = torch.cumsum(counts, 0).int()
    grouped = feats.index_select(0, order).contiguous()
    if method == "barycenter":
        idx, cnt = ops.herd_barycenter(grouped, offsets, int(nb_per_class))
    elif method == "cluster":
        idx, cnt = ops.herd_cluster(grouped, offsets, int(nb_per_class))
    else:
        raise NotImplementedError(f"Unknown rehearsal method {method}.")
    # row-local -> original indices, trimmed to the per-class counts
    k, q = idx.shape
    valid = torch.arange(q, device=device)[None, :] < cnt[:, None]
    glob = order[(idx.clamp_min(0).long() + offsets[:-1, None].long())]
    return glob[valid]


def _take(arrs, idx):
    """Index numpy arrays or torch tensors (host or device) with a device index tensor; only indices cross to the host."""
    out = []
    for a in arrs:
        if torch.is_tensor(a):
            out.append(a.index_select(0, idx.to(a.device)))
        else:
            out.append(a[idx.cpu().numpy()])
    return tuple(out)


def _unique_classes(y):
    """Ascending class ids as python ints (a handful of integers: the only thing that leaves the device)."""
    if torch.is_tensor(y):
        return [int(c) for c in torch.unique(y).tolist()]
    return [int(c) for c in np.unique(y)]


def _cat(a, b):
    if torch.is_tensor(a) or torch.is_tensor(b):
        a = a if torch.is_tensor(a) else torch.as_tensor(a)
        b = b if torch.is_tensor(b) else torch.as_tensor(b)
        return torch.cat([a, b.to(a.device)])
    return np.concatenate((a, b))


def herd_closest_to_barycenter(x, y, t, features, nb_per_class):
    """continuum's function signature: returns (x[idx], y[idx], t[idx]).  x / y / t may be numpy arrays (upstream) or
    torch tensors — device tensors stay on the device (template.main keeps the whole rehearsal path in HBM)."""
    return _take((x, y, t), herd_indices(features, y, nb_per_class))


def herd_closest_to_cluster(x, y, t, features, nb_per_class):
    """continuum.rehearsal.herd_closest_to_cluster (--herding_method cluster): distances and ranking on the GPU."""
    return _take((x, y, t), herd_indices(features, y, nb_per_class, method="cluster"))


def herd_random(x, y, t, features, nb_per_class):
    """continuum.rehearsal.herd_random (--herding_method random): index selection only (no arithmetic), drawn from
    NumPy's global RNG like upstream (the reference seeds it, template.py:54)."""
    yh = y.cpu().numpy() if torch.is_tensor(y) else y
    indexes = []
    for class_id in np.unique(yh):
        class_indexes = np.where(yh == class_id)[0]
        indexes.append(np.random.choice(class_indexes, size=min(nb_per_class, len(class_indexes)), replace=False))
    return _take((x, y, t), torch.from_numpy(np.concatenate(indexes)))


_HERDING = {"barycenter": herd_closest_to_barycenter, "cluster": herd_closest_to_cluster, "random": herd_random}


class RehearsalMemory:
    """`rehearsal.RehearsalMemory(memory_size, herding_method, fixed_memory)` (template.py:212-216).

    add(x, y, t, features) herds on the GPU.  With numpy inputs get() returns numpy arrays like upstream; with torch
    (device) inputs the exemplars, labels and features never leave HBM — only class ids and counts are read by the
    host.  `dedup_old=False` (default) reproduces the upstream behaviour the reference relies on: the task set passed
    at template.py:300-302 already contains the old exemplars, they are herded again and appended, so old classes
    appear twice.  `dedup_old=True` keeps each class once (the 2000-exemplar budget of BASELINE.md)."""

    def __init__(self, memory_size, herding_method="barycenter", fixed_memory=False, nb_total_classes=None,
                 dedup_old=False):
        if isinstance(herding_method, str):
            if herding_method not in _HERDING:
                raise NotImplementedError(f"Unknown rehearsal method {herding_method}.")
            herding_method = _HERDING[herding_method]
        elif not callable(herding_method):
            raise NotImplementedError("herding_method must be 'barycenter', 'cluster', 'random' or a callable")
        if fixed_memory and nb_total_classes is None:
            raise ValueError("When memory is fixed, precise the `nb_total_classes` argument.")
        self.memory_size, self.herding_method = memory_size, herding_method
        self.fixed_memory, self.nb_total_classes, self.dedup_old = fixed_memory, nb_total_classes, dedup_old
        self.seen_classes = set()
        self._x = self._y = self._t = None

    @property
    def nb_classes(self):
        return len(self.seen_classes)

    @property
    def memory_per_class(self):
        if self.fixed_memory:
            return self.memory_size // self.nb_total_classes
        elif self.nb_classes > 0:
            return self.memory_size // self.nb_classes
        return self.memory_size

    def __len__(self):
        return 0 if self._x is None else len(self._x)

    def get(self):
        return self._x, self._y, self._t

    def get_numpy(self):
        """(x, y, t) as numpy arrays whatever the storage (checkpoints, upstream-style consumers)."""
        return tuple(None if a is None else (a.cpu().numpy() if torch.is_tensor(a) else np.asarray(a))
                     for a in (self._x, self._y, self._t))

    def save(self, path="memory.npz"):
        x, y, t = self.get_numpy()
        np.savez(path, x=x, y=y, t=t)

    def load(self, path="memory.npz"):
        data = np.load(path)
        self._x, self._y, self._t = data["x"], data["y"], data["t"]
        assert len(self) <= self.memory_size, len(self)
        self.seen_classes = set(int(c) for c in np.unique(self._y))

    def _reduce(self):
        """Keep the first memory_per_class stored items of every class (stored order = herding rank order)."""
        q = self.memory_per_class
        if torch.is_tensor(self._y):
            y = self._y
            order = torch.argsort(y, stable=True)
            ys = y[order]
            first = torch.searchsorted(ys, ys)  # index of the first element of each item's class in the sorted order
            keep = order[(torch.arange(len(ys), device=y.device) - first) < q]
            self._x, self._y, self._t = _take((self._x, self._y, self._t), keep)
            return
        xs, ys, ts = [], [], []
        for class_id in np.unique(self._y):
            indexes = np.where(self._y == class_id)[0][:q]
            xs.append(self._x[indexes]); ys.append(self._y[indexes]); ts.append(self._t[indexes])
        self._x, self._y, self._t = np.concatenate(xs), np.concatenate(ys), np.concatenate(ts)

    def add(self, x, y, t, z):
        """z = features of (x, y, t) (template.py:300-302)."""
        for c in _unique_classes(y):
            self.seen_classes.add(c)
        mem_x, mem_y, mem_t = self.herding_method(x, y, t, z, nb_per_class=self.memory_per_class)
        if self._x is None:
            self._x, self._y, self._t = mem_x, mem_y, mem_t
        else:
            if not self.fixed_memory:
                self._reduce()
            if self.dedup_old:
                if torch.is_tensor(mem_y):
                    old = self._y if torch.is_tensor(self._y) else torch.as_tensor(self._y)
                    keep = torch.nonzero(~torch.isin(mem_y, torch.unique(old).to(mem_y.device))).flatten()
                    mem_x, mem_y, mem_t = _take((mem_x, mem_y, mem_t), keep)
                else:
                    keep = ~np.isin(mem_y, np.unique(self.get_numpy()[1]))
                    mem_x, mem_y, mem_t = mem_x[keep], mem_y[keep], mem_t[keep]
            self._x = _cat(self._x, mem_x)
            self._y = _cat(self._y, mem_y)
            self._t = _cat(self._t, mem_t)
