"""Device-resident data path: the reference's continuum scenario + torchvision/timm transform + 10-worker DataLoader
(utils.py:192-251, template.py:196-197,223-239) reduced to what the WA loop needs, with the samples held as uint8 in
HBM and the CIFAR transform (RandomCrop(32, padding=4), RandomHorizontalFlip, ToTensor, Normalize) executed by one CUDA
kernel per batch (wacil_augment_u8).  RandAugment / colour jitter (timm create_transform, utils.py:217-226) are host
PIL operations outside this path: the device path is the `--aa none` pipeline.

    scenario_train, nb_classes = build_dataset(True, args)      utils.py:192-208
    task = scenario_train[task_id]                               ClassIncremental.__getitem__   (template.py:226)
    task.add_samples(*memory.get())                              TaskSet.add_samples            (template.py:231)
    x, y, t = task.get_raw_samples()                             TaskSet.get_raw_samples        (template.py:301)
    scenario_val[:task_id + 1]                                   slice = union of tasks         (template.py:229)
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch

from . import _lib, ops

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
CIFAR_MEAN, CIFAR_STD = (0.5071, 0.4867, 0.4408), (0.2675, 0.2565, 0.2761)

# template.py:201-202
CIFAR100_CLASS_ORDER = [68, 56, 78, 8, 23, 84, 90, 65, 74, 76, 40, 89, 3, 92, 55, 9, 26, 80, 43, 38, 58, 70, 77, 1, 85,
                        19, 17, 50, 28, 53, 13, 81, 45, 82, 6, 59, 83, 16, 15, 44, 91, 41, 72, 60, 79, 52, 20, 10, 31,
                        54, 37, 95, 14, 71, 96, 98, 97, 2, 64, 66, 42, 22, 35, 86, 24, 34, 87, 21, 99, 0, 88, 27, 18,
                        94, 11, 12, 47, 25, 30, 46, 62, 69, 36, 61, 7, 63, 75, 5, 32, 4, 51, 48, 73, 93, 39, 67, 29,
                        49, 57, 33]


def transform_stats(args):
    """Normalize() statistics the reference ends up with (utils.py:228-230,246-251): CIFAR's only when
    args.data_set == 'CIFAR' (upper case); the default `--data_set cifar` keeps the ImageNet defaults."""
    if getattr(args, "input_size", 32) == 32 and getattr(args, "data_set", "cifar") == "CIFAR":
        return CIFAR_MEAN, CIFAR_STD
    return IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD


class DeviceTaskSet:
    """A task's samples as uint8 [N,H,W,3] + int64 labels + task ids, resident on the GPU (continuum TaskSet surface:
    add_samples / get_raw_samples / __len__)."""

    def __init__(self, x, y, t=None, device=None, mean=IMAGENET_DEFAULT_MEAN, std=IMAGENET_DEFAULT_STD, pad=4):
        if device is None:
            if not torch.cuda.is_available():
                raise _lib.WacilError("DeviceTaskSet needs a CUDA device (the data path has no CPU fallback)")
            device = torch.device("cuda", torch.cuda.current_device())
        x = torch.as_tensor(x)
        if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[-1] != 3:
            raise ValueError(f"expected uint8 [N,H,W,3] samples, got {tuple(x.shape)} {x.dtype}")
        self.x = x.to(device).contiguous()
        self.y = torch.as_tensor(y).long().to(device)
        self.t = (torch.zeros_like(self.y) if t is None else torch.as_tensor(t).long().to(device))
        self.mean, self.std, self.pad = tuple(mean), tuple(std), int(pad)

    def __len__(self):
        return self.x.shape[0]

    def add_samples(self, x, y, t=None):
        """TaskSet.add_samples (template.py:231): append the rehearsal memory."""
        if x is None or len(x) == 0:
            return
        dev = self.x.device
        self.x = torch.cat([self.x, torch.as_tensor(x).to(dev)]).contiguous()
        self.y = torch.cat([self.y, torch.as_tensor(y).long().to(dev)])
        tt = torch.zeros(len(x), dtype=torch.long) if t is None else torch.as_tensor(t).long()
        self.t = torch.cat([self.t, tt.to(dev)])

    def get_raw_samples(self, device=False):
        """TaskSet.get_raw_samples (template.py:301): numpy (x uint8 [N,H,W,3], y, t) like upstream, or — device=True —
        the device tensors themselves (RehearsalMemory.add herds on the GPU and keeps the exemplars in HBM)."""
        if device:
            return self.x, self.y, self.t
        return self.x.cpu().numpy(), self.y.cpu().numpy(), self.t.cpu().numpy()

    def draw_aug(self, n, seed=0, epoch=0, rank=0):
        """Random draws of RandomCrop(H, padding=pad) (top-left corner uniform in [0, 2*pad]^2) and
        RandomHorizontalFlip(0.5) for n samples, packed dy | dx << 8 | flip << 16 (int32, on the device).  Seeded per
        (seed, epoch, rank) so a run is reproducible and resumable."""
        g = torch.Generator()
        g.manual_seed((int(seed) * 1000003 + int(epoch)) * 4099 + int(rank) + 17)
        dy = torch.randint(0, 2 * self.pad + 1, (n,), generator=g, dtype=torch.int32)
        dx = torch.randint(0, 2 * self.pad + 1, (n,), generator=g, dtype=torch.int32)
        fl = torch.randint(0, 2, (n,), generator=g, dtype=torch.int32)
        return (dy | (dx << 8) | (fl << 16)).to(self.x.device)

    def batch(self, idx, aug=None, out=None):
        """fp32 NCHW batch of samples `idx` (int64, device): train transform when `aug` is given, eval transform
        (ToTensor + Normalize only, utils.py:245-251) when it is None."""
        return ops.augment_u8(self.x, idx, aug, self.mean, self.std, self.pad, out=out)


class ClassIncrementalScenario:
    """continuum.ClassIncremental as the reference uses it (utils.py:200-206): classes are re-labelled by their
    position in class_order, task 0 holds `initial_increment` classes, every further task `increment`."""

    def __init__(self, x, y, initial_increment, increment, class_order=None, device=None, mean=IMAGENET_DEFAULT_MEAN,
                 std=IMAGENET_DEFAULT_STD):
        y = np.asarray(y)
        classes = np.unique(y)
        order = list(classes) if class_order is None else list(class_order)
        if sorted(order) != sorted(classes.tolist()):
            raise ValueError("class_order must be a permutation of the dataset's classes")
        remap = np.zeros(int(max(order)) + 1, dtype=np.int64)
        remap[np.asarray(order)] = np.arange(len(order))
        self.x, self.y = np.asarray(x), remap[y]
        self.nb_classes = len(order)
        if initial_increment <= 0:
            initial_increment = increment
        self.increments = [initial_increment]
        while sum(self.increments) < self.nb_classes:
            self.increments.append(min(increment, self.nb_classes - sum(self.increments)))
        self.bounds = np.concatenate([[0], np.cumsum(self.increments)])
        self.task_of = np.searchsorted(self.bounds, self.y, side="right") - 1
        self.device, self.mean, self.std = device, mean, std

    def __len__(self):
        return len(self.increments)

    def _make(self, mask):
        return DeviceTaskSet(self.x[mask], self.y[mask], self.task_of[mask], device=self.device, mean=self.mean,
                             std=self.std)

    def __getitem__(self, i):
        if isinstance(i, slice):
            ids = list(range(len(self)))[i]
            return self._make(np.isin(self.task_of, ids))
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return self._make(self.task_of == i)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


def synthetic_cifar(nb_classes=100, per_class=500, seed=0, noise=48.0):
    """Synthetic stand-in for CIFAR-100 (no dataset can be downloaded here): uint8 [N,32,32,3] images = a smooth
    class prototype + per-sample noise, so that a classifier can actually learn the classes.  Class prototypes depend
    on (seed) only; samples on (seed, split) through `per_class` and the split offset the caller folds into seed."""
    rng = np.random.RandomState(1234)  # prototypes are shared by the train and validation splits
    coarse = rng.uniform(40, 215, size=(nb_classes, 4, 4, 3))
    proto = np.repeat(np.repeat(coarse, 8, axis=1), 8, axis=2)  # [K,32,32,3]
    srng = np.random.RandomState(seed)
    xs, ys = [], []
    for k in range(nb_classes):
        img = proto[k][None] + srng.normal(0.0, noise, size=(per_class, 32, 32, 3))
        shift = srng.randint(-2, 3, size=(per_class, 2))
        img = np.stack([np.roll(im, tuple(sh), axis=(0, 1)) for im, sh in zip(img, shift)])
        xs.append(np.clip(img, 0, 255).astype(np.uint8))
        ys.append(np.full(per_class, k, dtype=np.int64))
    return np.concatenate(xs), np.concatenate(ys)


def load_cifar100(data_path, train):
    """The python-pickle CIFAR-100 files continuum's CIFAR100 dataset reads (utils.py:195-196); there is no network
    here, so the files must already exist under data_path/cifar-100-python/."""
    f = os.path.join(data_path, "cifar-100-python", "train" if train else "test")
    if not os.path.exists(f):
        raise FileNotFoundError(f"{f} not found (no download possible): use --data_set synthetic")
    with open(f, "rb") as fh:
        d = pickle.load(fh, encoding="latin1")
    x = np.asarray(d["data"], dtype=np.uint8).reshape(-1, 3, 32, 32).transpose(0, 2, 3, 1)
    return np.ascontiguousarray(x), np.asarray(d["fine_labels"], dtype=np.int64)


def build_dataset(is_train, args, device=None):
    """utils.py:192-208 -> (scenario, nb_classes)."""
    name = args.data_set.lower()
    mean, std = transform_stats(args)
    if name == "cifar":
        x, y = load_cifar100(args.data_path, is_train)
    elif name == "synthetic":
        per = args.synthetic_per_class if is_train else args.synthetic_val_per_class
        x, y = synthetic_cifar(getattr(args, "nb_classes", 100) or 100, per, seed=11 if is_train else 29)
    else:
        raise ValueError(f"Unknown dataset {args.data_set}.")
    order = getattr(args, "class_order", None)
    if order is not None and sorted(order) != sorted(np.unique(y).tolist()):
        order = None
    scenario = ClassIncrementalScenario(x, y, args.num_bases, args.increment, class_order=order, device=device,
                                        mean=mean, std=std)
    return scenario, scenario.nb_classes
