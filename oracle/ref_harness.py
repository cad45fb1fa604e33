"""Import the UNMODIFIED reference (/root/reference/{resnet,utils,template}.py) on a CUDA-less host.

TEST INFRASTRUCTURE ONLY; works only where /root/reference exists (this build container, never the GPU box).
Used by oracle/make_golden.py to mint the golden vectors that pin oracle/wa_oracle.py.

The reference imports `timm` and `continuum` at module top (utils.py:11-14, template.py:8-9); neither is
installed, so empty stub modules are injected first.  The script body is under `if __name__ == "__main__"`
(template.py:191), so importing is side-effect free.  `.cuda()` is hard-coded at template.py:91,104,148 — on a
host without CUDA it is shimmed to identity for the duration of the harness.
"""
from __future__ import annotations

import contextlib
import importlib
import sys
import types

import torch.nn as nn

REFERENCE_ROOT = "/root/reference"


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs() -> None:
    if "timm" not in sys.modules:
        timm = _stub("timm")
        timm.utils = _stub("timm.utils")
        timm.data = _stub("timm.data", create_transform=lambda *a, **k: None)
        timm.data.constants = _stub("timm.data.constants", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406),
                                    IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
    if "continuum" not in sys.modules:
        cont = _stub("continuum", ClassIncremental=object)
        cont.rehearsal = _stub("continuum.rehearsal")
        cont.datasets = _stub("continuum.datasets", CIFAR100=object, ImageFolderDataset=object)


@contextlib.contextmanager
def cuda_shim():
    """Make nn.Module.cuda() the identity while the reference builds modules on a CPU-only host."""
    orig = nn.Module.cuda
    nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        nn.Module.cuda = orig


def load_reference():
    """Returns the reference's (resnet, utils, template) modules."""
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    mods = []
    for name in ("resnet", "utils", "template"):
        mods.append(importlib.import_module(name))
    return tuple(mods)


class Args:
    """The fields the reference's model code reads (template.py:72-84,152-154)."""

    def __init__(self, backbone="resnet32", task_id=0):
        self.backbone = backbone
        self.task_id = task_id
