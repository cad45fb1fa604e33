"""Arena layout + launch plans of the native ResNet engine (wacil_net_* in include/wacil.h).

`NetArena` owns the flat fp32 arenas (parameters, gradients, BatchNorm running statistics) as torch tensors — torch is
the allocator, nothing more — and exposes reference-shaped views of them (state_dict keys of resnet.py / template.py).
`Plan` binds one (mode, batch) launch plan to those arenas plus a workspace."""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict

import torch

from . import _lib
from ._lib import check, lib

STAGE_PLANES = (16, 32, 64)


def layer_names(depth):
    """(conv_name, bn_name) per conv layer in forward order — the reference's module names (resnet.py:71-77)."""
    n = (depth - 2) // 6
    out = [("backbone.conv_1_3x3", "backbone.bn_1")]
    for s in (1, 2, 3):
        for b in range(n):
            out.append((f"backbone.stage_{s}.{b}.conv_a", f"backbone.stage_{s}.{b}.bn_a"))
            out.append((f"backbone.stage_{s}.{b}.conv_b", f"backbone.stage_{s}.{b}.bn_b"))
    return out


class NetArena:
    def __init__(self, depth, max_classes, device="cpu"):
        if (depth - 2) % 6 != 0:
            raise AssertionError("depth should be one of 20, 32, 44, 56, 110")  # resnet.py:66
        self.depth, self.max_classes = depth, int(max_classes)
        lay = _lib.NetLayout()
        check(lib().wacil_net_layout(depth, self.max_classes, C.byref(lay)), "wacil_net_layout")
        self.layout = lay
        self.layers = []
        for l in range(lay.n_layers):
            ll = _lib.LayerLayout()
            check(lib().wacil_net_layer_layout(depth, self.max_classes, l, C.byref(ll)), "wacil_net_layer_layout")
            self.layers.append(ll)
        self.names = layer_names(depth)
        self.feat_dim = lay.feat_dim
        self.params = torch.zeros(lay.param_elems, dtype=torch.float32, device=device)
        self.grads = torch.zeros(lay.param_elems, dtype=torch.float32, device=device)
        self.bn = torch.zeros(lay.bn_elems, dtype=torch.float32, device=device)
        for ll in self.layers:  # running_var starts at 1 (nn.BatchNorm2d)
            self.bn[ll.rvar_off: ll.rvar_off + ll.cout] = 1.0
            self.params[ll.gamma_off: ll.gamma_off + ll.cout] = 1.0

    # ---- reference-shaped views -------------------------------------------------------------------------
    def conv_weight(self, l, buf=None):
        """[cout, cin, 3, 3] strided view of the tap-major [9][cin_pad][cout] storage."""
        ll = self.layers[l]
        buf = self.params if buf is None else buf
        w = buf[ll.w_off: ll.w_off + 9 * ll.cin_pad * ll.cout].view(3, 3, ll.cin_pad, ll.cout)
        return w.permute(3, 2, 0, 1)[:, : ll.cin]

    def bn_weight(self, l, buf=None):
        ll = self.layers[l]
        return (self.params if buf is None else buf)[ll.gamma_off: ll.gamma_off + ll.cout]

    def bn_bias(self, l, buf=None):
        ll = self.layers[l]
        return (self.params if buf is None else buf)[ll.beta_off: ll.beta_off + ll.cout]

    def running_mean(self, l):
        ll = self.layers[l]
        return self.bn[ll.rmean_off: ll.rmean_off + ll.cout]

    def running_var(self, l):
        ll = self.layers[l]
        return self.bn[ll.rvar_off: ll.rvar_off + ll.cout]

    def fc_weight(self, buf=None):
        o = self.layout.fc_weight_off
        return (self.params if buf is None else buf)[o: o + self.max_classes * self.feat_dim].view(self.max_classes,
                                                                                                  self.feat_dim)

    def fc_bias(self, buf=None):
        o = self.layout.fc_bias_off
        return (self.params if buf is None else buf)[o: o + self.max_classes]

    def to(self, device):
        self.params, self.grads, self.bn = self.params.to(device), self.grads.to(device), self.bn.to(device)
        return self

    def clone(self):
        other = NetArena.__new__(NetArena)
        other.__dict__.update(self.__dict__)
        other.params, other.grads, other.bn = self.params.clone(), torch.zeros_like(self.grads), self.bn.clone()
        return other


class Plan:
    """One wacil_net_t: (mode, batch) bound to an arena.  Holds the workspace and static output buffers."""

    def __init__(self, arena: NetArena, mode: int, batch: int):
        if not arena.params.is_cuda:
            raise _lib.WacilError("the engine runs on CUDA tensors only; call .cuda() first (no CPU path)")
        dev = arena.params.device
        self.arena, self.mode, self.batch = arena, mode, batch
        nbytes = lib().wacil_net_workspace_bytes(arena.depth, batch, arena.max_classes, mode)
        self.workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        base = (self.workspace.data_ptr() + 1023) // 1024 * 1024
        self.handle = C.c_void_p()
        # tensors captured by the plan: keep references so the addresses stay valid
        self._keep = (arena.params, arena.grads, arena.bn)
        check(lib().wacil_net_create(C.byref(self.handle), arena.depth, batch, arena.max_classes, mode,
                                     arena.params.data_ptr(),
                                     arena.grads.data_ptr() if mode == _lib.NET_TRAIN else None,
                                     arena.bn.data_ptr(), base, nbytes), "wacil_net_create")
        self.feats = torch.empty((batch, arena.feat_dim), dtype=torch.float32, device=dev)
        self._logits = torch.empty(batch * max(arena.max_classes, 1), dtype=torch.float32, device=dev)

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            lib().wacil_net_destroy(h)
            self.handle = None

    def fold(self):
        check(lib().wacil_net_fold(self.handle, torch.cuda.current_stream().cuda_stream), "wacil_net_fold")

    def forward(self, x, c_seen):
        if x.dtype != torch.float32 or not x.is_contiguous() or tuple(x.shape) != (self.batch, 3, 32, 32):
            raise _lib.WacilError(f"expected contiguous fp32 input [{self.batch}, 3, 32, 32], got {tuple(x.shape)} "
                                  f"{x.dtype}")
        logits = self._logits[: self.batch * c_seen].view(self.batch, c_seen) if c_seen > 0 else None
        check(lib().wacil_net_forward(self.handle, x.data_ptr(), int(c_seen), self.feats.data_ptr(),
                                      logits.data_ptr() if logits is not None else None,
                                      torch.cuda.current_stream().cuda_stream), "wacil_net_forward")
        return logits, self.feats

    def backward(self, dlogits):
        if dlogits.dtype != torch.float32 or not dlogits.is_contiguous() or dlogits.shape[0] != self.batch:
            raise _lib.WacilError("dlogits must be contiguous fp32 [batch, C_seen]")
        check(lib().wacil_net_backward(self.handle, dlogits.data_ptr(), dlogits.shape[1],
                                       torch.cuda.current_stream().cuda_stream), "wacil_net_backward")

    def set_side_stream(self, stream):
        """Overlap the weight-gradient kernels of backward() on `stream` (a torch.cuda.Stream) or None."""
        self._side = stream  # keep it alive
        check(lib().wacil_net_set_side_stream(self.handle, stream.cuda_stream if stream is not None else None),
              "wacil_net_set_side_stream")

    def set_optimizer(self, mom, lr_dev, momentum, weight_decay, world=1, rank=0, grad_ptrs=None, pad_ptrs=None, epoch=None):
        """Fuse gradient mean (NVLink peer reads, world > 1) + SGD into backward(): the convolution weights of the last
        resolution group are updated while the rest of the backward still runs, everything else right after it
        (template.py:243-244,269-271).  mom=None detaches the optimiser."""
        self._opt_keep = (mom, lr_dev, grad_ptrs, pad_ptrs, epoch)
        check(lib().wacil_net_set_optimizer(self.handle, mom.data_ptr() if mom is not None else None,
                                            lr_dev.data_ptr() if lr_dev is not None else None, float(momentum),
                                            float(weight_decay), int(world), int(rank), grad_ptrs, pad_ptrs,
                                            epoch.data_ptr() if epoch is not None else None), "wacil_net_set_optimizer")

    def tensor(self, what, layer=0):
        """A view of a tensor the TRAIN plan keeps after a forward (parity tests): 0 packed input, 1 conv output,
        2 activation (all bf16 NHWC), 3 BatchNorm batch mean / rstd (fp32 [2, C])."""
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        check(lib().wacil_net_tensor(self.handle, what, layer, C.byref(ptr), C.byref(nbytes)), "wacil_net_tensor")
        off = ptr.value - self.workspace.data_ptr()
        raw = self.workspace[off: off + nbytes.value]
        if what == 3:
            return raw.view(torch.float32).view(2, -1)
        if what == 0:
            return raw.view(torch.bfloat16).view(self.batch, 32, 32, 16)
        ll = self.arena.layers[layer]
        return raw.view(torch.bfloat16).view(self.batch, ll.H_out, ll.H_out, ll.cout)

    @property
    def launches(self):
        return lib().wacil_net_last_launches(self.handle)
