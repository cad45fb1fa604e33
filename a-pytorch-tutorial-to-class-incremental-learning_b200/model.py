"""Drop-in mirror of the reference's model surface (template.py:61-166, utils.py:121-132) on the native engine.

    CilModel(args)            .forward(x) -> (logits, feats)   .extract_vector(x)   .feature_dim
                              .prev_model_adaption(n)  .after_model_adaption(n, args)  .weight_align(n)
                              .copy()  .freeze(names)  .parameters()  .train()/.eval()  .cuda()
                              .fc[i].weight / .bias, len(fc)   .state_dict() / .load_state_dict() (reference keys)
    CilClassifier             .forward / .adaption / __getitem__ / __len__
    SoftTarget(T)(out_s, out_t)

The arithmetic runs in libwacil.so; this file only owns tensors and call order.  Differences from the reference are
deliberate and listed in DESIGN.md (bf16 tensor-core convolutions, one flat parameter, eval-mode forward is
inference-only)."""
from __future__ import annotations

import copy as _copy
import math
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib, ops
from .engine import NetArena, Plan

_BACKBONE_DEPTH = {"resnet20": 20, "resnet32": 32, "resnet44": 44, "resnet56": 56, "resnet110": 110}


def freeze_parameters(m, requires_grad=False):
    """template.py:61-69."""
    if m is None:
        return
    if isinstance(m, nn.Parameter):
        m.requires_grad = requires_grad
    else:
        for p in m.parameters():
            p.requires_grad = requires_grad


# ---------------------------------------------------------------------------------------------------- losses
class _SoftTargetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out_s, out_t, T):
        loss, grad = ops.kd_fwd_bwd(out_s, out_t.detach(), T, need_grad=out_s.requires_grad)
        ctx.save_for_backward(grad)
        return loss.to(out_s.dtype) if out_s.dtype != torch.float32 else loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g if grad is not None else None), None, None


class SoftTarget(nn.Module):
    """utils.py:121-132: KL(softmax(t/T) || softmax(s/T)) * T^2, reduction 'batchmean'; gradient to out_s only.
    One fused CUDA kernel computes the loss and its gradient; `out_s` may be a strided column slice."""

    def __init__(self, T=2):
        super().__init__()
        self.T = T

    def forward(self, out_s, out_t):
        return _SoftTargetFn.apply(out_s, out_t, float(self.T))


# ---------------------------------------------------------------------------------------------------- classifier
class _LinearFn(torch.autograd.Function):
    """x @ W.T + b on the tcgen05 classifier kernels (wacil_fc_forward / wacil_fc_backward, kind::tf32)."""

    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(x, W)
        return ops.fc_forward(x.detach(), W.detach(), b.detach().contiguous())

    @staticmethod
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        need = ctx.needs_input_grad
        dx, dW, db = ops.fc_backward(g, x, W, need_dfeats=need[0], need_dW=need[1], need_db=need[2])
        return dx, dW, db


def _linear(x, W, b):
    lead = x.shape[:-1]
    out = _LinearFn.apply(x.reshape(-1, x.shape[-1]), W, b)
    return out.reshape(*lead, W.shape[0])


class _Head:
    """What `fc[i]` returns: live views of the classifier arena rows of task i (nn.Linear-like)."""

    def __init__(self, owner, lo, hi):
        self._owner, self.lo, self.hi = owner, lo, hi
        self.in_features, self.out_features = owner.embed_dim, hi - lo

    @property
    def weight(self):
        return self._owner._arena.fc_weight()[self.lo: self.hi]

    @property
    def bias(self):
        return self._owner._arena.fc_bias()[self.lo: self.hi]

    def __call__(self, x):
        return _linear(x, self.weight, self.bias)


class CilClassifier(nn.Module):
    """template.py:87-104.  Heads are consecutive row ranges of ONE contiguous [C_max, D] matrix, so the forward
    over all heads is a single product (the reference concatenates t+1 separate nn.Linear outputs)."""

    def __init__(self, embed_dim, nb_classes, _model=None):
        super().__init__()
        self.embed_dim = embed_dim
        object.__setattr__(self, "_model", _model)  # not a submodule: the model owns the classifier
        self._bounds = [0]
        self.adaption(nb_classes)

    @property
    def _arena(self):
        return self._model._arena

    @property
    def heads(self):
        return [self[i] for i in range(len(self))]

    @property
    def num_classes(self):
        return self._bounds[-1]

    def __getitem__(self, index):
        n = len(self)
        if index < 0:
            index += n
        if not 0 <= index < n:
            raise IndexError(index)
        return _Head(self, self._bounds[index], self._bounds[index + 1])

    def __len__(self):
        return len(self._bounds) - 1

    def forward(self, x):
        # convenience entry (not on the training hot path: CilModel.forward fuses pool + classifier natively)
        c = self.num_classes
        return _linear(x, self._arena.fc_weight()[:c], self._arena.fc_bias()[:c])

    def adaption(self, nb_classes):
        lo = self._bounds[-1]
        hi = lo + int(nb_classes)
        self._model._ensure_capacity(hi)
        lin = nn.Linear(self.embed_dim, nb_classes)  # same init and RNG consumption as template.py:91,104
        with torch.no_grad():
            self._arena.fc_weight()[lo:hi].copy_(lin.weight)
            self._arena.fc_bias()[lo:hi].copy_(lin.bias)
        self._bounds.append(hi)


# ---------------------------------------------------------------------------------------------------- model
class _Backbone:
    """`model.backbone`: callable feature extractor with `.out_dim` (resnet.py:80,107-116)."""

    def __init__(self, model):
        self._model = model
        self.out_dim = 64
        self.featureSize = 64

    def __call__(self, x):
        return self._model.extract_vector(x)


class _CilForwardFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, arena_param, model):
        plan = model._plan(_lib.NET_TRAIN, x.shape[0])
        logits, feats = plan.forward(x.contiguous(), model.num_classes)
        ctx.plan, ctx.model = plan, model
        ctx.set_materialize_grads(False)
        model._num_batches_tracked += 1
        return logits.clone(), feats.clone()

    @staticmethod
    def backward(ctx, g_logits, g_feats):
        if g_feats is not None:
            raise _lib.WacilError("gradients through the returned features are not supported by the fused backward")
        if g_logits is None:
            return None, None, None
        ctx.plan.backward(g_logits.contiguous().float())
        return None, ctx.model._arena.grads.clone(), None


class CilModel(nn.Module):
    """template.py:107-166 on the native engine.  `args.backbone` in {resnet20,32,44,56,110} (template.py:72-84)."""

    def __init__(self, args, max_classes=None):
        super().__init__()
        name = args if isinstance(args, str) else getattr(args, "backbone", "resnet32")
        if name not in _BACKBONE_DEPTH:
            raise NotImplementedError(f"Unknown backbone {name}")
        self.backbone_name = name
        self.depth = _BACKBONE_DEPTH[name]
        if max_classes is None:
            max_classes = getattr(args, "nb_classes", None) or 100
        self._arena = NetArena(self.depth, max_classes)
        self._init_backbone_like_reference()
        self.arena = nn.Parameter(self._arena.params)  # ONE flat parameter: the whole model
        self.register_buffer("bn_running", self._arena.bn)
        self.fc = None
        self.backbone = _Backbone(self)
        self._plans = {}
        self._param_version = 0
        self._num_batches_tracked = 0

    # ---- construction --------------------------------------------------------------------------------------
    def _init_backbone_like_reference(self):
        """He-normal conv weights, BN gamma=1 beta=0 (resnet.py:83-88), consuming the torch CPU RNG in the order
        the reference's constructor does (each nn.Conv2d first draws its default init, then the normal_ pass)."""
        a = self._arena
        convs = [nn.Conv2d(ll.cin, ll.cout, 3, stride=ll.stride, padding=1, bias=False) for ll in a.layers]
        with torch.no_grad():
            for l, (m, ll) in enumerate(zip(convs, a.layers)):
                m.weight.normal_(0, math.sqrt(2.0 / (9 * ll.cout)))
                a.conv_weight(l).copy_(m.weight)

    def _ensure_capacity(self, n_classes):
        if n_classes <= self._arena.max_classes:
            return
        old = self._arena
        new = NetArena(self.depth, max(n_classes, 2 * old.max_classes), device=old.params.device)
        with torch.no_grad():
            nconv = old.layout.fc_weight_off
            new.params[:nconv].copy_(old.params[:nconv])
            new.fc_weight()[: old.max_classes].copy_(old.fc_weight())
            new.fc_bias()[: old.max_classes].copy_(old.fc_bias())
            new.bn.copy_(old.bn)
        self._rebind(new)

    def _rebind(self, arena):
        self._arena = arena
        self.arena = nn.Parameter(arena.params, requires_grad=self.arena.requires_grad)
        self.bn_running = arena.bn
        self._plans = {}
        self._touch()

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        # nn.Module moved the Parameter/buffer storage: re-point the arena at it and drop stale plans
        self._arena.params = self.arena.data
        self._arena.bn = self.bn_running
        self._arena.grads = torch.zeros_like(self.arena.data)
        self._plans = {}
        self._touch()
        return self

    def _touch(self):
        """Parameters or statistics changed through a native kernel (no torch version bump): refold before inference."""
        self._param_version = getattr(self, "_param_version", 0) + 1

    # ---- reference surface ---------------------------------------------------------------------------------
    @property
    def feature_dim(self):
        return self.backbone.out_dim

    @property
    def num_classes(self):
        return 0 if self.fc is None else self.fc.num_classes

    def _plan(self, mode, batch):
        key = (mode, batch)
        p = self._plans.get(key)
        if p is None:
            p = Plan(self._arena, mode, batch)
            self._plans[key] = p
        return p

    def _infer(self, x, want_logits):
        plan = self._plan(_lib.NET_INFER, x.shape[0])
        if getattr(plan, "_folded_version", None) != self._fold_version():
            plan.fold()
            plan._folded_version = self._fold_version()
        return plan.forward(x.contiguous(), self.num_classes if want_logits else 0)

    def _fold_version(self):
        return (self.arena._version, self.bn_running._version, self._num_batches_tracked, self._param_version)

    def extract_vector(self, x):
        """template.py:117-118.  Training mode: batch-statistics forward (as the reference's backbone(x) would);
        eval mode: folded-BN inference kernels."""
        if self.training:
            return self.forward(x)[1]
        return self._infer(x, False)[1].clone()

    def forward(self, x):
        """template.py:120-123 -> (logits [B, C_seen], feats [B, 64])."""
        if self.fc is None:
            raise TypeError("'NoneType' object is not callable")  # what the reference raises before a head exists
        if self.training:
            if torch.is_grad_enabled() and self.arena.requires_grad:
                return _CilForwardFn.apply(x, self.arena, self)
            plan = self._plan(_lib.NET_TRAIN, x.shape[0])
            logits, feats = plan.forward(x.contiguous(), self.num_classes)
            self._num_batches_tracked += 1
            return logits.clone(), feats.clone()
        logits, feats = self._infer(x, True)
        return logits.clone(), feats.clone()

    def copy(self):
        """template.py:125-126 (deepcopy): independent parameters and statistics, no shared plans."""
        new = CilModel.__new__(CilModel)
        nn.Module.__init__(new)
        new.backbone_name, new.depth = self.backbone_name, self.depth
        new._arena = self._arena.clone()
        new.arena = nn.Parameter(new._arena.params, requires_grad=self.arena.requires_grad)
        new.register_buffer("bn_running", new._arena.bn)
        new.backbone = _Backbone(new)
        new._plans, new._param_version, new._num_batches_tracked = {}, 0, self._num_batches_tracked
        new._frozen = set(getattr(self, "_frozen", set()))
        new.fc = None
        if self.fc is not None:
            fc = CilClassifier.__new__(CilClassifier)
            nn.Module.__init__(fc)
            fc.embed_dim, fc._bounds = self.fc.embed_dim, list(self.fc._bounds)
            object.__setattr__(fc, "_model", new)
            new.fc = fc
        new.train(self.training)
        return new

    def __deepcopy__(self, memo):
        return self.copy()

    def freeze(self, names=["all"]):
        """template.py:128-144.  The model is one flat parameter, so partial freezing is kept as a trainable element range
        of the arena that the optimiser honours (TaskTrainer): "fc" pins the classifier rows, "backbone" the convolution
        and BatchNorm parameters.  The reference also switches the frozen part to eval(): a frozen backbone then
        normalises with its running statistics — that forward is the inference plan; training through it is not offered
        (the script only ever freezes "all", template.py:290)."""
        freeze_parameters(self, requires_grad=True)
        self.train()
        self._frozen = set()
        for name in names:
            if name in ("fc", "backbone"):
                self._frozen.add(name)
            elif name == "all":
                self._frozen = {"fc", "backbone"}
                freeze_parameters(self)
                self.eval()
            else:
                raise NotImplementedError(f"Unknown module name to freeze {name}")
        return self

    def trainable_range(self):
        """[lo, hi) element range of the flat arena the optimiser may update (see freeze)."""
        fz = getattr(self, "_frozen", set())
        n, cut = self._arena.params.numel(), self._arena.layout.fc_weight_off
        if "fc" in fz and "backbone" in fz:
            return (0, 0)
        if "fc" in fz:
            return (0, cut)
        if "backbone" in fz:
            return (cut, n)
        return (0, n)

    def prev_model_adaption(self, nb_classes):
        """template.py:146-150."""
        if self.fc is None:
            self.fc = CilClassifier(self.feature_dim, nb_classes, _model=self)
        else:
            self.fc.adaption(nb_classes)
        self._touch()

    def after_model_adaption(self, nb_classes, args):
        """template.py:152-154."""
        if args.task_id > 0:
            self.weight_align(nb_classes)

    @torch.no_grad()
    def weight_align(self, nb_new_classes, verbose=True):
        """template.py:156-166 as one fused kernel, in place on the classifier arena.  Returns gamma (0-dim tensor);
        the reference prints it (template.py:165) — `verbose` keeps that, at the cost of one host sync."""
        c = self.num_classes
        gamma = ops.weight_align_(self._arena.fc_weight()[:c], nb_new_classes)
        self._touch()
        if verbose:
            print(f"old norm / new norm ={gamma.item()}")
        return gamma

    # ---- state dict with the reference's keys ----------------------------------------------------------------
    def state_dict(self, *args, **kwargs):
        a = self._arena
        sd = OrderedDict()
        for l, (conv, bn) in enumerate(a.names):
            sd[conv + ".weight"] = a.conv_weight(l).detach().clone().contiguous()
            sd[bn + ".weight"] = a.bn_weight(l).detach().clone()
            sd[bn + ".bias"] = a.bn_bias(l).detach().clone()
            sd[bn + ".running_mean"] = a.running_mean(l).detach().clone()
            sd[bn + ".running_var"] = a.running_var(l).detach().clone()
            sd[bn + ".num_batches_tracked"] = torch.tensor(self._num_batches_tracked, dtype=torch.long)
        if self.fc is not None:
            for i in range(len(self.fc)):
                sd[f"fc.heads.{i}.weight"] = self.fc[i].weight.detach().clone()
                sd[f"fc.heads.{i}.bias"] = self.fc[i].bias.detach().clone()
        return sd

    @torch.no_grad()
    def load_state_dict(self, sd, strict=True):
        a = self._arena
        used = set()
        for l, (conv, bn) in enumerate(a.names):
            for key, dst in ((conv + ".weight", a.conv_weight(l)), (bn + ".weight", a.bn_weight(l)),
                             (bn + ".bias", a.bn_bias(l)), (bn + ".running_mean", a.running_mean(l)),
                             (bn + ".running_var", a.running_var(l))):
                if key in sd:
                    dst.copy_(sd[key])
                    used.add(key)
                elif strict:
                    raise KeyError(key)
            k = bn + ".num_batches_tracked"
            if k in sd:
                self._num_batches_tracked = int(sd[k])
                used.add(k)
        i = 0
        while f"fc.heads.{i}.weight" in sd:
            w, b = sd[f"fc.heads.{i}.weight"], sd[f"fc.heads.{i}.bias"]
            if self.fc is None or i >= len(self.fc):
                self.prev_model_adaption(w.shape[0])
            self.fc[i].weight.copy_(w)
            self.fc[i].bias.copy_(b)
            used.update((f"fc.heads.{i}.weight", f"fc.heads.{i}.bias"))
            i += 1
        if strict:
            extra = set(sd) - used
            if extra:
                raise KeyError(f"unexpected keys: {sorted(extra)[:4]}...")
        self._touch()
        return self
