"""CPU oracle for the WA class-incremental training step.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU in torch fp32/fp64, the arithmetic of the reference's hot path so the CUDA
path can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; the product package never does.

Pinning: ``oracle/make_golden.py`` runs the *real* reference (``/root/reference/{resnet,template,utils}.py``
imported unmodified through ``oracle/ref_harness.py``) on seeded inputs and stores its outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against those files.
The herding part follows ``continuum==1.2.2`` (requirements.txt:1), whose source is NOT in /root/reference:
it is restated from its published algorithm and is "parity unpinned" (no reference test or vector exists).

Reference map (file:line in /root/reference):
  backbone_forward      resnet.py:107-116 (CifarResNet.forward), :40-53 (basic block), :15-17 (DownsampleA)
  classifier_forward    template.py:99-101 (cat of per-task nn.Linear heads)
  soft_target           utils.py:127-132 (SoftTarget.forward)
  weight_align          template.py:156-166
  sgd_step / cosine_lr  template.py:246-249 (torch.optim.SGD momentum+wd, CosineAnnealingLR stepped per epoch)
  train_step            template.py:255-271
  accuracy              timm.utils.accuracy as called at template.py:267 (timm==0.5.4, not installed)
  herd_barycenter_np    continuum.rehearsal.herd_closest_to_barycenter (call site template.py:300-302)
  RehearsalMemoryOracle continuum.rehearsal.RehearsalMemory (call sites template.py:212-216,231,300-302)
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
STAGE_PLANES = (16, 32, 64)


# ----------------------------------------------------------------------------------------------------------
# parameter naming (identical to the reference's state_dict keys) and seeded construction
# ----------------------------------------------------------------------------------------------------------
def blocks_per_stage(depth: int) -> int:
    assert (depth - 2) % 6 == 0, "depth should be one of 20, 32, 44, 56, 110"  # resnet.py:66
    return (depth - 2) // 6


def conv_bn_names(depth: int) -> List[Tuple[str, str, int, int, int]]:
    """[(conv_name, bn_name, cin, cout, stride)] in forward order, names as in the reference state_dict."""
    n = blocks_per_stage(depth)
    out = [("backbone.conv_1_3x3", "backbone.bn_1", 3, 16, 1)]
    inpl = 16
    for s, planes in enumerate(STAGE_PLANES, start=1):
        for b in range(n):
            stride = 2 if (s > 1 and b == 0) else 1
            pre = f"backbone.stage_{s}.{b}"
            out.append((f"{pre}.conv_a", f"{pre}.bn_a", inpl, planes, stride))
            out.append((f"{pre}.conv_b", f"{pre}.bn_b", planes, planes, 1))
            inpl = planes
    return out


def make_state(depth: int, head_sizes: Sequence[int], seed: int, dtype=torch.float32,
               perturb_bn: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """Seeded parameters/buffers with the reference's shapes and init scales (resnet.py:83-88; nn.Linear default
    for heads, template.py:91,104).  numpy RandomState keeps the stream stable across library versions.
    ``perturb_bn`` moves gamma/beta/running stats off their init values so tests can see them."""
    rs = np.random.RandomState(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for conv, bn, cin, cout, _ in conv_bn_names(depth):
        std = math.sqrt(2.0 / (9 * cout))
        sd[conv + ".weight"] = torch.from_numpy(rs.normal(0, std, size=(cout, cin, 3, 3))).to(dtype)
        if perturb_bn:
            sd[bn + ".weight"] = torch.from_numpy(rs.uniform(0.6, 1.4, size=(cout,))).to(dtype)
            sd[bn + ".bias"] = torch.from_numpy(rs.normal(0, 0.1, size=(cout,))).to(dtype)
            sd[bn + ".running_mean"] = torch.from_numpy(rs.normal(0, 0.2, size=(cout,))).to(dtype)
            sd[bn + ".running_var"] = torch.from_numpy(rs.uniform(0.5, 1.5, size=(cout,))).to(dtype)
        else:
            sd[bn + ".weight"] = torch.ones(cout, dtype=dtype)
            sd[bn + ".bias"] = torch.zeros(cout, dtype=dtype)
            sd[bn + ".running_mean"] = torch.zeros(cout, dtype=dtype)
            sd[bn + ".running_var"] = torch.ones(cout, dtype=dtype)
        sd[bn + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    d = 64
    bound = 1.0 / math.sqrt(d)
    for i, nb in enumerate(head_sizes):
        sd[f"fc.heads.{i}.weight"] = torch.from_numpy(rs.uniform(-bound, bound, size=(nb, d))).to(dtype)
        sd[f"fc.heads.{i}.bias"] = torch.from_numpy(rs.uniform(-bound, bound, size=(nb,))).to(dtype)
    return sd


def trainable_names(sd: Dict[str, torch.Tensor]) -> List[str]:
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


def num_heads(sd: Dict[str, torch.Tensor]) -> int:
    return sum(1 for k in sd if k.startswith("fc.heads.") and k.endswith(".weight"))


# ----------------------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------------------
def batch_norm(x: torch.Tensor, sd: Dict[str, torch.Tensor], bn: str, train: bool,
               new_buffers: Optional[Dict[str, torch.Tensor]]) -> torch.Tensor:
    """nn.BatchNorm2d defaults (resnet.py:31,35,73): train = batch mean / biased var for normalisation,
    running stats updated with momentum 0.1 and the UNBIASED variance; eval = running stats."""
    g, b = sd[bn + ".weight"], sd[bn + ".bias"]
    if train:
        m = x.shape[0] * x.shape[2] * x.shape[3]
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if new_buffers is not None:
            with torch.no_grad():
                rm, rv = sd[bn + ".running_mean"], sd[bn + ".running_var"]
                new_buffers[bn + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.detach()
                new_buffers[bn + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var.detach() * (m / (m - 1))
                new_buffers[bn + ".num_batches_tracked"] = sd[bn + ".num_batches_tracked"] + 1
    else:
        mean, var = sd[bn + ".running_mean"], sd[bn + ".running_var"]
    xhat = (x - mean[None, :, None, None]) * torch.rsqrt(var + BN_EPS)[None, :, None, None]
    return xhat * g[None, :, None, None] + b[None, :, None, None]


def downsample_a(x: torch.Tensor) -> torch.Tensor:
    """resnet.py:15-17: AvgPool2d(kernel 1, stride 2) == take the even pixels; cat(x, 0*x) == zero-pad channels."""
    x = x[:, :, ::2, ::2]
    return torch.cat((x, torch.zeros_like(x)), dim=1)


def backbone_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, depth: int, train: bool,
                     new_buffers: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    layers = conv_bn_names(depth)
    conv, bn, _, _, _ = layers[0]
    h = F.relu(batch_norm(F.conv2d(x, sd[conv + ".weight"], padding=1), sd, bn, train, new_buffers))
    i = 1
    while i < len(layers):
        conv_a, bn_a, cin, cout, stride = layers[i]
        conv_b, bn_b, _, _, _ = layers[i + 1]
        y = F.conv2d(h, sd[conv_a + ".weight"], stride=stride, padding=1)
        y = F.relu(batch_norm(y, sd, bn_a, train, new_buffers))
        y = batch_norm(F.conv2d(y, sd[conv_b + ".weight"], padding=1), sd, bn_b, train, new_buffers)
        res = downsample_a(h) if (stride != 1 or cin != cout) else h  # resnet.py:95-97
        h = F.relu(res + y)
        i += 2
    assert h.shape[2] == 8 and h.shape[3] == 8, "AvgPool2d(8) + Linear(64) need 32x32 inputs (resnet.py:79-80)"
    return h.mean(dim=(2, 3))


def classifier_forward(sd: Dict[str, torch.Tensor], feats: torch.Tensor) -> torch.Tensor:
    outs = [feats @ sd[f"fc.heads.{i}.weight"].t() + sd[f"fc.heads.{i}.bias"] for i in range(num_heads(sd))]
    return torch.cat(outs, dim=1)


def cil_forward(sd, x, depth, train, new_buffers=None) -> Tuple[torch.Tensor, torch.Tensor]:
    feats = backbone_forward(sd, x, depth, train, new_buffers)
    return classifier_forward(sd, feats), feats


# ----------------------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------------------
def soft_target(out_s: torch.Tensor, out_t: torch.Tensor, T: float) -> torch.Tensor:
    """utils.py:127-132.  Closed form of kl_div(log_softmax(s/T), softmax(t/T), 'batchmean') * T^2:
    (T^2/B) * sum_b sum_c p_t (log p_t - log p_s), with 0*log 0 := 0."""
    log_ps = torch.log_softmax(out_s / T, dim=1)
    log_pt = torch.log_softmax(out_t / T, dim=1)
    pt = log_pt.exp()
    term = torch.where(pt > 0, pt * (log_pt - log_ps), torch.zeros_like(pt))
    return term.sum() / out_s.shape[0] * (T * T)


def soft_target_grad(out_s: torch.Tensor, out_t: torch.Tensor, T: float) -> torch.Tensor:
    """d soft_target / d out_s = (T/B) (softmax(s/T) - softmax(t/T))."""
    return (torch.softmax(out_s / T, dim=1) - torch.softmax(out_t / T, dim=1)) * (T / out_s.shape[0])


def cross_entropy(logits: torch.Tensor, target: torch.Tensor, smooth: float = 0.0) -> torch.Tensor:
    """nn.CrossEntropyLoss(label_smoothing=smooth), mean reduction (template.py:219,259)."""
    logp = torch.log_softmax(logits, dim=1)
    nll = -logp.gather(1, target[:, None]).squeeze(1)
    if smooth > 0:
        nll = (1 - smooth) * nll + smooth * (-logp.mean(dim=1))
    return nll.mean()


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1,)) -> List[torch.Tensor]:
    """timm.utils.accuracy (timm 0.5.4): percentage of rows whose target is among the k largest logits."""
    maxk = min(max(topk), output.shape[1])
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.reshape(1, -1).expand_as(pred.t()))
    return [correct[: min(k, maxk)].reshape(-1).float().sum(0) * 100.0 / target.shape[0] for k in topk]


# ----------------------------------------------------------------------------------------------------------
# weight align, SGD, schedule
# ----------------------------------------------------------------------------------------------------------
def weight_align(head_weights: Sequence[torch.Tensor], nb_new: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """template.py:156-166: gamma = mean ||w_old rows|| / mean ||w_new rows||; returns (gamma, gamma * w_new)."""
    w = torch.cat(list(head_weights), dim=0)
    norms = torch.sqrt((w * w).sum(dim=1))
    gamma = norms[:-nb_new].mean() / norms[-nb_new:].mean()
    return gamma, gamma * w[-nb_new:]


def sgd_step(params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], bufs: Dict[str, torch.Tensor],
             lr: float, momentum: float = 0.9, weight_decay: float = 5e-4) -> None:
    """torch.optim.SGD(momentum, weight_decay), dampening 0, no nesterov (template.py:246-247), in place.
    First step: buf = g (no momentum history)."""
    with torch.no_grad():
        for k, p in params.items():
            g = grads[k] + weight_decay * p
            if k in bufs:
                bufs[k] = momentum * bufs[k] + g
            else:
                bufs[k] = g.clone()
            p -= lr * bufs[k]


def cosine_lr(base_lr: float, epoch: int, t_max: int) -> float:
    """CosineAnnealingLR(T_max=num_epochs), eta_min 0, value in effect DURING `epoch` (template.py:248-249,278)."""
    return 0.5 * base_lr * (1.0 + math.cos(math.pi * epoch / t_max))


# ----------------------------------------------------------------------------------------------------------
# the training step (template.py:255-271)
# ----------------------------------------------------------------------------------------------------------
def train_step(sd: Dict[str, torch.Tensor], teacher_sd: Optional[Dict[str, torch.Tensor]], x: torch.Tensor,
               y: torch.Tensor, depth: int, known_classes: int, lr: float, mom_bufs: Dict[str, torch.Tensor],
               lambda_kd: float = 0.5, T: float = 2.0, momentum: float = 0.9, weight_decay: float = 5e-4,
               smooth: float = 0.0, world_size: int = 1, apply_update: bool = True):
    """One optimisation step on one rank's minibatch.  Returns a dict with logits, feats, losses and grads
    (grads divided by world_size as DDP's allreduce-mean would after summing identical replicas)."""
    names = trainable_names(sd)
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    work = dict(sd)
    work.update(leaves)
    new_buffers: Dict[str, torch.Tensor] = {}
    logits, feats = cil_forward(work, x, depth, train=True, new_buffers=new_buffers)
    loss_ce = cross_entropy(logits, y, smooth)
    if teacher_sd is not None:
        with torch.no_grad():
            t_logits, _ = cil_forward(teacher_sd, x, depth, train=False)
        loss_kd = lambda_kd * soft_target(logits[:, :known_classes], t_logits, T)
    else:
        t_logits = None
        loss_kd = torch.zeros((), dtype=logits.dtype)
    loss = loss_ce + loss_kd
    glist = torch.autograd.grad(loss, [leaves[k] for k in names])
    grads = {k: g for k, g in zip(names, glist)}
    if apply_update:
        params = {k: sd[k] for k in names}
        sgd_step(params, grads, mom_bufs, lr, momentum, weight_decay)
        for k, v in new_buffers.items():
            sd[k] = v
    return {"logits": logits.detach(), "feats": feats.detach(), "t_logits": t_logits, "loss_ce": loss_ce.detach(),
            "loss_kd": loss_kd.detach(), "loss": loss.detach(), "grads": grads, "new_buffers": new_buffers}


# ----------------------------------------------------------------------------------------------------------
# herding + rehearsal memory (continuum 1.2.2 restated; parity unpinned)
# ----------------------------------------------------------------------------------------------------------
def herd_barycenter_np(x, y, t, features: np.ndarray, nb_per_class: int):
    """iCaRL barycenter herding as continuum.rehearsal.herd_closest_to_barycenter does it (NumPy, float32
    features): per class, L2-normalise each sample's feature (+1e-8), mu = mean; greedily pick
    argmax(w . D), w += mu - D[:, pick]; a re-picked sample advances w but not the rank; at most 1000
    iterations; unpicked samples rank last in index order."""
    if features.ndim != 2:
        raise ValueError(f"Expected features to have 2 dimensions, not {features.ndim}d.")
    indexes = []
    for class_id in np.unique(y):
        class_indexes = np.where(y == class_id)[0]
        class_features = features[class_indexes]
        D = class_features.T
        D = D / (np.linalg.norm(D, axis=0) + 1e-8)
        mu = np.mean(D, axis=1)
        herding_matrix = np.zeros((class_features.shape[0],))
        w_t = mu
        iter_herding, iter_herding_eff = 0, 0
        while not (np.sum(herding_matrix != 0) == min(nb_per_class, class_features.shape[0])) \
                and iter_herding_eff < 1000:
            tmp_t = np.dot(w_t, D)
            ind_max = np.argmax(tmp_t)
            iter_herding_eff += 1
            if herding_matrix[ind_max] == 0:
                herding_matrix[ind_max] = 1 + iter_herding
                iter_herding += 1
            w_t = w_t + mu - D[:, ind_max]
        herding_matrix[np.where(herding_matrix == 0)[0]] = 10000
        tmp_indexes = herding_matrix.argsort(kind="stable")[:nb_per_class]
        indexes.append(class_indexes[tmp_indexes])
    indexes = np.concatenate(indexes)
    return x[indexes], y[indexes], t[indexes], indexes


def herd_barycenter_seq(features: np.ndarray, y: np.ndarray, nb_per_class: int) -> np.ndarray:
    """Same selection with a FIXED float32 evaluation order (sequential sums over the feature dimension and over
    samples, separate multiply and add, first index wins ties) — the order oracle/herding.c and the CUDA kernel
    use, so indices can be compared bit-exactly.  Pure NumPy scalar loops: small cases only."""
    f32 = np.float32
    out = []
    for class_id in np.unique(y):
        ci = np.where(y == class_id)[0]
        n, d = len(ci), features.shape[1]
        Dn = np.zeros((n, d), dtype=f32)
        for i in range(n):
            ss = f32(0)
            for k in range(d):
                ss = f32(ss + f32(features[ci[i], k] * features[ci[i], k]))
            nrm = f32(f32(np.sqrt(ss)) + f32(1e-8))
            for k in range(d):
                Dn[i, k] = f32(features[ci[i], k] / nrm)
        mu = np.zeros(d, dtype=f32)
        for k in range(d):
            s = f32(0)
            for i in range(n):
                s = f32(s + Dn[i, k])
            mu[k] = f32(s / f32(n))
        rank = np.zeros(n, dtype=np.int64)
        w = mu.copy()
        it = eff = 0
        quota = min(nb_per_class, n)
        while it != quota and eff < 1000:
            best, bi = None, -1
            for i in range(n):
                s = f32(0)
                for k in range(d):
                    s = f32(s + f32(w[k] * Dn[i, k]))
                if best is None or s > best:
                    best, bi = s, i
            eff += 1
            if rank[bi] == 0:
                rank[bi] = 1 + it
                it += 1
            for k in range(d):
                w[k] = f32(f32(w[k] + mu[k]) - Dn[bi, k])
        rank[rank == 0] = 10000
        out.append(ci[np.argsort(rank, kind="stable")[:nb_per_class]])
    return np.concatenate(out)


def herd_cluster_np(x, y, t, features: np.ndarray, nb_per_class: int):
    """continuum.rehearsal.herd_closest_to_cluster (--herding_method cluster, template.py:34,214; continuum==1.2.2 is
    not vendored: PARITY UNPINNED): per class the samples closest to the class-mean feature, nearest first."""
    if features.ndim != 2:
        raise ValueError(f"Expected features to have 2 dimensions, not {features.ndim}d.")
    indexes = []
    for class_id in np.unique(y):
        class_indexes = np.where(y == class_id)[0]
        class_features = features[class_indexes]
        class_mean = np.mean(class_features, axis=0, keepdims=True)
        dist_to_mean = np.linalg.norm(class_mean - class_features, axis=1)
        tmp_indexes = dist_to_mean.argsort(kind="stable")[:nb_per_class]
        indexes.append(class_indexes[tmp_indexes])
    indexes = np.concatenate(indexes)
    return x[indexes], y[indexes], t[indexes], indexes


def herd_random_np(x, y, t, features, nb_per_class: int, rng=np.random):
    """continuum.rehearsal.herd_random (--herding_method random): nb_per_class samples per class, uniformly without
    replacement, from NumPy's global RNG (the reference seeds it, template.py:54)."""
    indexes = []
    for class_id in np.unique(y):
        class_indexes = np.where(y == class_id)[0]
        indexes.append(rng.choice(class_indexes, size=min(nb_per_class, len(class_indexes)), replace=False))
    indexes = np.concatenate(indexes)
    return x[indexes], y[indexes], t[indexes], indexes


# ---------------------------------------------------------------------------------------------------- data path
IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
CIFAR_MEAN, CIFAR_STD = (0.5071, 0.4867, 0.4408), (0.2675, 0.2565, 0.2761)


def transform_stats(data_set: str, input_size: int = 32):
    """The Normalize() the reference ends up with (utils.py:228-230,246-251): the CIFAR statistics only when
    args.data_set == 'CIFAR' (upper case) — with the default --data_set cifar the ImageNet defaults stay in place."""
    if input_size == 32 and data_set == "CIFAR":
        return CIFAR_MEAN, CIFAR_STD
    return IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD


def augment_u8_np(store: np.ndarray, idx, aug, mean, std, pad: int = 4) -> np.ndarray:
    """CPU restatement of the CIFAR transform chain of utils.py:210-251 for explicit random draws:
    RandomCrop(32, padding=4) (zero fill; top-left corner (dy, dx) in the padded image), RandomHorizontalFlip,
    ToTensor (uint8 -> float32 / 255), Normalize ((x - mean) / std, float32).  store uint8 [n,H,W,3]; idx [B] or None;
    aug int32 [B] = dy | dx << 8 | flip << 16 or None (centre crop, no flip = the eval transform, utils.py:245-251).
    RandAugment / colour jitter of timm's create_transform are host-side PIL operations outside this path."""
    n, H, W, _ = store.shape
    idx = np.arange(n) if idx is None else np.asarray(idx)
    B = len(idx)
    aug = np.full(B, pad | (pad << 8), dtype=np.int32) if aug is None else np.asarray(aug)
    out = np.empty((B, 3, H, W), dtype=np.float32)
    m = np.asarray(mean, dtype=np.float32).reshape(3, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(3, 1, 1)
    for b in range(B):
        dy, dx, flip = int(aug[b]) & 0xFF, (int(aug[b]) >> 8) & 0xFF, (int(aug[b]) >> 16) & 1
        padded = np.zeros((H + 2 * pad, W + 2 * pad, 3), dtype=np.uint8)
        padded[pad:pad + H, pad:pad + W] = store[idx[b]]
        crop = padded[dy:dy + H, dx:dx + W]
        if flip:
            crop = crop[:, ::-1]
        ten = crop.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
        out[b] = (ten - m) / s
    return out


def eval_meter(logits: torch.Tensor, target: torch.Tensor, topk: int = 5) -> List[float]:
    """What one iteration of eval() adds to its meters (template.py:174-184; utils.py:77-86): [samples, batches, batch
    mean CrossEntropy, top-1 hits, top-min(topk, C) hits]; acc = 100 * hits / samples at the end (global_avg)."""
    b, c = logits.shape
    k = min(topk, c)
    loss = F.cross_entropy(logits, target)
    a1, ak = accuracy(logits, target, topk=(1, k))
    return [float(b), 1.0, float(loss), float(a1) * b / 100.0, float(ak) * b / 100.0]


class RehearsalMemoryOracle:
    """continuum.rehearsal.RehearsalMemory restated (ctor template.py:212-216; get :231; add :300-302)."""

    def __init__(self, memory_size: int, herding_method="barycenter", fixed_memory: bool = False,
                 nb_total_classes: Optional[int] = None):
        if fixed_memory and nb_total_classes is None:
            raise ValueError("When memory is fixed, precise the `nb_total_classes` argument.")
        if herding_method != "barycenter" and not callable(herding_method):
            raise NotImplementedError(herding_method)
        self.memory_size, self.fixed_memory, self.nb_total_classes = memory_size, fixed_memory, nb_total_classes
        self.herding_method = herding_method
        self.seen_classes = set()
        self._x = self._y = self._t = None

    @property
    def nb_classes(self) -> int:
        return len(self.seen_classes)

    @property
    def memory_per_class(self) -> int:
        if self.fixed_memory:
            return self.memory_size // self.nb_total_classes
        if self.nb_classes > 0:
            return self.memory_size // self.nb_classes
        return self.memory_size

    def __len__(self):
        return 0 if self._x is None else len(self._x)

    def get(self):
        return self._x, self._y, self._t

    def _reduce(self):
        xs, ys, ts = [], [], []
        for class_id in np.unique(self._y):
            idx = np.where(self._y == class_id)[0][: self.memory_per_class]
            xs.append(self._x[idx]); ys.append(self._y[idx]); ts.append(self._t[idx])
        self._x, self._y, self._t = np.concatenate(xs), np.concatenate(ys), np.concatenate(ts)

    def add(self, x, y, t, z):
        for c in np.unique(y):
            self.seen_classes.add(int(c))
        if callable(self.herding_method):
            mx, my, mt = self.herding_method(x, y, t, z, nb_per_class=self.memory_per_class)
        else:
            mx, my, mt, _ = herd_barycenter_np(x, y, t, z, self.memory_per_class)
        if self._x is None:
            self._x, self._y, self._t = mx, my, mt
        else:
            if not self.fixed_memory:
                self._reduce()
            self._x = np.concatenate((self._x, mx))
            self._y = np.concatenate((self._y, my))
            self._t = np.concatenate((self._t, mt))
