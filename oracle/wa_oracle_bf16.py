"""Mixed-precision model of the training step: the oracle's arithmetic with bf16 rounding inserted at the points
where the CUDA path stores bf16.  TEST INFRASTRUCTURE ONLY.

Why it exists: the reference computes in fp32; the B200 path keeps fp32 master weights, BatchNorm statistics,
classifier and loss but feeds the tensor cores bf16 (BASELINE.json config 2).  Against the pure-fp32 oracle the
*forward* agrees to ~0.3 %, but parameter gradients differ by several per cent at small batch because a ReLU whose
pre-activation is within rounding distance of zero flips its mask and injects a full-size gradient element.  That is a
property of the precision, not of the kernels — so gradient parity is asserted against THIS model (same math, same
rounding points, explicit backward formulas), and this model is tied to the fp32 oracle on the CPU: with rounding
disabled it must reproduce `wa_oracle.train_step`'s autograd gradients (tests/test_oracle_golden.py).

Rounding points (DESIGN.md "Numerics"): network input, conv weights (operand copies), every conv output `y`, every
post-ReLU activation, every gradient wrt a conv output (`gy`) and the shortcut gradient (`gz`); the data-gradient
convolution's result is consumed in fp32 (fused with the BatchNorm backward of the layer below).
fp32: accumulators, batch statistics, x-hat, d-gamma/d-beta, weight gradients, features, logits, loss.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch.nn.grad import conv2d_input, conv2d_weight

import wa_oracle as wo


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(t.dtype)


def identity(t: torch.Tensor) -> torch.Tensor:
    return t


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, depth: int, rnd=bf16_round):
    """Train-mode forward.  Returns the state the backward needs: per layer the conv input, raw output y, batch
    mean / rstd and post-ReLU activation (NCHW), plus feats, logits and the new running statistics."""
    layers = wo.conv_bn_names(depth)
    a_in = rnd(x)
    st = {"ins": [], "ys": [], "means": [], "rstds": [], "acts": [], "new_buffers": {}}
    block_in = None
    for l, (conv, bn, cin, cout, stride) in enumerate(layers):
        w = rnd(sd[conv + ".weight"])
        y = rnd(F.conv2d(a_in, w, stride=stride, padding=1))
        m = y.shape[0] * y.shape[2] * y.shape[3]
        mean = y.mean(dim=(0, 2, 3))
        var = ((y * y).mean(dim=(0, 2, 3)) - mean * mean).clamp_min(0)
        rstd = torch.rsqrt(var + wo.BN_EPS)
        st["new_buffers"][bn + ".running_mean"] = 0.9 * sd[bn + ".running_mean"] + 0.1 * mean
        st["new_buffers"][bn + ".running_var"] = 0.9 * sd[bn + ".running_var"] + 0.1 * var * (m / (m - 1))
        xhat = (y - mean[None, :, None, None]) * rstd[None, :, None, None]
        z = xhat * sd[bn + ".weight"][None, :, None, None] + sd[bn + ".bias"][None, :, None, None]
        is_b = l >= 2 and l % 2 == 0
        if is_b:
            ca = layers[l - 1]
            z = z + (wo.downsample_a(block_in) if (ca[4] != 1 or ca[2] != ca[3]) else block_in)
        a = rnd(F.relu(z))
        st["ins"].append(a_in); st["ys"].append(y); st["means"].append(mean); st["rstds"].append(rstd)
        st["acts"].append(a)
        if l == 0 or is_b:
            block_in = a
        a_in = a
    st["feats"] = st["acts"][-1].mean(dim=(2, 3))
    n_heads = wo.num_heads(sd)
    W = torch.cat([sd[f"fc.heads.{i}.weight"] for i in range(n_heads)])
    b = torch.cat([sd[f"fc.heads.{i}.bias"] for i in range(n_heads)])
    st["logits"] = st["feats"] @ W.t() + b
    return st


def backward(sd: Dict[str, torch.Tensor], st, depth: int, dlogits: torch.Tensor, rnd=bf16_round,
             round_dgrad: bool = False):
    """Backward with explicit formulas from a forward state `st` (from `forward`, or read back from the GPU plan:
    with the masks fixed by the given activations the map dlogits -> gradients is linear, so the comparison with the
    kernels is free of the rounding avalanche of a deep bf16 forward).  `round_dgrad=False`: the data-gradient conv
    is fused with the ReLU mask + BatchNorm backward of the layer below and consumed in fp32 straight from TMEM."""
    layers = wo.conv_bn_names(depth)
    nl = len(layers)
    feats = st["feats"]
    n_heads = wo.num_heads(sd)
    W = torch.cat([sd[f"fc.heads.{i}.weight"] for i in range(n_heads)])
    grads = {}
    dW = dlogits.t() @ feats
    db = dlogits.sum(0)
    lo = 0
    for i in range(n_heads):
        nb = sd[f"fc.heads.{i}.weight"].shape[0]
        grads[f"fc.heads.{i}.weight"], grads[f"fc.heads.{i}.bias"] = dW[lo: lo + nb], db[lo: lo + nb]
        lo += nb
    dfeats = dlogits @ W
    acts, ys, ins = st["acts"], st["ys"], st["ins"]
    hw = acts[-1].shape[2] * acts[-1].shape[3]
    g = (dfeats / hw)[:, :, None, None].expand_as(acts[-1])  # fp32, as the kernel reads it
    gz_next, gz_next_mode = None, 0  # bf16 masked gradient of the following block's bn_b (shortcut gradient)
    for l in range(nl - 1, -1, -1):
        conv, bn, cin, cout, stride = layers[l]
        g_total = g
        if ((l == 0) or (l % 2 == 0)) and l + 2 < nl and gz_next is not None:
            if gz_next_mode == 1:
                g_total = g_total + gz_next
            else:  # option-A shortcut transposed: lands on the even pixels of the first C channels
                up = torch.zeros_like(g_total)
                up[:, :, ::2, ::2] = gz_next[:, : g_total.shape[1]]
                g_total = g_total + up
        gz = g_total * (acts[l] > 0)
        m = gz.shape[0] * gz.shape[2] * gz.shape[3]
        xhat = (ys[l] - st["means"][l][None, :, None, None]) * st["rstds"][l][None, :, None, None]
        dbeta = gz.sum(dim=(0, 2, 3))
        dgamma = (gz * xhat).sum(dim=(0, 2, 3))
        grads[bn + ".bias"], grads[bn + ".weight"] = dbeta, dgamma
        k = sd[bn + ".weight"] * st["rstds"][l]
        gy = rnd(k[None, :, None, None] * (gz - dbeta[None, :, None, None] / m - xhat * dgamma[None, :, None, None] / m))
        if l >= 2 and l % 2 == 0:
            gz_next = rnd(gz)
            ca = layers[l - 1]
            gz_next_mode = 2 if (ca[4] != 1 or ca[2] != ca[3]) else 1
        grads[conv + ".weight"] = conv2d_weight(ins[l], sd[conv + ".weight"].shape, gy, stride=stride, padding=1)
        if l > 0:
            g = conv2d_input(ins[l].shape, rnd(sd[conv + ".weight"]), gy, stride=stride, padding=1)
            if round_dgrad:
                g = rnd(g)
    return grads


def forward_backward(sd: Dict[str, torch.Tensor], x: torch.Tensor, depth: int, dlogits_fn, rnd=bf16_round):
    st = forward(sd, x, depth, rnd)
    dlogits = dlogits_fn(st["logits"])
    grads = backward(sd, st, depth, dlogits, rnd)
    return {"logits": st["logits"], "feats": st["feats"], "grads": grads, "new_buffers": st["new_buffers"],
            "dlogits": dlogits}


def ce_kd_dlogits(y: torch.Tensor, t_logits: Optional[torch.Tensor], known: int, lambda_kd: float = 0.5,
                  T: float = 2.0):
    """dlogits of CE(mean) + lambda * SoftTarget on the first `known` columns (template.py:259-266)."""
    def fn(logits):
        B = logits.shape[0]
        d = torch.softmax(logits, dim=1)
        d[torch.arange(B), y] -= 1.0
        d = d / B
        if t_logits is not None and known > 0:
            d[:, :known] += lambda_kd * wo.soft_target_grad(logits[:, :known], t_logits, T)
        return d
    return fn
