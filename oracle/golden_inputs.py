"""Seeded inputs shared by oracle/make_golden.py (which feeds them to the real reference) and the tests
(which feed them to the oracle and to the CUDA path).  TEST INFRASTRUCTURE ONLY.

Everything is generated from numpy RandomState streams so the golden .npz files only need to store outputs."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import wa_oracle  # noqa: E402

# name -> case.  `heads` are the student's head sizes; the teacher has all but the last.
TRAIN_CASES = {
    "r32_task1": dict(depth=32, batch=16, heads=[10, 10], known=10, steps=2, lr=0.02, lambda_kd=0.5, seed=11),
    "r20_task2": dict(depth=20, batch=12, heads=[10, 10, 10], known=20, steps=2, lr=0.05, lambda_kd=0.5, seed=23),
}


def FULL_GRAD_KEYS(depth):
    n = wa_oracle.blocks_per_stage(depth)
    return ["backbone.conv_1_3x3.weight", "backbone.bn_1.weight", "backbone.bn_1.bias",
            "backbone.stage_1.0.conv_a.weight", "backbone.stage_2.0.conv_a.weight",
            f"backbone.stage_3.{n - 1}.bn_b.weight", f"backbone.stage_3.{n - 1}.bn_b.bias",
            "fc.heads.0.weight", "fc.heads.0.bias", "fc.heads.1.weight", "fc.heads.1.bias"]


def FULL_PARAM_KEYS(depth):
    n = wa_oracle.blocks_per_stage(depth)
    return ["backbone.conv_1_3x3.weight", "backbone.bn_1.running_mean", "backbone.bn_1.running_var",
            f"backbone.stage_3.{n - 1}.bn_b.running_mean", f"backbone.stage_3.{n - 1}.bn_b.running_var",
            "fc.heads.0.weight", "fc.heads.1.weight", "fc.heads.1.bias"]


def student_state(case, dtype=torch.float32):
    return wa_oracle.make_state(case["depth"], case["heads"], case["seed"], dtype)


def teacher_state(case, dtype=torch.float32):
    return wa_oracle.make_state(case["depth"], case["heads"][:-1], case["seed"] + 1000, dtype)


def batch(case, step, dtype=torch.float32):
    rs = np.random.RandomState(case["seed"] * 7919 + step)
    x = rs.normal(0, 1, size=(case["batch"], 3, 32, 32))
    y = rs.randint(0, sum(case["heads"]), size=(case["batch"],))
    return torch.from_numpy(x).to(dtype), torch.from_numpy(y).long()


# (B, C, T, stride_pad): stride_pad > 0 means the student logits are the first C columns of a [B, C+pad] matrix
KD_CASES = [(8, 10, 2.0, 10), (16, 90, 2.0, 10), (5, 513, 4.0, 0), (3, 1000, 1.0, 100), (128, 50, 2.0, 10),
            (2, 8192, 8.0, 0)]


def kd_inputs(i, dtype=torch.float32):
    B, C, T, pad = KD_CASES[i]
    rs = np.random.RandomState(500 + i)
    s = rs.normal(0, 2.0, size=(B, C + pad))
    t = rs.normal(0, 2.0, size=(B, C))
    return torch.from_numpy(s).to(dtype), torch.from_numpy(t).to(dtype)


# (head sizes, feature dim)
WA_CASES = [([10, 10], 64), ([10] * 10, 64), ([50, 10, 10], 64), ([100, 100, 100], 256)]


def wa_inputs(i, dtype=torch.float32):
    heads, d = WA_CASES[i]
    rs = np.random.RandomState(900 + i)
    # old heads smaller norm than the new one, as in imgs/norm.png of the reference README
    ws = []
    for j, nb in enumerate(heads):
        scale = 0.2 if j < len(heads) - 1 else 0.35
        ws.append(torch.from_numpy(rs.normal(0, scale, size=(nb, d))).to(dtype))
    return ws


# (samples per class, classes, feature dim, quota)
HERD_CASES = [(6, 2, 4, 3), (40, 3, 16, 10), (500, 4, 64, 200), (37, 5, 64, 50), (700, 2, 64, 20)]


def herd_inputs(i):
    n, c, d, _ = HERD_CASES[i]
    rs = np.random.RandomState(1300 + i)
    # post-ReLU, post-avgpool like features: non-negative, class-dependent mean
    y = np.repeat(np.arange(c), n)
    rs.shuffle(y)
    centers = rs.uniform(0.2, 1.0, size=(c, d))
    feats = np.maximum(centers[y] + rs.normal(0, 0.3, size=(len(y), d)), 0).astype(np.float32)
    return feats, y.astype(np.int64)
