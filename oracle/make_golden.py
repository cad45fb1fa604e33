"""Mint the golden vectors under tests/golden/ by running the REAL reference on seeded inputs.

Run in the build container only (needs /root/reference):  python oracle/make_golden.py
The outputs are committed; tests never need /root/reference.  Inputs are regenerated from the seeds stored in
each file by `oracle/golden_inputs.py`, so only outputs (and small summaries of large tensors) are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import golden_inputs as gi  # noqa: E402
import ref_harness  # noqa: E402
import wa_oracle  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def summarize(t: torch.Tensor) -> np.ndarray:
    t = t.detach().double().reshape(-1)
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()], dtype=np.float64)


def golden_train_steps(resnet, utils, template, case: dict) -> dict:
    torch.manual_seed(0)
    depth, B, heads, known = case["depth"], case["batch"], case["heads"], case["known"]
    with ref_harness.cuda_shim():
        model = template.CilModel(ref_harness.Args(backbone=f"resnet{depth}"))
        for nb in heads:
            model.prev_model_adaption(nb)
        teacher = template.CilModel(ref_harness.Args(backbone=f"resnet{depth}"))
        for nb in heads[:-1]:
            teacher.prev_model_adaption(nb)
    sd = gi.student_state(case)
    tsd = gi.teacher_state(case)
    model.load_state_dict(sd, strict=True)
    teacher.load_state_dict(tsd, strict=True)
    teacher = teacher.copy().freeze()  # template.py:290
    criterion = torch.nn.CrossEntropyLoss(label_smoothing=0.0)
    kd = utils.SoftTarget(T=2)
    opt = torch.optim.SGD(model.parameters(), lr=case["lr"], momentum=0.9, weight_decay=5e-4)
    out = {}
    model.train()
    for step in range(case["steps"]):
        x, y = gi.batch(case, step)
        logits, feats = model(x)  # template.py:258
        loss_ce = criterion(logits, y)
        t_logits, _ = teacher(x)
        loss_kd = case["lambda_kd"] * kd(logits[:, :known], t_logits)
        loss = loss_ce + loss_kd
        opt.zero_grad()
        loss.backward()
        named = dict(model.named_parameters())
        if step == 0:
            out["t_logits"] = t_logits.detach().numpy()
            for k in gi.FULL_GRAD_KEYS(depth):
                out["grad/" + k] = named[k].grad.detach().numpy().copy()
            for k, p in named.items():
                out["gradsum/" + k] = summarize(p.grad)
        opt.step()
        out[f"logits{step}"] = logits.detach().numpy()
        out[f"feats{step}"] = feats.detach().numpy()
        out[f"losses{step}"] = np.array([loss_ce.item(), loss_kd.item(), loss.item()], dtype=np.float64)
    final = model.state_dict()
    for k, v in final.items():
        if v.dtype.is_floating_point:
            out["paramsum/" + k] = summarize(v)
    for k in gi.FULL_PARAM_KEYS(depth):
        out["param/" + k] = final[k].detach().numpy().copy()
    # eval-mode forward after the updates (exercises running statistics)
    model.eval()
    with torch.no_grad():
        x, _ = gi.batch(case, 0)
        lg, ft = model(x)
        out["eval_logits"] = lg.numpy()
        out["eval_feats"] = ft.numpy()
        out["eval_extract_vector"] = model.extract_vector(x).numpy()
    return out


def golden_kd(utils) -> dict:
    out = {}
    for i, (B, C, T, stride_pad) in enumerate(gi.KD_CASES):
        s, t = gi.kd_inputs(i)
        s = s.clone().requires_grad_(True)
        s_view = s[:, :C] if stride_pad else s
        loss = utils.SoftTarget(T=T)(s_view, t)
        (g,) = torch.autograd.grad(loss, s)
        out[f"loss{i}"] = np.array(loss.item(), dtype=np.float64)
        out[f"grad{i}"] = g.numpy()
    return out


def golden_weight_align(template) -> dict:
    out = {}
    for i, (heads, d) in enumerate(gi.WA_CASES):
        with ref_harness.cuda_shim():
            model = template.CilModel(ref_harness.Args())
            model.fc = template.CilClassifier(d, heads[0])
            for nb in heads[1:]:
                model.fc.adaption(nb)
        ws = gi.wa_inputs(i)
        for h, w in zip(model.fc.heads, ws):
            h.weight.data = w.clone()
        import io, contextlib
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            model.weight_align(heads[-1])
        out[f"new{i}"] = model.fc[-1].weight.data.numpy().copy()
        out[f"gamma{i}"] = np.array(float(buf.getvalue().strip().split("=")[-1]), dtype=np.float64)
        for j in range(len(heads) - 1):
            assert torch.equal(model.fc[j].weight.data, ws[j])  # old heads untouched
    return out


def golden_herding() -> dict:
    """continuum is not available: these come from the NumPy restatement (parity unpinned)."""
    out = {}
    for i, (n_per_class, n_classes, d, quota) in enumerate(gi.HERD_CASES):
        feats, y = gi.herd_inputs(i)
        x = np.arange(len(y))
        _, _, _, idx = wa_oracle.herd_barycenter_np(x, y, y.copy(), feats, quota)
        out[f"idx{i}"] = idx.astype(np.int64)
        if n_per_class * n_classes <= 400:
            seq = wa_oracle.herd_barycenter_seq(feats, y, quota)
            assert np.array_equal(seq, idx), "sequential-order restatement disagrees with the NumPy one"
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    resnet, utils, template = ref_harness.load_reference()
    torch.set_num_threads(8)
    for name, case in gi.TRAIN_CASES.items():
        out = golden_train_steps(resnet, utils, template, case)
        np.savez_compressed(os.path.join(OUT, f"train_{name}.npz"), **out)
        print(name, {k: float(v[2]) for k, v in out.items() if k.startswith("losses")})
    np.savez_compressed(os.path.join(OUT, "kd.npz"), **golden_kd(utils))
    np.savez_compressed(os.path.join(OUT, "weight_align.npz"), **golden_weight_align(template))
    np.savez_compressed(os.path.join(OUT, "herding.npz"), **golden_herding())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
