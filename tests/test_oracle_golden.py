"""Pins oracle/wa_oracle.py (the CPU restatement) against golden vectors produced by the REAL reference
(oracle/make_golden.py imported /root/reference unmodified).  CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

import golden_inputs as gi
import wa_oracle as wo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _summ(t):
    t = t.detach().double().reshape(-1)
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])


@pytest.mark.parametrize("name", list(gi.TRAIN_CASES))
def test_train_steps_match_reference(name, golden_dir):
    case = gi.TRAIN_CASES[name]
    g = np.load(os.path.join(golden_dir, f"train_{name}.npz"))
    sd, tsd = gi.student_state(case), gi.teacher_state(case)
    bufs = {}
    for step in range(case["steps"]):
        x, y = gi.batch(case, step)
        r = wo.train_step(sd, tsd, x, y, case["depth"], case["known"], case["lr"], bufs,
                          lambda_kd=case["lambda_kd"], T=2.0)
        np.testing.assert_allclose(r["logits"].numpy(), g[f"logits{step}"], rtol=2e-4, atol=2e-4)
        np.testing.assert_allclose(r["feats"].numpy(), g[f"feats{step}"], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose([r["loss_ce"].item(), r["loss_kd"].item(), r["loss"].item()],
                                   g[f"losses{step}"], rtol=1e-5)
        if step == 0:
            np.testing.assert_allclose(r["t_logits"].numpy(), g["t_logits"], rtol=2e-4, atol=2e-4)
            for k in gi.FULL_GRAD_KEYS(case["depth"]):
                ref = g["grad/" + k]
                np.testing.assert_allclose(r["grads"][k].numpy(), ref, rtol=2e-3, atol=2e-5 * np.abs(ref).max())
            for k, gr in r["grads"].items():
                ref = g["gradsum/" + k]
                assert abs(_summ(gr)[1] - ref[1]) <= 2e-3 * ref[1] + 1e-7, k
    for k in gi.FULL_PARAM_KEYS(case["depth"]):
        np.testing.assert_allclose(sd[k].numpy(), g["param/" + k], rtol=1e-4, atol=1e-5)
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            ref = g["paramsum/" + k]
            assert abs(_summ(v)[1] - ref[1]) <= 1e-4 * ref[1] + 1e-7, k
    x, _ = gi.batch(case, 0)
    with torch.no_grad():
        lg, ft = wo.cil_forward(sd, x, case["depth"], train=False)
    np.testing.assert_allclose(lg.numpy(), g["eval_logits"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(ft.numpy(), g["eval_extract_vector"], rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("i", range(len(gi.KD_CASES)))
def test_soft_target_matches_reference(i, golden_dir):
    g = np.load(os.path.join(golden_dir, "kd.npz"))
    B, C, T, pad = gi.KD_CASES[i]
    s, t = gi.kd_inputs(i)
    loss = wo.soft_target(s[:, :C], t, T)
    np.testing.assert_allclose(loss.item(), g[f"loss{i}"], rtol=2e-5)  # fp32 summation-order noise at C=8192
    grad = torch.zeros_like(s)
    grad[:, :C] = wo.soft_target_grad(s[:, :C], t, T)
    np.testing.assert_allclose(grad.numpy(), g[f"grad{i}"], rtol=1e-4, atol=1e-8)
    # fp64 closed form agrees with the fp32 reference output to fp32 round-off
    s64, t64 = gi.kd_inputs(i, torch.float64)
    np.testing.assert_allclose(wo.soft_target(s64[:, :C], t64, T).item(), g[f"loss{i}"], rtol=2e-5)


@pytest.mark.parametrize("i", range(len(gi.WA_CASES)))
def test_weight_align_matches_reference(i, golden_dir):
    g = np.load(os.path.join(golden_dir, "weight_align.npz"))
    heads, _ = gi.WA_CASES[i]
    ws = gi.wa_inputs(i)
    gamma, new = wo.weight_align(ws, heads[-1])
    np.testing.assert_allclose(gamma.item(), g[f"gamma{i}"], rtol=1e-5)
    np.testing.assert_allclose(new.numpy(), g[f"new{i}"], rtol=1e-6, atol=1e-8)
    # post-condition of WA: mean norm of new rows == mean norm of old rows
    w_all = torch.cat(ws[:-1] + [new])
    n = w_all.norm(dim=1)
    np.testing.assert_allclose(n[:-heads[-1]].mean().item(), n[-heads[-1]:].mean().item(), rtol=1e-5)


def test_cosine_lr_matches_torch_scheduler():
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=0.1)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=7)
    for e in range(7):
        assert abs(opt.param_groups[0]["lr"] - wo.cosine_lr(0.1, e, 7)) < 1e-9
        opt.step()
        sch.step()


def test_accuracy_topk():
    out = torch.tensor([[0.1, 0.9, 0.0], [0.8, 0.1, 0.3], [0.2, 0.3, 0.5]])
    tgt = torch.tensor([1, 2, 0])
    a1, a2 = wo.accuracy(out, tgt, topk=(1, 2))
    assert abs(a1.item() - 100 / 3) < 1e-4 and abs(a2.item() - 200 / 3) < 1e-4


# ---------------------------------------------------------------------------------------------- herding
def _c_oracle():
    so = os.path.join(ROOT, "oracle", "_build", "libherd_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so)
    lib.herd_barycenter_c.restype = ctypes.c_int64
    lib.herd_barycenter_c.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_void_p]
    return lib


def herd_c(feats, y, quota):
    lib = _c_oracle()
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.int64)
    out = np.zeros(len(y), dtype=np.int64)
    n = lib.herd_barycenter_c(feats.ctypes.data, y.ctypes.data, len(y), feats.shape[1], quota, out.ctypes.data)
    assert n >= 0
    return out[:n]


@pytest.mark.parametrize("i", range(len(gi.HERD_CASES)))
def test_herding_c_oracle_matches_numpy_restatement(i, golden_dir):
    g = np.load(os.path.join(golden_dir, "herding.npz"))
    feats, y = gi.herd_inputs(i)
    quota = gi.HERD_CASES[i][3]
    idx_np = wo.herd_barycenter_np(np.arange(len(y)), y, y, feats, quota)[3]
    assert np.array_equal(idx_np, g[f"idx{i}"])
    assert np.array_equal(herd_c(feats, y, quota), g[f"idx{i}"])


def test_herding_toy_by_hand():
    # two samples on the axes and one on the diagonal: the diagonal one is closest to the mean direction
    feats = np.array([[1, 0], [0, 1], [1, 1], [2, 0.1]], dtype=np.float32)
    y = np.zeros(4, dtype=np.int64)
    idx = herd_c(feats, y, 2)
    assert idx[0] == 2
    assert np.array_equal(idx, wo.herd_barycenter_seq(feats, y, 2))


def test_rehearsal_memory_quota_and_reduce():
    rs = np.random.RandomState(0)
    mem = wo.RehearsalMemoryOracle(memory_size=40, herding_method="barycenter")
    y0 = np.repeat(np.arange(2), 30)
    x0 = np.arange(60)
    mem.add(x0, y0, np.zeros(60, int), rs.rand(60, 8).astype(np.float32))
    assert len(mem) == 40 and mem.memory_per_class == 20
    y1 = np.repeat(np.arange(2, 4), 30)
    mem.add(np.arange(60, 120), y1, np.ones(60, int), rs.rand(60, 8).astype(np.float32))
    x, y, t = mem.get()
    assert mem.memory_per_class == 10 and len(x) == 40
    assert all((y == c).sum() == 10 for c in range(4))


# ------------------------------------------------------------------------- mixed-precision model of the same step
@pytest.mark.parametrize("depth,heads", [(20, [10, 10]), (32, [10])])
def test_manual_backward_model_equals_autograd_oracle_when_rounding_is_off(depth, heads):
    """oracle/wa_oracle_bf16.py with rounding disabled must reproduce the fp32 oracle (which is pinned to the
    reference): that ties the bf16 model used for GPU gradient parity back to the reference."""
    import wa_oracle_bf16 as wb
    sd = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in wo.make_state(depth, heads, 7).items()}
    tsd = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in wo.make_state(depth, heads[:1], 8).items()}
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(6, 3, 32, 32, generator=gen).double()
    y = torch.randint(0, sum(heads), (6,), generator=gen)
    known = 10 if len(heads) > 1 else 0
    r = wo.train_step(dict(sd), tsd if known else None, x, y, depth, known, 0.1, {}, apply_update=False)
    t_logits = r["t_logits"]
    m = wb.forward_backward(sd, x, depth, wb.ce_kd_dlogits(y, t_logits, known), rnd=wb.identity)
    np.testing.assert_allclose(m["logits"].numpy(), r["logits"].numpy(), rtol=1e-9, atol=1e-10)
    for k, g in r["grads"].items():
        np.testing.assert_allclose(m["grads"][k].numpy(), g.numpy(), rtol=1e-7, atol=1e-10 * (1 + g.abs().max().item()),
                                   err_msg=k)
    for k, v in r["new_buffers"].items():
        if v.dtype.is_floating_point:
            np.testing.assert_allclose(m["new_buffers"][k].numpy(), v.numpy(), rtol=1e-9, atol=1e-12)
