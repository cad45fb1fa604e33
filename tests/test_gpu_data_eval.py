"""GPU parity of the SURVEY 8f rows through the C ABI: device data path (bit-exact with the CPU restatement, which
tests/test_data_oracle.py pins against torchvision), eval meter, cluster herding (bit-exact with oracle/herding.c),
the weighted CE+KD objective, the end-to-end runner and checkpoint/resume."""
import argparse
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
import wa_oracle as wo
from test_data_oracle import aug_cases, herd_cluster_c

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def w():
    import wacil_b200
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return wacil_b200


# ------------------------------------------------------------------------------------------------ data path
@pytest.mark.parametrize("seed,n_store,B", [(1, 12, 9), (2, 300, 128), (3, 5, 1)])
def test_augment_bit_exact(w, seed, n_store, B):
    store, idx, aug = aug_cases(seed, n_store, B)
    for mean, std in (wo.transform_stats("cifar"), wo.transform_stats("CIFAR")):
        ref = wo.augment_u8_np(store, idx, aug, mean, std)
        got = w.ops.augment_u8(torch.from_numpy(store).cuda(), torch.from_numpy(idx).cuda(),
                               torch.from_numpy(aug).cuda(), mean, std)
        assert np.array_equal(got.cpu().numpy(), ref)  # bit-exact fp32
    # eval transform (no idx, no aug): ToTensor + Normalize of the whole store
    ref = wo.augment_u8_np(store, None, None, mean, std)
    got = w.ops.augment_u8(torch.from_numpy(store).cuda(), None, None, mean, std)
    assert np.array_equal(got.cpu().numpy(), ref)


def test_augment_rejects_bad_arguments(w):
    store = torch.zeros(4, 32, 32, 3, dtype=torch.uint8, device="cuda")
    with pytest.raises(w.WacilError):
        w.ops.augment_u8(store.float(), None, None)
    with pytest.raises(w.WacilError):
        w.ops.augment_u8(store, torch.zeros(2, dtype=torch.int32, device="cuda"), None)
    with pytest.raises(w.WacilError):
        w.ops.augment_u8(store, None, None, std=(0.2, 0.0, 0.2))
    with pytest.raises(w.WacilError):
        w.ops.augment_u8(store.cpu(), None, None)


def test_device_taskset_draws_and_batches(w):
    x, y = w.data.synthetic_cifar(nb_classes=4, per_class=10, seed=1)
    ts = w.DeviceTaskSet(x, y)
    a1, a2 = ts.draw_aug(40, seed=0, epoch=3, rank=0), ts.draw_aug(40, seed=0, epoch=3, rank=0)
    assert torch.equal(a1, a2) and not torch.equal(a1, ts.draw_aug(40, seed=0, epoch=4, rank=0))
    a = a1.cpu().numpy()
    assert ((a & 0xFF) <= 8).all() and (((a >> 8) & 0xFF) <= 8).all() and ((a >> 16) <= 1).all()
    idx = torch.tensor([5, 0, 39, 17], device="cuda")
    got = ts.batch(idx, a1[:4])
    ref = wo.augment_u8_np(x, idx.cpu().numpy(), a[:4], ts.mean, ts.std)
    assert np.array_equal(got.cpu().numpy(), ref)


# ------------------------------------------------------------------------------------------------ eval meter
@pytest.mark.parametrize("B,C", [(128, 10), (100, 100), (7, 3), (64, 1000), (16, 2048)])
def test_eval_rows_matches_oracle(w, B, C):
    g = torch.Generator().manual_seed(B * 7 + C)
    meter = torch.zeros(8, dtype=torch.float64, device="cuda")
    want = np.zeros(5)
    for it in range(3):
        z = torch.randn(B, C, generator=g) * 3
        y = torch.randint(0, C, (B,), generator=g)
        if it == 2:
            z[0, :] = 1.5  # a fully tied row: torch.topk keeps the lowest columns
        want += np.array(wo.eval_meter(z.double(), y, 5))
        w.ops.eval_rows_(z.cuda(), y.cuda(), meter, topk=5)
    got = meter.cpu().numpy()[:5]
    assert got[0] == want[0] and got[1] == want[1]
    assert abs(got[2] - want[2]) <= 5e-6 * abs(want[2])        # fp32 CE vs the fp64 closed form
    # hits are integers; a tied row may legitimately differ in which tied column torch.topk reports: none here except row 0
    assert abs(got[3] - want[3]) <= 1 and abs(got[4] - want[4]) <= 1
    z = torch.randn(B, C, generator=g)
    y = torch.randint(0, C, (B,), generator=g)
    m2 = torch.zeros(8, dtype=torch.float64, device="cuda")
    w.ops.eval_rows_(z.cuda(), y.cuda(), m2)
    ref = wo.eval_meter(z.double(), y, 5)
    assert round(m2[3].item()) == round(ref[3]) and round(m2[4].item()) == round(ref[4])


def test_eval_function_matches_reference_semantics(w):
    """eval(model, loader) (template.py:169-188): acc1 over all samples, loss averaged per batch."""
    torch.manual_seed(0)
    model = w.CilModel("resnet20").cuda()
    model.prev_model_adaption(10)
    model.eval()
    g = torch.Generator().manual_seed(1)
    batches = [(torch.randn(n, 3, 32, 32, generator=g), torch.randint(0, 10, (n,), generator=g), None) for n in (16, 16, 5)]
    acc1 = w.eval(model, batches)
    hits, n = 0, 0
    for x, y, _ in batches:
        logits = model(x.cuda())[0]
        hits += int((logits.argmax(1).cpu() == y).sum())
        n += len(y)
    assert abs(acc1 - 100.0 * hits / n) < 1e-9


# ------------------------------------------------------------------------------------------------ cluster herding
@pytest.mark.parametrize("i", range(len(gi.HERD_CASES)))
def test_cluster_herding_bit_exact_vs_c_oracle(w, i):
    feats, y = gi.herd_inputs(i)
    quota = gi.HERD_CASES[i][3]
    idx = w.rehearsal.herd_indices(feats, y, quota, method="cluster").cpu().numpy()
    assert np.array_equal(idx, herd_cluster_c(feats, y, quota))


def test_herding_methods_through_memory(w):
    feats, y = gi.herd_inputs(2)
    x = np.arange(len(y))
    for name in ("cluster", "random"):
        np.random.seed(0)
        mem = w.RehearsalMemory(40, name)
        mem.add(x, y, np.zeros(len(y), int), feats)
        mx, my, _ = mem.get()
        assert len(mem) == 40 and all((my == c).sum() == 10 for c in np.unique(y)) and np.array_equal(y[mx], my)


# ------------------------------------------------------------------------------------------------ weighted CE + KD
def test_ce_kd_weighted_objective(w):
    g = torch.Generator().manual_seed(3)
    B, Cs, Ck, lam = 32, 30, 20, 20 / 30.0
    z = torch.randn(B, Cs, generator=g) * 2
    y = torch.randint(0, Cs, (B,), generator=g)
    tz = torch.randn(B, Ck, generator=g) * 2
    z64 = z.double().requires_grad_(True)
    ce = (1 - lam) * wo.cross_entropy(z64, y, 0.0)
    kd = lam * wo.soft_target(z64[:, :Ck], tz.double(), 2.0)
    (ce + kd).backward()
    losses, dl = w.ops.ce_kd_fwd_bwd(z.cuda(), y.cuda(), tz.cuda(), Ck, 2.0, lam, 0.0, ce_weight=1 - lam)
    got = losses.cpu().double()
    assert abs(got[0] - ce.item()) <= 5e-6 * abs(ce.item()) and abs(got[1] - kd.item()) <= 5e-6 * abs(kd.item())
    assert (dl.cpu().double() - z64.grad).abs().max().item() <= 1e-4 * z64.grad.abs().max().item()
    from wacil_b200.train import kd_weights
    a = argparse.Namespace(dynamic_lambda_kd=True, lambda_kd=0.5)
    assert kd_weights(a, 20, 10, True) == (1 - lam, lam) and kd_weights(a, 0, 10, False) == (1.0, 0.5)
    a.dynamic_lambda_kd = False
    assert kd_weights(a, 20, 10, True) == (1.0, 0.5)


# ------------------------------------------------------------------------------------------------ runner + resume
def _runner_args(w, tmp, extra=()):
    p = argparse.ArgumentParser(parents=[w.get_args_parser()])
    w.train.add_runner_args(p)
    return p.parse_args(["--data_set", "synthetic", "--nb_classes", "6", "--num_bases", "2", "--increment", "2",
                         "--backbone", "resnet20", "--batch_size", "32", "--num_epochs", "10", "--lr", "0.05",
                         "--memory_size", "60", "--synthetic_per_class", "128", "--synthetic_val_per_class", "32",
                         "--eval_every_epoch", "100", "--checkpoint_dir", str(tmp), *extra])


def test_runner_end_to_end_and_resume(w, tmp_path):
    """template.main on a 3-task synthetic protocol (>= 80 optimisation steps per task, the last ones at a small cosine
    learning rate, so that the BatchNorm running statistics the evaluation uses have caught up with the weights):
    learns (accuracy far above chance), writes a checkpoint per task, and a run resumed from the task-0 checkpoint
    completes the protocol from there with bit-identical results: every reduction of the step has a fixed order, the
    augmentation draws are seeded per (seed, epoch, rank), and the checkpoint restores the RNG streams."""
    logs = []
    acc_full = w.template.main(_runner_args(w, tmp_path / "a"), log=logs.append)
    assert len(acc_full) == 3 and acc_full[0] > 80.0 and acc_full[-1] > 40.0, acc_full  # chance: 50 % / 16.7 %
    ck0 = tmp_path / "a" / "task_00.pt"
    assert ck0.exists() and (tmp_path / "a" / "task_02.pt").exists()
    acc_res = w.template.main(_runner_args(w, tmp_path / "b", ("--resume", str(ck0))), log=lambda *a: None)
    assert len(acc_res) == 3 and acc_res[0] == acc_full[0]
    assert acc_res[1] > 50.0 and acc_res[2] > 40.0, acc_res  # chance: 25 % / 16.7 %
    s_full = torch.load(tmp_path / "a" / "task_02.pt", weights_only=True)  # tensors + plain python only
    s_res = torch.load(tmp_path / "b" / "task_02.pt", weights_only=True)
    assert s_full["known_classes"] == 6 and sorted(s_full["seen_classes"]) == list(range(6))
    assert len(s_full["memory"]["y"]) <= 60 + 40  # dedup_old=False keeps the upstream duplicates
    # the resumed run IS the uninterrupted run: same accuracies, same parameters, same exemplars
    assert acc_res == acc_full, (acc_res, acc_full)
    assert all(torch.equal(s_full["model"][k], s_res["model"][k]) for k in s_full["model"])
    assert torch.equal(s_full["memory"]["x"], s_res["memory"]["x"]) and torch.equal(s_full["memory"]["y"], s_res["memory"]["y"])
    with pytest.raises(ValueError):  # resuming into a different protocol is refused up front
        w.template.main(_runner_args(w, tmp_path / "c", ("--resume", str(ck0), "--backbone", "resnet32")), log=lambda *a: None)


def test_device_resident_memory_matches_numpy_memory(w):
    """RehearsalMemory fed with device tensors (template.main's path: exemplars, labels and features never leave HBM)
    keeps exactly what the upstream-style numpy memory keeps — barycenter and cluster, growing, dedup on/off, and
    --fixed_memory (quota = memory_size // nb_total_classes, no reduction of old classes)."""
    rs = np.random.RandomState(3)
    for method in ("barycenter", "cluster"):
        for dedup, fixed in ((False, False), (True, False), (False, True)):
            mn = w.RehearsalMemory(24, method, fixed_memory=fixed, nb_total_classes=6, dedup_old=dedup)
            md = w.RehearsalMemory(24, method, fixed_memory=fixed, nb_total_classes=6, dedup_old=dedup)
            for task, classes in enumerate(([0, 1], [2, 3], [4, 5])):
                y = np.repeat(np.asarray(classes), 20)
                x = rs.randint(0, 256, (len(y), 4, 4, 3)).astype(np.uint8)
                t = np.full(len(y), task)
                if task > 0:  # template.py:231: the task set contains the old exemplars again
                    ox, oy, ot = mn.get()
                    x, y, t = np.concatenate([x, ox]), np.concatenate([y, oy]), np.concatenate([t, ot])
                feats = np.maximum(rs.normal(0.5, 0.3, (len(y), 64)), 0).astype(np.float32)
                mn.add(x, y, t, feats)
                md.add(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(t).cuda(),
                       torch.from_numpy(feats).cuda())
                dx, dy, dt = md.get()
                assert dx.is_cuda and dy.is_cuda
                nx, ny, nt = mn.get()
                assert np.array_equal(dx.cpu().numpy(), nx) and np.array_equal(dy.cpu().numpy(), ny) \
                    and np.array_equal(dt.cpu().numpy(), nt), (method, dedup, fixed, task)
            if fixed:
                assert mn.memory_per_class == 4 and all((mn.get()[1] == c).sum() >= 4 for c in range(6))


def test_checkpoint_round_trip_is_exact(w, tmp_path):
    torch.manual_seed(0)
    model = w.CilModel("resnet20", max_classes=10).cuda()
    model.prev_model_adaption(4)
    model.prev_model_adaption(3)
    model.train()
    model(torch.randn(8, 3, 32, 32, device="cuda"))  # move the running statistics
    mem = w.RehearsalMemory(20, "barycenter")
    mem._x, mem._y, mem._t = np.arange(12), np.repeat(np.arange(4), 3), np.zeros(12, int)
    mem.seen_classes = {0, 1, 2, 3}
    args = argparse.Namespace(known_classes=4)
    path = w.checkpoint.save_checkpoint(str(tmp_path / "c.pt"), model, mem, args, 1, [91.0, 80.5])
    mem2, args2 = w.RehearsalMemory(20, "barycenter"), argparse.Namespace(known_classes=0)
    m2, teacher, task_id, acc1s = w.checkpoint.load_checkpoint(path, mem2, args2, device="cuda")
    assert task_id == 1 and acc1s == [91.0, 80.5] and args2.known_classes == 4 and mem2.seen_classes == {0, 1, 2, 3}
    assert np.array_equal(mem2.get_numpy()[1], mem.get()[1]) and mem2.get()[1].is_cuda  # resumed straight into HBM
    sd1, sd2 = model.state_dict(), m2.state_dict()
    assert sd1.keys() == sd2.keys() and all(torch.equal(sd1[k].cpu(), sd2[k].cpu()) for k in sd1)
    x = torch.randn(5, 3, 32, 32, device="cuda")
    assert torch.equal(model.eval()(x)[0], m2.eval()(x)[0]) and torch.equal(teacher(x)[0], m2(x)[0])
    assert not teacher.arena.requires_grad
