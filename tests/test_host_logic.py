"""CPU tests of the Python host layer (no CUDA kernels run here)."""
import math
import os

import numpy as np
import pytest
import torch

import wa_oracle as wo

REF = "/root/reference"


@pytest.fixture(scope="module")
def w():
    import wacil_b200
    return wacil_b200


def test_state_dict_roundtrip_with_reference_keys(w):
    sd = wo.make_state(20, [10, 10], 5)
    m = w.CilModel("resnet20")
    m.load_state_dict(sd)
    out = m.state_dict()
    assert set(out) == set(sd)
    for k in sd:
        assert torch.equal(out[k].float(), sd[k].float()) or k.endswith("num_batches_tracked"), k
    # tap-major storage really is the permuted view
    a = m._arena
    assert torch.equal(a.conv_weight(3), sd[a.names[3][0] + ".weight"])


def test_copy_freeze_and_adaption(w):
    m = w.CilModel("resnet20")
    m.prev_model_adaption(10)
    t = m.copy().freeze()
    assert not t.training and not t.arena.requires_grad and m.arena.requires_grad and m.training
    with torch.no_grad():
        m.arena.add_(1.0)
    assert not torch.equal(t.arena, m.arena)
    assert len(list(m.parameters())) == 1 and len(m.fc) == 1 and len(t.fc) == 1
    with pytest.raises(NotImplementedError):
        w.CilModel("resnet18")
    import copy
    t2 = copy.deepcopy(m)
    assert torch.equal(t2.arena, m.arena) and t2.arena.data_ptr() != m.arena.data_ptr()


@pytest.mark.skipif(not os.path.exists(REF), reason="needs the reference checkout (build container only)")
def test_init_consumes_rng_like_the_reference(w):
    import ref_harness
    resnet, utils, template = ref_harness.load_reference()
    torch.manual_seed(0)
    with ref_harness.cuda_shim():
        ref = template.CilModel(ref_harness.Args("resnet32"))
        ref.prev_model_adaption(10)
    torch.manual_seed(0)
    ours = w.CilModel("resnet32")
    ours.prev_model_adaption(10)
    rsd, osd = ref.state_dict(), ours.state_dict()
    assert set(rsd) == set(osd)
    for k in rsd:
        assert torch.equal(rsd[k], osd[k]), k


@pytest.mark.parametrize("n,world", [(7000, 1), (7000, 4), (5001, 8), (3, 8)])
def test_sampler_indices_equal_distributed_sampler(w, n, world):
    from torch.utils.data import DistributedSampler
    ds = list(range(n))
    for epoch in (0, 3):
        seen = []
        for rank in range(world):
            s = DistributedSampler(ds, num_replicas=world, rank=rank, shuffle=True)
            s.set_epoch(epoch)
            ours = w.train.sampler_indices(n, epoch, rank, world)
            assert list(s) == ours.tolist()
            seen += ours.tolist()
        assert set(seen) == set(range(n))


def test_cosine_lr_and_flags(w):
    assert abs(w.train.cosine_lr(0.1, 0, 140) - 0.1) < 1e-12
    assert abs(w.train.cosine_lr(0.1, 70, 140) - 0.05) < 1e-12
    args = w.get_args_parser().parse_args([])
    assert (args.batch_size, args.memory_size, args.lr, args.momentum, args.weight_decay, args.num_epochs,
            args.lambda_kd, args.backbone, args.num_bases, args.increment) == (128, 2000, 0.1, 0.9, 5e-4, 140, 0.5,
                                                                              "resnet32", 50, 10)


def test_rehearsal_memory_bookkeeping_with_callable_herding(w):
    def first_k(x, y, t, z, nb_per_class):
        idx = np.concatenate([np.where(y == c)[0][:nb_per_class] for c in np.unique(y)])
        return x[idx], y[idx], t[idx]

    ora = wo.RehearsalMemoryOracle(40, first_k)
    for dedup in (False, True):
        mem = w.RehearsalMemory(40, first_k, dedup_old=dedup)
        y0 = np.repeat(np.arange(2), 30)
        mem.add(np.arange(60), y0, np.zeros(60, int), None)
        assert len(mem) == 40
        # second task set contains the old exemplars again, as template.py:231,300-302 does
        mx, my, mt = mem.get()
        y1 = np.concatenate([np.repeat(np.arange(2, 4), 30), my])
        x1 = np.concatenate([np.arange(60, 120), mx])
        mem.add(x1, y1, np.ones(len(y1), int), None)
        if dedup:
            assert len(mem) == 40 and all((mem.get()[1] == c).sum() == 10 for c in range(4))
        else:
            assert len(mem) == 60  # upstream: old classes re-herded and appended again
    ora.add(np.arange(60), np.repeat(np.arange(2), 30), np.zeros(60, int), None)
    assert len(ora) == 40
    with pytest.raises(NotImplementedError):
        w.RehearsalMemory(10, "kmeans")
    for name in ("barycenter", "cluster", "random"):  # the three methods continuum ships (template.py:34)
        assert callable(w.RehearsalMemory(10, name).herding_method)
    with pytest.raises(ValueError):
        w.RehearsalMemory(10, "barycenter", fixed_memory=True)


def test_checkpoint_round_trip_on_the_host(tmp_path):
    """checkpoint.py is host logic (torch.save of reference-keyed state dicts + numpy memory): the round trip is exact
    without a GPU."""
    import argparse
    import wacil_b200 as w
    torch.manual_seed(1)
    model = w.CilModel("resnet20", max_classes=12)
    model.prev_model_adaption(5)
    model.prev_model_adaption(4)
    mem = w.RehearsalMemory(18, "barycenter")
    mem._x, mem._y, mem._t = np.arange(18).reshape(18, 1), np.repeat(np.arange(9), 2), np.zeros(18, int)
    mem.seen_classes = set(range(9))
    args = argparse.Namespace(known_classes=9)
    path = w.checkpoint.save_checkpoint(str(tmp_path / "t.pt"), model, mem, args, 1, [88.0, 70.25])
    assert not os.path.exists(path + ".tmp")
    mem2, args2 = w.RehearsalMemory(18, "barycenter"), argparse.Namespace(known_classes=0)
    m2, teacher, task_id, acc1s = w.checkpoint.load_checkpoint(path, mem2, args2, device=None, restore_rng=False)
    assert (task_id, acc1s, args2.known_classes, mem2.seen_classes) == (1, [88.0, 70.25], 9, set(range(9)))
    assert np.array_equal(mem2.get()[0], mem.get()[0]) and len(m2.fc) == 2 and m2.num_classes == 9
    sd1, sd2, sdt = model.state_dict(), m2.state_dict(), teacher.state_dict()
    assert list(sd1) == list(sd2) and all(torch.equal(sd1[k], sd2[k]) and torch.equal(sd1[k], sdt[k]) for k in sd1)
    assert not teacher.arena.requires_grad and not teacher.training and m2.arena.requires_grad


def test_device_taskset_draws_and_runner_flags_on_the_host():
    import argparse
    import wacil_b200 as w
    x, y = w.data.synthetic_cifar(nb_classes=3, per_class=4, seed=2)
    ts = w.DeviceTaskSet(x, y, device="cpu", pad=4)
    a, b = ts.draw_aug(50, seed=1, epoch=2, rank=0), ts.draw_aug(50, seed=1, epoch=2, rank=1)
    assert torch.equal(a, ts.draw_aug(50, seed=1, epoch=2, rank=0)) and not torch.equal(a, b)
    an = a.numpy()
    assert ((an & 0xFF) <= 8).all() and (((an >> 8) & 0xFF) <= 8).all() and ((an >> 16) <= 1).all() and (an >= 0).all()
    # the Normalize() statistics follow the reference's data_set == 'CIFAR' quirk (utils.py:228-230)
    ns = argparse.Namespace(input_size=32, data_set="cifar")
    assert w.data.transform_stats(ns) == (w.data.IMAGENET_DEFAULT_MEAN, w.data.IMAGENET_DEFAULT_STD)
    ns.data_set = "CIFAR"
    assert w.data.transform_stats(ns) == (w.data.CIFAR_MEAN, w.data.CIFAR_STD)
    p = argparse.ArgumentParser(parents=[w.get_args_parser()])
    w.train.add_runner_args(p)
    args = p.parse_args(["--data_set", "synthetic", "--dynamic_lambda_kd", "--herding_method", "cluster"])
    assert args.lambda_kd == 0.5 and args.memory_size == 2000 and args.num_bases == 50 and args.batch_size == 128
    from wacil_b200.train import kd_weights
    assert kd_weights(args, 50, 10, True) == (1 - 50 / 60, 50 / 60) and kd_weights(args, 0, 50, False) == (1.0, 0.5)
    with pytest.raises(FileNotFoundError):
        args.data_set, args.data_path = "cifar", "/nonexistent"
        w.build_dataset(True, args, device="cpu")
