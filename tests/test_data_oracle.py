"""CPU checks of the oracle restatements added for the SURVEY 8f rows (eval meter, device data path, cluster / random
herding) and of the host-side scenario logic.  The augmentation restatement is pinned against torchvision's own
functional transforms (the library the reference's transform chain is made of, utils.py:210-251)."""
import ctypes
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
import wa_oracle as wo
from test_oracle_golden import _c_oracle


def herd_cluster_c(feats, y, quota):
    lib = _c_oracle()
    lib.herd_cluster_c.restype = ctypes.c_int64
    lib.herd_cluster_c.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_void_p]
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.int64)
    out = np.zeros(len(y), dtype=np.int64)
    n = lib.herd_cluster_c(feats.ctypes.data, y.ctypes.data, len(y), feats.shape[1], quota, out.ctypes.data)
    assert n >= 0
    return out[:n]


def aug_cases(seed, n_store=12, B=9, pad=4):
    rs = np.random.RandomState(seed)
    store = rs.randint(0, 256, size=(n_store, 32, 32, 3)).astype(np.uint8)
    idx = rs.randint(0, n_store, size=B).astype(np.int64)
    dy, dx, fl = rs.randint(0, 2 * pad + 1, B), rs.randint(0, 2 * pad + 1, B), rs.randint(0, 2, B)
    dy[0], dx[0], fl[0] = 0, 0, 1          # top-left corner: 4 rows / columns of padding visible
    if B > 1:
        dy[1], dx[1], fl[1] = 2 * pad, 2 * pad, 0  # bottom-right corner
    aug = (dy | (dx << 8) | (fl << 16)).astype(np.int32)
    return store, idx, aug


@pytest.mark.parametrize("stats", ["imagenet", "cifar"])
def test_augment_restatement_matches_torchvision(stats):
    tvf = pytest.importorskip("torchvision.transforms.functional")
    from PIL import Image
    mean, std = wo.transform_stats("CIFAR" if stats == "cifar" else "cifar")
    store, idx, aug = aug_cases(5)
    got = wo.augment_u8_np(store, idx, aug, mean, std)
    for b in range(len(idx)):
        dy, dx, flip = int(aug[b]) & 0xFF, (int(aug[b]) >> 8) & 0xFF, (int(aug[b]) >> 16) & 1
        img = Image.fromarray(store[idx[b]])
        img = tvf.crop(tvf.pad(img, 4, fill=0), dy, dx, 32, 32)  # RandomCrop(32, padding=4) with these draws
        if flip:
            img = tvf.hflip(img)
        ref = tvf.normalize(tvf.to_tensor(img), mean, std).numpy()
        assert np.array_equal(got[b], ref), f"sample {b}"
    # eval transform = ToTensor + Normalize
    ev = wo.augment_u8_np(store, None, None, mean, std)
    ref = tvf.normalize(tvf.to_tensor(Image.fromarray(store[3])), mean, std).numpy()
    assert np.array_equal(ev[3], ref)


def test_transform_stats_quirk():
    # utils.py:228-230: the CIFAR statistics are only used for data_set == 'CIFAR' (the default is 'cifar')
    assert wo.transform_stats("cifar") == (wo.IMAGENET_DEFAULT_MEAN, wo.IMAGENET_DEFAULT_STD)
    assert wo.transform_stats("CIFAR") == (wo.CIFAR_MEAN, wo.CIFAR_STD)


@pytest.mark.parametrize("i", range(len(gi.HERD_CASES)))
def test_cluster_herding_c_oracle_matches_numpy(i):
    feats, y = gi.herd_inputs(i)
    quota = gi.HERD_CASES[i][3]
    idx_np = wo.herd_cluster_np(np.arange(len(y)), y, y, feats, quota)[3]
    idx_c = herd_cluster_c(feats, y, quota)
    assert len(idx_np) == len(idx_c)
    # float32 sequential sums (C) vs NumPy's pairwise float32: the selected SET per class must agree except where two
    # distances tie within rounding; on these seeded cases they agree exactly
    assert np.array_equal(idx_np, idx_c)


def test_cluster_herding_toy_by_hand():
    feats = np.array([[0, 0], [4, 0], [1.9, 0], [2.2, 0]], dtype=np.float32)  # mean x = 2.025
    y = np.zeros(4, dtype=np.int64)
    assert herd_cluster_c(feats, y, 2).tolist() == [2, 3]
    assert wo.herd_cluster_np(y, y, y, feats, 2)[3].tolist() == [2, 3]


def test_random_herding_quota_and_classes():
    y = np.repeat(np.arange(3), [5, 50, 20])
    x = np.arange(len(y))
    mx, my, mt, idx = wo.herd_random_np(x, y, y, None, 10, rng=np.random.RandomState(0))
    assert [int((my == c).sum()) for c in range(3)] == [5, 10, 10]
    assert len(set(idx.tolist())) == len(idx) and np.array_equal(y[idx], my)


@pytest.mark.parametrize("B,C", [(7, 3), (64, 10), (33, 100)])
def test_eval_meter_restatement(B, C):
    g = torch.Generator().manual_seed(B + C)
    z = torch.randn(B, C, generator=g) * 2
    y = torch.randint(0, C, (B,), generator=g)
    m = wo.eval_meter(z, y, 5)
    k = min(5, C)
    rank = (z > z.gather(1, y[:, None])).sum(1)
    assert m[0] == B and m[1] == 1.0
    assert abs(m[2] - torch.nn.functional.cross_entropy(z, y).item()) < 1e-6
    assert round(m[3]) == int((rank < 1).sum()) and round(m[4]) == int((rank < k).sum())


def test_scenario_and_taskset_host_logic():
    import wacil_b200 as w
    x, y = w.data.synthetic_cifar(nb_classes=10, per_class=6, seed=3)
    assert x.shape == (60, 32, 32, 3) and x.dtype == np.uint8
    order = [3, 1, 4, 0, 5, 9, 2, 6, 8, 7]
    sc = w.data.ClassIncrementalScenario(x, y, 4, 3, class_order=order, device="cpu")
    assert len(sc) == 3 and sc.increments == [4, 3, 3] and sc.nb_classes == 10
    t0 = sc[0]
    # task 0 holds the first four classes of the order, re-labelled 0..3
    assert sorted(np.unique(t0.y.numpy()).tolist()) == [0, 1, 2, 3] and len(t0) == 24
    raw_x, raw_y, raw_t = t0.get_raw_samples()
    assert np.array_equal(np.unique(y[np.isin(y, order[:4])]), sorted(order[:4])) and (raw_t == 0).all()
    # original class 3 is label 0
    first = np.where(y == 3)[0][0]
    assert any(np.array_equal(raw_x[i], x[first]) and raw_y[i] == 0 for i in range(len(raw_y)))
    val = sc[:2]
    assert len(val) == 42 and sorted(np.unique(val.y.numpy()).tolist()) == list(range(7))
    t1 = sc[1]
    t1.add_samples(raw_x[:5], raw_y[:5], raw_t[:5])   # template.py:231
    assert len(t1) == 18 + 5
    with pytest.raises(ValueError):
        w.data.ClassIncrementalScenario(x, y, 4, 3, class_order=[0, 1, 2], device="cpu")
