"""CPU-side checks of the C ABI: the library loads, exports every symbol include/wacil.h declares, validates
arguments without touching a GPU, and reports the arena layout the host code relies on."""
import ctypes as C
import os
import re

import pytest
import torch

import wa_oracle as wo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def w():
    import wacil_b200
    return wacil_b200


def test_every_declared_symbol_is_exported_and_bound(w):
    hdr = open(os.path.join(ROOT, "include", "wacil.h")).read()
    declared = set(re.findall(r"\b(wacil_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = C.CDLL(w._lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in wacil.h but not exported"
    assert declared == set(w._lib.SIGNATURES), declared ^ set(w._lib.SIGNATURES)
    assert w._lib.lib().wacil_version() == 100


def test_argument_errors_do_not_need_a_gpu(w):
    lib = w._lib.lib()
    lay = w._lib.NetLayout()
    assert lib.wacil_net_layout(33, 100, C.byref(lay)) == -1
    assert b"6n+2" in lib.wacil_last_error()
    assert lib.wacil_kd_fwd_bwd(None, 0, None, 0, 4, 4, 2.0, 1.0, None, None, 0, 0, 0, None, None) == -4
    assert lib.wacil_weight_align(None, 10, 64, 5, None, 0, None, None) == -4
    assert lib.wacil_kd_workspace_bytes(128) >= 128 * 12
    with pytest.raises(w.WacilError):
        w.ops.kd_fwd_bwd(torch.zeros(2, 4), torch.zeros(2, 4), 2.0)  # CPU tensors: no fallback


@pytest.mark.parametrize("depth", [20, 32, 56])
def test_layout_matches_reference_parameter_count(w, depth):
    arena = w.engine.NetArena(depth, 100)
    sd = wo.make_state(depth, [], 0)
    n_ref = sum(v.numel() for k, v in sd.items() if k.endswith("weight") or k.endswith("bias"))
    n_live = sum(ll.cin * ll.cout * 9 + 2 * ll.cout for ll in arena.layers)
    assert n_live == n_ref  # 463,504 - 650 head params for ResNet-32
    assert arena.layout.n_layers == depth - 1 and arena.layout.fc_weight_off % 4 == 0
    assert arena.conv_weight(0).shape == (16, 3, 3, 3) and arena.conv_weight(depth - 2).shape == (64, 64, 3, 3)
