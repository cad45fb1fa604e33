"""ctypes binding of libwacil.so (declared in include/wacil.h).  The CUDA library is the only compute path:
if it is missing or fails to load, importing/using the package raises — there is no CPU or PyTorch fallback."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libwacil.so")

F32, BF16 = 0, 1
NET_TRAIN, NET_INFER = 1, 2


class WacilError(RuntimeError):
    pass


class NetLayout(C.Structure):
    _fields_ = [("depth", C.c_int), ("n_layers", C.c_int), ("feat_dim", C.c_int), ("param_elems", C.c_int64),
                ("bn_elems", C.c_int64), ("fc_weight_off", C.c_int64), ("fc_bias_off", C.c_int64)]


class LayerLayout(C.Structure):
    _fields_ = [("cin", C.c_int), ("cin_pad", C.c_int), ("cout", C.c_int), ("stride", C.c_int), ("H_in", C.c_int),
                ("H_out", C.c_int), ("w_off", C.c_int64), ("gamma_off", C.c_int64), ("beta_off", C.c_int64),
                ("rmean_off", C.c_int64), ("rvar_off", C.c_int64)]


P, I, L, F, Z = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/wacil.h (tests/test_abi.py checks it)
SIGNATURES = {
    "wacil_version": (I, []),
    "wacil_last_error": (C.c_char_p, []),
    "wacil_fc_forward": (I, [P, L, P, L, P, P, L, I, I, I, P]),
    "wacil_fc_backward": (I, [P, L, P, L, P, L, P, L, P, L, P, I, I, I, P]),
    "wacil_kd_workspace_bytes": (Z, [I]),
    "wacil_kd_fwd_bwd": (I, [P, L, P, L, I, I, F, F, P, P, L, I, I, P, P]),
    "wacil_ce_kd_fwd_bwd": (I, [P, L, P, P, L, I, I, I, F, F, F, P, P, L, P, P]),
    "wacil_ce_kd_fwd_bwd_w": (I, [P, L, P, P, L, I, I, I, F, F, F, F, P, P, L, P, P]),
    "wacil_eval_workspace_bytes": (Z, [I]),
    "wacil_eval_rows": (I, [P, L, P, I, I, I, P, P, P]),
    "wacil_weight_align_workspace_bytes": (Z, [I]),
    "wacil_weight_align": (I, [P, I, I, I, P, I, P, P]),
    "wacil_sgd_step": (I, [P, P, P, L, P, F, F, F, P]),
    "wacil_allreduce_sgd": (I, [P, P, P, P, I, I, L, P, F, F, P, P]),
    "wacil_herd_workspace_bytes": (Z, [L, I]),
    "wacil_herd_barycenter": (I, [P, L, P, I, I, I, P, P, P, P]),
    "wacil_herd_cluster": (I, [P, L, P, I, I, I, P, P, P, P]),
    "wacil_augment_u8": (I, [P, P, P, I, I, I, I, C.POINTER(F), C.POINTER(F), P, P]),
    "wacil_net_layout": (I, [I, I, C.POINTER(NetLayout)]),
    "wacil_net_layer_layout": (I, [I, I, I, C.POINTER(LayerLayout)]),
    "wacil_net_workspace_bytes": (Z, [I, I, I, I]),
    "wacil_net_create": (I, [C.POINTER(P), I, I, I, I, P, P, P, P, Z]),
    "wacil_net_destroy": (None, [P]),
    "wacil_net_fold": (I, [P, P]),
    "wacil_net_forward": (I, [P, P, I, P, P, P]),
    "wacil_net_backward": (I, [P, P, I, P]),
    "wacil_net_last_launches": (I, [P]),
    "wacil_net_set_side_stream": (I, [P, P]),
    "wacil_net_tensor": (I, [P, I, I, C.POINTER(P), C.POINTER(Z)]),
    "wacil_conv3x3_fprop": (I, [P, P, P, P, I, I, I, I, I, I, P]),
    "wacil_conv3x3_bn_relu_workspace_bytes": (Z, [I]),
    "wacil_conv3x3_bn_relu": (I, [P, P, P, P, P, I, P, P, P, P, P, I, I, I, I, I, I, P, P]),
    "wacil_conv3x3_wgrad_workspace_bytes": (Z, [I, I, I, I, I]),
    "wacil_conv3x3_wgrad": (I, [P, P, P, I, I, I, I, I, I, P, P]),
    "wacil_net_set_optimizer": (I, [P, P, P, F, F, I, I, P, P, P]),
    "wacil_net_chain_count": (I, [P, I]),
    "wacil_net_chain_info": (I, [P, I, I, C.POINTER(I), C.POINTER(I), C.POINTER(I), C.POINTER(I), C.POINTER(Z)]),
    "wacil_net_run_chain": (I, [P, I, I, P, P]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libwacil.so (and the tcgen05 self-test) for sm_100a with nvcc; returns the library path."""
    out = subprocess.run(["make", "-C", CSRC, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise WacilError("building libwacil.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WacilError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(nvcc, sm_100a).  There is no fallback path.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().wacil_last_error().decode(errors="replace")
        raise WacilError(f"{what} failed (rc={rc}): {msg}")
