"""allreduce_sgd_kernel on ONE GPU for ncu (a multi-rank command cannot run under ncu): world = 2 with both "ranks" mapped
to this device's own buffers — the same code path (flag handshake, rank-ordered 16-byte reads of every rank's gradient
arena, SGD) with local instead of NVLink reads.  ResNet-32 arena size (464,154 floats padded to a multiple of 4)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402

import wacil_b200 as w  # noqa: E402

n = 464156
p = torch.randn(n, device="cuda")
v = torch.zeros(n, device="cuda")
g = [torch.randn(n, device="cuda") for _ in range(2)]
pad = torch.zeros(64, dtype=torch.int32, device="cuda")  # both "ranks" share the pad: each slot is written by its own rank id
lr = torch.tensor([0.05], device="cuda")
ep = torch.zeros(2, dtype=torch.int32, device="cuda")
gp = (C.c_void_p * 2)(g[0].data_ptr(), g[1].data_ptr())
pp = (C.c_void_p * 2)(pad.data_ptr(), pad.data_ptr())
# rank 0 of a world of 2 whose peer flags are pre-set: emulate the peer by writing its flags ahead of every launch
for it in range(4):
    pad[1] = it + 1            # peer "ready" flag for this epoch (slot rank 1)
    pad[16 + 1] = it + 1       # peer "done reading" flag (kMaxRanks = 16 slots further)
    w.ops.allreduce_sgd_(p, v, gp, pp, 2, 0, lr, ep)
torch.cuda.synchronize()
print("done", float(p.abs().mean()))
