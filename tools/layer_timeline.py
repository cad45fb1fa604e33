"""Per-CTA timelines (clock64 stamps, cycles since kernel entry) of single fused launches of a TRAIN plan at the
benchmarked size, per-layer engine: forward (conv + BatchNorm + shortcut + ReLU) and backward (data gradient + ReLU
mask + BatchNorm backward) elements of each resolution.  Usage: python tools/layer_timeline.py [depth] [batch]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402

import wacil_b200 as w  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 32
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
lib = w._lib.lib()
m = w.CilModel(f"resnet{depth}").cuda()
m.prev_model_adaption(10)
m.prev_model_adaption(10)
m.train()
x = torch.randn(batch, 3, 32, 32, device="cuda")
y = torch.randint(0, 20, (batch,), device="cuda")
plan = m._plan(w._lib.NET_TRAIN, batch)
st = torch.cuda.current_stream().cuda_stream
logits, feats = plan.forward(x, 20)
losses, dl = w.ops.ce_kd_fwd_bwd(logits, y, None, 0)
plan.backward(dl)
torch.cuda.synchronize()
NAMES = {1: "setup", 2: "tma-issued", 3: "w-landed", 4: "a-landed", 5: "mma-issued", 6: "acc0-done", 7: "phase1",
         8: "sums-seen", 13: "constants", 9: "phase2", 10: "exit"}


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


for bwd in (0, 1):
    nch = lib.wacil_net_chain_count(plan.handle, bwd)
    seen = set()
    for i in range(nch):
        first, nlay, grid, ew, smem = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        lib.wacil_net_chain_info(plan.handle, bwd, i, C.byref(first), C.byref(nlay), C.byref(grid), C.byref(ew), C.byref(smem))
        if first.value not in (3, 15, 25, 6, 16, 26):  # one regular (stride-1) layer per resolution
            continue
        us = timed(lambda: w._lib.check(lib.wacil_net_run_chain(plan.handle, bwd, i, None, st), "run_chain"))
        dbg = torch.zeros((grid.value, 16), dtype=torch.int64, device="cuda")
        w._lib.check(lib.wacil_net_run_chain(plan.handle, bwd, i, dbg.data_ptr(), st), "run_chain")
        torch.cuda.synchronize()
        h = dbg.cpu()
        slow = int(torch.argmax(h[:, 10] - h[:, 0]))
        print(f"{'backward' if bwd else 'forward '} element {i:2d} (layer {first.value:2d}) grid {grid.value:3d} smem {smem.value:6d}: "
              f"{us:6.2f} us back-to-back")
        for b in (0, slow):
            print(f"    cta {b:3d}: " + " ".join(f"{NAMES[k]} {int(h[b, k] - h[b, 0])}" for k in (1, 3, 4, 2, 5, 7, 8, 13, 9, 10) if h[b, k]))
