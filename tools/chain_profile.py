"""Per-chain timing of a TRAIN / INFER plan at the benchmarked size: CUDA events around stand-alone chain launches
(wacil_net_run_chain) and the clock64 stamps of the epilogue phases (phase 1 done / grid barrier passed / phase 2 done)
per layer.  Usage: python tools/chain_profile.py [depth] [batch]"""
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
side = torch.cuda.Stream()
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=20):
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


logits, feats = plan.forward(x, 20)
losses, dl = w.ops.ce_kd_fwd_bwd(logits, y, None, 0)
plan.backward(dl)
torch.cuda.synchronize()
print(f"resnet{depth} batch {batch}: forward {timed(lambda: plan.forward(x, 20)):.1f} us, "
      f"backward (one stream) {timed(lambda: plan.backward(dl)):.1f} us")
plan.set_side_stream(side)
print(f"backward (weight gradients on a side stream) {timed(lambda: plan.backward(dl)):.1f} us")
tplan = m.copy().freeze()._plan(w._lib.NET_INFER, batch)
tplan.fold()
print(f"inference forward {timed(lambda: tplan.forward(x, 20)):.1f} us")
for pl, name in ((plan, "train"), (tplan, "infer")):
    for bwd in ((0, 1) if pl is plan else (0,)):
        nch = lib.wacil_net_chain_count(pl.handle, bwd)
        for i in range(nch):
            first, nlay, grid, ew, smem = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
            lib.wacil_net_chain_info(pl.handle, bwd, i, C.byref(first), C.byref(nlay), C.byref(grid), C.byref(ew), C.byref(smem))
            dbg = torch.zeros((grid.value, nlay.value, 48), dtype=torch.int64, device="cuda")
            us = timed(lambda: w._lib.check(lib.wacil_net_run_chain(pl.handle, bwd, i, None, st), "run_chain"))
            w._lib.check(lib.wacil_net_run_chain(pl.handle, bwd, i, dbg.data_ptr(), st), "run_chain")
            torch.cuda.synchronize()
            d = dbg.cpu()
            line = f"{name} {'bwd' if bwd else 'fwd'} chain {i}: first layer {first.value}, {nlay.value} layers, grid {grid.value}, " \
                   f"ew {ew.value}, smem {smem.value}: {us:.1f} us ({us / nlay.value:.2f} us/layer)"
            print(line)
            if d.abs().sum() > 0:
                for cta in (0, grid.value // 2):
                    c0 = d[cta]
                    base = c0[0, 4].item() or c0[0, 0].item()
                    names = ["p1", "bar", "p2", "pub", "dep", "ld", "op", "mma", "sum", "cst", "p2t0", "acc0", "p1e"]
                    order = [4, 6, 5, 11, 7, 12, 0, 1, 8, 9, 10, 2, 3]
                    rows = []
                    for l in range(min(nlay.value, 4)):
                        rows.append(f"L{l}: " + " ".join(f"{names[k]} {c0[l, k].item() - base}" for k in order if c0[l, k].item()))
                    print(f"   cta{cta} cycles since its first dependency check:  " + " | ".join(rows))
                    if nlay.value > 1:
                        b1 = c0[1, 4].item()
                        for nm, off in (("producer slot free", 16), ("mma operands landed", 24), ("mma issued+commit", 32)):
                            print(f"      layer 1 per tile, {nm}: " + " ".join(str(c0[1, off + k].item() - b1) for k in range(8) if c0[1, off + k].item()))

# ---- do CTAs of two kernels share an SM?  a backward chain launch on the main stream, weight-gradient launches on a side stream
import wacil_b200.ops as ops  # noqa: E402
xa = torch.randn(128, 32, 32, 16, device="cuda").bfloat16()
dya = torch.randn(128, 32, 32, 16, device="cuda").bfloat16()
s2 = torch.cuda.Stream()
nb = lib.wacil_net_chain_count(plan.handle, 1)


def both():
    s2.wait_stream(torch.cuda.current_stream())
    w._lib.check(lib.wacil_net_run_chain(plan.handle, 1, nb - 2, None, st), "run_chain")
    with torch.cuda.stream(s2):
        for _ in range(4):
            ops.conv3x3_wgrad(xa, dya, 1)
    torch.cuda.current_stream().wait_stream(s2)


def only_chain():
    w._lib.check(lib.wacil_net_run_chain(plan.handle, 1, nb - 2, None, st), "run_chain")


def only_wgrad():
    for _ in range(4):
        ops.conv3x3_wgrad(xa, dya, 1)


print(f"co-residency probe: chain alone {timed(only_chain):.1f} us, 4 stage-1 wgrads alone {timed(only_wgrad):.1f} us, "
      f"both on two streams {timed(both):.1f} us")

# ---- the single-layer fused kernel (conv_gemm_kernel, cooperative) beside the weight-gradient kernel, and two chains side by side
gam, bet = torch.ones(16, device="cuda"), torch.zeros(16, device="cuda")
rm_, rv_ = torch.zeros(16, device="cuda"), torch.ones(16, device="cuda")
wt = (torch.randn(9, 16, 16, device="cuda") * 0.1).bfloat16()
yb, ab = torch.empty_like(xa), torch.empty_like(xa)
savedb = torch.empty(32, device="cuda")
wsb = torch.zeros(lib.wacil_conv3x3_bn_relu_workspace_bytes(16), dtype=torch.uint8, device="cuda")


def fused4(stream):
    for _ in range(4):
        rc = lib.wacil_conv3x3_bn_relu(xa.data_ptr(), wt.data_ptr(), gam.data_ptr(), bet.data_ptr(), dya.data_ptr(), 1,
                                       yb.data_ptr(), ab.data_ptr(), savedb.data_ptr(), rm_.data_ptr(), rv_.data_ptr(),
                                       128, 32, 32, 16, 16, 1, wsb.data_ptr(), stream)
        assert rc == 0


def both_old():
    s2.wait_stream(torch.cuda.current_stream())
    fused4(st)
    with torch.cuda.stream(s2):
        only_wgrad()
    torch.cuda.current_stream().wait_stream(s2)


print(f"co-residency probe (single-layer fused kernel): 4 fused layers alone {timed(lambda: fused4(st)):.1f} us, "
      f"with 4 wgrads on a second stream {timed(both_old):.1f} us")


def two_chains():
    s2.wait_stream(torch.cuda.current_stream())
    w._lib.check(lib.wacil_net_run_chain(plan.handle, 0, 0, None, st), "run_chain")
    with torch.cuda.stream(s2):
        w._lib.check(lib.wacil_net_run_chain(tplan.handle, 0, 0, None, s2.cuda_stream), "run_chain")
    torch.cuda.current_stream().wait_stream(s2)


print(f"co-residency probe (train chain 0 + inference chain 0 on two streams): {timed(two_chains):.1f} us")
