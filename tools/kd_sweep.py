"""BASELINE.json config 5: KD-loss and weight-align kernel microbench on one B200, next to PyTorch eager running the
reference's own formulas (utils.py:127-132, template.py:156-166) on the same GPU.  Writes a markdown table to stdout.

CUDA events around `reps` back-to-back launches; buffers rotate so the working set exceeds the 126 MB L2 for the large
shapes (small shapes are launch-latency bound either way and are reported as such)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import torch.nn.functional as F
import wacil_b200 as w

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = "cuda"


def timeit(fn, reps):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps  # us


def kd_case(B, C, T, dtype):
    esz = 4 if dtype == torch.float32 else 2
    nbytes = 3 * B * C * esz + 4 * B
    nb = max(2, min(12, int(300e6 // nbytes) + 1))
    s = [(torch.randn(B, C, device=dev) * 2).to(dtype) for _ in range(nb)]
    t = [(torch.randn(B, C, device=dev) * 2).to(dtype) for _ in range(nb)]
    g = [torch.empty(B, C, device=dev, dtype=dtype) for _ in range(nb)]
    lib, st = w._lib.lib(), torch.cuda.current_stream().cuda_stream
    loss = torch.empty((), device=dev)
    ws = torch.zeros(lib.wacil_kd_workspace_bytes(B), dtype=torch.uint8, device=dev)
    code = 0 if dtype == torch.float32 else 1

    def ours(i):
        k = i % nb
        lib.wacil_kd_fwd_bwd(s[k].data_ptr(), C, t[k].data_ptr(), C, B, C, float(T), 1.0, loss.data_ptr(), g[k].data_ptr(), C, 0,
                             code, ws.data_ptr(), st)

    sg = [x.clone().requires_grad_(True) for x in s]

    def eager(i):  # the reference's SoftTarget forward + backward
        k = i % nb
        sg[k].grad = None
        l = F.kl_div(F.log_softmax(sg[k] / T, dim=1), F.softmax(t[k] / T, dim=1), reduction="batchmean") * T * T
        l.backward()

    reps = 200 if nbytes < 8e6 else 60
    us = timeit(ours, reps)
    us_e = timeit(eager, max(10, reps // 4))
    return nbytes, us, us_e


def wa_case(C, D, n_new, dtype):
    esz = 4 if dtype == torch.float32 else 2
    nbytes = (C + n_new) * D * esz
    nb = max(2, min(10, int(300e6 // (C * D * esz)) + 1))
    Ws = [(torch.randn(C, D, device=dev) * 0.1).to(dtype) for _ in range(nb)]

    lib, st = w._lib.lib(), torch.cuda.current_stream().cuda_stream
    gamma = torch.empty((), device=dev)
    ws = torch.zeros(lib.wacil_weight_align_workspace_bytes(C), dtype=torch.uint8, device=dev)
    code = 0 if dtype == torch.float32 else 1

    def ours(i):
        lib.wacil_weight_align(Ws[i % nb].data_ptr(), C, D, n_new, gamma.data_ptr(), code, ws.data_ptr(), st)

    def eager(i):  # template.py:156-166
        W = Ws[i % nb]
        norms = torch.norm(W.float(), dim=1)
        gamma = norms[:-n_new].mean() / norms[-n_new:].mean()
        W[-n_new:] = (gamma * W[-n_new:].float()).to(dtype)

    reps = 100
    return nbytes, timeit(ours, reps), timeit(eager, 30)


print(f"# cfg-5 microbench (1x B200, HBM peak {PEAK:.0f} GB/s measured)\n")
print("## KD loss fused forward+backward: algorithmic bytes 3*B*C*sizeof + 4*B\n")
print("| B | C | T | dtype | bytes (MB) | ours us | ours GB/s | frac of peak | torch eager us | speed-up |")
print("|---|---|---|---|---|---|---|---|---|---|")
for dtype in (torch.float32, torch.bfloat16):
    for (B, C) in [(32, 512), (128, 512), (1024, 512), (128, 1024), (512, 2048), (1024, 2048), (256, 4096), (1024, 4096),
                   (64, 8192), (512, 8192), (1024, 8192)]:
        for T in ((2.0,) if (B, C) != (1024, 8192) else (1.0, 2.0, 4.0, 8.0)):
            nbytes, us, us_e = kd_case(B, C, T, dtype)
            gbs = nbytes / us / 1e3
            print(f"| {B} | {C} | {T:g} | {'fp32' if dtype == torch.float32 else 'bf16'} | {nbytes / 1e6:.2f} | {us:.2f} | {gbs:.0f} | "
                  f"{gbs / PEAK:.3f} | {us_e:.1f} | {us_e / us:.1f}x |")
print("\n## Weight align (row norms -> gamma -> scale new rows): algorithmic bytes (C_total + n_new)*D*sizeof\n")
print("| C_total | D | n_new | dtype | bytes (MB) | ours us | ours GB/s | frac of peak | torch eager us | speed-up |")
print("|---|---|---|---|---|---|---|---|---|---|")
for dtype in (torch.float32, torch.bfloat16):
    for (C, D, n) in [(100, 64, 10), (1000, 512, 100), (1000, 2048, 100), (8192, 64, 1024), (8192, 512, 1024), (8192, 2048, 1024),
                      (8192, 2048, 10)]:
        nbytes, us, us_e = wa_case(C, D, n, dtype)
        gbs = nbytes / us / 1e3
        print(f"| {C} | {D} | {n} | {'fp32' if dtype == torch.float32 else 'bf16'} | {nbytes / 1e6:.2f} | {us:.2f} | {gbs:.0f} | "
              f"{gbs / PEAK:.3f} | {us_e:.1f} | {us_e / us:.1f}x |")
