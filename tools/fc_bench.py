"""Classifier-head GEMMs: the tcgen05 tf32 kernels (wacil_fc_forward / wacil_fc_backward) next to torch (cuBLAS fp32 and
TF32) at the head sizes of BASELINE configs 2-5.  CUDA events over 50 calls after 10 warm-ups."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wacil_b200 as w


def timed(fn, n=50, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for B, D, C in [(128, 64, 20), (128, 512, 100), (1024, 512, 1000), (1024, 2048, 1000), (1024, 2048, 8192)]:
    f = torch.randn(B, D, device="cuda")
    W = torch.randn(C, D, device="cuda") * 0.1
    b = torch.randn(C, device="cuda")
    g = torch.randn(B, C, device="cuda")
    out = torch.empty(B, C, device="cuda")
    t_f = timed(lambda: w.ops.fc_forward(f, W, b, out=out))
    t_b = timed(lambda: w.ops.fc_backward(g, f, W))
    torch.backends.cuda.matmul.allow_tf32 = False
    t_tf = timed(lambda: torch.nn.functional.linear(f, W, b))
    t_tb = timed(lambda: (g @ W, g.t() @ f, g.sum(0)))
    torch.backends.cuda.matmul.allow_tf32 = True
    t_tf32 = timed(lambda: torch.nn.functional.linear(f, W, b))
    t_tb32 = timed(lambda: (g @ W, g.t() @ f, g.sum(0)))
    fl = 2.0 * B * D * C
    print(f"B={B:5d} D={D:5d} C={C:5d}: forward {t_f:7.2f} us ({fl / t_f * 1e-6:7.1f} TFLOP/s) | torch fp32 {t_tf:7.2f} "
          f"tf32 {t_tf32:7.2f} || backward {t_b:7.2f} us ({2 * fl / t_b * 1e-6:7.1f} TFLOP/s) | torch fp32 {t_tb:7.2f} "
          f"tf32 {t_tb32:7.2f}")
