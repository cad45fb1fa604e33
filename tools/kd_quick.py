"""A/B timing of the KD kernel at the BASELINE shape (B=1024, C=8192), min over 7 rounds of 40 launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import wacil_b200 as w
lib, st = w._lib.lib(), torch.cuda.current_stream().cuda_stream
for dtype, code, esz in ((torch.float32, 0, 4), (torch.bfloat16, 1, 2)):
    B, C, nb = 1024, 8192, 6
    s = [(torch.randn(B, C, device="cuda") * 2).to(dtype) for _ in range(nb)]
    t = [(torch.randn(B, C, device="cuda") * 2).to(dtype) for _ in range(nb)]
    g = [torch.empty(B, C, device="cuda", dtype=dtype) for _ in range(nb)]
    loss = torch.empty((), device="cuda")
    ws = torch.zeros(lib.wacil_kd_workspace_bytes(B), dtype=torch.uint8, device="cuda")
    best = 1e9
    for r in range(7):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(40):
            k = i % nb
            lib.wacil_kd_fwd_bwd(s[k].data_ptr(), C, t[k].data_ptr(), C, B, C, 2.0, 1.0, loss.data_ptr(), g[k].data_ptr(), C, 0, code,
                                 ws.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, 1e3 * e0.elapsed_time(e1) / 40)
    nbytes = 3 * B * C * esz + 4 * B
    print(f"STREAM={os.environ.get('WACIL_KD_STREAM', '1')} MATH={os.environ.get('WACIL_KD_MATH', '-')} {dtype}: {best:.2f} us  {nbytes / best / 1e3:.0f} GB/s  loss {loss.item():.6f}")
