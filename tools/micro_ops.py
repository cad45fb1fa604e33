"""Micro-benchmarks of the two HBM-bound north-star kernels (BASELINE.json config 5) with CUDA events over rotating buffer
sets larger than L2:  python tools/micro_ops.py [kd|wa|all] [quick]
  KD   : wacil_kd_fwd_bwd  B x C logits, fp32 and bf16 I/O         algorithmic bytes 3*B*C*sizeof + 4*B
  WA   : wacil_weight_align [C, D] fp32 / bf16, n_new new rows      algorithmic bytes (C + n_new)*D*sizeof"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import json  # noqa: E402

import torch  # noqa: E402

import wacil_b200 as w  # noqa: E402

PEAK = 6583.5
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
lib = w._lib.lib()
st = torch.cuda.current_stream().cuda_stream
what = sys.argv[1] if len(sys.argv) > 1 else "all"
quick = len(sys.argv) > 2


def timed(fn, reps):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


if what in ("kd", "all"):
    shapes = [(1024, 8192)] if quick else [(1024, 8192), (1024, 4096), (1024, 2048), (512, 8192), (256, 1024), (128, 512)]
    for dt, code in ((torch.float32, 0), (torch.bfloat16, 1)):
        for B, C in shapes:
            es = 4 if dt == torch.float32 else 2
            nb = max(3, int(300e6 // (3 * B * C * es)) + 1)
            ss = [torch.randn(B, C, device="cuda").to(dt) for _ in range(nb)]
            ts = [torch.randn(B, C, device="cuda").to(dt) for _ in range(nb)]
            gs = [torch.empty(B, C, device="cuda", dtype=dt) for _ in range(nb)]
            loss = torch.empty((), device="cuda")
            ws = torch.zeros(lib.wacil_kd_workspace_bytes(B), dtype=torch.uint8, device="cuda")

            def run(i):
                k = i % nb
                rc = lib.wacil_kd_fwd_bwd(ss[k].data_ptr(), C, ts[k].data_ptr(), C, B, C, 2.0, 1.0, loss.data_ptr(),
                                          gs[k].data_ptr(), C, 0, code, ws.data_ptr(), st)
                assert rc == 0
            us = timed(run, 60)
            nbytes = 3 * B * C * es + 4 * B
            print(f"KD {str(dt)[6:]:8s} B={B:5d} C={C:5d}: {us:7.2f} us  {nbytes / us / 1e3:7.1f} GB/s  {nbytes / us / 1e3 / PEAK:5.3f} of peak")
            del ss, ts, gs

if what in ("wa", "all"):
    shapes = [(8192, 2048, 1024)] if quick else [(8192, 2048, 1024), (8192, 512, 1024), (1000, 2048, 100), (4096, 64, 10), (100, 64, 10)]
    for dt, code in ((torch.float32, 0), (torch.bfloat16, 1)):
        for C, D, nn in shapes:
            es = 4 if dt == torch.float32 else 2
            nb = max(3, int(300e6 // (C * D * es)) + 1)
            Ws = [(torch.randn(C, D, device="cuda") * 0.1).to(dt) for _ in range(nb)]
            g = torch.empty((), device="cuda")
            ws = torch.zeros(lib.wacil_weight_align_workspace_bytes(C), dtype=torch.uint8, device="cuda")

            def run(i):
                rc = lib.wacil_weight_align(Ws[i % nb].data_ptr(), C, D, nn, g.data_ptr(), code, ws.data_ptr(), st)
                assert rc == 0, lib.wacil_last_error()
            us = timed(run, 60)
            nbytes = (C + nn) * D * es
            print(f"WA {str(dt)[6:]:8s} C={C:5d} D={D:5d} n_new={nn:5d}: {us:7.2f} us  {nbytes / us / 1e3:7.1f} GB/s  {nbytes / us / 1e3 / PEAK:5.3f} of peak")
            del Ws
