"""Run the first forward chain of a TRAIN plan a few times (for ncu captures): python tools/chain_one.py [bwd] [index]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402

import wacil_b200 as w  # noqa: E402

bwd = int(sys.argv[1]) if len(sys.argv) > 1 else 0
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lib = w._lib.lib()
m = w.CilModel("resnet32").cuda()
m.prev_model_adaption(10)
m.prev_model_adaption(10)
m.train()
x = torch.randn(128, 3, 32, 32, device="cuda")
y = torch.randint(0, 20, (128,), device="cuda")
plan = m._plan(w._lib.NET_TRAIN, 128)
logits, feats = plan.forward(x, 20)
losses, dl = w.ops.ce_kd_fwd_bwd(logits, y, None, 0)
plan.backward(dl)
torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream
for _ in range(4):
    w._lib.check(lib.wacil_net_run_chain(plan.handle, bwd, idx, None, st), "run_chain")
torch.cuda.synchronize()
print("done")
