"""One fused backward element (data gradient of a 32x32 layer + ReLU mask + BatchNorm backward of the layer below) of a
ResNet-32 TRAIN plan at batch 128, launched alone a few times, for ncu (-k regex:bn_bwd)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402

import wacil_b200 as w  # noqa: E402

lib = w._lib.lib()
m = w.CilModel("resnet32").cuda()
m.prev_model_adaption(10)
m.prev_model_adaption(10)
m.train()
x = torch.randn(128, 3, 32, 32, device="cuda")
y = torch.randint(0, 20, (128,), device="cuda")
plan = m._plan(w._lib.NET_TRAIN, 128)
st = torch.cuda.current_stream().cuda_stream
logits, feats = plan.forward(x, 20)
losses, dl = w.ops.ce_kd_fwd_bwd(logits, y, None, 0)
plan.backward(dl)
torch.cuda.synchronize()
for _ in range(4):
    w._lib.check(lib.wacil_net_run_chain(plan.handle, 1, 24, None, st), "run_chain")  # element 24: layer 6
torch.cuda.synchronize()
print("done")
