"""The dominant kernel of the step in isolation, for ncu: the fused stage-1 layer (wacil_conv3x3_bn_relu: 3x3 conv 16->16
over 128x32x32 + batch statistics + exchange of the partial sums + BatchNorm + shortcut + ReLU) and the weight gradient of the same layer
(wacil_conv3x3_wgrad), five launches each over rotating buffers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402

import wacil_b200 as w  # noqa: E402

lib = w._lib.lib()
st = torch.cuda.current_stream().cuda_stream
B = 128
nb = 12
xs = [torch.randn(B, 32, 32, 16, device="cuda").bfloat16() for _ in range(nb)]
rs = [torch.randn(B, 32, 32, 16, device="cuda").bfloat16() for _ in range(nb)]
wt = (torch.randn(9, 16, 16, device="cuda") * 0.1).bfloat16()
gamma, beta = torch.ones(16, device="cuda"), torch.zeros(16, device="cuda")
rm, rv = torch.zeros(16, device="cuda"), torch.ones(16, device="cuda")
ys = torch.empty((nb, B, 32, 32, 16), dtype=torch.bfloat16, device="cuda")
acts = torch.empty((nb, B, 32, 32, 16), dtype=torch.bfloat16, device="cuda")
saved = torch.empty(32, device="cuda")
wsb = torch.zeros(lib.wacil_conv3x3_bn_relu_workspace_bytes(16), dtype=torch.uint8, device="cuda")
for k in range(5):
    rc = lib.wacil_conv3x3_bn_relu(xs[k].data_ptr(), wt.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rs[k].data_ptr(), 1,
                                   ys[k].data_ptr(), acts[k].data_ptr(), saved.data_ptr(), rm.data_ptr(), rv.data_ptr(),
                                   B, 32, 32, 16, 16, 1, wsb.data_ptr(), st)
    assert rc == 0, lib.wacil_last_error()
for k in range(5):
    w.ops.conv3x3_wgrad(xs[k], rs[k + 5], 1)
torch.cuda.synchronize()
print("done")
