"""Repeat the ResNet-110 forward layer check and report where the conv outputs deviate by more than a bf16 ulp."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import wa_oracle as wo  # noqa: E402
import wacil_b200 as w  # noqa: E402

depth, batch = int(sys.argv[1]) if len(sys.argv) > 1 else 110, 32
sd = wo.make_state(depth, [10], 7, perturb_bn=True)
layers = wo.conv_bn_names(depth)
for rep in range(3):
    m = w.CilModel(f"resnet{depth}", max_classes=100)
    m.load_state_dict(sd)
    m = m.cuda().train()
    x = torch.randn(batch, 3, 32, 32, generator=torch.Generator().manual_seed(2))
    logits, feats = m(x.cuda())
    plan = m._plan(w._lib.NET_TRAIN, batch)
    nchw = lambda t: t.float().cpu().permute(0, 3, 1, 2).contiguous()
    prev = nchw(plan.tensor(0))[:, :3]
    nbad = 0
    for l, (conv, bn, cin, cout, stride) in enumerate(layers):
        wq = sd[conv + ".weight"].bfloat16().float()
        yref = F.conv2d(prev, wq, stride=stride, padding=1)
        yg = nchw(plan.tensor(1, l))
        ulp = yref.abs().clamp_min(1e-3) * 2.0 ** -7
        bad = (yg - yref).abs() > ulp
        if bad.any():
            nbad += 1
            idx = bad.nonzero()
            worst = ((yg - yref).abs() / ulp).max().item()
            print(f"rep {rep} layer {l} {conv}: {int(bad.sum())} of {bad.numel()} outside 1 ulp, worst {worst:.2f} ulp, "
                  f"max|yref| {yref.abs().max().item():.3g}, first at (n,c,h,w) {idx[0].tolist()}, "
                  f"rows hit {sorted(set(idx[:, 2].tolist()))[:12]}")
        prev = nchw(plan.tensor(2, l))
    print(f"rep {rep}: {nbad} layers with deviations; feats finite {torch.isfinite(feats).all().item()}")
