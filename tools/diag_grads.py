"""Per-layer gradient error profile of the engine backward vs the CPU oracle (diagnostic, GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import wa_oracle as wo
import wacil_b200 as w

depth, batch, heads = int(sys.argv[1]) if len(sys.argv) > 1 else 20, 16, [10, 10]
perturb = (sys.argv[2] != "plain") if len(sys.argv) > 2 else True
sd = wo.make_state(depth, heads, 7, perturb_bn=perturb)
m = w.CilModel(f"resnet{depth}"); m.load_state_dict(sd); m = m.cuda().train()
gen = torch.Generator().manual_seed(2)
x = torch.randn(batch, 3, 32, 32, generator=gen); y = torch.randint(0, 20, (batch,), generator=gen)
r = wo.train_step(dict(sd), None, x, y, depth, 0, 0.1, {}, apply_update=False)
r64 = wo.train_step({k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}, None, x.double(), y, depth, 0, 0.1, {}, apply_update=False)
logits, feats = m(x.cuda())
loss = torch.nn.functional.cross_entropy(logits, y.cuda()); loss.backward()
a, g = m._arena, m.arena.grad
rel = lambda p, q: ((p.double() - q.double()).norm() / (q.double().norm() + 1e-30)).item()
print("loss", loss.item(), r["loss"].item(), "feat relerr", rel(feats.detach().cpu(), r["feats"]))
for l, (conv, bn) in enumerate(a.names):
    print(f"{l:3d} {conv:32s} W {rel(a.conv_weight(l, g).cpu(), r['grads'][conv+'.weight']):.4f} "
          f"gamma {rel(a.bn_weight(l, g).cpu(), r['grads'][bn+'.weight']):.4f} beta {rel(a.bn_bias(l, g).cpu(), r['grads'][bn+'.bias']):.4f}"
          f"   | fp32-vs-fp64 oracle W {rel(r['grads'][conv+'.weight'], r64['grads'][conv+'.weight']):.1e}")
