"""CUDA-graph timings of the parts of one training step (ResNet-32, batch 128, task 1): student forward, teacher forward,
both on two branches, backward with / without the weight-gradient branch, whole step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402

import wa_oracle as wo  # noqa: E402
import wacil_b200 as w  # noqa: E402

sd, tsd = wo.make_state(32, [10, 10], 0, perturb_bn=False), wo.make_state(32, [10], 1, perturb_bn=False)
m = w.CilModel("resnet32"); m.load_state_dict(sd); m = m.cuda().train()
t = w.CilModel("resnet32"); t.load_state_dict(tsd); t = t.cuda().copy().freeze()
x = torch.randn(128, 3, 32, 32, device="cuda")
y = torch.randint(0, 20, (128,), device="cuda")
ps = m._plan(w._lib.NET_TRAIN, 128)
pt = t._plan(w._lib.NET_INFER, 128)
pt.fold()
side, side2 = torch.cuda.Stream(), torch.cuda.Stream()


def graph_time(fn, reps=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


logits, _ = ps.forward(x, 20)
losses, dl = w.ops.ce_kd_fwd_bwd(logits, y, None, 0)
ps.backward(dl)
print(f"student forward          {graph_time(lambda: ps.forward(x, 20)):8.1f} us")
print(f"teacher forward          {graph_time(lambda: pt.forward(x, 10)):8.1f} us")


def both():
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        pt.forward(x, 10)
    ps.forward(x, 20)
    main.wait_stream(side)


print(f"student || teacher       {graph_time(both):8.1f} us")
print(f"backward, one stream     {graph_time(lambda: ps.backward(dl)):8.1f} us")
ps.set_side_stream(side2)
print(f"backward, wgrad branch   {graph_time(lambda: ps.backward(dl)):8.1f} us")
tr = w.TaskTrainer(m, t, 10, lr=0.05, lambda_kd=0.5, use_graph=True)
for _ in range(4):
    tr.step(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    tr.step(x, y)
e1.record()
torch.cuda.synchronize()
print(f"whole step (TaskTrainer) {1e3 * e0.elapsed_time(e1) / 50:8.1f} us, {tr.launches_per_step} launches")
