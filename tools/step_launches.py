"""One WA training step (ResNet-32, batch 128, task 1) inside a cudaProfiler range, for an ncu launch list:
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/step_launches.py"""
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
tr = w.TaskTrainer(m, t, 10, lr=0.05, lambda_kd=0.5, use_graph=False)
x = torch.randn(128, 3, 32, 32, device="cuda")
y = torch.randint(0, 20, (128,), device="cuda")
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches in the step:", tr.launches_per_step)
