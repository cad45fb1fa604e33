import os, sys
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import golden_inputs as gi
import wacil_b200 as w
from test_gpu_engine import _model_from
for name in gi.TRAIN_CASES:
    case = gi.TRAIN_CASES[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", f"train_{name}.npz"))
    for rep in range(4):
        m = _model_from(w, case["depth"], gi.student_state(case)).train()
        t = _model_from(w, case["depth"], gi.teacher_state(case)).copy().freeze()
        tr = w.TaskTrainer(m, t, case["known"], lr=case["lr"], lambda_kd=case["lambda_kd"], use_graph=False)
        for step in range(case["steps"]):
            x, y = gi.batch(case, step)
            tr.step(x.cuda(), y.cuda())
        sd = m.state_dict()
        worst = (0, None)
        for k in gi.FULL_PARAM_KEYS(case["depth"]):
            ref = g["param/" + k]
            r = np.abs(sd[k].cpu().numpy() - ref).max() / max(np.abs(ref).max(), 1e-3)
            if r > worst[0]: worst = (r, k)
        print(name, rep, "worst rel", round(float(worst[0]), 5), worst[1], flush=True)
