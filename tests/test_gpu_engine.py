"""GPU parity of the ResNet engine, the drop-in model API and the training step against the CPU oracle.

Tolerances (bf16 tensor-core convolutions with fp32 accumulation, fp32 BatchNorm statistics / classifier / loss,
versus the reference's pure-fp32 arithmetic).  Calibrated on B200 and stated per assertion:
  features / logits        |err| <= 3e-2 * max|ref|   (31 bf16 layers deep)
  loss                     2e-2 relative
  parameter gradients      ||g - g_ref|| / ||g_ref|| <= 2e-2 per tensor vs the mixed-precision model of the same math
                           (oracle/wa_oracle_bf16.py); <= 0.5 vs the pure-fp32 oracle (ReLU-mask flips, see that file)
  running statistics       2e-3 vs the mixed-precision model, 2e-2 vs fp32, relative to the tensor's max
"""
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
import wa_oracle as wo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def w():
    import wacil_b200
    assert torch.cuda.is_available()
    return wacil_b200


def _mild_state(depth, heads, seed):
    """Seeded state with BN close to a trained regime (activations O(1)) so bf16 error is representative."""
    sd = wo.make_state(depth, heads, seed, perturb_bn=True)
    return sd


def _model_from(w, depth, sd, max_classes=100):
    m = w.CilModel(f"resnet{depth}", max_classes=max_classes)
    m.load_state_dict(sd)
    return m.cuda()


def relerr(a, b):
    return (a.double() - b.double()).norm().item() / (b.double().norm().item() + 1e-30)


@pytest.mark.parametrize("depth,batch", [(20, 8), (32, 16), (32, 7)])
def test_infer_forward_matches_oracle_eval(w, depth, batch):
    sd = _mild_state(depth, [10, 10], 3)
    m = _model_from(w, depth, sd).eval()
    x = torch.randn(batch, 3, 32, 32, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref_logits, ref_feats = wo.cil_forward(sd, x, depth, train=False)
    logits, feats = m(x.cuda())
    assert (feats.cpu() - ref_feats).abs().max().item() <= 3e-2 * ref_feats.abs().max().item()
    assert (logits.cpu() - ref_logits).abs().max().item() <= 3e-2 * max(ref_logits.abs().max().item(), 1.0)
    fv = m.extract_vector(x.cuda())
    assert torch.equal(fv, feats)


def _gpu_state(m, plan, depth):
    """Read the TRAIN plan's saved forward tensors back as the NCHW fp32 state oracle/wa_oracle_bf16.backward takes."""
    nl = depth - 1
    nchw = lambda t: t.float().cpu().permute(0, 3, 1, 2).contiguous()
    st = {"ins": [], "ys": [], "means": [], "rstds": [], "acts": []}
    a0 = nchw(plan.tensor(0))
    for l in range(nl):
        st["ins"].append(a0[:, :3] if l == 0 else st["acts"][l - 1])
        st["ys"].append(nchw(plan.tensor(1, l)))
        saved = plan.tensor(3, l).cpu()
        st["means"].append(saved[0].clone()); st["rstds"].append(saved[1].clone())
        st["acts"].append(nchw(plan.tensor(2, l)))
    return st, a0


@pytest.mark.parametrize("depth,batch,heads", [(20, 16, [10, 10]), (32, 32, [10, 10, 10]), (32, 9, [10]),
                                               (32, 128, [10, 10]),      # the benchmarked configuration (task 1)
                                               (32, 128, [10] * 10),     # task 9: 100 seen classes
                                               (56, 32, [10, 10]), (110, 32, [10])])
def test_train_forward_backward_matches_oracle(w, depth, batch, heads):
    """Three layers of evidence (a deep bf16 ReLU network is chaotic in its roundings, so end-to-end agreement between
    two bf16 evaluations is ~3e-3 at best and gradients disagree by several % through ReLU-mask flips):
      1. end to end vs the fp32 oracle (pinned to the reference): features/logits 3e-2, loss 2e-2;
      2. every layer's kernel outputs vs CPU math applied to the GPU's OWN layer input: conv within 1 bf16 ulp,
         BatchNorm statistics 1e-4, BN+residual+ReLU within 1 ulp;
      3. the whole backward vs oracle/wa_oracle_bf16.backward run on the GPU's saved forward state (fixed masks make
         it linear): every parameter gradient within 2e-2 relative."""
    import wa_oracle_bf16 as wb
    import torch.nn.functional as F
    sd = _mild_state(depth, heads, 7)
    m = _model_from(w, depth, sd).train()
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(batch, 3, 32, 32, generator=gen)
    y = torch.randint(0, sum(heads), (batch,), generator=gen)
    r = wo.train_step(dict(sd), None, x, y, depth, 0, 0.1, {}, apply_update=False)
    logits, feats = m(x.cuda())
    # 1. end to end vs fp32
    assert (feats.detach().cpu() - r["feats"]).abs().max().item() <= 3e-2 * r["feats"].abs().max().item()
    assert (logits.detach().cpu() - r["logits"]).abs().max().item() <= 3e-2 * max(r["logits"].abs().max().item(), 1.0)
    loss = torch.nn.functional.cross_entropy(logits, y.cuda())
    assert abs(loss.item() - r["loss"].item()) <= 2e-2 * abs(r["loss"].item())
    # 2. layer by layer from the GPU's own inputs
    plan = m._plan(w._lib.NET_TRAIN, batch)
    st, a0 = _gpu_state(m, plan, depth)
    assert torch.equal(a0[:, :3], x.bfloat16().float()) and a0[:, 3:].abs().max().item() == 0.0
    layers = wo.conv_bn_names(depth)
    block_in = None
    for l, (conv, bn, cin, cout, stride) in enumerate(layers):
        wq = sd[conv + ".weight"].bfloat16().float()
        yref = F.conv2d(st["ins"][l], wq, stride=stride, padding=1)
        ulp = yref.abs().clamp_min(1e-3) * 2.0 ** -7
        dev = (st["ys"][l] - yref).abs() / ulp  # fp32 accumulation order differs from the CPU's: a value sitting on a
        # rounding boundary may land one bf16 step away (1 ulp of the larger neighbour = up to 2 of this scale)
        assert (dev <= 2.0).all() and (dev > 1.0).float().mean().item() <= 1e-4, (conv, dev.max())
        yq = st["ys"][l]
        mean = yq.mean(dim=(0, 2, 3))
        var = (yq * yq).mean(dim=(0, 2, 3)) - mean * mean
        assert (st["means"][l] - mean).abs().max().item() <= 1e-4 * (mean.abs().max().item() + var.sqrt().max().item())
        assert (st["rstds"][l] - torch.rsqrt(var + 1e-5)).abs().max().item() <= 1e-4 * st["rstds"][l].abs().max().item()
        z = (yq - st["means"][l][None, :, None, None]) * st["rstds"][l][None, :, None, None] \
            * sd[bn + ".weight"][None, :, None, None] + sd[bn + ".bias"][None, :, None, None]
        is_b = l >= 2 and l % 2 == 0
        if is_b:
            ca = layers[l - 1]
            z = z + (wo.downsample_a(block_in) if (ca[4] != 1 or ca[2] != ca[3]) else block_in)
        aref = F.relu(z)
        assert ((st["acts"][l] - aref).abs() <= aref.abs().clamp_min(1e-2) * 2.0 ** -7).all(), bn
        if l == 0 or is_b:
            block_in = st["acts"][l]
    st["feats"] = st["acts"][-1].mean(dim=(2, 3))
    assert (feats.detach().cpu() - st["feats"]).abs().max().item() <= 1e-5 * st["feats"].abs().max().item()
    # 3. backward from the GPU's forward state
    loss.backward()
    dlogits = wb.ce_kd_dlogits(y, None, 0)(logits.detach().cpu())
    ref_grads = wb.backward(sd, st, depth, dlogits)
    a = m._arena
    g = m.arena.grad
    worst_model, worst_fp32 = 0.0, 0.0
    for l, (conv, bn) in enumerate(a.names):
        for key, view in ((conv + ".weight", a.conv_weight(l, g)), (bn + ".weight", a.bn_weight(l, g)),
                          (bn + ".bias", a.bn_bias(l, g))):
            e = relerr(view.cpu(), ref_grads[key])
            worst_model = max(worst_model, e)
            # bf16 gradient tensors between the layers: the error of the chain grows with its length (calibrated at
            # ResNet-32: 2e-2; ResNet-110 has 3.4x the layers)
            assert e <= 2e-2 * max(1.0, (depth - 1) / 31.0), (key, e)
            worst_fp32 = max(worst_fp32, relerr(view.cpu(), r["grads"][key]))
    # vs the pure-fp32 oracle the per-tensor figure is dominated by ReLU-mask flips (0.3 .. 0.8 for these tiny
    # batches at random init); what must hold is that the overall descent direction is the same
    flat_ours = torch.cat([a.conv_weight(l, g).reshape(-1).cpu() for l in range(len(a.names))])
    flat_ref = torch.cat([r["grads"][conv + ".weight"].reshape(-1) for conv, _ in a.names])
    cos = torch.nn.functional.cosine_similarity(flat_ours.double(), flat_ref.double(), dim=0).item()
    # (109 bf16 layers decorrelate further from the fp32 evaluation than 31 do: the direction check is for depth <= 56)
    assert cos >= (0.85 if depth <= 56 else 0.4), cos
    assert worst_fp32 <= 2.0, worst_fp32
    c = sum(heads)
    gw_ref = torch.cat([ref_grads[f"fc.heads.{i}.weight"] for i in range(len(heads))])
    gb_ref = torch.cat([ref_grads[f"fc.heads.{i}.bias"] for i in range(len(heads))])
    assert relerr(a.fc_weight(g)[:c].cpu(), gw_ref) <= 1e-4 and relerr(a.fc_bias(g)[:c].cpu(), gb_ref) <= 1e-4
    # running statistics: momentum 0.1, unbiased variance, from the batch statistics the kernels saved
    for l, (conv, bn) in enumerate(a.names):
        mcount = batch * st["ys"][l].shape[2] * st["ys"][l].shape[3]
        rm = 0.9 * sd[bn + ".running_mean"] + 0.1 * st["means"][l]
        var = 1.0 / st["rstds"][l] ** 2 - 1e-5
        rv = 0.9 * sd[bn + ".running_var"] + 0.1 * var * (mcount / (mcount - 1))
        assert (a.running_mean(l).cpu() - rm).abs().max().item() <= 1e-5 * max(rm.abs().max().item(), 0.1)
        assert (a.running_var(l).cpu() - rv).abs().max().item() <= 1e-4 * rv.abs().max().item()
        rm, rv = r["new_buffers"][bn + ".running_mean"], r["new_buffers"][bn + ".running_var"]
        assert (a.running_mean(l).cpu() - rm).abs().max().item() <= 2e-2 * max(rm.abs().max().item(), 0.1)
        assert (a.running_var(l).cpu() - rv).abs().max().item() <= 2e-2 * rv.abs().max().item()
    print(f"worst per-tensor gradient relative error: {worst_model:.4f} vs the model on the GPU's forward state, "
          f"{worst_fp32:.4f} vs the fp32 oracle (cosine of all conv-weight gradients {cos:.4f})")


@pytest.mark.parametrize("name", list(gi.TRAIN_CASES))
def test_two_training_steps_match_reference_golden(w, name, golden_dir):
    """The full step (student fwd/bwd, teacher fwd, CE + 0.5*KD(T=2), SGD) against outputs of the REAL reference."""
    case = gi.TRAIN_CASES[name]
    g = np.load(os.path.join(golden_dir, f"train_{name}.npz"))
    m = _model_from(w, case["depth"], gi.student_state(case)).train()
    t = _model_from(w, case["depth"], gi.teacher_state(case)).copy().freeze()
    tr = w.TaskTrainer(m, t, case["known"], lr=case["lr"], lambda_kd=case["lambda_kd"], use_graph=False)
    for step in range(case["steps"]):
        x, y = gi.batch(case, step)
        losses = tr.step(x.cuda(), y.cuda()).cpu()
        ref = g[f"losses{step}"]
        assert abs(losses[0].item() - ref[0]) <= 3e-2 * abs(ref[0]), (step, losses, ref)
        assert abs(losses[1].item() - ref[1]) <= 5e-2 * abs(ref[1]) + 1e-3, (step, losses, ref)
    sd = m.state_dict()
    for k in gi.FULL_PARAM_KEYS(case["depth"]):
        ref = g["param/" + k]
        # bf16 activations vs the fp32 reference: the worst tensor (the first conv) lands 1-2.6 % off; the BatchNorm sums
        # and weight gradients are reduced in a fixed order (no atomics), so the figure does not vary run to run
        # (test_training_is_bit_reproducible) -> 3 % bound
        assert np.abs(sd[k].cpu().numpy() - ref).max() <= 3e-2 * max(np.abs(ref).max(), 1e-3), k
    m.eval()
    lg, ft = m(gi.batch(case, 0)[0].cuda())
    assert np.abs(ft.cpu().numpy() - g["eval_feats"]).max() <= 5e-2 * np.abs(g["eval_feats"]).max()


def test_fused_trainer_equals_autograd_path_and_graph_replay(w):
    """Same step three ways: reference-style autograd (model(x); loss.backward(); torch SGD), fused eager, fused
    CUDA graph.  They run the same kernels, so parameters agree to fp32 round-off."""
    depth, heads = 20, [10, 10]
    sd, tsd = _mild_state(depth, heads, 11), _mild_state(depth, heads[:1], 12)
    gen = torch.Generator().manual_seed(4)
    xs = [torch.randn(16, 3, 32, 32, generator=gen).cuda() for _ in range(4)]
    ys = [torch.randint(0, 20, (16,), generator=gen).cuda() for _ in range(4)]
    teacher = _model_from(w, depth, tsd).copy().freeze()
    # (a) the reference's own step sequence on the drop-in classes (template.py:258-271)
    ma = _model_from(w, depth, sd).train()
    opt = torch.optim.SGD(ma.parameters(), lr=0.05, momentum=0.9, weight_decay=5e-4)
    kd = w.SoftTarget(T=2)
    for x, y in zip(xs, ys):
        logits, _ = ma(x)
        loss = torch.nn.functional.cross_entropy(logits, y) + 0.5 * kd(logits[:, :10], teacher(x)[0])
        opt.zero_grad()
        loss.backward()
        opt.step()
    results = []
    for use_graph in (False, True):
        mb = _model_from(w, depth, sd).train()
        tr = w.TaskTrainer(mb, teacher, 10, lr=0.05, lambda_kd=0.5, use_graph=use_graph)
        for x, y in zip(xs, ys):
            last = tr.step(x, y).clone()
        results.append((mb, last))
        assert relerr(mb.arena.data[: mb._arena.layout.fc_weight_off], ma.arena.data[: ma._arena.layout.fc_weight_off]) < 2e-2
        assert relerr(mb.bn_running, ma.bn_running) < 5e-3
    assert abs(results[0][1][2].item() - results[1][1][2].item()) <= 1e-2 * abs(results[0][1][2].item())
    assert abs(results[0][1][2].item() - loss.item()) <= 1e-2 * abs(loss.item())


def test_full_batch_properties(w):
    """BASELINE-size batch (128 images, ResNet-32, 20 classes): size-independent properties instead of a CPU replay.
    backward is linear in dlogits; classifier-bias gradient of CE sums to ~0; frozen teacher is unchanged by training."""
    m = w.CilModel("resnet32").cuda()
    m.prev_model_adaption(10)
    m.prev_model_adaption(10)
    m.train()
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(128, 3, 32, 32, generator=gen).cuda()
    y = torch.randint(0, 20, (128,), generator=gen).cuda()
    plan = m._plan(w._lib.NET_TRAIN, 128)
    logits, feats = plan.forward(x, 20)
    assert torch.isfinite(logits).all() and torch.isfinite(feats).all()
    losses, dl = w.ops.ce_kd_fwd_bwd(logits, y, None, 0)
    plan.backward(dl)
    g1 = m._arena.grads.clone()
    plan.backward(dl * 2)
    g2 = m._arena.grads.clone()
    assert relerr(g2, 2 * g1) < 2e-2
    assert abs(m._arena.fc_bias(g1)[:20].sum().item()) < 1e-4
    assert g1[m._arena.layout.fc_weight_off + 20 * 64: m._arena.layout.fc_bias_off].abs().max().item() == 0.0
    teacher = m.copy().freeze()
    before = teacher.arena.data.clone()
    tr = w.TaskTrainer(m, None, 0, lr=0.02)
    l0 = tr.step(x, y)[2].item()
    for _ in range(12):
        l1 = tr.step(x, y)[2].item()
    assert l1 < l0, (l0, l1)  # the same batch twelve more times: the loss goes down
    assert torch.equal(teacher.arena.data, before)


def test_model_api_surface(w):
    m = w.CilModel("resnet20")
    assert m.feature_dim == 64 and m.fc is None
    m = m.cuda()
    m.prev_model_adaption(10)
    m.prev_model_adaption(5)
    assert len(m.fc) == 2 and m.fc[1].weight.shape == (5, 64) and m.fc[-1].bias.shape == (5,)
    x = torch.randn(4, 3, 32, 32).cuda()
    logits, feats = m(x)
    assert logits.shape == (4, 15) and feats.shape == (4, 64)
    # weight_align scales only the newest head (template.py:166) and returns gamma
    w_old, w_new = m.fc[0].weight.clone(), m.fc[1].weight.clone()
    gamma = m.weight_align(5, verbose=False)
    ref_gamma = w_old.norm(dim=1).mean() / w_new.norm(dim=1).mean()
    assert abs(gamma.item() - ref_gamma.item()) < 1e-5 * ref_gamma.item()
    assert torch.equal(m.fc[0].weight, w_old)
    assert torch.allclose(m.fc[1].weight, gamma * w_new, rtol=1e-6, atol=1e-8)
    # state_dict round trip under the reference's keys
    sd = m.state_dict()
    m2 = w.CilModel("resnet20").cuda()
    m2.load_state_dict(sd)
    m.eval(); m2.eval()
    assert torch.equal(m(x)[0], m2(x)[0])
    # capacity growth beyond max_classes keeps the parameters
    m3 = w.CilModel("resnet20", max_classes=10).cuda()
    m3.prev_model_adaption(10)
    keep = m3.fc[0].weight.clone()
    m3.prev_model_adaption(10)
    assert m3.num_classes == 20 and torch.equal(m3.fc[0].weight, keep)
    assert m3.eval()(x)[0].shape == (4, 20)


def test_mini_class_incremental_run(w):
    """Two tasks end to end through train_task / update_model / herding on separable synthetic classes."""
    import argparse
    torch.manual_seed(0)
    args = w.get_args_parser().parse_args([])
    args.backbone, args.batch_size, args.num_epochs, args.lr, args.memory_size = "resnet20", 64, 4, 0.05, 40
    args.rank, args.world_size = 0, 1
    gen = torch.Generator().manual_seed(0)
    protos = torch.randn(4, 3, 32, 32, generator=gen)

    def task_data(classes, n=96):
        y = torch.tensor(classes).repeat_interleave(n)
        x = protos[y] + 0.3 * torch.randn(len(y), 3, 32, 32, generator=gen)
        return x.cuda(), y.cuda()

    model = w.CilModel(args).cuda()
    memory = w.RehearsalMemory(args.memory_size, "barycenter", False, dedup_old=True)
    teacher, args.known_classes = None, 0
    for task_id, classes in enumerate([[0, 1], [2, 3]]):
        args.task_id = task_id
        x, y = task_data(classes)
        if task_id > 0:
            mx, my, _ = memory.get()
            x = torch.cat([x, torch.from_numpy(mx).cuda()]); y = torch.cat([y, torch.from_numpy(my).cuda()])
        w.update_model(model, args, 2, "before")
        hist = w.train_task(model, teacher, (x, y), args, log=None)
        assert hist[-1]["loss"] < hist[0]["loss"]
        teacher = w.update_model(model, args, 2, "after")
        from wacil_b200.train import extract_features
        feats = extract_features(model, x, 64)
        memory.add(x.cpu().numpy(), y.cpu().numpy(), np.full(len(y), task_id), feats)
        args.known_classes += 2
    assert len(memory) == 40 and memory.memory_per_class == 10
    model.eval()
    xt, yt = task_data([0, 1, 2, 3], n=32)
    acc = (model(xt)[0].argmax(1) == yt).float().mean().item()
    assert acc > 0.7, acc


@pytest.mark.parametrize("heads,known", [([10, 10], 10), ([10] * 10, 90)])
def test_full_batch_step_matches_oracle(w, heads, known):
    """ONE full TaskTrainer.step at the benchmarked size (ResNet-32, batch 128, teacher + CE + 0.5*KD(T=2) + SGD) against
    wo.train_step (template.py:255-271 restated; pinned to the unmodified reference by tests/test_oracle_golden.py), for
    task 1 (20 seen / 10 known classes) and task 9 (100 / 90).  Tolerances: CE 2e-2 relative, KD 5e-2 relative + 1e-3,
    student logits 3e-2 of max|ref|, teacher-driven KD through the folded-BN inference path included."""
    depth, batch = 32, 128
    sd, tsd = _mild_state(depth, heads, 21), _mild_state(depth, heads[:-1], 22)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(batch, 3, 32, 32, generator=gen)
    y = torch.randint(0, sum(heads), (batch,), generator=gen)
    ref = wo.train_step(dict(sd), tsd, x, y, depth, known, 0.05, {}, lambda_kd=0.5, T=2.0, apply_update=False)
    m = _model_from(w, depth, sd).train()
    t = _model_from(w, depth, tsd).copy().freeze()
    tr = w.TaskTrainer(m, t, known, lr=0.05, lambda_kd=0.5, T=2.0, use_graph=False)
    losses = tr.step(x.cuda(), y.cuda()).cpu()
    ce, kd = losses[0].item(), losses[1].item()
    assert abs(ce - ref["loss_ce"].item()) <= 2e-2 * abs(ref["loss_ce"].item()), (ce, ref["loss_ce"])
    assert abs(kd - ref["loss_kd"].item()) <= 5e-2 * abs(ref["loss_kd"].item()) + 1e-3, (kd, ref["loss_kd"])
    st = tr._state(batch)
    logits = st.plan_s._logits[: batch * sum(heads)].view(batch, sum(heads)).cpu()
    assert (logits - ref["logits"]).abs().max().item() <= 3e-2 * max(ref["logits"].abs().max().item(), 1.0)
    top1_ref = (ref["logits"].argmax(1) == y).sum().item()
    assert abs(losses[3].item() - top1_ref) <= 2  # logits within 3e-2 can flip a near-tie
    # the update moved the parameters along the oracle's gradient: cosine of the conv-weight step vs -grad
    a = m._arena
    step = torch.cat([(a.conv_weight(l).cpu() - sd[conv + ".weight"]).reshape(-1) for l, (conv, _) in enumerate(a.names)])
    gref = torch.cat([(ref["grads"][conv + ".weight"] + 5e-4 * sd[conv + ".weight"]).reshape(-1) for conv, _ in a.names])
    cos = torch.nn.functional.cosine_similarity(step.double(), -gref.double(), dim=0).item()
    assert cos >= 0.85, cos


def test_training_is_bit_reproducible(w):
    """The reference runs with cudnn.deterministic = True (template.py:57).  Here every reduction has a fixed order
    (per-CTA BatchNorm partial sums, per-CTA weight-gradient slots, last-CTA loss finish): two runs of the same four
    steps — eager and CUDA-graph replay, teacher included — must give bit-identical parameters, statistics and losses."""
    depth, heads = 32, [10, 10]
    sd, tsd = _mild_state(depth, heads, 31), _mild_state(depth, heads[:1], 32)
    gen = torch.Generator().manual_seed(6)
    xs = [torch.randn(128, 3, 32, 32, generator=gen).cuda() for _ in range(4)]
    ys = [torch.randint(0, 20, (128,), generator=gen).cuda() for _ in range(4)]
    outs = []
    for run in range(3):
        m = _model_from(w, depth, sd).train()
        t = _model_from(w, depth, tsd).copy().freeze()
        tr = w.TaskTrainer(m, t, 10, lr=0.05, lambda_kd=0.5, use_graph=(run == 2))
        ls = [tr.step(x, y).clone() for x, y in zip(xs, ys)]
        torch.cuda.synchronize()
        outs.append((m.arena.data.clone(), m.bn_running.clone(), torch.stack(ls)))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    # the captured graph launches the same kernels with the same grids
    assert torch.equal(outs[0][0], outs[2][0]) and torch.equal(outs[0][1], outs[2][1])


def test_host_fed_steps_equal_device_fed_steps(w):
    """TaskTrainer.step with pinned host batches (the reference's loader hands the loop host tensors,
    template.py:255-257) stages them through two device buffers on a copy stream; five steps — more than the two
    staging buffers, so both are reused — must leave exactly the parameters the device-fed steps leave."""
    depth, heads = 20, [10, 10]
    sd, tsd = _mild_state(depth, heads, 51), _mild_state(depth, heads[:1], 52)
    gen = torch.Generator().manual_seed(9)
    xs = [torch.randn(64, 3, 32, 32, generator=gen).pin_memory() for _ in range(5)]
    ys = [torch.randint(0, 20, (64,), generator=gen).pin_memory() for _ in range(5)]
    outs = []
    for host in (False, True):
        m = _model_from(w, depth, sd).train()
        t = _model_from(w, depth, tsd).copy().freeze()
        tr = w.TaskTrainer(m, t, 10, lr=0.05, lambda_kd=0.5, use_graph=True)
        ls = [tr.step(x if host else x.cuda(), y if host else y.cuda()).clone() for x, y in zip(xs, ys)]
        torch.cuda.synchronize()
        outs.append((m.arena.data.clone(), torch.stack(ls)))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("names", [["fc"], ["backbone"]])
def test_partial_freeze_pins_its_range(w, names):
    """CilModel.freeze(["fc"]) / freeze(["backbone"]) (template.py:128-144): the optimiser leaves the frozen element range
    of the flat arena bit-identical and moves the rest."""
    depth, heads = 20, [10, 10]
    sd = _mild_state(depth, heads, 41)
    m = _model_from(w, depth, sd)
    m.freeze(names)
    assert m.training and m.arena.requires_grad
    cut = m._arena.layout.fc_weight_off
    before = m.arena.detach().clone()
    tr = w.TaskTrainer(m, None, 0, lr=0.05, use_graph=False)
    assert not tr.fused_opt
    gen = torch.Generator().manual_seed(8)
    for _ in range(3):
        tr.step(torch.randn(16, 3, 32, 32, generator=gen).cuda(), torch.randint(0, 20, (16,), generator=gen).cuda())
    after = m.arena.detach()
    frozen = slice(cut, None) if names == ["fc"] else slice(0, cut)
    moving = slice(0, cut) if names == ["fc"] else slice(cut, cut + 20 * 64)
    assert torch.equal(after[frozen], before[frozen])
    assert not torch.equal(after[moving], before[moving])
    assert m.copy().trainable_range() == m.trainable_range()
    assert m.freeze(["all"]).trainable_range() == (0, 0) and not m.training


def test_chain_engine_stays_under_parity():
    """WACIL_ENGINE=chain (several layers per persistent conv_chain_kernel launch instead of one cooperative launch per
    layer; the default is the per-layer engine, profiles/r02_chain_experiments.md) runs the layer-by-layer parity, the
    golden two-step and the bit-reproducibility tests in a subprocess (the switch is read once per process)."""
    import subprocess
    import sys
    env = dict(os.environ, WACIL_ENGINE="chain")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", __file__, "-k",
                        "(train_forward_backward and not 110 and not 56) or two_training_steps or infer_forward or "
                        "bit_reproducible or full_batch_step"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
