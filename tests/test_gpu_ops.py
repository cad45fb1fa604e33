"""GPU parity of the single-operator entry points against the CPU oracle / golden vectors (through the C ABI)."""
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
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return wacil_b200


# ------------------------------------------------------------------------------------------------ KD
@pytest.mark.parametrize("i", range(len(gi.KD_CASES)))
def test_kd_golden(w, i, golden_dir):
    """fp32 tolerance: loss within 2e-5 relative of the fp32 reference output and within 5e-6 of the fp64 closed form;
    gradient within 2e-7 absolute + 1e-4 relative."""
    g = np.load(os.path.join(golden_dir, "kd.npz"))
    B, C, T, pad = gi.KD_CASES[i]
    s, t = gi.kd_inputs(i)
    s_d, t_d = s.cuda(), t.cuda()
    loss, grad = w.ops.kd_fwd_bwd(s_d[:, :C], t_d, T)  # strided student slice when pad > 0
    s64, t64 = gi.kd_inputs(i, torch.float64)
    ref64 = wo.soft_target(s64[:, :C], t64, T).item()
    assert abs(loss.item() - float(g[f"loss{i}"])) <= 2e-5 * abs(ref64)
    assert abs(loss.item() - ref64) <= 5e-6 * abs(ref64)
    np.testing.assert_allclose(grad.cpu().numpy(), g[f"grad{i}"][:, :C], rtol=1e-4, atol=2e-7)


def test_kd_autograd_module_and_accumulate(w):
    torch.manual_seed(0)
    s = torch.randn(32, 40, device="cuda", requires_grad=True)
    t = torch.randn(32, 30, device="cuda")
    loss = 0.5 * w.SoftTarget(T=2)(s[:, :30], t)
    loss.backward()
    s2 = s.detach().cpu().double().requires_grad_(True)
    ref = 0.5 * wo.soft_target(s2[:, :30], t.cpu().double(), 2.0)
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    np.testing.assert_allclose(s.grad.cpu().numpy(), s2.grad.numpy(), rtol=1e-4, atol=1e-7)
    # accumulate=True adds into an existing gradient buffer
    buf = torch.ones(32, 30, device="cuda")
    w.ops.kd_fwd_bwd(s.detach()[:, :30], t, 2.0, grad_out=buf, accumulate=True)
    np.testing.assert_allclose(buf.cpu().numpy(), 1.0 + 2.0 * s2.grad.numpy()[:, :30], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("B,C,T", [(32, 512, 1.0), (256, 1024, 2.0), (64, 2048, 4.0), (1024, 4096, 8.0),
                                   (128, 8192, 2.0), (7, 130, 2.0), (3, 1021, 2.0)])
def test_kd_microbench_shapes_fp32_and_bf16(w, B, C, T):
    gen = torch.Generator().manual_seed(B * 131 + C)
    s = torch.randn(B, C, generator=gen) * 4
    t = torch.randn(B, C, generator=gen) * 4
    ref = wo.soft_target(s.double(), t.double(), T)
    gref = wo.soft_target_grad(s.double(), t.double(), T)
    loss, grad = w.ops.kd_fwd_bwd(s.cuda(), t.cuda(), T)
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert (grad.cpu().double() - gref).abs().max().item() <= 1e-4 * gref.abs().max().item() + 1e-9
    # property: KD(s, s) == 0 and zero gradient
    l0, g0 = w.ops.kd_fwd_bwd(s.cuda(), s.cuda(), T)
    assert abs(l0.item()) < 1e-5 and g0.abs().max().item() < 1e-7
    # bf16 I/O: same math on the rounded inputs, gradient rounded to bf16
    sb, tb = s.bfloat16(), t.bfloat16()
    refb = wo.soft_target(sb.double(), tb.double(), T)
    lb, gb = w.ops.kd_fwd_bwd(sb.cuda(), tb.cuda(), T)
    assert gb.dtype == torch.bfloat16
    assert abs(lb.item() - refb.item()) <= 1e-5 * abs(refb.item())
    grefb = wo.soft_target_grad(sb.double(), tb.double(), T)
    assert (gb.cpu().double() - grefb).abs().max().item() <= 1e-2 * grefb.abs().max().item()


# ------------------------------------------------------------------------------------------------ CE + KD
@pytest.mark.parametrize("B,Cs,Ck,smooth", [(128, 20, 10, 0.0), (16, 100, 90, 0.0), (33, 10, 0, 0.0),
                                            (64, 1000, 900, 0.1), (8, 2000, 1000, 0.0)])
def test_ce_kd_matches_oracle(w, B, Cs, Ck, smooth):
    gen = torch.Generator().manual_seed(Cs + B)
    z = (torch.randn(B, Cs, generator=gen) * 3).requires_grad_(True)
    y = torch.randint(0, Cs, (B,), generator=gen)
    tz = torch.randn(B, Ck, generator=gen) * 3 if Ck else None
    z64 = z.detach().double().requires_grad_(True)
    ce = wo.cross_entropy(z64, y, smooth)
    kd = 0.5 * wo.soft_target(z64[:, :Ck], tz.double(), 2.0) if Ck else torch.zeros((), dtype=torch.float64)
    (ce + kd).backward()
    losses, dl = w.ops.ce_kd_fwd_bwd(z.detach().cuda(), y.cuda(), tz.cuda() if Ck else None, Ck, 2.0, 0.5, smooth)
    got = losses.cpu().double()
    assert abs(got[0] - ce.item()) <= 5e-6 * abs(ce.item())
    assert abs(got[1] - kd.item()) <= 5e-6 * abs(kd.item()) + 1e-9
    assert abs(got[2] - (ce + kd).item()) <= 5e-6 * abs((ce + kd).item())
    assert int(got[3].item()) == int((z.detach().argmax(1) == y).sum().item())
    assert (dl.cpu().double() - z64.grad).abs().max().item() <= 1e-4 * z64.grad.abs().max().item() + 1e-9
    # CE gradient rows sum to zero (softmax - onehot), KD rows too
    assert dl.sum(dim=1).abs().max().item() < 1e-6


# ------------------------------------------------------------------------------------------------ weight align
@pytest.mark.parametrize("i", range(len(gi.WA_CASES)))
def test_weight_align_golden(w, i, golden_dir):
    g = np.load(os.path.join(golden_dir, "weight_align.npz"))
    heads, d = gi.WA_CASES[i]
    ws = gi.wa_inputs(i)
    W = torch.cat(ws).cuda().contiguous()
    before = W.clone()
    gamma = w.ops.weight_align_(W, heads[-1])
    assert abs(gamma.item() - float(g[f"gamma{i}"])) <= 2e-6 * float(g[f"gamma{i}"])
    np.testing.assert_allclose(W[-heads[-1]:].cpu().numpy(), g[f"new{i}"], rtol=2e-6, atol=1e-8)
    assert torch.equal(W[:-heads[-1]], before[:-heads[-1]])  # old rows untouched, bit for bit


@pytest.mark.parametrize("C,D,n_new,dtype", [(8192, 2048, 1024, torch.float32), (1000, 2048, 100, torch.float32),
                                             (4096, 64, 10, torch.float32), (1000, 512, 125, torch.bfloat16),
                                             (37, 63, 5, torch.float32)])
def test_weight_align_sizes_and_invariant(w, C, D, n_new, dtype):
    gen = torch.Generator().manual_seed(C + D)
    W0 = torch.randn(C, D, generator=gen) * 0.1
    W0[-n_new:] *= 1.7
    W0 = W0.to(dtype)
    gref, nref = wo.weight_align([W0[:-n_new].double(), W0[-n_new:].double()], n_new)
    W = W0.cuda().contiguous()
    gamma = w.ops.weight_align_(W, n_new)
    assert abs(gamma.item() - gref.item()) <= 1e-5 * gref.item()
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert (W[-n_new:].cpu().double() - nref).abs().max().item() <= tol * nref.abs().max().item()
    if dtype == torch.float32:  # WA post-condition: mean row norm of the new rows equals that of the old rows
        n = W.double().norm(dim=1)
        assert abs(n[:-n_new].mean().item() - n[-n_new:].mean().item()) <= 1e-5 * n[:-n_new].mean().item()


# ------------------------------------------------------------------------------------------------ SGD
def test_sgd_three_steps_match_oracle(w):
    gen = torch.Generator().manual_seed(5)
    n = 100003
    p0 = torch.randn(n, generator=gen)
    params, bufs = {"p": p0.clone().double()}, {}
    p, v = p0.cuda(), torch.zeros(n, device="cuda")
    lr_dev = torch.zeros(1, device="cuda")
    for step, lr in enumerate((0.1, 0.05, 0.02)):
        g = torch.randn(n, generator=gen)
        wo.sgd_step(params, {"p": g.double() * 0.5}, bufs, lr)  # grad_scale 0.5 = mean over 2 ranks
        lr_dev.fill_(lr)
        w.ops.sgd_step_(p, g.cuda(), v, lr_dev, 0.9, 5e-4, 0.5)
        np.testing.assert_allclose(p.cpu().numpy(), params["p"].numpy(), rtol=2e-6, atol=2e-7)
        np.testing.assert_allclose(v.cpu().numpy(), bufs["p"].numpy(), rtol=2e-6, atol=2e-7)


# ------------------------------------------------------------------------------------------------ herding
def _herd_c(feats, y, quota):
    from test_oracle_golden import herd_c
    return herd_c(feats, y, quota)


@pytest.mark.parametrize("i", range(len(gi.HERD_CASES)))
def test_herding_bit_exact_vs_c_oracle_and_golden(w, i, golden_dir):
    g = np.load(os.path.join(golden_dir, "herding.npz"))
    feats, y = gi.herd_inputs(i)
    quota = gi.HERD_CASES[i][3]
    idx = w.rehearsal.herd_indices(feats, y, quota).cpu().numpy()
    assert np.array_equal(idx, _herd_c(feats, y, quota))
    assert np.array_equal(idx, g[f"idx{i}"])


def test_herding_cifar_sized_and_edge_cases(w):
    rs = np.random.RandomState(3)
    # 20 classes x 500 samples (+ ragged), D = 64, the quota of task 1 in the 10x10 protocol
    y = np.concatenate([np.full(500 - 7 * (c % 3), c) for c in range(20)])
    rs.shuffle(y)
    feats = np.maximum(rs.normal(0.5, 0.4, size=(len(y), 64)), 0).astype(np.float32)
    feats[5] = 0.0  # an all-zero feature vector (norm 0 + 1e-8)
    idx = w.rehearsal.herd_indices(feats, y, 100).cpu().numpy()
    assert np.array_equal(idx, _herd_c(feats, y, 100))
    assert len(idx) == 20 * 100 and len(np.unique(idx)) == len(idx)
    assert all((y[idx] == c).sum() == 100 for c in range(20))
    # quota larger than a class: every sample is returned, picked ones first
    y2 = np.array([0] * 5 + [1] * 3)
    f2 = rs.rand(8, 64).astype(np.float32)
    idx2 = w.rehearsal.herd_indices(f2, y2, 10).cpu().numpy()
    assert np.array_equal(idx2, _herd_c(f2, y2, 10)) and len(idx2) == 8
    # rehearsal memory bookkeeping through the continuum-style API
    mem = w.RehearsalMemory(memory_size=2000, herding_method="barycenter", dedup_old=True)
    x = np.arange(len(y))
    mem.add(x, y, np.zeros_like(y), feats)
    mx, my, mt = mem.get()
    assert len(mx) == 2000 and mem.memory_per_class == 100 and np.array_equal(np.sort(mx), np.sort(idx))


# ------------------------------------------------------------------------------------------------ conv ops
@pytest.mark.parametrize("N,H,cin,cout,stride", [(8, 32, 16, 16, 1), (8, 32, 16, 32, 2), (8, 16, 32, 32, 1),
                                                 (8, 16, 32, 64, 2), (8, 8, 64, 64, 1), (5, 8, 64, 64, 1)])
def test_conv3x3_fprop_and_wgrad_vs_torch(w, N, H, cin, cout, stride):
    """bf16 operands, fp32 accumulate: compared with torch's CPU conv on the same bf16-rounded operands.
    Tolerance: output 2e-2 abs + 1e-2 rel (one bf16 rounding of the result); wgrad 2e-3 relative to the max."""
    gen = torch.Generator().manual_seed(N + H + cin + cout)
    x = torch.randn(N, H, H, cin, generator=gen).bfloat16()
    wt = (torch.randn(9, cout, cin, generator=gen) * 0.2).bfloat16()
    y, stats = w.ops.conv3x3_fprop(x.cuda(), wt.cuda(), stride, want_stats=True)
    x_nchw = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    w_oihw = wt.float().view(3, 3, cout, cin).permute(2, 3, 0, 1).contiguous().requires_grad_(True)
    ref = torch.nn.functional.conv2d(x_nchw, w_oihw, stride=stride, padding=1)
    got = y.float().cpu().permute(0, 3, 1, 2)
    assert ((got - ref.detach()).abs() <= 2e-2 + 1e-2 * ref.detach().abs()).all()
    r = ref.detach().bfloat16().float()
    np.testing.assert_allclose(stats[:cout].cpu().numpy(), r.sum(dim=(0, 2, 3)).numpy(), rtol=2e-2, atol=0.5)
    np.testing.assert_allclose(stats[cout:].cpu().numpy(), (r * r).sum(dim=(0, 2, 3)).numpy(), rtol=2e-2)
    dy = (torch.randn(N, H // stride, H // stride, cout, generator=gen) * 0.25).bfloat16()
    ref.backward(dy.float().permute(0, 3, 1, 2))
    dw = w.ops.conv3x3_wgrad(x.cuda(), dy.cuda(), stride)  # [9][cin][cout]
    dw_ref = w_oihw.grad.permute(2, 3, 1, 0).reshape(9, cin, cout)
    assert (dw.cpu() - dw_ref).abs().max().item() <= 2e-3 * dw_ref.abs().max().item() + 1e-3


@pytest.mark.parametrize("N,H,cin,cout,stride,res_mode", [(16, 32, 16, 16, 1, 1), (128, 32, 16, 16, 1, 0), (16, 32, 16, 32, 2, 0),
                                                         (16, 16, 32, 32, 1, 2), (32, 8, 64, 64, 1, 1), (7, 8, 64, 64, 1, 1)])
def test_fused_conv_bn_relu_vs_torch(w, N, H, cin, cout, stride, res_mode):
    """The fused layer kernel (conv -> batch statistics -> grid barrier -> BN + shortcut + ReLU) against CPU math on the same
    bf16 operands: y within 1 bf16 ulp, statistics 1e-4, a within 1 ulp of BN applied to the kernel's own y."""
    gen = torch.Generator().manual_seed(N * 7 + H + cin + cout + res_mode)
    x = torch.randn(N, H, H, cin, generator=gen).bfloat16()
    wt = (torch.randn(9, cout, cin, generator=gen) * 0.15).bfloat16()
    gamma = torch.rand(cout, generator=gen) + 0.5
    beta = torch.randn(cout, generator=gen) * 0.2
    rm, rv = torch.randn(cout, generator=gen) * 0.1, torch.rand(cout, generator=gen) + 0.5
    Ho = H // stride
    res = None
    if res_mode == 1:
        res = torch.randn(N, Ho, Ho, cout, generator=gen).bfloat16()
    elif res_mode == 2:
        res = torch.randn(N, 2 * Ho, 2 * Ho, cout // 2, generator=gen).bfloat16()
    rm_d, rv_d = rm.cuda(), rv.cuda()
    y, a, saved = w.ops.conv3x3_bn_relu(x.cuda(), wt.cuda(), gamma.cuda(), beta.cuda(), rm_d, rv_d,
                                        res.cuda() if res is not None else None, res_mode, stride)
    w_oihw = wt.float().view(3, 3, cout, cin).permute(2, 3, 0, 1).contiguous()
    yref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w_oihw, stride=stride, padding=1)
    yg = y.float().cpu().permute(0, 3, 1, 2)
    assert ((yg - yref).abs() <= yref.abs().clamp_min(1e-3) * 2.0 ** -7).all()
    m = N * Ho * Ho
    mean, var = yg.mean(dim=(0, 2, 3)), (yg * yg).mean(dim=(0, 2, 3)) - yg.mean(dim=(0, 2, 3)) ** 2
    sv = saved.cpu()
    assert (sv[0] - mean).abs().max() <= 1e-4 * (mean.abs().max() + var.sqrt().max())
    assert (sv[1] - torch.rsqrt(var + 1e-5)).abs().max() <= 1e-4 * sv[1].abs().max()
    z = (yg - sv[0][None, :, None, None]) * sv[1][None, :, None, None] * gamma[None, :, None, None] + beta[None, :, None, None]
    if res_mode == 1:
        z = z + res.float().permute(0, 3, 1, 2)
    elif res_mode == 2:
        r = res.float().permute(0, 3, 1, 2)[:, :, ::2, ::2]
        z = z + torch.cat([r, torch.zeros_like(r)], dim=1)
    aref = torch.relu(z)
    ag = a.float().cpu().permute(0, 3, 1, 2)
    assert ((ag - aref).abs() <= aref.abs().clamp_min(1e-2) * 2.0 ** -7).all()
    np.testing.assert_allclose(rm_d.cpu().numpy(), (0.9 * rm + 0.1 * mean).numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rv_d.cpu().numpy(), (0.9 * rv + 0.1 * var * m / (m - 1)).numpy(), rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------------------------------------ classifier head
@pytest.mark.parametrize("B,D,C", [(128, 64, 20), (128, 512, 100), (1024, 2048, 1000), (96, 520, 36), (7, 64, 10),
                                   (200, 2048, 300)])
def test_fc_tensor_core_head_vs_fp64(w, B, D, C):
    """wacil_fc_forward / wacil_fc_backward (tcgen05 kind::tf32, fp32 accumulation) against the fp64 product of
    CilClassifier.forward (template.py:99-101) and its autograd backward.  tf32 keeps 10 mantissa bits of each
    operand, so every term carries a relative error <= 2^-10 (two truncations): the bound asserted per element is
    1.5e-3 * sum_k |a_k||b_k| (a layout bug is an O(1) error).  Covers ragged tiles (B, D, C not multiples of the
    128 x 128 x 32 tile) and a dlogits pitch that is not TMA-addressable (C = 10: the wrapper re-pitches it)."""
    gen = torch.Generator().manual_seed(B * 7 + D + C)
    feats = torch.randn(B, D, generator=gen)
    W = torch.randn(C, D, generator=gen) * 0.1
    b = torch.randn(C, generator=gen)
    g = torch.randn(B, C, generator=gen) * 0.05
    fd, Wd, gd = feats.double(), W.double(), g.double()
    tol = 1.5e-3

    got = w.ops.fc_forward(feats.cuda(), W.cuda(), b.cuda()).cpu().double()
    ref = fd @ Wd.T + b.double()
    assert ((got - ref).abs() <= tol * (fd.abs() @ Wd.abs().T) + 1e-6).all()

    dx, dW, db = w.ops.fc_backward(g.cuda(), feats.cuda(), W.cuda())
    assert ((dx.cpu().double() - gd @ Wd).abs() <= tol * (gd.abs() @ Wd.abs()) + 1e-6).all()
    assert ((dW.cpu().double() - gd.T @ fd).abs() <= tol * (gd.abs().T @ fd.abs()) + 1e-6).all()
    np.testing.assert_allclose(db.cpu().numpy(), g.sum(0).numpy(), rtol=1e-5, atol=1e-5)


def test_classifier_module_runs_on_the_tensor_core_head(w):
    """CilClassifier.forward / fc[i](x) (template.py:99-104) go through the same kernels with autograd: the gradient
    lands in the flat arena, heads are row ranges."""
    torch.manual_seed(3)
    model = w.CilModel("resnet20")
    model.prev_model_adaption(10)
    model.prev_model_adaption(10)
    model = model.cuda()
    x = torch.randn(33, 64, device="cuda", requires_grad=True)
    out = model.fc(x)
    assert out.shape == (33, 20)
    Wt, bt = model.fc[1].weight.detach(), model.fc[1].bias.detach()
    head1 = model.fc[1](x)
    ref = torch.nn.functional.linear(x.detach().double(), Wt.double(), bt.double())
    assert (head1.detach().double() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-5
    assert torch.equal(out[:, 10:], head1)
    out.square().sum().backward()
    full_W = torch.cat([model.fc[0].weight, model.fc[1].weight]).detach().double()
    dlog = 2 * out.detach().double()
    assert (x.grad.double() - dlog @ full_W).abs().max().item() <= 2e-3 * (dlog.abs() @ full_W.abs()).max().item()
