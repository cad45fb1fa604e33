"""Tensor-level wrappers over the C ABI: each passes data_ptr()s and the current CUDA stream, nothing else.
Every function raises if its tensors are not CUDA tensors — there is no CPU path."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, lib

_ws_cache = {}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.WacilError("wacil ops run on CUDA tensors only (no CPU fallback)")


def _workspace(key, nbytes, device):
    """Zero-initialised scratch, cached per (kind, size, device); kernels leave their counters re-armed."""
    k = (key, device)
    ws = _ws_cache.get(k)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[k] = ws
    return ws


def _dtype_code(t):
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    raise _lib.WacilError(f"unsupported dtype {t.dtype}")


def kd_fwd_bwd(out_s, out_t, T, scale=1.0, grad_out=None, accumulate=False, need_grad=True, workspace=None):
    """SoftTarget (utils.py:127-132) forward + backward in one kernel.
    out_s [B,C] (any row stride, unit column stride), out_t [B,C].  Returns (loss 0-dim fp32, grad or None)."""
    _need_cuda(out_s, out_t)
    if out_s.dim() != 2 or out_s.shape != out_t.shape or out_s.dtype != out_t.dtype:
        raise _lib.WacilError("kd: out_s/out_t must be 2-D with equal shape and dtype")
    if out_s.stride(1) != 1:
        out_s = out_s.contiguous()
    if out_t.stride(1) != 1:
        out_t = out_t.contiguous()
    B, Cn = out_s.shape
    loss = torch.empty((), dtype=torch.float32, device=out_s.device)
    g = None
    if need_grad:
        g = grad_out if grad_out is not None else torch.empty((B, Cn), dtype=out_s.dtype, device=out_s.device)
        if g.stride(1) != 1:
            raise _lib.WacilError("kd: grad_out needs unit column stride")
    ws = workspace if workspace is not None else _workspace("kd", lib().wacil_kd_workspace_bytes(B), out_s.device)
    check(lib().wacil_kd_fwd_bwd(out_s.data_ptr(), out_s.stride(0), out_t.data_ptr(), out_t.stride(0), B, Cn, float(T),
                                 float(scale), loss.data_ptr(), g.data_ptr() if g is not None else None,
                                 g.stride(0) if g is not None else 0, int(bool(accumulate)), _dtype_code(out_s),
                                 ws.data_ptr(), _stream()), "wacil_kd_fwd_bwd")
    return loss, g


def ce_kd_fwd_bwd(logits, targets, t_logits, known, T=2.0, lambda_kd=0.5, smooth=0.0, losses=None, dlogits=None,
                  workspace=None, ce_weight=1.0):
    """ce_weight*CE + lambda*KD (template.py:259-266; README.md:172-176 for ce_weight = 1 - lambda), forward and
    backward.  Returns (losses[4] = [ce_weight*CE, lambda*KD, sum, top-1 count], dlogits[B,C_seen])."""
    _need_cuda(logits, targets, t_logits)
    B, Cs = logits.shape
    if logits.dtype != torch.float32 or logits.stride(1) != 1 or targets.dtype != torch.int64:
        raise _lib.WacilError("ce_kd: logits fp32 with unit column stride, targets int64")
    if losses is None:
        losses = torch.empty(4, dtype=torch.float32, device=logits.device)
    if dlogits is None:
        dlogits = torch.empty((B, Cs), dtype=torch.float32, device=logits.device)
    ws = workspace if workspace is not None else _workspace("cekd", lib().wacil_kd_workspace_bytes(B), logits.device)
    tl = None if t_logits is None else t_logits
    check(lib().wacil_ce_kd_fwd_bwd_w(logits.data_ptr(), logits.stride(0), targets.data_ptr(),
                                      tl.data_ptr() if tl is not None else None, tl.stride(0) if tl is not None else 0,
                                      B, Cs, int(known) if tl is not None else 0, float(T), float(ce_weight),
                                      float(lambda_kd), float(smooth), losses.data_ptr(), dlogits.data_ptr(),
                                      dlogits.stride(0), ws.data_ptr(), _stream()), "wacil_ce_kd_fwd_bwd_w")
    return losses, dlogits


def _rows_f32(t, what):
    """fp32 2-D tensor whose rows TMA can address: unit column stride, pitch a multiple of 4 floats, 16-byte base."""
    if t.dim() != 2 or t.dtype != torch.float32:
        raise _lib.WacilError(f"{what}: 2-D fp32 tensor expected")
    if t.stride(1) != 1 or t.stride(0) % 4 or t.stride(0) < t.shape[1] or t.data_ptr() % 16:
        pitch = (t.shape[1] + 3) // 4 * 4
        buf = torch.zeros((t.shape[0], pitch), dtype=torch.float32, device=t.device)
        buf[:, : t.shape[1]].copy_(t)
        t = buf[:, : t.shape[1]]
    return t


def fc_forward(feats, W, bias=None, out=None):
    """CilClassifier.forward (template.py:99-101) as tcgen05 tf32 tiles: logits = feats @ W.T + bias.
    feats [B,D], W [C,D] (a row range of the classifier arena), bias [C] or None.  Returns logits [B,C] fp32."""
    _need_cuda(feats, W, bias)
    feats, W = _rows_f32(feats, "feats"), _rows_f32(W, "W")
    B, D = feats.shape
    Cn = W.shape[0]
    if W.shape[1] != D or (bias is not None and (bias.dtype != torch.float32 or bias.numel() != Cn or not bias.is_contiguous())):
        raise _lib.WacilError("fc_forward: W [C,D] and contiguous fp32 bias [C] expected")
    if out is None:
        out = torch.empty((B, Cn), dtype=torch.float32, device=feats.device)
    check(lib().wacil_fc_forward(feats.data_ptr(), feats.stride(0), W.data_ptr(), W.stride(0),
                                 bias.data_ptr() if bias is not None else None, out.data_ptr(), out.stride(0), B, D, Cn,
                                 _stream()), "wacil_fc_forward")
    return out


def fc_backward(dlogits, feats, W, need_dfeats=True, need_dW=True, need_db=True):
    """Backward of fc_forward: returns (dfeats [B,D] | None, dW [C,D] | None, db [C] | None)."""
    _need_cuda(dlogits, feats, W)
    dlogits, feats, W = _rows_f32(dlogits, "dlogits"), _rows_f32(feats, "feats"), _rows_f32(W, "W")
    B, D = feats.shape
    Cn = W.shape[0]
    if dlogits.shape != (B, Cn) or W.shape[1] != D:
        raise _lib.WacilError("fc_backward: dlogits [B,C], feats [B,D], W [C,D] expected")
    dev = feats.device
    df = torch.empty((B, D), dtype=torch.float32, device=dev) if need_dfeats else None
    dW = torch.empty((Cn, D), dtype=torch.float32, device=dev) if need_dW else None
    db = torch.empty((Cn,), dtype=torch.float32, device=dev) if need_db else None
    check(lib().wacil_fc_backward(dlogits.data_ptr(), dlogits.stride(0), feats.data_ptr(), feats.stride(0), W.data_ptr(),
                                  W.stride(0), df.data_ptr() if df is not None else None, D,
                                  dW.data_ptr() if dW is not None else None, D,
                                  db.data_ptr() if db is not None else None, B, D, Cn, _stream()), "wacil_fc_backward")
    return df, dW, db


def weight_align_(W, n_new, gamma_out=None):
    """In-place WA on the contiguous classifier matrix W [C_total, D] (template.py:156-166); returns gamma (0-dim)."""
    _need_cuda(W)
    if W.dim() != 2 or not W.is_contiguous():
        raise _lib.WacilError("weight_align: W must be a contiguous [C_total, D] matrix")
    if gamma_out is None:
        gamma_out = torch.empty((), dtype=torch.float32, device=W.device)
    ws = _workspace("wa", lib().wacil_weight_align_workspace_bytes(W.shape[0]), W.device)
    check(lib().wacil_weight_align(W.data_ptr(), W.shape[0], W.shape[1], int(n_new), gamma_out.data_ptr(),
                                   _dtype_code(W), ws.data_ptr(), _stream()), "wacil_weight_align")
    return gamma_out


def sgd_step_(p, g, v, lr_dev, momentum=0.9, weight_decay=5e-4, grad_scale=1.0):
    _need_cuda(p, g, v, lr_dev)
    for t in (p, g, v):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise _lib.WacilError("sgd: flat contiguous fp32 arenas")
    check(lib().wacil_sgd_step(p.data_ptr(), g.data_ptr(), v.data_ptr(), p.numel(), lr_dev.data_ptr(), float(momentum),
                               float(weight_decay), float(grad_scale), _stream()), "wacil_sgd_step")


def allreduce_sgd_(p, v, peer_grad_ptrs, peer_pad_ptrs, world, rank, lr_dev, epoch_ws, momentum=0.9, weight_decay=5e-4):
    """Fused DDP mean + SGD over NVLink peer memory (wacil_allreduce_sgd).  peer_*_ptrs: ctypes arrays of `world` device
    pointers (symmetric-memory / IPC mappings of every rank's gradient arena and flag pad)."""
    _need_cuda(p, v, lr_dev, epoch_ws)
    check(lib().wacil_allreduce_sgd(p.data_ptr(), v.data_ptr(), peer_grad_ptrs, peer_pad_ptrs, int(world), int(rank), p.numel(),
                                    lr_dev.data_ptr(), float(momentum), float(weight_decay), epoch_ws.data_ptr(), _stream()),
          "wacil_allreduce_sgd")


def herd_barycenter(features, class_offsets, quota):
    """features [n,D] fp32 grouped by class; class_offsets int32 [n_classes+1].  Returns (idx [n_classes,quota]
    int32 row-local, counts [n_classes] int32)."""
    _need_cuda(features, class_offsets)
    if features.dtype != torch.float32 or not features.is_contiguous() or class_offsets.dtype != torch.int32:
        raise _lib.WacilError("herd: features fp32 contiguous, class_offsets int32")
    n, D = features.shape
    k = class_offsets.numel() - 1
    idx = torch.empty((k, quota), dtype=torch.int32, device=features.device)
    cnt = torch.empty((k,), dtype=torch.int32, device=features.device)
    ws = _workspace("herd", lib().wacil_herd_workspace_bytes(n, D), features.device)
    check(lib().wacil_herd_barycenter(features.data_ptr(), n, class_offsets.data_ptr(), k, D, int(quota), idx.data_ptr(),
                                      cnt.data_ptr(), ws.data_ptr(), _stream()), "wacil_herd_barycenter")
    return idx, cnt


def herd_cluster(features, class_offsets, quota):
    """continuum's herd_closest_to_cluster on the GPU: same layout as herd_barycenter."""
    _need_cuda(features, class_offsets)
    if features.dtype != torch.float32 or not features.is_contiguous() or class_offsets.dtype != torch.int32:
        raise _lib.WacilError("herd: features fp32 contiguous, class_offsets int32")
    n, D = features.shape
    k = class_offsets.numel() - 1
    idx = torch.empty((k, quota), dtype=torch.int32, device=features.device)
    cnt = torch.empty((k,), dtype=torch.int32, device=features.device)
    ws = _workspace("herdc", n * 4, features.device)
    check(lib().wacil_herd_cluster(features.data_ptr(), n, class_offsets.data_ptr(), k, D, int(quota), idx.data_ptr(),
                                   cnt.data_ptr(), ws.data_ptr(), _stream()), "wacil_herd_cluster")
    return idx, cnt


def eval_rows_(logits, targets, meter, topk=5):
    """Adds one validation batch to the float64[8] device meter (template.py:169-188): [samples, batches, sum of batch
    mean CE, top-1 count, top-k count]; k = min(topk, C) like timm.utils.accuracy(topk=(1, min(5, C)))."""
    _need_cuda(logits, targets, meter)
    B, Cn = logits.shape
    if logits.dtype != torch.float32 or logits.stride(1) != 1 or targets.dtype != torch.int64:
        raise _lib.WacilError("eval_rows: logits fp32 with unit column stride, targets int64")
    if meter.dtype != torch.float64 or meter.numel() < 5 or not meter.is_contiguous():
        raise _lib.WacilError("eval_rows: meter is a contiguous float64 tensor of >= 5 elements")
    ws = _workspace("eval", lib().wacil_eval_workspace_bytes(B), logits.device)
    check(lib().wacil_eval_rows(logits.data_ptr(), logits.stride(0), targets.data_ptr(), B, Cn, min(int(topk), Cn),
                                meter.data_ptr(), ws.data_ptr(), _stream()), "wacil_eval_rows")
    return meter


def augment_u8(store, idx=None, aug=None, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), pad=4, out=None):
    """Device data path (utils.py:210-251): store uint8 [n,H,W,3] -> fp32 NCHW [B,3,H,W]: gather idx, RandomCrop(H, pad)
    + horizontal flip from the packed `aug` words (dy | dx << 8 | flip << 16), ToTensor, Normalize(mean, std)."""
    _need_cuda(store, idx, aug)
    if store.dtype != torch.uint8 or store.dim() != 4 or store.shape[3] != 3 or not store.is_contiguous():
        raise _lib.WacilError("augment_u8: store must be a contiguous uint8 [n,H,W,3] tensor")
    n, H, W_, _ = store.shape
    B = n if idx is None else idx.numel()
    if idx is not None and (idx.dtype != torch.int64 or not idx.is_contiguous()):
        raise _lib.WacilError("augment_u8: idx int64 contiguous")
    if aug is not None and (aug.dtype != torch.int32 or aug.numel() != B or not aug.is_contiguous()):
        raise _lib.WacilError("augment_u8: aug int32 [B] contiguous")
    if out is None:
        out = torch.empty((B, 3, H, W_), dtype=torch.float32, device=store.device)
    elif out.dtype != torch.float32 or tuple(out.shape) != (B, 3, H, W_) or not out.is_contiguous():
        raise _lib.WacilError("augment_u8: out must be contiguous fp32 [B,3,H,W]")
    import ctypes as C
    m3, s3 = (C.c_float * 3)(*[float(v) for v in mean]), (C.c_float * 3)(*[float(v) for v in std])
    check(lib().wacil_augment_u8(store.data_ptr(), idx.data_ptr() if idx is not None else None,
                                 aug.data_ptr() if aug is not None else None, B, H, W_, int(pad), m3, s3,
                                 out.data_ptr(), _stream()), "wacil_augment_u8")
    return out


def conv3x3_fprop(x, w, stride=1, want_stats=False):
    """x [N,H,W,cin] bf16 NHWC, w [9,cout,cin] bf16 -> y [N,H/s,W/s,cout] bf16 (+ stats [2*cout] fp32)."""
    _need_cuda(x, w)
    N, H, W_, cin = x.shape
    cout = w.shape[1]
    y = torch.empty((N, H // stride, W_ // stride, cout), dtype=torch.bfloat16, device=x.device)
    stats = torch.zeros(2 * cout, dtype=torch.float32, device=x.device) if want_stats else None
    check(lib().wacil_conv3x3_fprop(x.data_ptr(), w.data_ptr(), y.data_ptr(),
                                    stats.data_ptr() if stats is not None else None, N, H, W_, cin, cout, stride,
                                    _stream()), "wacil_conv3x3_fprop")
    return (y, stats) if want_stats else y


def conv3x3_bn_relu(x, w, gamma, beta, running_mean, running_var, res=None, res_mode=0, stride=1):
    """One fused launch: y = conv3x3(x, w); a = relu(batchnorm_train(y) + shortcut).  x [N,H,W,cin] bf16, w [9,cout,cin]
    bf16, gamma/beta/running_* fp32 [cout] (running statistics updated in place).  Returns (y, a, mean_rstd[2,cout])."""
    _need_cuda(x, w, gamma, beta, running_mean, running_var, res)
    N, H, W_, cin = x.shape
    cout = w.shape[1]
    y = torch.empty((N, H // stride, W_ // stride, cout), dtype=torch.bfloat16, device=x.device)
    a = torch.empty_like(y)
    saved = torch.empty((2, cout), dtype=torch.float32, device=x.device)
    # one workspace per problem shape: the kernel's published-sum rows are laid out by (grid, cout) and only ever
    # cleared by the launches themselves
    ws = _workspace(("convbn", N, H, cin, cout, stride), lib().wacil_conv3x3_bn_relu_workspace_bytes(cout), x.device)
    check(lib().wacil_conv3x3_bn_relu(x.data_ptr(), w.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                      res.data_ptr() if res is not None else None, int(res_mode), y.data_ptr(),
                                      a.data_ptr(), saved.data_ptr(), running_mean.data_ptr(), running_var.data_ptr(),
                                      N, H, W_, cin, cout, stride, ws.data_ptr(), _stream()), "wacil_conv3x3_bn_relu")
    return y, a, saved


def conv3x3_wgrad(x, dy, stride=1):
    """x [N,H,W,cin] bf16, dy [N,H/s,W/s,cout] bf16 -> dW [9,cin,cout] fp32."""
    _need_cuda(x, dy)
    N, H, W_, cin = x.shape
    cout = dy.shape[3]
    dw = torch.empty((9, cin, cout), dtype=torch.float32, device=x.device)
    ws = torch.empty(lib().wacil_conv3x3_wgrad_workspace_bytes(N, H, cin, cout, stride), dtype=torch.uint8, device=x.device)
    check(lib().wacil_conv3x3_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), N, H, W_, cin, cout, stride, ws.data_ptr(),
                                    _stream()), "wacil_conv3x3_wgrad")
    return dw
