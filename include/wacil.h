/* wacil.h — C ABI of libwacil.so: the B200-native (sm_100a) WA class-incremental training step.
 *
 * The reference (G-U-N/a-PyTorch-Tutorial-to-Class-Incremental-Learning) is pure Python/PyTorch and has no
 * FFI layer; each entry point below names the reference code it replaces (file:line under /root/reference).
 * The Python host (package a-pytorch-tutorial-to-class-incremental-learning_b200) binds these with ctypes and
 * mirrors the reference's class/method names on top.
 *
 * Conventions
 *  - Every function returns int: 0 = OK; < 0 = argument error (WACIL_E_*); > 0 = cudaError_t of a failed launch.
 *    No exception crosses the ABI.  wacil_last_error() returns a thread-local message for the last failure.
 *  - Ownership: the caller allocates every device buffer (inputs, outputs, arenas, workspaces).  The library
 *    never allocates or frees device memory and keeps no pointer beyond what a wacil_net_t was bound to.
 *  - Stream-ordered on the `stream` argument (a cudaStream_t passed as void*), no host synchronisation, no
 *    hidden allocation: every call can be captured into a CUDA graph.  Device = the current device.
 *  - Plain pointers and sizes only; no torch types.  dtype codes: WACIL_F32 = 0, WACIL_BF16 = 1.
 */
#ifndef WACIL_H_
#define WACIL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WACIL_F32 0
#define WACIL_BF16 1

#define WACIL_E_SHAPE (-1)     /* bad size / shape */
#define WACIL_E_ALIGN (-2)     /* pointer or stride alignment */
#define WACIL_E_DTYPE (-3)     /* unsupported dtype code */
#define WACIL_E_NULL (-4)      /* required pointer is NULL */
#define WACIL_E_WORKSPACE (-5) /* workspace too small */
#define WACIL_E_DRIVER (-6)    /* cuTensorMapEncodeTiled unavailable / failed */
#define WACIL_E_STATE (-7)     /* call order (e.g. backward before forward) */

int wacil_version(void);
const char* wacil_last_error(void);

/* ------------------------------------------------------------------------------------------------------
 * Losses
 * --------------------------------------------------------------------------------------------------- */

/* SoftTarget.forward (utils.py:127-132) fused with its backward:
 *   loss = scale * (T^2 / B) * sum_b sum_c p_t (log p_t - log p_s),  p = softmax(. / T) over the C columns
 *   grad_s[b][c] (+)= scale * (T / B) * (p_s - p_t)
 * s: student logits, row stride s_ld elements (the reference passes the strided slice logits[:, :known],
 * template.py:263); t: teacher logits, row stride t_ld; grad row stride g_ld; dtype applies to s, t and grad.
 * loss_out: one float.  accumulate != 0 adds into grad_s instead of overwriting.  grad_s may be NULL
 * (forward only).  workspace: wacil_kd_workspace_bytes(B) bytes (per-row partials + a counter), zero on first use
 * (the kernel re-zeroes it).  One pass over HBM: reads s and t once, writes grad once. */
size_t wacil_kd_workspace_bytes(int B);
int wacil_kd_fwd_bwd(const void* s, int64_t s_ld, const void* t, int64_t t_ld, int B, int C, float T, float scale,
                     float* loss_out, void* grad_s, int64_t g_ld, int accumulate, int dtype, void* workspace,
                     void* stream);

/* nn.CrossEntropyLoss(label_smoothing) + lambda_kd * SoftTarget (template.py:259-266) in one pass over the
 * logits, forward and backward:
 *   losses_out[0] = CE (mean over B), [1] = lambda_kd * KD, [2] = their sum, [3] = number of rows whose argmax
 *   equals the target (top-1 count, template.py:267)
 *   dlogits[b][c] = d(losses_out[2]) / d logits[b][c]  for c < C_seen (fp32, row stride d_ld)
 * logits: fp32 [B][C_seen] row stride l_ld; targets: int64 [B]; t_logits: fp32 [B][C_known] row stride t_ld or
 * NULL (then C_known is ignored: task 0, template.py:264-265).  KD acts on columns [0, C_known). */
int wacil_ce_kd_fwd_bwd(const float* logits, int64_t l_ld, const int64_t* targets, const float* t_logits,
                        int64_t t_ld, int B, int C_seen, int C_known, float T, float lambda_kd, float smooth,
                        float* losses_out, float* dlogits, int64_t d_ld, void* workspace, void* stream);

/* The same with a weight on the CrossEntropy term: loss = ce_weight * CE + lambda_kd * KD.  README.md:172-176 states
 * the WA objective as (1 - lambda) * CE + lambda * KD with lambda = n_old / (n_old + n_new) ("--dynamic_lambda_kd",
 * template.py:48, a flag the reference declares but never reads); losses_out[0] = ce_weight * CE. */
int wacil_ce_kd_fwd_bwd_w(const float* logits, int64_t l_ld, const int64_t* targets, const float* t_logits,
                          int64_t t_ld, int B, int C_seen, int C_known, float T, float ce_weight, float lambda_kd,
                          float smooth, float* losses_out, float* dlogits, int64_t d_ld, void* workspace, void* stream);

/* Validation meter of eval() (template.py:169-188): one pass over a batch of logits, accumulated on the device into
 * meter (float64[8], zeroed by the caller before the first batch):
 *   meter[0] += B   meter[1] += 1   meter[2] += mean CrossEntropy of the batch (MetricLogger.update(loss=..), utils.py:77-86)
 *   meter[3] += #rows whose target is the top-1 logit   meter[4] += #rows whose target is among the top `topk`
 * (timm.utils.accuracy(logits, target, topk=(1, min(5, C))), template.py:179-180: acc = 100 * count / samples).
 * logits fp32 [B][C] row stride l_ld, targets int64 [B].  workspace: wacil_eval_workspace_bytes(B), zero on first use.
 * The cross-rank reduction (MetricLogger.synchronize_between_processes, utils.py:36-43) is one all-reduce of meter. */
size_t wacil_eval_workspace_bytes(int B);
int wacil_eval_rows(const float* logits, int64_t l_ld, const int64_t* targets, int B, int C, int topk, double* meter,
                    void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Classifier head as tensor-core tiles  (CilClassifier.forward, template.py:99-101: the heads' nn.Linear outputs
 * concatenated = ONE product over the contiguous [C][D] arena; backward = what autograd derives for it)
 *   forward : logits[b][c] = sum_d feats[b][d] W[c][d] + bias[c]
 *   backward: dfeats[b][d] = sum_c dlogits[b][c] W[c][d];  dW[c][d] = sum_b dlogits[b][c] feats[b][d];
 *             db[c] = sum_b dlogits[b][c]      (each output optional: pass NULL to skip)
 * fp32 tensors, row-major with explicit row pitches in floats (*_ld; multiples of 4, bases 16-byte aligned: TMA).
 * Arithmetic: tcgen05.mma kind::tf32 (operands read as tf32, fp32 accumulation in TMEM) — relative error 1e-3 of
 * |feats|.|W|; tests/test_gpu_ops.py states the tolerance.  Any B, D, C (tiles are zero-filled at the edges); sized
 * for the large heads of BASELINE configs 3-5 (D = 512 / 2048, C = 100 ... 8192).  The D = 64 CIFAR head inside
 * wacil_net_forward stays a CUDA-core kernel fused with the 8x8 pooling. */
int wacil_fc_forward(const float* feats, int64_t f_ld, const float* W, int64_t w_ld, const float* bias, float* logits,
                     int64_t l_ld, int B, int D, int C, void* stream);
int wacil_fc_backward(const float* dlogits, int64_t g_ld, const float* feats, int64_t f_ld, const float* W,
                      int64_t w_ld, float* dfeats, int64_t df_ld, float* dW, int64_t dw_ld, float* db, int B, int D,
                      int C, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Weight aligning  (CilModel.weight_align, template.py:156-166)
 *   norms = ||W[r,:]||_2 ; gamma = mean(norms[:C_total-n_new]) / mean(norms[C_total-n_new:])
 *   W[C_total-n_new:, :] *= gamma     (in place; bias untouched)
 * W: [C_total][D] row-major contiguous (the classifier arena; heads are row ranges).  gamma_out: one float on the
 * device (the reference prints it, template.py:165; the host decides whether to read it back).
 * workspace: wacil_weight_align_workspace_bytes(C_total) bytes whose first 64 bytes are zero (grid barrier counters,
 * left zero). */
size_t wacil_weight_align_workspace_bytes(int C_total);
int wacil_weight_align(void* W, int C_total, int D, int n_new, float* gamma_out, int dtype, void* workspace,
                       void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Optimiser  (torch.optim.SGD(momentum, weight_decay), template.py:246-247,271; DDP mean, template.py:243-244)
 *   g = grad * grad_scale + wd * p ; v = mom * v + g ; p -= lr * v         over n fp32 elements
 * lr is read from device memory (lr_dev[0]) so that a captured graph follows the cosine schedule. */
int wacil_sgd_step(float* p, const float* g, float* v, int64_t n, const float* lr_dev, float momentum,
                   float weight_decay, float grad_scale, void* stream);

/* DDP gradient mean + SGD fused over NVLink peer memory (template.py:243-244,270-271): every rank reads all ranks'
 * gradient arenas (grad_ptrs[r] = rank r's arena mapped into this process: CUDA IPC / symmetric memory) in rank order —
 * bit-identical sums on all ranks — and updates its own p, v:  g = mean_r grad_r + wd*p; v = mom*v + g; p -= lr*v.
 * pad_ptrs[r]: rank r's peer-writable flag pad (>= 128 bytes, zero initially); epoch_ws: 8 bytes of LOCAL device memory,
 * zero initially.  Host arrays of `world` pointers (world <= 16); n must be a multiple of 4.  All ranks must call it the
 * same number of times; the kernel returns only when every peer has finished reading this rank's gradients, so the
 * caller may overwrite them right after (stream order).  Capturable: one graph holds the whole step. */
int wacil_allreduce_sgd(float* p, float* v, const void* const* grad_ptrs, void* const* pad_ptrs, int world, int rank,
                        int64_t n, const float* lr_dev, float momentum, float weight_decay, void* epoch_ws, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Herding  (continuum.rehearsal.herd_closest_to_barycenter, call site template.py:300-302)
 * features: fp32 [n][D] row-major; class_offsets: int32 [n_classes + 1], samples of
 * class k are rows class_offsets[k] .. class_offsets[k+1]-1 (the host groups rows by ascending class id);
 * quota = exemplars per class.  out_idx: int32 [n_classes][quota] row-local indices in herding-rank order
 * (entries beyond out_count[k] are -1); out_count: int32 [n_classes].
 * Bit-exact with oracle/herding.c (same float32 evaluation order).  Limits: D <= 2048, class size <= 65535.
 * workspace: wacil_herd_workspace_bytes(n, D) bytes. */
size_t wacil_herd_workspace_bytes(int64_t n, int D);
int wacil_herd_barycenter(const float* features, int64_t n, const int32_t* class_offsets, int n_classes, int D,
                          int quota, int32_t* out_idx, int32_t* out_count, void* workspace, void* stream);

/* continuum.rehearsal.herd_closest_to_cluster (--herding_method cluster, template.py:34,214): per class the `quota`
 * samples whose feature is closest (L2) to the class mean, nearest first.  Same argument layout as
 * wacil_herd_barycenter; bit-exact with oracle/herding.c:herd_cluster.  workspace: n * 4 bytes. */
int wacil_herd_cluster(const float* features, int64_t n, const int32_t* class_offsets, int n_classes, int D, int quota,
                       int32_t* out_idx, int32_t* out_count, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Device data path  (the CIFAR transform of utils.py:210-251 without the 10-worker host loader, template.py:236-239)
 * store: uint8 [n][H][W][3] raw samples resident in HBM (what dataset.get_raw_samples() / memory.get() hold,
 * template.py:231,301); idx: int64 [B] sample of each batch row (NULL: rows 0..B-1); aug: int32 [B] packed
 * dy | dx << 8 | flip << 16 with 0 <= dy, dx <= 2*pad (NULL: centre crop, no flip = the eval transform).
 *   out[b][c][y][x] = (padded(store[idx[b]])[y + dy][(flip ? W-1-x : x) + dx][c] / 255 - mean[c]) / std[c]
 * i.e. RandomCrop(W, padding=pad) with zero fill, RandomHorizontalFlip, ToTensor, Normalize in torchvision's order
 * and rounding (IEEE fp32 division): bit-exact with the CPU restatement.  out: fp32 NCHW [B][3][H][W] — the tensor
 * wacil_net_forward consumes.  mean3/std3 are HOST arrays of 3 floats. */
int wacil_augment_u8(const uint8_t* store, const int64_t* idx, const int32_t* aug, int B, int H, int W, int pad,
                     const float* mean3, const float* std3, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * The ResNet engine  (CifarResNet + CilClassifier forward/backward: resnet.py:107-116, template.py:99-123,270)
 *
 * A wacil_net_t is a launch plan bound to caller-owned buffers for one (depth, batch) pair; it holds the TMA
 * tensor maps of every layer.  Several plans may share the same parameter arenas (e.g. the 128-image plan and
 * the plan of the last, smaller batch of an epoch).
 *
 * Arena layouts (fp32 elements), L = conv layers in forward order, layer l has cin_l' (3 padded to 16 for the
 * first layer) and cout_l channels:
 *   params / grads / momentum : for l in L: W_l [9][cin_l'][cout_l]   (tap-major "RSCK"; the reference's
 *                                 [cout][cin][3][3] tensor is the permuted view W_l.view(3,3,cin',cout).permute(3,2,0,1))
 *                               for l in L: gamma_l [cout_l], beta_l [cout_l]
 *                               fc_weight [max_classes][64], fc_bias [max_classes]
 *   bn_running                : for l in L: running_mean_l [cout_l], running_var_l [cout_l]
 * wacil_net_layout() reports every offset.
 * --------------------------------------------------------------------------------------------------- */
typedef struct wacil_net wacil_net_t;

typedef struct {
  int depth;           /* 20 / 32 / 44 / 56 / 110 */
  int n_layers;        /* 6n+1 conv layers */
  int feat_dim;        /* 64 */
  int64_t param_elems; /* length of params / grads / momentum arenas */
  int64_t bn_elems;    /* length of the bn_running arena */
  int64_t fc_weight_off, fc_bias_off;
} wacil_net_layout_t;

typedef struct {
  int cin, cin_pad, cout, stride, H_in, H_out;
  int64_t w_off, gamma_off, beta_off; /* into params */
  int64_t rmean_off, rvar_off;        /* into bn_running */
} wacil_layer_layout_t;

int wacil_net_layout(int depth, int max_classes, wacil_net_layout_t* out);
int wacil_net_layer_layout(int depth, int max_classes, int layer, wacil_layer_layout_t* out);

#define WACIL_NET_TRAIN 1 /* forward keeps activations, batch statistics; backward available */
#define WACIL_NET_INFER 2 /* BatchNorm folded into bf16 weights + bias (teacher / feature extraction / eval) */

size_t wacil_net_workspace_bytes(int depth, int batch, int max_classes, int mode);

/* Binds a plan.  params/grads: fp32 arenas (grads may be NULL in INFER mode); bn_running: fp32 arena;
 * workspace: wacil_net_workspace_bytes() bytes, 1024-byte aligned.  Any batch >= 1 (ragged last tiles are
 * zero-filled by TMA and masked on store). */
int wacil_net_create(wacil_net_t** out, int depth, int batch, int max_classes, int mode, float* params,
                     float* grads, float* bn_running, void* workspace, size_t workspace_bytes);
void wacil_net_destroy(wacil_net_t* net);

/* INFER mode: (re)fold BatchNorm running statistics into the bf16 weight copies (call after the parameters or
 * running statistics change: CilModel.copy().freeze(), template.py:290; model.eval(), template.py:173). */
int wacil_net_fold(wacil_net_t* net, void* stream);

/* Forward.  x: fp32 NCHW [batch][3][32][32] (what the reference's loader yields, template.py:256);
 * feats: fp32 [batch][64] (CilModel.extract_vector, template.py:117-118); logits: fp32 [batch][C_seen] row stride
 * C_seen (CilModel.forward, template.py:120-123), may be NULL with C_seen = 0 for features only.
 * TRAIN mode uses batch statistics and updates the running ones (momentum 0.1, unbiased variance). */
int wacil_net_forward(wacil_net_t* net, const float* x, int C_seen, float* feats, float* logits, void* stream);

/* Backward of the last TRAIN forward: dlogits fp32 [batch][C_seen] -> grads arena (overwritten, not
 * accumulated: the arena is zeroed inside).  feats/logits of that forward must still be intact. */
int wacil_net_backward(wacil_net_t* net, const float* dlogits, int C_seen, void* stream);

/* Optional fused optimiser.  Once set, wacil_net_backward also applies the DDP gradient mean (template.py:243-244,270)
 * and the SGD update (template.py:246-247,271) of wacil_allreduce_sgd / wacil_sgd_step itself, as the last launch of the
 * backward (one launch, inside the same CUDA graph; an earlier instalment for the weights that are complete early was
 * measured slower, profiles/r02_early_instalment_ab.log).  momentum_buf: fp32 arena like params
 * (NULL detaches the optimiser); lr_dev: device float; world / rank / grad_ptrs / pad_ptrs / epoch_ws as in
 * wacil_allreduce_sgd (world = 1: no exchange, the pointers may be NULL).  All ranks must call backward the same
 * number of times. */
int wacil_net_set_optimizer(wacil_net_t* net, float* momentum_buf, const float* lr_dev, float momentum,
                            float weight_decay, int world, int rank, const void* const* grad_ptrs,
                            void* const* pad_ptrs, void* epoch_ws);

/* Optional second stream for wacil_net_backward: weight-gradient kernels are issued on it and ordered against the
 * data-gradient chain with events (captured as parallel graph branches).  NULL restores single-stream order. */
int wacil_net_set_side_stream(wacil_net_t* net, void* side_stream);

/* Introspection for the parity tests: device pointer + size of a tensor the TRAIN plan keeps in its workspace
 * after a forward: WACIL_T_INPUT packed input bf16 [B][32][32][16] (layer ignored); WACIL_T_CONV_OUT raw conv
 * output y_l bf16 [B][H][W][C]; WACIL_T_ACT post-BN/ReLU activation a_l bf16; WACIL_T_BN_SAVED fp32 [2C] batch mean
 * and 1/sqrt(var+eps) of layer l. */
#define WACIL_T_INPUT 0
#define WACIL_T_CONV_OUT 1
#define WACIL_T_ACT 2
#define WACIL_T_BN_SAVED 3
int wacil_net_tensor(const wacil_net_t* net, int what, int layer, void** ptr, size_t* bytes);

/* Number of kernels the last forward / backward call launched (for bench.py's gpu_launches). */
int wacil_net_last_launches(const wacil_net_t* net);

/* The convolution layers of a plan are a list of launch elements: one layer each in the default per-layer engine
 * (conv_gemm_kernel*, csrc/tc_conv.cuh), several layers per persistent cooperative launch with WACIL_ENGINE=chain
 * (csrc/tc_chain.cuh, kept under the parity tests; slower, profiles/r02_chain_experiments.md).  Introspection +
 * profiling: number of elements of the forward (backward = 0) or of the data-gradient pass (backward = 1); layout of
 * one of them; and a stand-alone run of ONE element on the plan's own buffers (after a forward, and for backward
 * elements after a backward, so that its inputs hold real data) — tools/layer_timeline.py and the ncu drivers use it.
 * dbg: NULL, or clock64 stamps: per-layer engine [grid][16] int64 (slots: 1 setup, 2 last load issued, 3 weights landed,
 * 4 first operands landed, 5 last MMA issued, 7 phase 1 done, 8 all partial sums seen, 13 constants ready, 9 phase 2
 * done, 10 exit); chain engine [grid][n_layers][48]. */
int wacil_net_chain_count(const wacil_net_t* net, int backward);
int wacil_net_chain_info(const wacil_net_t* net, int backward, int index, int* first_layer, int* n_layers, int* grid,
                         int* epilogue_warps, size_t* smem_bytes);
int wacil_net_run_chain(wacil_net_t* net, int backward, int index, long long* dbg, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Single-operator entry points kept for the parity tests (same kernels the engine launches).
 * --------------------------------------------------------------------------------------------------- */
/* 3x3 pad-1 convolution, NHWC bf16, weights [9][cout][cin] bf16, out bf16 [N][H/stride][W/stride][cout];
 * stats (may be NULL): fp32 [2*cout] per-channel sum and sum of squares of the stored outputs, accumulated. */
int wacil_conv3x3_fprop(const void* x, const void* w, void* y, float* stats, int N, int H, int W, int cin,
                        int cout, int stride, void* stream);
/* The fused forward kernel of one layer (resnet.py:43-53): y = conv3x3(x, w) (bf16), batch statistics of y,
 * a = ReLU(BatchNorm(y) + shortcut) (bf16), running statistics updated (momentum 0.1, unbiased variance), in ONE
 * cooperative launch (the CTAs exchange their per-channel partial sums between the statistics and the normalisation;
 * accumulators stay in TMEM).
 * res_mode: 0 none, 1 res has the shape of a, 2 option-A shortcut (res is [N][2Ho][2Wo][cout/2], resnet.py:15-17).
 * mean_rstd: fp32 [2*cout] batch mean and 1/sqrt(var+eps) out.  workspace: wacil_conv3x3_bn_relu_workspace_bytes(cout)
 * bytes, ALL zero before the first call and afterwards used by calls of this function with the same problem shape
 * only (a launch counter and two buffers of published partial-sum rows; every launch clears the buffer the next one
 * writes, so nothing is cleared per call). */
size_t wacil_conv3x3_bn_relu_workspace_bytes(int cout);
int wacil_conv3x3_bn_relu(const void* x, const void* w, const float* gamma, const float* beta, const void* res,
                          int res_mode, void* y, void* a, float* mean_rstd, float* running_mean, float* running_var,
                          int N, int H, int W, int cin, int cout, int stride, void* workspace, void* stream);
/* dW[9][cin][cout] fp32 = sum over pixels of x (shifted) * dy  (the weight gradient autograd forms at template.py:270);
 * x: [N][H][W][cin], dy: [N][H/s][W/s][cout], bf16 NHWC.  Overwrites dw.  workspace:
 * wacil_conv3x3_wgrad_workspace_bytes() bytes, 16-byte aligned (per-CTA partial sums, added in CTA order: the result is
 * bit-reproducible — the reference runs with cudnn.deterministic = True, template.py:57). */
size_t wacil_conv3x3_wgrad_workspace_bytes(int N, int H, int cin, int cout, int stride);
int wacil_conv3x3_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int cin, int cout,
                        int stride, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WACIL_H_ */
