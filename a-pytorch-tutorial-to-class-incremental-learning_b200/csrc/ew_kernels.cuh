// ew_kernels.cuh — the memory-bound kernels around the tcgen05 convolutions:
//   pack_input        fp32 NCHW image batch -> bf16 NHWC, 3 channels zero-padded to 16      (template.py:256)
//   cast_weights      fp32 master weights [9][cin][cout] -> bf16 fprop ([9][cout][cin]) and dgrad
//                     ([9 flipped][cin][cout]) operand copies, all layers in one launch
//   fold_bn           eval-mode BatchNorm folded into bf16 weights + fp32 bias (teacher / feature extraction)
//   (BatchNorm forward / backward are fused into the conv kernels, tc_conv.cuh; the last layer's backward together with
//    the classifier's data gradient is head_kernels.cuh)
//   pool_fc_fwd / fc_bwd           8x8 average pool + classifier heads                      (resnet.py:114-115,
//                                                                                          template.py:99-101)
//   sgd_step          momentum SGD with weight decay on the flat arena                      (template.py:246-247,271)
//   weight_align      fused row-norm + two-group mean + scale                               (template.py:156-166)
//   herd_barycenter   iCaRL herding, one CTA per class, float32 order of oracle/herding.c   (template.py:300-302)
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace wacil {

struct LayerDesc {
  int cin, cin_pad, cout, stride, H_in, H_out;
  long long w_off, gamma_off, beta_off;  // params arena (fp32 elements)
  long long rmean_off, rvar_off;         // bn_running arena
  long long wf_off, wd_off;              // bf16 weight buffer (elements)
  long long bias_off;                    // folded bias buffer (fp32 elements)
};

__device__ __forceinline__ void bf16x8_to_float(const uint4& q, float (&v)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __low2float(p[i]);
    v[2 * i + 1] = __high2float(p[i]);
  }
}
__device__ __forceinline__ uint4 float_to_bf16x8(const float (&v)[8]) {
  uint4 q;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return q;
}

// ------------------------------------------------------------------------------------------------ pack
__global__ void pack_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int HW) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;  // pixel index
  if (i >= static_cast<long long>(N) * HW) return;
  const long long n = i / HW, p = i - n * HW;
  const float* xp = x + n * 3 * HW + p;
  float v[8] = {xp[0], xp[HW], xp[2 * HW], 0.f, 0.f, 0.f, 0.f, 0.f};
  uint4* op = reinterpret_cast<uint4*>(out + i * 16);
  op[0] = float_to_bf16x8(v);
  op[1] = make_uint4(0u, 0u, 0u, 0u);
}

// ------------------------------------------------------------------------------------------------ weights
// Operand copies of the fp32 master weights [tap = 3r+s][ci][co]:
//   forward [tap][n = co][k = ci];   dgrad [tap' = 8 - tap][n = ci][k = co]  (the flipped taps of the transposed conv)
__global__ void cast_weights_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ wbuf,
                                    const LayerDesc* __restrict__ layers) {
  const LayerDesc L = layers[blockIdx.y];
  const float* w = params + L.w_off;
  const int tot = 9 * L.cout * L.cin_pad;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 2 * tot; e += gridDim.x * blockDim.x) {
    if (e < tot) {
      const int k = e % L.cin_pad, n = (e / L.cin_pad) % L.cout, t = e / (L.cin_pad * L.cout);
      wbuf[L.wf_off + e] = __float2bfloat16_rn(w[(t * L.cin_pad + k) * L.cout + n]);
    } else {
      const int f = e - tot;
      const int k = f % L.cout, n = (f / L.cout) % L.cin_pad, t = f / (L.cout * L.cin_pad);
      wbuf[L.wd_off + f] = __float2bfloat16_rn(w[((8 - t) * L.cin_pad + n) * L.cout + k]);
    }
  }
}

__global__ void fold_bn_kernel(const float* __restrict__ params, const float* __restrict__ bn_running,
                               __nv_bfloat16* __restrict__ wbuf, float* __restrict__ bias,
                               const LayerDesc* __restrict__ layers, float eps) {
  const LayerDesc L = layers[blockIdx.y];
  const float* w = params + L.w_off;
  const float* g = params + L.gamma_off;
  const float* b = params + L.beta_off;
  const float* rm = bn_running + L.rmean_off;
  const float* rv = bn_running + L.rvar_off;
  const int tot = 9 * L.cout * L.cin_pad;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += gridDim.x * blockDim.x) {
    const int k = e % L.cin_pad, n = (e / L.cin_pad) % L.cout, t = e / (L.cin_pad * L.cout);
    const float sc = g[n] * rsqrtf(rv[n] + eps);
    wbuf[L.wf_off + e] = __float2bfloat16_rn(w[(t * L.cin_pad + k) * L.cout + n] * sc);
    if (e < L.cout) {
      const float s2 = g[e] * rsqrtf(rv[e] + eps);
      bias[L.bias_off + e] = b[e] - rm[e] * s2;
    }
  }
}

// ------------------------------------------------------------------------------------------------ pool + FC
// One CTA (64 threads) per image: feats = mean over the 8x8 map; logits = feats . W^T + b over C_seen rows.
__global__ void __launch_bounds__(64) pool_fc_fwd_kernel(const __nv_bfloat16* __restrict__ act, int HW, int D,
                                                         const float* __restrict__ W, const float* __restrict__ b,
                                                         int C_seen, float* __restrict__ feats,
                                                         float* __restrict__ logits) {
  __shared__ float s_f[64];
  const int n = blockIdx.x, c = threadIdx.x;
  const __nv_bfloat16* ap = act + static_cast<long long>(n) * HW * D + c;
  float s = 0.f;
  for (int i = 0; i < HW; ++i) s += __bfloat162float(ap[static_cast<long long>(i) * D]);
  s *= 1.f / static_cast<float>(HW);
  s_f[c] = s;
  feats[static_cast<long long>(n) * D + c] = s;
  __syncthreads();
  if (logits == nullptr) return;
  for (int j = c; j < C_seen; j += blockDim.x) {
    const float4* wp = reinterpret_cast<const float4*>(W + static_cast<long long>(j) * D);
    float acc = 0.f;
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
      const float4 q = __ldg(wp + k);
      acc += q.x * s_f[4 * k] + q.y * s_f[4 * k + 1] + q.z * s_f[4 * k + 2] + q.w * s_f[4 * k + 3];
    }
    logits[static_cast<long long>(n) * C_seen + j] = acc + __ldg(b + j);
  }
}

// One block per class row j: dW[j][:] = sum_n dlogits[n][j] feats[n][:] and db[j] = sum_n dlogits[n][j].  256 threads =
// 4 groups of 64 (one per feature column) that split the batch 4 ways; fixed-order combine, no atomics.  (The data
// gradient dfeats = dlogits . W is computed by head_bwd_kernel, head_kernels.cuh.)
__global__ void __launch_bounds__(256) fc_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ feats,
                                                     int B, int C_seen, int D, float* __restrict__ dW,
                                                     float* __restrict__ db) {
  __shared__ float s_acc[4][64], s_accb[4];
  const int c = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int j = blockIdx.x;
  float acc = 0.f, accb = 0.f;
#pragma unroll 8
  for (int n = grp; n < B; n += 4) {
    const float g = __ldg(dlogits + static_cast<long long>(n) * C_seen + j);
    acc = fmaf(g, __ldg(feats + static_cast<long long>(n) * D + c), acc);
    accb += g;
  }
  s_acc[grp][c] = acc;
  if (c == 0) s_accb[grp] = accb;
  __syncthreads();
  if (grp == 0) {
    dW[static_cast<long long>(j) * D + c] = (s_acc[0][c] + s_acc[1][c]) + (s_acc[2][c] + s_acc[3][c]);
    if (c == 0) db[j] = (s_accb[0] + s_accb[1]) + (s_accb[2] + s_accb[3]);
  }
}

// ------------------------------------------------------------------------------------------------ SGD
__global__ void __launch_bounds__(256) sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ v, long long n,
                                                       const float* __restrict__ lr_dev, float mom, float wd,
                                                       float gscale) {
  const float lr = __ldg(lr_dev);
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 vv = reinterpret_cast<float4*>(v)[i];
    vv.x = mom * vv.x + (gg.x * gscale + wd * pp.x); pp.x -= lr * vv.x;
    vv.y = mom * vv.y + (gg.y * gscale + wd * pp.y); pp.y -= lr * vv.y;
    vv.z = mom * vv.z + (gg.z * gscale + wd * pp.z); pp.z -= lr * vv.z;
    vv.w = mom * vv.w + (gg.w * gscale + wd * pp.w); pp.w -= lr * vv.w;
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0) {
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      const float vv = mom * v[i] + (g[i] * gscale + wd * p[i]);
      v[i] = vv;
      p[i] -= lr * vv;
    }
  }
}

// ------------------------------------------------------------------------------------------------ fused all-reduce + SGD
// DDP gradient mean (template.py:243-244,270) + SGD (template.py:271) in ONE kernel over NVLink peer memory:
// every rank reads all ranks' gradient arenas (peer-mapped, "symmetric" allocations) in rank order — the sum is bit-
// identical on every rank, so replicas stay in lock step without a broadcast — and applies the update to its own
// parameters.  Cross-GPU ordering uses epoch-stamped flags in a peer-writable pad:
//   pad[r]          <- E+1 written by rank r when ITS gradients are complete ("ready")
//   pad[kMaxRanks+r]<- E+1 written by rank r when it has finished READING ("done": its peers may overwrite their grads)
// The epoch E lives in device memory and is advanced by the last block, so a captured graph can be replayed.
constexpr int kMaxRanks = 16;
struct AllreduceSgdParams {
  const float* grads[kMaxRanks];  // peer-mapped gradient arenas, index = rank
  unsigned* pads[kMaxRanks];      // peer-mapped flag pads, index = rank (2 * kMaxRanks words each)
  int world, rank;
  float* p;
  float* v;
  long long n;            // elements of segment 0 (starting at seg_off[0])
  long long seg_off[2];   // up to two segments of the arenas per launch (a partially frozen model updates the ranges
  long long seg_n1;       // on either side of the frozen one)
  const float* lr_dev;
  float mom, wd;
  unsigned* epoch;       // [0] epoch, [1] blocks-done counter (local device memory, zero initially)
};

__device__ __forceinline__ void st_release_sys(unsigned* ptr, unsigned val) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ptr), "r"(val) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* ptr) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void wait_flags(const unsigned* pad, int world, unsigned target, const char* what) {
  const long long t0 = clock64();
  for (int r = 0; r < world; ++r) {
    while (static_cast<int>(ld_acquire_sys(pad + r) - target) < 0) {
      if (clock64() - t0 > 20000000000LL) {  // ~10 s: a missing peer must trap, not hang the GPU
        printf("wacil: allreduce_sgd timeout waiting for rank %d (%s)\n", r, what);
        __trap();
      }
    }
  }
}

__global__ void __launch_bounds__(256) allreduce_sgd_kernel(const AllreduceSgdParams q) {
  __shared__ unsigned s_epoch, s_last;
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned*>(q.epoch);
  __syncthreads();
  const unsigned target = s_epoch + 1u;
  // my gradients were completed by earlier kernels of this stream: announce them to every rank (block 0 only)
  if (blockIdx.x == 0 && threadIdx.x < q.world) {
    __threadfence_system();
    st_release_sys(q.pads[threadIdx.x] + q.rank, target);
  }
  if (threadIdx.x == 0) wait_flags(q.pads[q.rank], q.world, target, "ready");
  __syncthreads();
  const float lr = __ldg(q.lr_dev);
  const float gscale = 1.f / static_cast<float>(q.world);
  for (int seg = 0; seg < 2; ++seg) {
    const long long n4 = (seg == 0 ? q.n : q.seg_n1) / 4, o4 = q.seg_off[seg] / 4;
    for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < n4;
         j += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long i = o4 + j;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int r = 0; r < q.world; ++r) {
        const float4 t = __ldcv(reinterpret_cast<const float4*>(q.grads[r]) + i);  // peer read, never from a stale L1 line
        g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
      }
      float4 pp = reinterpret_cast<float4*>(q.p)[i];
      float4 vv = reinterpret_cast<float4*>(q.v)[i];
      vv.x = q.mom * vv.x + (g.x * gscale + q.wd * pp.x); pp.x -= lr * vv.x;
      vv.y = q.mom * vv.y + (g.y * gscale + q.wd * pp.y); pp.y -= lr * vv.y;
      vv.z = q.mom * vv.z + (g.z * gscale + q.wd * pp.z); pp.z -= lr * vv.z;
      vv.w = q.mom * vv.w + (g.w * gscale + q.wd * pp.w); pp.w -= lr * vv.w;
      reinterpret_cast<float4*>(q.p)[i] = pp;
      reinterpret_cast<float4*>(q.v)[i] = vv;
    }
  }
  // n is a multiple of 4 by construction of the arena layout (wacil_net_layout)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(q.epoch + 1, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  // last block: every block of this rank has finished reading -> tell the peers, wait until they have finished
  // reading MY gradients (the next backward overwrites them), then open the next epoch
  if (threadIdx.x < q.world) st_release_sys(q.pads[threadIdx.x] + kMaxRanks + q.rank, target);
  if (threadIdx.x == 0) {
    wait_flags(q.pads[q.rank] + kMaxRanks, q.world, target, "done");
    q.epoch[1] = 0u;
    __threadfence();
    q.epoch[0] = target;
  }
}

// ------------------------------------------------------------------------------------------------ weight align
// CilModel.weight_align (template.py:156-166): gamma = mean ||w_old row|| / mean ||w_new row||, new rows *= gamma.
// One persistent cooperative launch, one 1024-thread CTA per SM:
//   phase 1  a warp streams a row with 16-byte loads, eight in flight per lane (131 KB in flight per SM); the OLD rows
//            are dealt round-robin over all warps with evict-first loads (read once), the NEW rows are dealt round-robin
//            over the CTAs and parked in shared memory while their norm is taken (8.4 MB of [8192,2048] fp32 / 148 SMs =
//            57 KB per CTA), so that phase 3 neither re-reads them from L2 nor from HBM;
//   phase 2  per-CTA partial sums of the two groups of norms (fixed order inside the CTA) -> [grid][2] table -> grid
//            barrier (release add + acquire poll) -> every CTA adds the table in CTA order: gamma is bit-identical
//            everywhere and reproducible (the first version had every CTA re-read all C_total norms: 38 MB of L2 reads
//            at [8192,2048] behind an atomicAdd-polling barrier — 33 % of the HBM roofline);
//   phase 3  scale the parked rows, streaming stores.
// Rows that do not fit the shared-memory budget (n_new / grid > cap) are re-read in phase 3.
// bar[0], bar[1]: grid barrier counters, zero at rest.
constexpr int kWaThreads = 1024;

__device__ __forceinline__ void wa_grid_barrier(unsigned* bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned seen;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
      if (seen < gridDim.x && clock64() - t0 > 4000000000LL) {
        printf("wacil: weight_align grid barrier timeout (block %d)\n", blockIdx.x);
        __trap();
      }
    } while (seen < gridDim.x);
    if (atomicAdd(bar + 1, 1u) == gridDim.x - 1) {  // last one out: every CTA has left the polling loop
      bar[1] = 0u;
      __threadfence();
      bar[0] = 0u;
    }
  }
  __syncthreads();
}

template <typename T> __device__ __forceinline__ float wa_ld(const T* p);
template <> __device__ __forceinline__ float wa_ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float wa_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void wa_st(T* p, float v);
template <> __device__ __forceinline__ void wa_st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void wa_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T> __device__ __forceinline__ float wa_sumsq16(const uint4& q) {
  constexpr int VW = 16 / sizeof(T);
  const T* e = reinterpret_cast<const T*>(&q);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VW; ++i) {
    const float f = wa_ld<T>(e + i);
    ss += f * f;
  }
  return ss;
}

// Sum of squares of one row by one warp.  PARK: also copy the row to `park` (shared memory).  VEC: 16-byte path.
template <typename T, bool PARK, bool STREAM>
__device__ __forceinline__ float wa_row_sumsq(const T* __restrict__ row, int D, int lane, bool vec, T* park) {
  constexpr int VW = 16 / sizeof(T);
  float ss = 0.f;
  if (vec) {
    const int nv = D / VW;  // 16-byte vectors in the row
    const uint4* src = reinterpret_cast<const uint4*>(row);
    for (int v0 = lane; v0 < nv; v0 += 32 * 8) {
      uint4 q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int v = v0 + 32 * j;
        if (v < nv) q[j] = __ldcs(src + v);  // every row is read exactly once
        else q[j] = make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int v = v0 + 32 * j;
        ss += wa_sumsq16<T>(q[j]);
        if (PARK && v < nv) reinterpret_cast<uint4*>(park)[v] = q[j];
      }
    }
  } else {
    for (int k = lane; k < D; k += 32) {
      const float f = wa_ld<T>(row + k);
      ss += f * f;
      if (PARK) park[k] = row[k];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  return ss;
}

template <typename T>
__global__ void __launch_bounds__(kWaThreads, 1) weight_align_kernel(T* __restrict__ W, int C_total, int D, int n_new,
                                                                    float* __restrict__ partial, unsigned* bar,
                                                                    float* __restrict__ gamma_out, int park_rows) {
  extern __shared__ __align__(16) uint8_t wa_smem[];
  __shared__ float s_old[32], s_new[32], s_tot[2][32];
  T* parked = reinterpret_cast<T*>(wa_smem);  // [park_rows][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  constexpr int VW = 16 / sizeof(T);
  const bool vec = (D % VW) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0;
  const int C_old = C_total - n_new;
  const size_t rowlen = static_cast<size_t>(D);
  // programmatic dependent launch: everything above ran under the previous kernel's tail; its writes (W may have been
  // produced by it) are visible from here on
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // ---- phase 1a: the NEW rows of this CTA (row C_old + b + k*G), parked in shared memory while they fit
  float sn = 0.f;
  {
    int k = warp;
    for (int i = static_cast<int>(blockIdx.x) + warp * G; i < n_new; i += 32 * G, k += 32) {
      const T* row = W + (static_cast<size_t>(C_old) + i) * rowlen;
      const float ss = k < park_rows ? wa_row_sumsq<T, true, false>(row, D, lane, vec, parked + static_cast<size_t>(k) * rowlen)
                                     : wa_row_sumsq<T, false, false>(row, D, lane, vec, nullptr);
      sn += sqrtf(ss);  // every lane holds the row total after the butterfly: lane 0's copy is used below
    }
  }
  // ---- phase 1b: the OLD rows, round-robin over all warps of the grid, read once (evict-first)
  float so = 0.f;
  for (int r = static_cast<int>(blockIdx.x) * 32 + warp; r < C_old; r += G * 32)
    so += sqrtf(wa_row_sumsq<T, false, true>(W + static_cast<size_t>(r) * rowlen, D, lane, vec, nullptr));
  if (lane == 0) { s_old[warp] = so; s_new[warp] = sn; }
  __syncthreads();
  if (threadIdx.x == 0) {  // fixed order inside the CTA
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 32; ++i) { a += s_old[i]; b += s_new[i]; }
    partial[2 * blockIdx.x] = a;
    partial[2 * blockIdx.x + 1] = b;
  }
  wa_grid_barrier(bar);
  // ---- phase 2: every CTA adds the [G][2] table in CTA order (warp 0: lanes take CTAs lane, lane + 32, ... then a
  // fixed shuffle tree) -> the same gamma everywhere
  if (warp == 0) {
    float a = 0.f, b = 0.f;
    for (int c0 = lane; c0 < G; c0 += 32 * 4) {
      float2 t[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + 32 * j;
        t[j] = c < G ? __ldcg(reinterpret_cast<const float2*>(partial) + c) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { a += t[j].x; b += t[j].y; }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) { s_tot[0][0] = a; s_tot[1][0] = b; }
  }
  __syncthreads();
  const float gamma = (s_tot[0][0] / static_cast<float>(C_old)) / (s_tot[1][0] / static_cast<float>(n_new));
  if (blockIdx.x == 0 && threadIdx.x == 0) *gamma_out = gamma;
  // ---- phase 3: scale this CTA's new rows (from shared memory where parked), streaming stores
  {
    int k = warp;
    for (int i = static_cast<int>(blockIdx.x) + warp * G; i < n_new; i += 32 * G, k += 32) {
      T* row = W + (static_cast<size_t>(C_old) + i) * rowlen;
      const T* src = k < park_rows ? parked + static_cast<size_t>(k) * rowlen : row;
      if (vec) {
        const int nv = D / VW;
        for (int v = lane; v < nv; v += 32) {
          uint4 q = reinterpret_cast<const uint4*>(src)[v];
          T* e = reinterpret_cast<T*>(&q);
#pragma unroll
          for (int j = 0; j < VW; ++j) wa_st<T>(e + j, gamma * wa_ld<T>(e + j));
          __stcs(reinterpret_cast<uint4*>(row) + v, q);
        }
      } else {
        for (int j = lane; j < D; j += 32) wa_st<T>(row + j, gamma * wa_ld<T>(src + j));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ herding
// One CTA per class.  Every float32 operation is an explicit round-to-nearest intrinsic in the order of
// oracle/herding.c, so the selected indices are bit-identical.  dnT: workspace [D][n] (feature-major) so that
// thread i streams sample i's coordinates with coalesced loads.
__global__ void __launch_bounds__(256) herd_barycenter_kernel(const float* __restrict__ features,
                                                              const int* __restrict__ class_offsets, int D, int quota,
                                                              long long n_total, float* __restrict__ dnT,
                                                              int* __restrict__ rank, int* __restrict__ out_idx,
                                                              int* __restrict__ out_count) {
  extern __shared__ float s_mem[];  // mu[D], w[D]
  float* s_mu = s_mem;
  float* s_w = s_mem + D;
  __shared__ float s_best[8];
  __shared__ int s_bidx[8];
  __shared__ int s_pick, s_it, s_eff;
  const int k_class = blockIdx.x;
  const int lo = class_offsets[k_class], hi = class_offsets[k_class + 1];
  const int m = hi - lo;
  const int take = quota < m ? quota : m;
  int* oidx = out_idx + static_cast<long long>(k_class) * quota;
  for (int i = threadIdx.x; i < quota; i += blockDim.x) oidx[i] = -1;
  if (m <= 0) {
    if (threadIdx.x == 0) out_count[k_class] = 0;
    return;
  }
  // 1. normalise each sample (sequential over the feature dimension)
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const float* f = features + static_cast<long long>(lo + i) * D;
    float ss = 0.f;
    for (int k = 0; k < D; ++k) ss = __fadd_rn(ss, __fmul_rn(f[k], f[k]));
    const float nrm = __fadd_rn(__fsqrt_rn(ss), 1e-8f);
    for (int k = 0; k < D; ++k) dnT[static_cast<long long>(k) * n_total + lo + i] = __fdiv_rn(f[k], nrm);
    rank[lo + i] = 0;
  }
  __syncthreads();
  // 2. mean (sequential over samples)
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    const float* col = dnT + static_cast<long long>(k) * n_total + lo;
    float s = 0.f;
    for (int i = 0; i < m; ++i) s = __fadd_rn(s, col[i]);
    const float mu = __fdiv_rn(s, static_cast<float>(m));
    s_mu[k] = mu;
    s_w[k] = mu;
  }
  if (threadIdx.x == 0) { s_it = 0; s_eff = 0; }
  __syncthreads();
  // 3. greedy selection
  while (s_it != take && s_eff < 1000) {
    float best = 0.f;
    int bi = -1;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const float* col = dnT + lo + i;
      float s = 0.f;
      for (int k = 0; k < D; ++k) s = __fadd_rn(s, __fmul_rn(s_w[k], col[static_cast<long long>(k) * n_total]));
      if (bi < 0 || s > best) { best = s; bi = i; }
    }
    // first maximum wins: order by (value desc, index asc)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_best[threadIdx.x >> 5] = best; s_bidx[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float b = s_best[0];
      int bidx = s_bidx[0];
      for (int wv = 1; wv < static_cast<int>(blockDim.x >> 5); ++wv) {
        const float ob = s_best[wv];
        const int oi = s_bidx[wv];
        if (oi >= 0 && (bidx < 0 || ob > b || (ob == b && oi < bidx))) { b = ob; bidx = oi; }
      }
      s_pick = bidx;
      s_eff = s_eff + 1;
      if (rank[lo + bidx] == 0) {
        rank[lo + bidx] = 1 + s_it;
        oidx[s_it] = bidx;
        s_it = s_it + 1;
      }
    }
    __syncthreads();
    const int pick = s_pick;
    for (int k = threadIdx.x; k < D; k += blockDim.x)
      s_w[k] = __fsub_rn(__fadd_rn(s_w[k], s_mu[k]), dnT[static_cast<long long>(k) * n_total + lo + pick]);
    __syncthreads();
  }
  // 4. iteration cap hit before the quota was filled: unpicked samples follow in index order
  if (threadIdx.x == 0) {
    int got = s_it;
    for (int i = 0; i < m && got < take; ++i)
      if (rank[lo + i] == 0) oidx[got++] = i;
    out_count[k_class] = got;
  }
}

// ------------------------------------------------------------------------------------------------ cluster herding
// continuum's herd_closest_to_cluster: per class the samples closest (L2) to the class-mean feature, nearest first.
// One CTA per class; float32 operations in the order of oracle/herding.c (herd_cluster) so the indices are bit-exact.
// dist: workspace [n].
__global__ void __launch_bounds__(256) herd_cluster_kernel(const float* __restrict__ features,
                                                           const int* __restrict__ class_offsets, int D, int quota,
                                                           float* __restrict__ dist, int* __restrict__ out_idx,
                                                           int* __restrict__ out_count) {
  extern __shared__ float s_mem[];  // mean[D]
  const int k_class = blockIdx.x;
  const int lo = class_offsets[k_class], hi = class_offsets[k_class + 1];
  const int m = hi - lo;
  const int take = quota < m ? quota : m;
  int* oidx = out_idx + static_cast<long long>(k_class) * quota;
  for (int i = threadIdx.x; i < quota; i += blockDim.x) oidx[i] = -1;
  if (threadIdx.x == 0) out_count[k_class] = take > 0 ? take : 0;
  if (m <= 0) return;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {  // mean, sequential over the samples
    float s = 0.f;
    for (int i = 0; i < m; ++i) s = __fadd_rn(s, features[static_cast<long long>(lo + i) * D + k]);
    s_mem[k] = __fdiv_rn(s, static_cast<float>(m));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += blockDim.x) {  // distance, sequential over the feature dimension
    const float* f = features + static_cast<long long>(lo + i) * D;
    float ss = 0.f;
    for (int k = 0; k < D; ++k) {
      const float d = __fsub_rn(s_mem[k], f[k]);
      ss = __fadd_rn(ss, __fmul_rn(d, d));
    }
    dist[lo + i] = __fsqrt_rn(ss);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += blockDim.x) {  // stable rank = position in the ascending order
    const float di = dist[lo + i];
    int r = 0;
    for (int j = 0; j < m; ++j) {
      const float dj = dist[lo + j];
      r += (dj < di || (dj == di && j < i)) ? 1 : 0;
    }
    if (r < take) oidx[r] = i;
  }
}

// ------------------------------------------------------------------------------------------------ device data path
// The CIFAR train/eval transform of utils.py:210-251 without the host loader: gather sample idx[b] from the uint8 HWC
// store, RandomCrop(32, padding=4) (zero fill) at offset (dy, dx), RandomHorizontalFlip, ToTensor (x / 255),
// Normalize ((x - mean) / std) -> fp32 NCHW.  aug[b] = dy | dx << 8 | flip << 16 (NULL: centre crop dy = dx = pad, no
// flip = the eval transform).  IEEE division in the order torchvision applies it: bit-exact with the CPU restatement.
__global__ void augment_u8_kernel(const uint8_t* __restrict__ store, const int64_t* __restrict__ idx,
                                  const int32_t* __restrict__ aug, int B, int H, int W, int pad, float m0, float m1,
                                  float m2, float s0, float s1, float s2, float* __restrict__ out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;  // output pixel
  const long long hw = static_cast<long long>(H) * W;
  if (i >= B * hw) return;
  const int b = static_cast<int>(i / hw);
  const int p = static_cast<int>(i - b * hw);
  const int y = p / W, x = p - y * W;
  const int a = aug ? aug[b] : (pad | (pad << 8));
  const int dy = a & 0xff, dx = (a >> 8) & 0xff, flip = (a >> 16) & 1;
  const int sy = y + dy - pad;
  const int sx = (flip ? (W - 1 - x) : x) + dx - pad;
  float r = 0.f, g = 0.f, bl = 0.f;
  if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
    const long long n = idx ? idx[b] : b;
    const uint8_t* px = store + (n * hw + static_cast<long long>(sy) * W + sx) * 3;
    r = static_cast<float>(px[0]);
    g = static_cast<float>(px[1]);
    bl = static_cast<float>(px[2]);
  }
  float* o = out + static_cast<long long>(b) * 3 * hw + p;
  o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn(r, 255.f), m0), s0);
  o[hw] = __fdiv_rn(__fsub_rn(__fdiv_rn(g, 255.f), m1), s1);
  o[2 * hw] = __fdiv_rn(__fsub_rn(__fdiv_rn(bl, 255.f), m2), s2);
}

}  // namespace wacil
