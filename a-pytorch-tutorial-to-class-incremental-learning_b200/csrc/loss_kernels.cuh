// loss_kernels.cuh — fused softmax-T + KL-divergence distillation loss (forward + backward in one pass) and the
// fused CrossEntropy + KD step loss.  Reference: utils.py:127-132 (SoftTarget), template.py:259-266.
//
// Both are HBM-bound row kernels: a row (one sample's logits) is owned by a warp (C <= 1024) or by a 256-thread
// CTA (C <= 8192), lives in registers between the reduction and the gradient write, and is read from HBM exactly
// once.  Algorithmic traffic of the KD kernel: 3*B*C*sizeof(dtype) (+4B per row).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace wacil {

constexpr float kLog2e = 1.4426950408889634f;

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 q = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
  static __device__ __forceinline__ float ld1(const float* p) { return __ldcs(p); }
  static __device__ __forceinline__ void st1(float* p, float v) { *p = v; }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 q = __ldcs(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&q.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&q.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 q;
    q.x = *reinterpret_cast<uint32_t*>(&a);
    q.y = *reinterpret_cast<uint32_t*>(&b);
    __stcs(reinterpret_cast<uint2*>(p), q);
  }
  static __device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// Reductions over the TPR threads that own a row (TPR = 32: one warp; TPR = 256: the whole CTA).
template <int TPR> struct RowReduce {
  // red: shared scratch of 64 floats (TPR > 32 only)
  static __device__ __forceinline__ void max2(float& a, float& b, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if (TPR > 32) {
      const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
      __syncthreads();
      if (l == 0) { red[w] = a; red[32 + w] = b; }
      __syncthreads();
      a = red[0]; b = red[32];
#pragma unroll
      for (int i = 1; i < TPR / 32; ++i) { a = fmaxf(a, red[i]); b = fmaxf(b, red[32 + i]); }
    }
  }
  static __device__ __forceinline__ void sum2(float& a, float& b, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (TPR > 32) {
      const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
      __syncthreads();
      if (l == 0) { red[w] = a; red[32 + w] = b; }
      __syncthreads();
      a = red[0]; b = red[32];
#pragma unroll
      for (int i = 1; i < TPR / 32; ++i) { a += red[i]; b += red[32 + i]; }
    }
  }
};

// Deterministic finish: the last CTA to arrive adds the per-row partials in row order.
// part layout: float part[NPARTS][B].  Returns true on exactly one thread of the grid (thread 0 of the last CTA)
// with acc[] holding the ordered sums; the counter is re-armed for the next launch.
template <int NPARTS>
__device__ __forceinline__ bool finish_rows(float* part, unsigned* counter, int B, float (&acc)[NPARTS], bool wrote_part) {
  __shared__ unsigned s_last;
  if (wrote_part) __threadfence();  // only the threads that published a row partial; the gradient stores need no fence
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last || threadIdx.x >= 32) return false;
  __threadfence();
#pragma unroll
  for (int k = 0; k < NPARTS; ++k) acc[k] = 0.f;
  // fixed order: lane l sums rows l, l+32, ... then a fixed shuffle tree
  for (int r = threadIdx.x; r < B; r += 32)
#pragma unroll
    for (int k = 0; k < NPARTS; ++k) acc[k] += __ldcg(part + k * B + r);
#pragma unroll
  for (int k = 0; k < NPARTS; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  if (threadIdx.x == 0) *counter = 0u;
  return threadIdx.x == 0;
}

// ------------------------------------------------------------------------------------------------ KD
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Row arithmetic in the log2 domain: u = x * (log2(e) / T), so e^(x/T - max) = 2^(u - max_u) is one FADD + one MUFU;
// padding lanes hold a large negative FINITE sentinel (2^sentinel = 0 and no inf - inf), which removes every select.
constexpr float kRowPad = -1.0e30f;

template <typename T, int TPR, int kRowCache, bool VEC>
__global__ void __launch_bounds__(TPR > 256 ? TPR : 256, TPR > 256 ? 1 : 3) kd_rows_kernel(const T* __restrict__ s, int64_t s_ld, const T* __restrict__ t,
                                                      int64_t t_ld, int B, int C, float invT, float gscale,
                                                      float lscale, float* __restrict__ loss_out, T* __restrict__ g,
                                                      int64_t g_ld, int accumulate, float* part, unsigned* counter) {
  __shared__ float red[64];
  constexpr int RPB = TPR >= 256 ? 1 : 256 / TPR;  // rows per block
  const int tid = threadIdx.x % TPR;
  const int row = blockIdx.x * RPB + threadIdx.x / TPR;
  const bool live = row < B;
  const T* sp = s + static_cast<int64_t>(live ? row : 0) * s_ld;
  const T* tp = t + static_cast<int64_t>(live ? row : 0) * t_ld;
  const float k2 = invT * kLog2e;

  float a[kRowCache], b[kRowCache];
  float ma = kRowPad, mb = kRowPad;
  if (VEC) {
#pragma unroll
    for (int i = 0; i < kRowCache / 4; ++i) {
      const int col = (i * TPR + tid) * 4;
      float va[4], vb[4];
      const bool in = live && col < C;
      if (in) { Vec4<T>::load(sp + col, va); Vec4<T>::load(tp + col, vb); }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[i * 4 + j] = in ? va[j] * k2 : kRowPad;
        b[i * 4 + j] = in ? vb[j] * k2 : kRowPad;
        ma = fmaxf(ma, a[i * 4 + j]); mb = fmaxf(mb, b[i * 4 + j]);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < kRowCache; ++i) {
      const int col = i * TPR + tid;
      const bool in = live && col < C;
      a[i] = in ? Vec4<T>::ld1(sp + col) * k2 : kRowPad;
      b[i] = in ? Vec4<T>::ld1(tp + col) * k2 : kRowPad;
      ma = fmaxf(ma, a[i]); mb = fmaxf(mb, b[i]);
    }
  }
  RowReduce<TPR>::max2(ma, mb, red);
  // One exponential per element: the cache is overwritten with 2^(u - max); the KL sum is accumulated unnormalised,
  //   sum_i p_t,i (log p_t,i - log p_s,i) = ln2 * (sum_i e_b,i (db_i - da_i)) / S_b - ln S_b + ln S_a
  // (da, db = shifted log2-domain logits); entries whose e_b underflows contribute 0, the xlogy(0, .) = 0 of F.kl_div.
  float sa = 0.f, sb = 0.f, acc = 0.f;
#pragma unroll
  for (int i = 0; i < kRowCache; ++i) {
    const float da = a[i] - ma, db = b[i] - mb;
    const float ea = ex2_approx(da), eb = ex2_approx(db);
    sa += ea;
    sb += eb;
    acc = fmaf(eb, db - da, acc);
    a[i] = ea;
    b[i] = eb;
  }
  RowReduce<TPR>::sum2(sa, sb, red);
  float dummy = 0.f;
  RowReduce<TPR>::sum2(acc, dummy, red);
  const float inv_sa = 1.f / sa, inv_sb = 1.f / sb;
  const float kl = 0.6931471805599453f * acc * inv_sb - logf(sb) + logf(sa);
  const float ca = inv_sa * gscale, cb = inv_sb * gscale;
#pragma unroll
  for (int i = 0; i < kRowCache; ++i) a[i] = fmaf(a[i], ca, -(b[i] * cb));  // gradient, reusing the cache
  if (g != nullptr && live) {
    T* gp = g + static_cast<int64_t>(row) * g_ld;
    if (VEC) {
#pragma unroll
      for (int i = 0; i < kRowCache / 4; ++i) {
        const int col = (i * TPR + tid) * 4;
        if (col < C) {
          float v[4] = {a[i * 4], a[i * 4 + 1], a[i * 4 + 2], a[i * 4 + 3]};
          if (accumulate) {
            float o[4];
            Vec4<T>::load(gp + col, o);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] += o[j];
          }
          Vec4<T>::store(gp + col, v);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < kRowCache; ++i) {
        const int col = i * TPR + tid;
        if (col < C) Vec4<T>::st1(gp + col, accumulate ? Vec4<T>::ld1(gp + col) + a[i] : a[i]);
      }
    }
  }
  if (live && tid == 0) part[row] = kl;
  float tot[1];
  if (finish_rows<1>(part, counter, B, tot, live && tid == 0)) *loss_out = tot[0] * lscale;
}

// Variant with natural-domain arithmetic and the accurate exp2f (kept for A/B measurement, WACIL_KD_MATH=0).
template <typename T, int TPR, int kRowCache, bool VEC>
__global__ void __launch_bounds__(256, 3) kd_rows_kernel_v2(const T* __restrict__ s, int64_t s_ld, const T* __restrict__ t,
                                                      int64_t t_ld, int B, int C, float invT, float gscale,
                                                      float lscale, float* __restrict__ loss_out, T* __restrict__ g,
                                                      int64_t g_ld, int accumulate, float* part, unsigned* counter) {
  __shared__ float red[64];
  constexpr int RPB = 256 / TPR;  // rows per block
  const int tid = threadIdx.x % TPR;
  const int row = blockIdx.x * RPB + threadIdx.x / TPR;
  const bool live = row < B;
  const T* sp = s + static_cast<int64_t>(live ? row : 0) * s_ld;
  const T* tp = t + static_cast<int64_t>(live ? row : 0) * t_ld;

  float a[kRowCache], b[kRowCache];
  float ma = -INFINITY, mb = -INFINITY;
  if (VEC) {
#pragma unroll
    for (int i = 0; i < kRowCache / 4; ++i) {
      const int col = (i * TPR + tid) * 4;
      float va[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, vb[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (live && col < C) { Vec4<T>::load(sp + col, va); Vec4<T>::load(tp + col, vb); }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[i * 4 + j] = va[j] * invT; b[i * 4 + j] = vb[j] * invT;
        ma = fmaxf(ma, a[i * 4 + j]); mb = fmaxf(mb, b[i * 4 + j]);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < kRowCache; ++i) {
      const int col = i * TPR + tid;
      const bool in = live && col < C;
      a[i] = in ? Vec4<T>::ld1(sp + col) * invT : -INFINITY;
      b[i] = in ? Vec4<T>::ld1(tp + col) * invT : -INFINITY;
      ma = fmaxf(ma, a[i]); mb = fmaxf(mb, b[i]);
    }
  }
  RowReduce<TPR>::max2(ma, mb, red);
  if (!live) { ma = 0.f; mb = 0.f; }
  // One exponential per element: the cache is overwritten with e^(x - max); the KL sum is accumulated unnormalised,
  //   sum_i p_t,i (log p_t,i - log p_s,i) = (sum_i e_b,i (b_i - a_i)) / S_b - log S_b + log S_a     (shifted logits)
  // with 0 * log 0 := 0 (xlogy semantics of F.kl_div) for underflowed / padded entries.
  float sa = 0.f, sb = 0.f, acc = 0.f;
#pragma unroll
  for (int i = 0; i < kRowCache; ++i) {
    const float as = a[i] - ma, bs = b[i] - mb;  // <= 0; -inf for padding
    const float ea = exp2f(as * kLog2e), eb = exp2f(bs * kLog2e);
    sa += ea;
    sb += eb;
    acc += (eb > 0.f) ? eb * (bs - as) : 0.f;
    a[i] = ea;
    b[i] = eb;
  }
  RowReduce<TPR>::sum2(sa, sb, red);
  float dummy = 0.f;
  RowReduce<TPR>::sum2(acc, dummy, red);
  const float inv_sa = 1.f / sa, inv_sb = 1.f / sb;
  const float kl = acc * inv_sb - logf(sb) + logf(sa);
#pragma unroll
  for (int i = 0; i < kRowCache; ++i) a[i] = (a[i] * inv_sa - b[i] * inv_sb) * gscale;  // gradient, reusing the cache
  if (g != nullptr && live) {
    T* gp = g + static_cast<int64_t>(row) * g_ld;
    if (VEC) {
#pragma unroll
      for (int i = 0; i < kRowCache / 4; ++i) {
        const int col = (i * TPR + tid) * 4;
        if (col < C) {
          float v[4] = {a[i * 4], a[i * 4 + 1], a[i * 4 + 2], a[i * 4 + 3]};
          if (accumulate) {
            float o[4];
            Vec4<T>::load(gp + col, o);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] += o[j];
          }
          Vec4<T>::store(gp + col, v);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < kRowCache; ++i) {
        const int col = i * TPR + tid;
        if (col < C) Vec4<T>::st1(gp + col, accumulate ? Vec4<T>::ld1(gp + col) + a[i] : a[i]);
      }
    }
  }
  if (live && tid == 0) part[row] = kl;
  float tot[1];
  if (finish_rows<1>(part, counter, B, tot, live && tid == 0)) *loss_out = tot[0] * lscale;
}


// ------------------------------------------------------------------------------------------------ CE + KD
// One pass over the student logits of a row: softmax over C_seen (CE) and softmax-T over the first C_known
// columns (KD) share the loaded values.  fp32 in, fp32 out.
template <int TPR, int kRowCache>
__global__ void __launch_bounds__(256)
    ce_kd_rows_kernel(const float* __restrict__ z, int64_t z_ld, const int64_t* __restrict__ tgt,
                      const float* __restrict__ tz, int64_t t_ld, int B, int Cs, int Ck, float invT, float kd_gscale,
                      float kd_lscale, float smooth, float* __restrict__ losses, float* __restrict__ dz, int64_t d_ld,
                      float* part, unsigned* counter) {
  __shared__ float red[64];
  constexpr int RPB = 256 / TPR;
  const int tid = threadIdx.x % TPR;
  const int row = blockIdx.x * RPB + threadIdx.x / TPR;
  const bool live = row < B;
  const float* zp = z + static_cast<int64_t>(live ? row : 0) * z_ld;
  const float* tp = tz ? tz + static_cast<int64_t>(live ? row : 0) * t_ld : nullptr;
  const int y = live ? static_cast<int>(tgt[row]) : 0;
  const float invB = 1.f / static_cast<float>(B);

  float v[kRowCache], b[kRowCache];
  float m1 = -INFINITY, mk = -INFINITY, mt = -INFINITY;
  int amax = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < kRowCache; ++i) {
    const int col = i * TPR + tid;
    v[i] = (live && col < Cs) ? __ldg(zp + col) : -INFINITY;
    b[i] = (tp && live && col < Ck) ? __ldg(tp + col) * invT : -INFINITY;
    if (v[i] > m1) { m1 = v[i]; amax = col; }
    if (col < Ck) mk = fmaxf(mk, v[i]);
    mt = fmaxf(mt, b[i]);
  }
  // row max with arg (first index wins ties, as torch.topk/argmax report the lowest index)
  {
    float mm = m1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mm, o);
      const int oa = __shfl_xor_sync(0xffffffffu, amax, o);
      if (om > mm || (om == mm && oa < amax)) { mm = om; amax = oa; }
    }
    if (TPR > 32) {
      __shared__ float smax[8];
      __shared__ int sarg[8];
      const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
      if (l == 0) { smax[w] = mm; sarg[w] = amax; }
      __syncthreads();
      mm = smax[0]; amax = sarg[0];
      for (int i = 1; i < TPR / 32; ++i)
        if (smax[i] > mm || (smax[i] == mm && sarg[i] < amax)) { mm = smax[i]; amax = sarg[i]; }
      __syncthreads();
    }
    m1 = mm;
  }
  RowReduce<TPR>::max2(mk, mt, red);
  if (!live) { m1 = 0.f; }
  const bool has_kd = tp != nullptr && Ck > 0;
  if (!has_kd || !live) { mk = 0.f; mt = 0.f; }
  mk *= invT;  // max of z/T over the known columns
  float s1 = 0.f, sk = 0.f, st = 0.f, sumz = 0.f, zy = 0.f;
#pragma unroll
  for (int i = 0; i < kRowCache; ++i) {
    const int col = i * TPR + tid;
    s1 += exp2f((v[i] - m1) * kLog2e);
    if (has_kd) {
      sk += (col < Ck) ? exp2f((v[i] * invT - mk) * kLog2e) : 0.f;
      st += exp2f((b[i] - mt) * kLog2e);
    }
    if (live && col < Cs) { sumz += v[i]; if (col == y) zy = v[i]; }
  }
  RowReduce<TPR>::sum2(s1, sk, red);
  RowReduce<TPR>::sum2(st, sumz, red);
  float dummy = 0.f;
  RowReduce<TPR>::sum2(zy, dummy, red);
  const float ls1 = logf(s1);
  // CE with label smoothing eps:  (1-eps) * (-log p_y) + eps * mean_c(-log p_c)
  const float nll = -(zy - m1 - ls1);
  const float mean_nlp = -(sumz / static_cast<float>(Cs) - m1 - ls1);
  const float ce_row = (1.f - smooth) * nll + smooth * mean_nlp;
  const float inv_s1 = 1.f / s1;
  float kl = 0.f;
  const float lsk = has_kd ? logf(sk) : 0.f, lst = has_kd ? logf(st) : 0.f;
  const float inv_sk = has_kd ? 1.f / sk : 0.f, inv_st = has_kd ? 1.f / st : 0.f;
  float* dp = dz ? dz + static_cast<int64_t>(live ? row : 0) * d_ld : nullptr;
#pragma unroll
  for (int i = 0; i < kRowCache; ++i) {
    const int col = i * TPR + tid;
    if (live && col < Cs) {
      const float p = exp2f((v[i] - m1) * kLog2e) * inv_s1;
      float gr = (p - ((col == y) ? (1.f - smooth) : 0.f) - smooth / static_cast<float>(Cs)) * invB;
      if (has_kd && col < Ck) {
        const float as = v[i] * invT - mk;
        const float ps = exp2f(as * kLog2e) * inv_sk;
        const float pt = exp2f((b[i] - mt) * kLog2e) * inv_st;
        kl += (pt > 0.f) ? pt * ((b[i] - mt - lst) - (as - lsk)) : 0.f;
        gr += (ps - pt) * kd_gscale;
      }
      if (dp) dp[col] = gr;
    }
  }
  RowReduce<TPR>::sum2(kl, dummy, red);
  if (live && tid == 0) {
    part[row] = ce_row;
    part[B + row] = kl;
    part[2 * B + row] = (amax == y) ? 1.f : 0.f;
  }
  float acc[3];
  if (finish_rows<3>(part, counter, B, acc, live && tid == 0)) {
    const float ce = acc[0] * invB, kd = acc[1] * kd_lscale;
    losses[0] = ce;
    losses[1] = kd;
    losses[2] = ce + kd;
    losses[3] = acc[2];
  }
}

}  // namespace wacil
