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

#include "tc_prims.cuh"

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

// Deterministic finish: the last CTA to arrive adds the per-row partials in a fixed order (thread t sums rows
// t, t + blockDim, ...; then a fixed shuffle tree and a fixed cross-warp tree), every thread of that CTA taking part so
// that the tail costs one L2 round trip instead of B/32 dependent ones.
// part layout: float part[NPARTS][B].  Returns true on exactly one thread of the grid (thread 0 of the last CTA)
// with acc[] holding the ordered sums; the counter is re-armed for the next launch.  counter == nullptr: no finish.
template <int NPARTS>
__device__ __forceinline__ bool finish_rows(float* part, unsigned* counter, int B, float (&acc)[NPARTS], bool wrote_part) {
  __shared__ unsigned s_last;
  __shared__ float s_fin[NPARTS][32];
  if (counter == nullptr) return false;
  (void)wrote_part;
  // bar.sync orders every thread's partial stores before thread 0's release-arrive (cumulative at gpu scope); the last
  // arriver acquires, and the second bar.sync carries that to the CTA's loads: two L2 round trips in total, no fences.
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
    s_last = (old == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return false;
#pragma unroll
  for (int k = 0; k < NPARTS; ++k) acc[k] = 0.f;
  for (int r = threadIdx.x; r < B; r += blockDim.x)
#pragma unroll
    for (int k = 0; k < NPARTS; ++k) acc[k] += __ldcg(part + k * B + r);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NPARTS; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (lane == 0) s_fin[k][warp] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x >= 32) return false;
#pragma unroll
  for (int k = 0; k < NPARTS; ++k) {
    acc[k] = lane < nwarps ? s_fin[k][lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  }
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

// ------------------------------------------------------------------------------------------------ KD, streamed
// Persistent TMA-fed variant for wide rows (1024 < C <= 8192, 16-byte aligned rows): one 512-thread CTA per SM;
// a row pair (student, teacher) is brought into a shared-memory ring with two 1-D bulk copies
// (cp.async.bulk + mbarrier complete_tx) issued rows ahead of their use, copied once to registers (the stage is
// released right after that copy, so almost the whole ring is in flight at any time), reduced, and the gradient is
// written straight from registers with streaming 16-byte stores.  A CTA owns rows blockIdx + k*gridDim; its G row
// groups of TPR threads (G = 512 / TPR) take k = g, g + G, ... so that narrow rows keep every warp busy.
// HBM traffic is the algorithmic 3*B*C*sizeof(T); loads carry an evict-first L2 policy (nothing is re-read).
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                             uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

template <typename T> struct Vec16;  // one 16-byte vector = VE elements
template <> struct Vec16<float> {
  static constexpr int VE = 4;
  static __device__ __forceinline__ void lds(const float* p, float* v) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
  static __device__ __forceinline__ void ldg(const float* p, float* v) {
    const float4 q = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
  static __device__ __forceinline__ void stg(float* p, const float* v) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int VE = 8;
  static __device__ __forceinline__ void unpack(const uint4& q, float* v) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void lds(const __nv_bfloat16* p, float* v) {
    unpack(*reinterpret_cast<const uint4*>(p), v);
  }
  static __device__ __forceinline__ void ldg(const __nv_bfloat16* p, float* v) {
    unpack(__ldcs(reinterpret_cast<const uint4*>(p)), v);
  }
  static __device__ __forceinline__ void stg(__nv_bfloat16* p, const float* v) {
    uint4 q;
    uint32_t* w = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    __stcs(reinterpret_cast<uint4*>(p), q);
  }
};

template <int TPR> __device__ __forceinline__ void group_sync(int g) {
  if (TPR > 32) asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(TPR) : "memory");
}

template <typename T, int THREADS, int TPR, int EPT, int NS> struct KdStreamCfg {
  static constexpr int G = THREADS / TPR;
  static constexpr uint32_t kRowBytes = static_cast<uint32_t>(TPR) * EPT * sizeof(T);
  static constexpr size_t kDataBytes = static_cast<size_t>(NS) * 2 * kRowBytes;
  static constexpr size_t kSmemBytes = kDataBytes + 2 * NS * sizeof(uint64_t) + G * 5 * 32 * sizeof(float) + 128;
};

// THREADS per CTA, TPR threads per row group (G = THREADS / TPR rows are in flight in registers at once), EPT
// elements per thread and array, NS ring stages (one stage = one student + one teacher row).
template <typename T, int THREADS, int TPR, int EPT, int NS>
__global__ void __launch_bounds__(THREADS, 1)
    kd_stream_kernel(const T* __restrict__ s, int64_t s_ld, const T* __restrict__ t, int64_t t_ld, int B, int C,
                     float invT, float gscale, float lscale, float* __restrict__ loss_out, T* __restrict__ g,
                     int64_t g_ld, int accumulate, float* part, unsigned* counter) {
  using Cfg = KdStreamCfg<T, THREADS, TPR, EPT, NS>;
  using V = Vec16<T>;
  constexpr int G = Cfg::G, VE = V::VE, CH = EPT / VE, NW = TPR / 32;
  extern __shared__ uint8_t kd_smem_raw[];
  uint8_t* data = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(kd_smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* full = reinterpret_cast<uint64_t*>(data + Cfg::kDataBytes);  // [NS]
  uint64_t* empty = full + NS;                                            // [NS]
  float* red_all = reinterpret_cast<float*>(empty + NS);                  // [G][5][32]

  const int grp = threadIdx.x / TPR, tid = threadIdx.x % TPR, lane = threadIdx.x & 31, wig = tid >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], NW);
    }
    fence_mbar_init();
  }
  __syncthreads();

  // Programmatic dependent launch (when the launch carries the attribute): the barrier setup above ran under the tail
  // of the previous kernel of the stream; its results (possibly our inputs, and the shared workspace) are touched only
  // after this wait.  The next launch may start its own setup as soon as our CTAs retire.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const uint32_t row_bytes = static_cast<uint32_t>(C) * sizeof(T);
  const uint64_t policy = l2_evict_first_policy();
  float* red = red_all + grp * 5 * 32;
  const float k2 = invT * kLog2e;

  // CTA-local row sequence k = 0, 1, ...: row blockIdx + k * gridDim, ring stage k % NS, row group k % G
  auto issue = [&](int stage, int row) {
    mbar_expect_tx(&full[stage], 2 * row_bytes);
    bulk_load_1d(data + (stage * 2 + 0) * Cfg::kRowBytes, s + static_cast<int64_t>(row) * s_ld, row_bytes,
                 &full[stage], policy);
    bulk_load_1d(data + (stage * 2 + 1) * Cfg::kRowBytes, t + static_cast<int64_t>(row) * t_ld, row_bytes,
                 &full[stage], policy);
  };
  if (threadIdx.x == 0) {
#pragma unroll 1
    for (int k = 0; k < NS; ++k) {
      const int row = blockIdx.x + k * gridDim.x;
      if (row < B) issue(k, row);
    }
  }

  bool wrote = false;
#pragma unroll 1
  for (int k = grp;; k += G) {
    const int row = blockIdx.x + k * gridDim.x;
    if (row >= B) break;
    const int stage = k % NS;
    const uint32_t parity = (k / NS) & 1;
    mbar_wait(&full[stage], parity);
    const T* ss = reinterpret_cast<const T*>(data + (stage * 2 + 0) * Cfg::kRowBytes);
    const T* ts = reinterpret_cast<const T*>(data + (stage * 2 + 1) * Cfg::kRowBytes);
    float a[EPT], b[EPT];
    float ma = kRowPad, mb = kRowPad;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int col = (i * TPR + tid) * VE;
      if (col < C) {
        V::lds(ss + col, a + i * VE);
        V::lds(ts + col, b + i * VE);
#pragma unroll
        for (int j = 0; j < VE; ++j) {  // raw logits; the 1/T * log2(e) scale (> 0) is folded into the exponent FMA
          ma = fmaxf(ma, a[i * VE + j]);
          mb = fmaxf(mb, b[i * VE + j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < VE; ++j) a[i * VE + j] = b[i * VE + j] = kRowPad;
      }
    }
    // the row now lives in registers: hand the stage back and (the group's first thread) refill it NS rows ahead
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
    if (tid == 0) {
      const int nrow = blockIdx.x + (k + NS) * gridDim.x;
      if (nrow < B) {
        mbar_wait(&empty[stage], parity);
        issue(stage, nrow);
      }
    }
    // ---- row max (log2 domain)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o));
      mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
    }
    if (NW > 1) {
      if (lane == 0) { red[wig] = ma; red[32 + wig] = mb; }
      group_sync<TPR>(grp);
      ma = red[0]; mb = red[32];
#pragma unroll
      for (int i = 1; i < NW; ++i) { ma = fmaxf(ma, red[i]); mb = fmaxf(mb, red[32 + i]); }
    }
    // ---- one exponential per element; KL accumulated unnormalised (see kd_rows_kernel)
    float sa = 0.f, sb = 0.f, acc = 0.f;
    const float na = -ma * k2, nb = -mb * k2;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      const float da = fmaf(a[i], k2, na), db = fmaf(b[i], k2, nb);  // log2-domain shifted logits (<= 0; pad -> -huge)
      const float ea = ex2_approx(da), eb = ex2_approx(db);
      sa += ea;
      sb += eb;
      acc = fmaf(eb, db - da, acc);
      a[i] = ea;
      b[i] = eb;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sa += __shfl_xor_sync(0xffffffffu, sa, o);
      sb += __shfl_xor_sync(0xffffffffu, sb, o);
      acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (NW > 1) {  // scratch rows separate from the max rows: one barrier per reduction is enough (loop order)
      if (lane == 0) { red[64 + wig] = sa; red[96 + wig] = sb; red[128 + wig] = acc; }
      group_sync<TPR>(grp);
      sa = red[64]; sb = red[96]; acc = red[128];
#pragma unroll
      for (int i = 1; i < NW; ++i) { sa += red[64 + i]; sb += red[96 + i]; acc += red[128 + i]; }
    }
    const float inv_sa = 1.f / sa, inv_sb = 1.f / sb;
    const float kl = 0.6931471805599453f * acc * inv_sb - logf(sb) + logf(sa);
    const float ca = inv_sa * gscale, cb = inv_sb * gscale;
    if (g != nullptr) {
      T* gp = g + static_cast<int64_t>(row) * g_ld;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int col = (i * TPR + tid) * VE;
        if (col < C) {
          float v[VE];
#pragma unroll
          for (int j = 0; j < VE; ++j) v[j] = fmaf(a[i * VE + j], ca, -(b[i * VE + j] * cb));
          if (accumulate) {
            float o[VE];
            V::ldg(gp + col, o);
#pragma unroll
            for (int j = 0; j < VE; ++j) v[j] += o[j];
          }
          V::stg(gp + col, v);
        }
      }
    }
    if (tid == 0) { part[row] = kl; wrote = true; }
  }
  float tot[1];
  if (finish_rows<1>(part, counter, B, tot, wrote)) *loss_out = tot[0] * lscale;
}

// ------------------------------------------------------------------------------------------------ CE + KD
// One pass over the student logits of a row: softmax over C_seen (CE) and softmax-T over the first C_known
// columns (KD) share the loaded values.  fp32 in, fp32 out.
template <int TPR, int kRowCache>
__global__ void __launch_bounds__(256)
    ce_kd_rows_kernel(const float* __restrict__ z, int64_t z_ld, const int64_t* __restrict__ tgt,
                      const float* __restrict__ tz, int64_t t_ld, int B, int Cs, int Ck, float invT, float kd_gscale,
                      float kd_lscale, float smooth, float ce_w, float* __restrict__ losses, float* __restrict__ dz,
                      int64_t d_ld, float* part, unsigned* counter) {
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
      float gr = (p - ((col == y) ? (1.f - smooth) : 0.f) - smooth / static_cast<float>(Cs)) * (invB * ce_w);
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
    const float ce = acc[0] * invB * ce_w, kd = acc[1] * kd_lscale;
    losses[0] = ce;
    losses[1] = kd;
    losses[2] = ce + kd;
    losses[3] = acc[2];
  }
}

// ------------------------------------------------------------------------------------------------ eval meter
// One pass over the logits of a validation batch (template.py:169-188): per row the CrossEntropy term and the rank of
// the target among the logits (timm.utils.accuracy top-1 / top-k); the last CTA adds the batch to a float64 meter
//   meter[0] += B (samples)   meter[1] += 1 (batches)   meter[2] += mean CE of this batch (MetricLogger.update(loss=..))
//   meter[3] += rows with rank < 1   meter[4] += rows with rank < topk
// The rank counts logits strictly greater than the target's, ties broken towards the lower column (torch.topk order).
template <int TPR, int kRowCache>
__global__ void __launch_bounds__(256)
    eval_rows_kernel(const float* __restrict__ z, int64_t z_ld, const int64_t* __restrict__ tgt, int B, int C, int topk,
                     double* __restrict__ meter, float* part, unsigned* counter) {
  __shared__ float red[64];
  constexpr int RPB = 256 / TPR;
  const int tid = threadIdx.x % TPR;
  const int row = blockIdx.x * RPB + threadIdx.x / TPR;
  const bool live = row < B;
  const float* zp = z + static_cast<int64_t>(live ? row : 0) * z_ld;
  const int y = live ? static_cast<int>(tgt[row]) : 0;
  const float zy = (live && y >= 0 && y < C) ? __ldg(zp + y) : INFINITY;
  float v[kRowCache];
  float m1 = -INFINITY, dummy = -INFINITY;
  float ahead = 0.f;
#pragma unroll
  for (int i = 0; i < kRowCache; ++i) {
    const int col = i * TPR + tid;
    v[i] = (live && col < C) ? __ldg(zp + col) : -INFINITY;
    m1 = fmaxf(m1, v[i]);
    if (v[i] > zy || (v[i] == zy && col < y)) ahead += 1.f;
  }
  RowReduce<TPR>::max2(m1, dummy, red);
  if (!live) m1 = 0.f;
  float s1 = 0.f;
#pragma unroll
  for (int i = 0; i < kRowCache; ++i) s1 += exp2f((v[i] - m1) * kLog2e);
  RowReduce<TPR>::sum2(s1, ahead, red);
  if (live && tid == 0) {
    part[row] = -(zy - m1 - logf(s1));
    part[B + row] = ahead < 1.f ? 1.f : 0.f;
    part[2 * B + row] = ahead < static_cast<float>(topk) ? 1.f : 0.f;
  }
  float acc[3];
  if (finish_rows<3>(part, counter, B, acc, live && tid == 0)) {
    meter[0] += static_cast<double>(B);
    meter[1] += 1.0;
    meter[2] += static_cast<double>(acc[0]) / static_cast<double>(B);
    meter[3] += static_cast<double>(acc[1]);
    meter[4] += static_cast<double>(acc[2]);
  }
}

}  // namespace wacil
