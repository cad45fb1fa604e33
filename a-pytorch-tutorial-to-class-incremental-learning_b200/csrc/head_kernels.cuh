// head_kernels.cuh — the start of the backward as ONE cooperative launch (template.py:270 through the pooled features):
//   dfeats[n][c] = sum_j dlogits[n][j] W[j][c]            (the classifier's data gradient, resnet.py:114-115 backward)
//   g = dfeats / (H*W) under the ReLU mask of the last activation; d-beta = sum g, d-gamma = sum g * x-hat  (BatchNorm
//   backward of the last layer, per-CTA partial sums exchanged as self-validating words like the fused conv kernels)
//   gy = gamma * rstd * (g - d-beta / M - x-hat * d-gamma / M)   (+ gz = g for the last block's shortcut)
// It replaces the dfeats half of fc_bwd_kernel + bn_bwd_reduce_kernel + bn_bwd_apply_kernel, three dependent launches
// (24.7 us of the step's critical path, profiles/r02_head_knockout.log); the classifier's weight gradient (the other
// half of fc_bwd_kernel) moves to the weight-gradient branch.
// A CTA owns a contiguous range of pixels (whole images when the batch fits the grid), so the data gradient of its
// images lives in shared memory and the second pass re-reads activations it has just read (L1 / L2 hits).
#pragma once
#include "ew_kernels.cuh"
#include "tc_conv.cuh"  // published_row_sum4

namespace wacil {

constexpr int kHeadThreads = 256;
constexpr int kHeadMaxImgs = 8;  // images a CTA's pixel range may touch (batch <= 8 * grid)

struct HeadBwdParams {
  const float* dlogits;   // [B][C_seen]
  const float* W;         // [C_seen][64] classifier rows
  int B, C_seen, HW;      // HW = pixels per image (64)
  const __nv_bfloat16* act;  // last activation [B*HW][64]
  const __nv_bfloat16* y;    // last raw conv output
  const float* saved;     // mean, rstd of the last BatchNorm
  const float* gamma;
  float* dgamma;
  float* dbeta;
  float* dfeats;          // [B][64] out (kept for callers that inspect it)
  __nv_bfloat16* gy;      // out
  __nv_bfloat16* gz;      // out or nullptr
  unsigned* table;        // two buffers of [grid][128] complemented words
  unsigned* counter;      // {launch counter, check-out counter}
  int pix_per_cta;
  long long num_pixels;
};

__global__ void __launch_bounds__(kHeadThreads) head_bwd_kernel(const HeadBwdParams p) {
  constexpr int C = 64, CV = 8;                 // channels, 8-channel vectors per pixel
  constexpr int LANES = kHeadThreads / CV;      // pixels per pass
  __shared__ float s_df[kHeadMaxImgs][C];
  __shared__ float s_acc[8][2 * C];
  __shared__ __align__(16) float s_scr[8 * 2 * C];  // [S = 8][V = 128]
  __shared__ float s_k[C], s_a[C], s_b[C], s_mean[C], s_rstd[C];
  const int tid = threadIdx.x;
  const long long pix0 = static_cast<long long>(blockIdx.x) * p.pix_per_cta;
  const long long pix1 = min(pix0 + p.pix_per_cta, p.num_pixels);
  const int img0 = static_cast<int>(pix0 / p.HW);
  const int nimg = pix1 > pix0 ? static_cast<int>((pix1 - 1) / p.HW) - img0 + 1 : 0;
  const float inv_hw = 1.f / static_cast<float>(p.HW);
  // programmatic dependent launch: this grid may have been started under the tail of the loss kernel, and lets the
  // first fused backward launch run its prologue (barriers, TMEM, weight TMA) under this one
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // ---- the classifier's data gradient of this CTA's images (4 groups of 64 threads take images round robin)
  for (int i = tid >> 6; i < nimg; i += 4) {
    const int c = tid & 63, n = img0 + i;
    const float* dl = p.dlogits + static_cast<long long>(n) * p.C_seen;
    float acc = 0.f;
#pragma unroll 4
    for (int j = 0; j < p.C_seen; ++j) acc = fmaf(__ldg(dl + j), __ldg(p.W + static_cast<long long>(j) * C + c), acc);
    s_df[i][c] = acc;
    p.dfeats[static_cast<long long>(n) * C + c] = acc;
  }
  if (tid < C) {
    s_mean[tid] = p.saved[tid];
    s_rstd[tid] = p.saved[C + tid];
  }
  __syncthreads();
  // ---- pass 1: per-channel sums of g and g * x-hat over this CTA's pixels
  const int myc = tid % CV, c0 = myc * 8, lane_px = tid / CV;
  float sg[8], sx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sg[i] = sx[i] = 0.f;
  for (long long pix = pix0 + lane_px; pix < pix1; pix += LANES) {
    const long long v = pix * CV + myc;
    const uint4 ra = __ldg(reinterpret_cast<const uint4*>(p.act) + v);
    const uint4 ry = __ldg(reinterpret_cast<const uint4*>(p.y) + v);
    float a[8], yv[8];
    bf16x8_to_float(ra, a);
    bf16x8_to_float(ry, yv);
    const float* df = s_df[static_cast<int>(pix / p.HW) - img0] + c0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float g = a[i] > 0.f ? df[i] * inv_hw : 0.f;
      sg[i] += g;
      sx[i] += g * (yv[i] - s_mean[c0 + i]) * s_rstd[c0 + i];
    }
  }
  // threads of a warp that own the same channel vector: lane ids congruent mod CV
  for (int o = 16; o >= CV; o >>= 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sg[i] += __shfl_xor_sync(0xffffffffu, sg[i], o);
      sx[i] += __shfl_xor_sync(0xffffffffu, sx[i], o);
    }
  }
  if ((tid & 31) < CV) {
    float* mine = s_acc[tid >> 5];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mine[c0 + i] = sg[i];
      mine[C + c0 + i] = sx[i];
    }
  }
  __syncthreads();
  // ---- publish this CTA's row (complemented words: zero = not written), clear its row of the other buffer
  unsigned launch_no;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(launch_no) : "l"(p.counter) : "memory");
  const size_t table_words = static_cast<size_t>(gridDim.x) * (2 * C);
  unsigned* const rows_cur = p.table + (launch_no & 1u) * table_words;
  if (tid < 2 * C) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += s_acc[wv][tid];
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(rows_cur + static_cast<size_t>(blockIdx.x) * (2 * C) + tid),
                 "r"(~__float_as_uint(t))
                 : "memory");
    p.table[((launch_no & 1u) ^ 1u) * table_words + static_cast<size_t>(blockIdx.x) * (2 * C) + tid] = 0u;
  }
  // ---- every CTA adds all rows in CTA order
  {
    constexpr int V = 2 * C, Q = V / 4, S = kHeadThreads / Q;  // 32 column groups, 8 row slices
    const int cg = tid % Q, sl = tid / Q;
    *reinterpret_cast<float4*>(s_scr + sl * V + 4 * cg) =
        published_row_sum4<10>(rows_cur, V, cg, sl, S, static_cast<int>(gridDim.x));
  }
  __syncthreads();
  unsigned ticket = 0;
  if (tid == 0) ticket = atomicAdd(p.counter + 1, 1u);
  if (tid < C) {
    float db = 0.f, dg = 0.f;
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) {
      db += s_scr[sl * 2 * C + tid];
      dg += s_scr[sl * 2 * C + C + tid];
    }
    const float inv_m = 1.f / static_cast<float>(p.num_pixels);
    s_k[tid] = p.gamma[tid] * s_rstd[tid];  // dL/dy = k * (g - a - x-hat * b)
    s_a[tid] = db * inv_m;
    s_b[tid] = dg * inv_m;
    if (blockIdx.x == 0) {
      p.dbeta[tid] = db;
      p.dgamma[tid] = dg;
    }
  }
  __syncthreads();
  // ---- pass 2: the gradient wrt the conv output (and the masked gradient for the last block's shortcut)
  for (long long pix = pix0 + lane_px; pix < pix1; pix += LANES) {
    const long long v = pix * CV + myc;
    const uint4 ra = __ldg(reinterpret_cast<const uint4*>(p.act) + v);
    const uint4 ry = __ldg(reinterpret_cast<const uint4*>(p.y) + v);
    float a[8], yv[8], g[8], o[8];
    bf16x8_to_float(ra, a);
    bf16x8_to_float(ry, yv);
    const float* df = s_df[static_cast<int>(pix / p.HW) - img0] + c0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      g[i] = a[i] > 0.f ? df[i] * inv_hw : 0.f;
      const float xhat = (yv[i] - s_mean[c0 + i]) * s_rstd[c0 + i];
      o[i] = s_k[c0 + i] * (g[i] - s_a[c0 + i] - xhat * s_b[c0 + i]);
    }
    reinterpret_cast<uint4*>(p.gy)[v] = float_to_bf16x8(o);
    if (p.gz) reinterpret_cast<uint4*>(p.gz)[v] = float_to_bf16x8(g);
  }
  if (tid == 0 && ticket == gridDim.x - 1) {  // last CTA out: every CTA has read every row
    p.counter[1] = 0u;
    __threadfence();
    p.counter[0] = launch_no + 1u;
  }
}

}  // namespace wacil
