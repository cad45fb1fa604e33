// tc_conv.cuh — tcgen05/TMEM implicit-GEMM kernels for the 3x3 (and 1x1) convolutions of the
// CIFAR ResNet (reference: resnet.py:29-34,71-72 conv layers; autograd of template.py:270).
//
// Data layout in HBM: activations NHWC bf16 (channels 16/32/64, the 3-channel input is zero-padded
// to 16), weights tap-major bf16 ([tap][n][k], K contiguous), accumulators fp32 in TMEM.
//
//   conv_gemm_kernel  : out[pix][n] = sum_tap sum_k A_tap[pix][k] * Wt[tap][n][k]
//                       fprop  : A = x,  k = cin,  n = cout
//                       dgrad  : A = dy, k = cout, n = cin, taps flipped (host builds that weight copy)
//                       M tile = 128 output pixels (full-width rows of one or two images); every tap is one
//                       TMA box load of the shifted window (out-of-bounds rows/cols zero-filled = padding 1)
//                       into a K-major swizzled tile; tcgen05.mma M=128, N=n, K=16 steps.
//   wgrad_gemm_kernel : dW[tap*cin + ci][co] += sum_pix X_tap[pix][ci] * dY[pix][co]
//                       both operands MN-major straight from the NHWC tiles; the 9 taps are stacked along M
//                       (LBO = tap-slot stride), K = pixels, accumulators live in TMEM across all of a CTA's
//                       pixel tiles and are flushed once with vector fp32 reductions.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issue, warps 2..5 = epilogue
// (TMEM lane quarter = warp_idx % 4).
#pragma once
#include "tc_prims.cuh"

namespace wacil {

enum ConvEpilogue : int {
  EPI_RAW = 0,    // bf16 store of the accumulator
  EPI_STATS = 1,  // bf16 store + per-channel sum / sum-of-squares of the stored (rounded) values
  EPI_INFER = 2,  // + bias[c] (+ residual) -> ReLU -> bf16 store   (BatchNorm folded into weights/bias)
  // Fused modes: cooperative launch, every tile of the CTA keeps its accumulator in TMEM across the exchange of the
  // per-CTA partial sums (published_row_sum4).
  EPI_BN_TRAIN = 3,  // fprop: store y, batch statistics -> exchange -> BatchNorm + shortcut + ReLU -> store a
  EPI_BNBWD = 4,     // dgrad: gz = (acc + shortcut grad) * relu mask, d-gamma/d-beta -> exchange -> gy (BN backward)
};

struct ConvGemmParams {
  int kc;         // K per tap (input channels of this GEMM): 16 / 32 / 64
  int nc;         // N (output channels of this GEMM): 16 .. 128, multiple of 16
  int taps;       // 9 (3x3, pad 1) or 1 (1x1)
  int stride;     // 1 or 2 (2: A map is the parity-plane view, see make_act_map)
  int Ht, Nt;     // tile = Nt images x Ht rows x full width; Nt*Ht*Wo == 128
  int tiles_h;    // tiles per image along H (1 when Nt > 1)
  int num_tiles;  // ceil(N*Ho*Wo / 128)
  int num_pixels; // N*Ho*Wo (store guard)
  int Ho, Wo;     // output spatial size
  int stages;     // A ring depth (in stages)
  int tps;        // taps per pipeline stage (one mbarrier round trip per stage)
  int epi;
  __nv_bfloat16* out;        // [num_pixels][nc]
  float* stats;              // [2*nc] sum, sumsq (EPI_STATS, a single-op diagnostic mode: global atomics, zero before the
                             // launch).  Fused modes: two buffers of [gridDim.x][2*nc] complemented words, the per-CTA
                             // partial sums (BN_TRAIN: sum, sumsq; BNBWD: d-beta, d-gamma) that every CTA publishes, polls
                             // and adds in CTA order (deterministic); zero before the first launch
  const float* bias;         // [nc] (EPI_INFER)
  const __nv_bfloat16* res;  // residual (EPI_INFER) or nullptr
  int res_stride;            // 1: same geometry as out; 2: residual is [N][2Ho][2Wo][res_c] sampled at even pixels
  int res_c;                 // residual channels (<= nc); channels >= res_c get no residual (zero pad)
  // ---- fused modes ----
  int acc_stages;            // TMEM accumulator stages: 2 (ring) or >= tiles per CTA (fused modes)
  unsigned* grid_bar;        // [2] this layer's launch counter + check-out counter (zero before the first launch)
  // EPI_BN_TRAIN: out = y (raw), act_out = a; stats/gamma/beta/running/saved as in bn_apply_fwd
  __nv_bfloat16* act_out;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float* saved;              // [2*nc] batch mean, rstd (written in BN_TRAIN, read in BNBWD)
  float eps, momentum;
  // EPI_BNBWD (this launch is the dgrad producing the gradient wrt a_l; BN backward of layer l):
  const __nv_bfloat16* g_res;   // shortcut gradient source or nullptr
  int gres_mode;                // 1 same shape; 2 source [N][Ho/2][Wo/2][2*nc], lands on even pixels
  const __nv_bfloat16* act_in;  // a_l (ReLU mask)
  const __nv_bfloat16* y_in;    // y_l (x-hat)
  float* dgamma;
  float* dbeta;
  __nv_bfloat16* gy;            // out: gradient wrt y_l
  __nv_bfloat16* gy_up;         // optional zero-upsampled copy (stride-2 layer l)
  __nv_bfloat16* gz;            // optional masked gradient (shortcut of the previous block)
  long long* dbg;               // optional [gridDim.x][16] clock64 timestamps of pipeline milestones (profiling)
  int pdl;                      // launched with programmatic stream serialization: wait for the upstream grid here
  int w_early;                  // weights were written >= 2 launches upstream: their TMA may precede the wait
  int rmerge;                   // row-merged activation loads (stride 1, 3x3, tile inside one image): ONE box of Ht+2 image
                                // rows per horizontal tap offset s; the three vertical taps r are M-windows of that box
                                // shifted by r*Wo rows (= r KB: whole swizzle atoms).  3 loads of (Ht+2) rows instead of 9
                                // loads of Ht rows: the L2->SM traffic of the main loop halves (stage 1) / drops 2.4x (stage 2)
};

#define WACIL_STAMP(slot)                                                        \
  do {                                                                           \
    if (p.dbg != nullptr) p.dbg[blockIdx.x * 16 + (slot)] = clock64();           \
  } while (0)

// bf16 <-> fp32 in registers: a bf16 is the upper half of its fp32 (shift / mask to widen), and two floats round to a
// packed pair with one cvt.rn.bf16x2.f32 — no per-element conversions, no local-memory staging arrays.
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);  // .x (low half) = lo
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ void unpack16_bf16(const uint4& a, const uint4& b, float (&v)[16]) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[2 * i] = bf16_lo(w[i]);
    v[2 * i + 1] = bf16_hi(w[i]);
  }
}
__device__ __forceinline__ void pack16_bf16(const float (&v)[16], uint4& a, uint4& b) {
  a = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  b = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
}
__device__ __forceinline__ void load16_bf16(const __nv_bfloat16* p, float (&v)[16]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  const uint4 a = __ldg(q), b = __ldg(q + 1);
  unpack16_bf16(a, b, v);
}
__device__ __forceinline__ void store16_bf16(__nv_bfloat16* p, const float (&v)[16]) {
  uint4 a, b;
  pack16_bf16(v, a, b);
  uint4* op = reinterpret_cast<uint4*>(p);
  op[0] = a;
  op[1] = b;
}

__device__ __forceinline__ void tap_coords(int taps, int stride, int tap, int kc, int h0, int n0, int& c0, int& c1,
                                           int& c2, int& c3, int& c4) {
  int r = 1, s = 1;
  if (taps == 9) {
    r = tap / 3;
    s = tap - 3 * r;
  }
  if (stride == 1) {
    // dims {C, W, H, N, 1}
    c0 = 0;
    c1 = s - 1;
    c2 = h0 + r - 1;
    c3 = n0;
    c4 = 0;
  } else {
    // dims {2C, W/2, 2, H/2, N}: input row 2*ho + r - 1 = 2*(ho + dh) + p
    const int pp = (r == 1) ? 0 : 1, dh = (r == 0) ? -1 : 0;
    const int qq = (s == 1) ? 0 : 1, dw = (s == 0) ? -1 : 0;
    c0 = qq * kc;
    c1 = dw;
    c2 = pp;
    c3 = h0 + dh;
    c4 = n0;
  }
}

// Column sums over the 32 lanes of a warp for 16 per-lane values (recursive halving, 16 shuffles).
// On return lane L holds in v0 the total of column  8*b4 + 4*b3 + 2*b2 + b1  (b_i = bit i of L).
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < half) {
        const float send = upper ? v[i] : v[i + half];
        const float keep = upper ? v[i + half] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ int warp_colsum16_col(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// ---- MMA issue loops.  One thread issues every tcgen05.mma of the CTA, and at these channel counts an MMA is cheap
// (mma_probe.cu: 64 cycles per instruction in a minimal loop, independent of N <= 128) while every scalar instruction
// of that lone thread costs ~6 cycles of dependent-issue latency: a loop with runtime trip counts, flag arithmetic
// and descriptor rebuilding spent ~250-300 cycles per MMA — the main loop was bound by this thread, not by TMA or the
// tensor core.  These loops are fully unrolled per (K steps, taps per stage): per MMA one predicate set, the MMA, and
// immediate-offset descriptor adds.
// Row-merged tile: the three boxes (s = 0..2) of a stage, each feeding the three vertical taps.
template <int KS>
__device__ __forceinline__ void issue_rmerge_tile(bool leader, uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                                  uint32_t idesc, uint32_t box16, uint32_t win16, uint32_t wtap16) {
#pragma unroll
  for (int sft = 0; sft < 3; ++sft) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        if (sft == 0 && r == 0 && k == 0) { if (leader) umma_bf16(d_tmem, adesc, bdesc, idesc, 0u); }
        else if (leader)
          umma_bf16_acc(d_tmem, adesc + (sft * box16 + r * win16 + 2 * k), bdesc + ((3 * r + sft) * wtap16 + 2 * k), idesc);
      }
    }
  }
}
template <int KS, int NTAP>
__device__ __forceinline__ void issue_tap_group(bool leader, uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                                uint32_t idesc, uint32_t a16, uint32_t wtap16, bool first_of_tile) {
#pragma unroll
  for (int j = 0; j < NTAP; ++j) {
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      if (j == 0 && k == 0) { if (leader) umma_bf16(d_tmem, adesc, bdesc, idesc, first_of_tile ? 0u : 1u); }
      else if (leader) umma_bf16_acc(d_tmem, adesc + (j * a16 + 2 * k), bdesc + (j * wtap16 + 2 * k), idesc);
    }
  }
}

constexpr int kConvThreads = 192;

// Fixed-order sum of rows first, first + step, ... (< rows) of a [rows][V] fp32 table, columns [4 cg, 4 cg + 4).  The
// 16-byte loads of a batch of 10 rows are independent and issued together (one L2 round trip per batch: the whole
// 148 x 32 table of a 16-channel layer is ONE batch for 128 threads); the additions keep the row order, so the result
// is the same on every CTA and every run.
__device__ __forceinline__ float4 ordered_row_sum4(const float* __restrict__ table, int V, int cg, int first, int step,
                                                   int rows) {
  constexpr int NB = 10;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r0 = first; r0 < rows; r0 += step * NB) {
    float4 t[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int r = r0 + j * step;
      t[j] = r < rows ? __ldcg(reinterpret_cast<const float4*>(table + static_cast<size_t>(r) * V) + cg)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      acc.x += t[j].x;
      acc.y += t[j].y;
      acc.z += t[j].z;
      acc.w += t[j].w;
    }
  }
  return acc;
}

// The statistics exchange of the fused kernels without a barrier.  A CTA publishes its per-channel partial sums as
// complemented fp32 words (~bits: an all-zero word = "not written yet"; 32-bit stores are single-copy atomic, so every
// word validates itself) into its own row of a [CTAs][V] table, and every CTA polls all rows until no word is zero.
// No release fence, no arrival counter, no separate "barrier passed -> now read the data" round trip: publishing costs
// one store, the wait IS the read.  The table is double-buffered on the parity of the layer's launch counter: a launch
// clears its row of the OTHER buffer (last read one launch ago, next written one launch ahead).
// Fixed-order sum of rows first, first + step, ... (< rows), words [4 cg, 4 cg + 4); NB rows are in flight together
// (one L2 round trip per batch once they are published: the whole 148 x 32 table of a 16-channel layer is ONE batch).
template <int NB>
__device__ __forceinline__ float4 published_row_sum4(const unsigned* __restrict__ table, int V, int cg, int first,
                                                     int step, int rows) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r0 = first; r0 < rows; r0 += step * NB) {
    uint4 t[NB];
    const long long t0 = clock64();
    bool ok;
    do {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int r = r0 + j * step;
        if (r < rows) {
          asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(t[j].x), "=r"(t[j].y), "=r"(t[j].z), "=r"(t[j].w)
                       : "l"(table + static_cast<size_t>(r) * V + 4 * cg)
                       : "memory");
        } else {
          t[j] = make_uint4(~0u, ~0u, ~0u, ~0u);  // +0.0f, published
        }
      }
      ok = true;
#pragma unroll
      for (int j = 0; j < NB; ++j) ok = ok && t[j].x != 0u && t[j].y != 0u && t[j].z != 0u && t[j].w != 0u;
      if (!ok && clock64() - t0 > 4000000000LL) {  // a bug must trap, not hang
        printf("wacil: statistics exchange timeout (block %d)\n", blockIdx.x);
        __trap();
      }
    } while (!ok);
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      acc.x += __uint_as_float(~t[j].x);
      acc.y += __uint_as_float(~t[j].y);
      acc.z += __uint_as_float(~t[j].z);
      acc.w += __uint_as_float(~t[j].w);
    }
  }
  return acc;
}

// EPI_T >= 0: the epilogue mode as a compile-time constant (the two fused modes are their own kernels: each gets its own
// register allocation and a third of the code); EPI_T < 0: p.epi at run time (RAW / STATS / INFER).
template <int EPI_T>
__device__ __forceinline__ void conv_gemm_body(const CUtensorMap& mapA, const CUtensorMap& mapW, const ConvGemmParams& p) {
  // (the run-time flavour never runs a fused mode: launch_conv sends those to their own kernels)
  const int epi = EPI_T >= 0 ? EPI_T : (p.epi < EPI_BN_TRAIN ? p.epi : EPI_RAW);
  constexpr int EW = 4;        // epilogue warps (one per TMEM lane quarter)
  constexpr int ET = EW * 32;  // epilogue threads
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) WACIL_STAMP(0);  // kernel entry
  const int rowbytes = p.kc * 2;                 // bytes per K row of A and W tiles
  const uint32_t a_bytes = 128u * rowbytes;      // one tap tile of A
  const bool rm = p.rmerge != 0;
  const uint32_t win_bytes = static_cast<uint32_t>(p.Wo) * rowbytes;          // one image row of the tile (rmerge)
  const uint32_t box_bytes = static_cast<uint32_t>(p.Ht + 2) * win_bytes;      // rmerge: Ht+2 image rows per load
  // one pipeline stage = tps tap tiles, or (row-merged) one box holding the three vertical taps of one s
  const uint32_t stage_bytes = rm ? 3 * box_bytes : p.tps * a_bytes;  // row-merged: the whole tile (s = 0, 1, 2) per stage
  const int groups = rm ? 1 : (p.taps + p.tps - 1) / p.tps;
  const uint32_t w_tap_bytes = p.nc * rowbytes;  // one tap tile of W
  const uint32_t w_bytes = p.taps * w_tap_bytes;

  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + ((w_bytes + 1023) & ~1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + p.stages * stage_bytes);
  uint64_t* full = bars;                  // [stages]
  uint64_t* empty = bars + p.stages;      // [stages]
  uint64_t* wfull = bars + 2 * p.stages;  // [1]
  const int nacc = p.acc_stages;
  uint64_t* accfull = wfull + 1;          // [nacc]
  uint64_t* accempty = accfull + nacc;    // [nacc]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + nacc);
  // [(2*EW+5)*64] per-channel scratch, 16-byte aligned (read with 128-bit loads; the size budget has 64 bytes of slack)
  float* s_stats = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));

  // Fused modes keep a second TMEM region beside the accumulators (same tile/column indexing, offset stash_col): what the
  // post-barrier phase needs besides the accumulator (BN_TRAIN: the shortcut; BNBWD: x-hat) is staged there in phase 1,
  // while the main loop still runs, so that phase 2 touches no global memory but its stores.
  const bool stash = (epi == EPI_BNBWD) || (epi == EPI_BN_TRAIN && p.res != nullptr);
  const uint32_t stash_col = static_cast<uint32_t>(nacc * p.nc);
  // BNBWD with a shortcut gradient: a third region holds it (fp32) when three regions fit half of the tensor memory
  // (the 16x16 and 8x8 layers); the 32x32 layers (7 tiles x 16 columns x 3 > 256) carry it in registers instead
  const bool gres3 = epi == EPI_BNBWD && p.g_res != nullptr && 3u * static_cast<uint32_t>(nacc * p.nc) <= 256u;
  const uint32_t gres_col = 2u * stash_col;
  const uint32_t tmem_cols = next_pow2_cols(static_cast<uint32_t>(nacc * p.nc) * (gres3 ? 3u : (stash ? 2u : 1u)));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapW);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(wfull, 1);
    for (int i = 0; i < nacc; ++i) {
      mbar_init(&accfull[i], 1);
      mbar_init(&accempty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  for (int i = threadIdx.x; i < (2 * EW + 5) * 64; i += blockDim.x) s_stats[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) WACIL_STAMP(1);  // barriers initialised, TMEM allocated

  // Programmatic dependent launch: everything above (and the weight TMA when the weights are older than the upstream
  // grid) overlaps the tail of the previous kernel; its outputs are only touched after griddepcontrol.wait.
  if (p.pdl && p.w_early && warp == 0) {  // one lane per tap: loads of one thread would run one after the other
    if (lane == 0) mbar_expect_tx_relaxed(wfull, w_bytes);
    __syncwarp();
    if (lane < p.taps) tma_load_2d(smem_w + lane * w_tap_bytes, &mapW, wfull, 0, lane * p.nc);
  }
  if (p.pdl) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // (whole warp in the loop, lane 0 issues: see the MMA issuer)
    {
      const bool leader = elect_one_sync();
      const bool timing = p.dbg != nullptr;
      if (!(p.pdl && p.w_early)) {
        if (leader) mbar_expect_tx_relaxed(wfull, w_bytes);
        __syncwarp();
        if (lane < p.taps) tma_load_2d(smem_w + lane * w_tap_bytes, &mapW, wfull, 0, lane * p.nc);
      }
      int stage = 0, bx = 0;
      uint32_t phase = 0;
      long long dbg_wait = 0, dbg_t = 0;  // profiling only (p.dbg): cycles blocked on a free slot
      const int nstages = p.stages, ntiles = p.num_tiles, tstride = gridDim.x, tps = p.tps, taps = p.taps;
      for (int tile = blockIdx.x; tile < ntiles; tile += tstride) {
        const int n0 = (tile / p.tiles_h) * p.Nt;
        const int h0 = (tile % p.tiles_h) * p.Ht;
        for (int g = 0; g < groups; ++g) {
          if (timing) dbg_t = clock64();
          mbar_wait(&empty[stage], phase ^ 1);
          if (timing) dbg_wait += clock64() - dbg_t;
          uint8_t* dst = smem_a + stage * stage_bytes;
#ifdef WACIL_MULTILANE_TMA
          // every box from its own lane
          const int j = (lane - bx) & 31;
          if (rm) {  // per horizontal tap offset s: image rows h0-1 .. h0+Ht, columns shifted by s-1 (OOB = zero padding)
            if (leader) mbar_expect_tx_relaxed(&full[stage], 3 * box_bytes);
            __syncwarp();
            if (j < 3) tma_load_5d(dst + j * box_bytes, &mapA, &full[stage], 0, j - 1, h0 - 1, n0, 0);
            bx = (bx + 3) & 31;
          } else {
            const int tap_lo = g * tps;
            const int ntap = min(tps, taps - tap_lo);
            if (leader) mbar_expect_tx_relaxed(&full[stage], ntap * a_bytes);
            __syncwarp();
            if (j < ntap) {
              int c0, c1, c2, c3, c4;
              tap_coords(taps, p.stride, tap_lo + j, p.kc, h0, n0, c0, c1, c2, c3, c4);
              tma_load_5d(dst + j * a_bytes, &mapA, &full[stage], c0, c1, c2, c3, c4);
            }
            bx = (bx + ntap) & 31;
          }
#else
          // the elected lane issues the boxes back to back (~50 cycles each; the pipeline step is bound by the mbarrier
          // round trips around them, not by the issue: tma_probe3.cu)
          (void)bx;
          if (rm) {  // per horizontal tap offset s: image rows h0-1 .. h0+Ht, columns shifted by s-1 (OOB = zero padding)
            if (leader) {
              mbar_expect_tx_relaxed(&full[stage], 3 * box_bytes);
              tma_load_5d(dst, &mapA, &full[stage], 0, -1, h0 - 1, n0, 0);
              tma_load_5d(dst + box_bytes, &mapA, &full[stage], 0, 0, h0 - 1, n0, 0);
              tma_load_5d(dst + 2 * box_bytes, &mapA, &full[stage], 0, 1, h0 - 1, n0, 0);
            }
          } else {
            const int tap_lo = g * tps;
            const int ntap = min(tps, taps - tap_lo);
            if (leader) mbar_expect_tx_relaxed(&full[stage], ntap * a_bytes);
            for (int j = 0; j < ntap; ++j) {
              int c0, c1, c2, c3, c4;
              tap_coords(taps, p.stride, tap_lo + j, p.kc, h0, n0, c0, c1, c2, c3, c4);
              if (leader) tma_load_5d(dst + j * a_bytes, &mapA, &full[stage], c0, c1, c2, c3, c4);
            }
          }
#endif
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (leader) {
        WACIL_STAMP(2);  // producer: every TMA load issued
        if (timing) p.dbg[blockIdx.x * 16 + 11] = dbg_wait;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // The whole warp runs this loop (warp-uniform control flow keeps the descriptor arithmetic in uniform registers, where
    // UTCHMMA wants its operands); only the tcgen05 instructions themselves are predicated on the leader lane.
    {
      const bool leader = elect_one_sync();
      const uint32_t layout = umma_layout_for_rowbytes(rowbytes);
      const uint32_t sbo = 8u * rowbytes;
      const uint32_t idesc = umma_idesc_bf16(128, p.nc, 0, 0);
      const int ksteps = p.kc / 16;
      const uint64_t desc_hi = umma_smem_desc(0, 16, sbo, layout);
      // descriptors differ only in the start-address field ((addr >> 4) in the low bits): add, don't rebuild
      const uint64_t a_ring_desc = desc_hi | (smem_u32(smem_a) >> 4);
      const uint64_t w_desc = desc_hi | (smem_u32(smem_w) >> 4);
      const uint32_t stage16 = stage_bytes >> 4, a16 = a_bytes >> 4, win16 = win_bytes >> 4, wtap16 = w_tap_bytes >> 4;
      const uint32_t box16 = box_bytes >> 4;
      const int nstages = p.stages, ntiles = p.num_tiles, tstride = gridDim.x, nc = p.nc, tps = p.tps, taps = p.taps;
      const bool timing = p.dbg != nullptr;
      mbar_wait(wfull, 0);
      tc_fence_after();
      if (leader) WACIL_STAMP(3);  // weights landed
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      long long dbg_wait = 0, dbg_t = 0;  // profiling only (p.dbg): cycles blocked on operands
      bool first_stage = true;
      for (int tile = blockIdx.x; tile < ntiles; tile += tstride) {
        // an mbarrier try_wait costs ~270 cycles even when the phase has completed (tma_probe3.cu): the two waits of a
        // tile go out from the two halves of the warp at the same time
        if (timing) dbg_t = clock64();
#ifdef WACIL_SERIAL_WAITS
        mbar_wait(&accempty[as], aphase ^ 1);
        mbar_wait(&full[stage], phase);
#else
        if (lane < 16) mbar_wait(&accempty[as], aphase ^ 1);
        else mbar_wait(&full[stage], phase);
        __syncwarp();
#endif
        if (timing) dbg_wait += clock64() - dbg_t;
        const uint32_t d_tmem = tmem_base + as * nc;
        for (int g = 0; g < groups; ++g) {
          if (g > 0) mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (first_stage) { if (leader) WACIL_STAMP(4); first_stage = false; }  // first activation stage landed
          const uint64_t adesc = a_ring_desc + stage * stage16;
          if (rm) {  // taps (r, s): M-window of box s shifted by r image rows; weight tap 3r + s
            if (ksteps == 1) issue_rmerge_tile<1>(leader, d_tmem, adesc, w_desc, idesc, box16, win16, wtap16);
            else if (ksteps == 2) issue_rmerge_tile<2>(leader, d_tmem, adesc, w_desc, idesc, box16, win16, wtap16);
            else issue_rmerge_tile<4>(leader, d_tmem, adesc, w_desc, idesc, box16, win16, wtap16);
          } else {
            const int tap_lo = g * tps;
            const uint64_t bdesc = w_desc + tap_lo * wtap16;
            const bool first = g == 0;
            const int ntap = min(tps, taps - tap_lo);
            if (ntap == 9) {
              if (ksteps == 1) issue_tap_group<1, 9>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
              else if (ksteps == 2) issue_tap_group<2, 9>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
              else issue_tap_group<4, 9>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
            } else if (ntap == 3) {
              if (ksteps == 1) issue_tap_group<1, 3>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
              else if (ksteps == 2) issue_tap_group<2, 3>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
              else issue_tap_group<4, 3>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
            } else {
              if (ksteps == 1) issue_tap_group<1, 1>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
              else if (ksteps == 2) issue_tap_group<2, 1>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
              else issue_tap_group<4, 1>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
            }
          }
          if (leader) umma_commit(&empty[stage]);
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (leader) umma_commit(&accfull[as]);
        if (++as == nacc) {
          as = 0;
          aphase ^= 1;
        }
      }
      if (leader) {
        WACIL_STAMP(5);  // every MMA issued
        if (timing) p.dbg[blockIdx.x * 16 + 12] = dbg_wait;
      }
    }
  } else if (EPI_T >= 0 && epi >= EPI_BN_TRAIN) {
    // ------------------------------------------------------------ fused epilogues (accumulators stay in TMEM)
    const int quarter = warp & 3;
    const int ew = warp - 2;          // epilogue warp index
    const int m = quarter * 32 + lane;
    const int tid128 = threadIdx.x - 64;  // index among the ET epilogue threads
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    float* s_part = s_stats;                    // [EW warps][2][64] private per-warp partial sums (zeroed at kernel start)
    float* s_c0 = s_stats + EW * 128;           // [nc] scale    | k = gamma * rstd
    float* s_c1 = s_stats + EW * 128 + 64;      // [nc] shift    | d-beta / M
    float* s_c2 = s_stats + EW * 128 + 128;     // [nc]          | d-gamma / M
    float* s_mean = s_stats + EW * 128 + 192;   // [nc]
    float* s_rstd = s_stats + EW * 128 + 256;   // [nc]
    const int ncr = p.nc;
    if (epi == EPI_BNBWD) {
      for (int c = tid128; c < p.nc; c += ET) {
        s_mean[c] = p.saved[c % ncr];
        s_rstd[c] = p.saved[ncr + c % ncr];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
    }
    // ---------------- phase 1: per-channel reductions (and the raw conv output in BN_TRAIN); stages phase 2's operands
    int as = 0;
    if (epi == EPI_BNBWD) {
      // Everything phase 1 needs from global memory — the activation (ReLU mask), the raw conv output (x-hat) and the
      // shortcut gradient — does not depend on this launch's MMAs, so it is fetched ONE TILE AHEAD of the accumulator
      // (the first tile: under the operand loads and the main loop): the mask goes to a 64-bit register, x-hat to the
      // stash region, the shortcut gradient to the third TMEM region (or, for the single-chunk 32x32 layers, stays in
      // 8 registers).  What follows the accumulator then touches tensor memory only.  (Fetched after the
      // accumulator, chunk by chunk, these loads were 1.2-1.6 k cycles of exposed latency per 16-channel chunk: 6-8 k
      // cycles behind the last MMA of an 8x8 layer, profiles/r02_layer_timelines.log.)
      const int gmode = p.g_res == nullptr ? 0 : (gres3 ? 1 : (p.nc == 16 ? 2 : 3));  // 3: no room, load late
      auto gres_ptr_of = [&](int pix) -> const __nv_bfloat16* {
        if (p.g_res == nullptr || pix >= p.num_pixels) return nullptr;
        if (p.gres_mode == 1) return p.g_res + static_cast<size_t>(pix) * p.nc;
        const int w = pix % p.Wo, h = (pix / p.Wo) % p.Ho, n = pix / (p.Wo * p.Ho);
        if (((w | h) & 1) != 0) return nullptr;
        return p.g_res + ((static_cast<size_t>(n) * (p.Ho / 2) + h / 2) * (p.Wo / 2) + w / 2) * (2 * p.nc);
      };
      // one 16-channel chunk of a tile: raw loads, then their use.  (Processing the current tile between the two, so
      // that the next tile's load latency hides under it, needs 24 live registers more than the 128 that keep the side
      // branches co-resident: the spills made the 32x32 layers 1 us slower, profiles/r02_layer_timelines.log.)
      struct RawChunk { uint4 a0, a1, y0, y1, g0, g1; };
      auto issue_chunk = [&](int tile, int c0, RawChunk& rc) {
        const int pix = tile * 128 + m;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        rc.a0 = rc.a1 = rc.y0 = rc.y1 = rc.g0 = rc.g1 = z;
        if (pix < p.num_pixels) {
          const size_t base = static_cast<size_t>(pix) * p.nc + c0;
          const uint4* qa = reinterpret_cast<const uint4*>(p.act_in + base);
          const uint4* qy = reinterpret_cast<const uint4*>(p.y_in + base);
          rc.a0 = __ldg(qa); rc.a1 = __ldg(qa + 1);
          rc.y0 = __ldg(qy); rc.y1 = __ldg(qy + 1);
        }
        if (gmode == 1 || gmode == 2) {
          const __nv_bfloat16* gp = gres_ptr_of(pix);
          if (gp != nullptr) {
            const uint4* qg = reinterpret_cast<const uint4*>(gp + c0);
            rc.g0 = __ldg(qg); rc.g1 = __ldg(qg + 1);
          }
        }
      };
      auto consume_chunk = [&](int as_, int c0, const RawChunk& rc, unsigned long long& mask, uint4& g0, uint4& g1) {
        float a[16], yv[16];
        unpack16_bf16(rc.a0, rc.a1, a);
        unpack16_bf16(rc.y0, rc.y1, yv);
        unsigned bits = 0u;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          bits |= (a[i] > 0.f ? 1u : 0u) << i;                      // rows beyond the last pixel loaded zeros: mask 0
          yv[i] = (yv[i] - s_mean[c0 + i]) * s_rstd[c0 + i];        // x-hat (garbage in masked rows is never used: gz = 0)
        }
        mask |= static_cast<unsigned long long>(bits) << c0;
        tmem_st16(tmem_base + lane_base + stash_col + as_ * p.nc + c0, yv);
        if (gmode == 1) {
          float r[16];
          unpack16_bf16(rc.g0, rc.g1, r);
          tmem_st16(tmem_base + lane_base + gres_col + as_ * p.nc + c0, r);
        } else if (gmode == 2) {
          g0 = rc.g0;
          g1 = rc.g1;
        }
      };
      auto stage_tile = [&](int tile, int as_, unsigned long long& mask, uint4& g0, uint4& g1) {
        mask = 0ull;
        for (int c0 = 0; c0 < p.nc; c0 += 16) {
          RawChunk rc;
          issue_chunk(tile, c0, rc);
          consume_chunk(as_, c0, rc, mask, g0, g1);
        }
      };
      unsigned long long mask_cur = 0ull, mask_nxt = 0ull;
      uint4 gc0 = make_uint4(0u, 0u, 0u, 0u), gc1 = gc0, gn0 = gc0, gn1 = gc0;
      int tile = blockIdx.x;
      if (tile < p.num_tiles) stage_tile(tile, 0, mask_cur, gc0, gc1);
      for (; tile < p.num_tiles; tile += gridDim.x, ++as) {
        const int nt = tile + static_cast<int>(gridDim.x);
        if (nt < p.num_tiles) stage_tile(nt, as + 1, mask_nxt, gn0, gn1);
        mbar_wait(&accfull[as], 0);
        tc_fence_after();
        if (tid128 == 0 && as == 0) WACIL_STAMP(6);  // first accumulator complete
        for (int c0 = 0; c0 < p.nc; c0 += 16) {
          float v[16], xh[16], s2[16];
          tmem_ld16(tmem_base + lane_base + as * p.nc + c0, v);
          tmem_ld16(tmem_base + lane_base + stash_col + as * p.nc + c0, xh);
          if (gmode == 1) {
            float r[16];
            tmem_ld16(tmem_base + lane_base + gres_col + as * p.nc + c0, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += r[i];
          } else if (gmode == 2) {
            float r[16];
            unpack16_bf16(gc0, gc1, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += r[i];
          } else if (gmode == 3) {
            const __nv_bfloat16* gp = gres_ptr_of(tile * 128 + m);
            if (gp != nullptr) {
              float r[16];
              load16_bf16(gp + c0, r);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += r[i];
            }
          }
          const unsigned bits = static_cast<unsigned>(mask_cur >> c0) & 0xffffu;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v[i] = ((bits >> i) & 1u) ? v[i] : 0.f;  // gz (masked gradient; rows beyond the last pixel have mask 0)
            s2[i] = v[i] * xh[i];
          }
          tmem_st16(tmem_base + lane_base + as * p.nc + c0, v);  // phase 2 reads gz here, x-hat from the stash
          const float t1 = warp_colsum16(v, lane);
          const float t2 = warp_colsum16(s2, lane);
          if ((lane & 1) == 0) {
            const int c = c0 + warp_colsum16_col(lane);
            s_part[(ew * 2 + 0) * 64 + c] += t1;
            s_part[(ew * 2 + 1) * 64 + c] += t2;
          }
        }
        mask_cur = mask_nxt;
        gc0 = gn0;
        gc1 = gn1;
      }
    }
    for (int tile = blockIdx.x; epi == EPI_BN_TRAIN && tile < p.num_tiles; tile += gridDim.x, ++as) {
      const int pix = tile * 128 + m;
      const bool valid = pix < p.num_pixels;
      const size_t base = static_cast<size_t>(valid ? pix : 0) * p.nc;
      const __nv_bfloat16* res_ptr = nullptr;
      if (p.res != nullptr && valid) {
        if (p.res_stride == 1) {
          res_ptr = p.res + static_cast<size_t>(pix) * p.res_c;
        } else {
          const int w = pix % p.Wo, h = (pix / p.Wo) % p.Ho, n = pix / (p.Wo * p.Ho);
          res_ptr = p.res + ((static_cast<size_t>(n) * (2 * p.Ho) + 2 * h) * (2 * p.Wo) + 2 * w) * p.res_c;
        }
      }
      if (stash) {
        // the shortcut does not depend on this tile's accumulator: fetch it while the MMAs of the tile are in flight
        for (int c0 = 0; c0 < p.nc; c0 += 16) {
          float r[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = 0.f;
          if (res_ptr != nullptr && c0 < p.res_c) load16_bf16(res_ptr + c0, r);
          tmem_st16(tmem_base + lane_base + stash_col + as * p.nc + c0, r);
        }
      }
      mbar_wait(&accfull[as], 0);
      tc_fence_after();
      if (tid128 == 0 && as == 0) WACIL_STAMP(6);  // first accumulator complete
      for (int c0 = 0; c0 < p.nc; c0 += 16) {
        float v[16], s2[16];
        tmem_ld16(tmem_base + lane_base + as * p.nc + c0, v);
        {
          uint4 oa, ob;
          pack16_bf16(v, oa, ob);  // y as stored (bf16); the statistics are taken over the rounded values
          if (valid) {
            uint4* op = reinterpret_cast<uint4*>(p.out + base + c0);
            op[0] = oa;
            op[1] = ob;
          } else {
            oa = ob = make_uint4(0u, 0u, 0u, 0u);
          }
          unpack16_bf16(oa, ob, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) s2[i] = v[i] * v[i];
        }
        const float t1 = warp_colsum16(v, lane);
        const float t2 = warp_colsum16(s2, lane);
        if ((lane & 1) == 0) {  // this lane owns column c of this warp's private partial row: plain read-modify-write
          const int c = c0 + warp_colsum16_col(lane);
          s_part[(ew * 2 + 0) * 64 + c] += t1;
          s_part[(ew * 2 + 1) * 64 + c] += t2;
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
    if (tid128 == 0) WACIL_STAMP(7);  // phase 1 done
    // this layer's launch counter p.grid_bar[0] (bumped by the last CTA to leave, below) picks the table buffer
    unsigned launch_no;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(launch_no) : "l"(p.grid_bar) : "memory");
    const size_t table_words = static_cast<size_t>(gridDim.x) * (2 * ncr);
    unsigned* const rows_cur = reinterpret_cast<unsigned*>(p.stats) + (launch_no & 1u) * table_words;
    {
      // this CTA's partial sums -> its own row of the [grid][2*nc] table (complemented words, see published_row_sum4);
      // its row of the other buffer back to "not written".  No atomics anywhere on the BatchNorm sums, so the
      // statistics are bit-reproducible run to run.
      unsigned* g0 = rows_cur + static_cast<size_t>(blockIdx.x) * (2 * ncr);
      unsigned* z0 = reinterpret_cast<unsigned*>(p.stats) + ((launch_no & 1u) ^ 1u) * table_words +
                     static_cast<size_t>(blockIdx.x) * (2 * ncr);
      for (int c = tid128; c < p.nc; c += ET) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int wv = 0; wv < EW; ++wv) {
          a0 += s_part[(2 * wv + 0) * 64 + c];
          a1 += s_part[(2 * wv + 1) * 64 + c];
        }
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(g0 + c), "r"(~__float_as_uint(a0)) : "memory");
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(g0 + ncr + c), "r"(~__float_as_uint(a1)) : "memory");
        z0[c] = 0u;
        z0[ncr + c] = 0u;
      }
    }
    // what the per-channel constants need besides the grid-wide sums is fetched before the barrier (nc <= 64 <= 128
    // threads: one channel per thread)
    const int myc = tid128 < p.nc ? tid128 : 0;
    const float gam = p.gamma[myc % ncr];
    const float bet = epi == EPI_BN_TRAIN ? p.beta[myc % ncr] : 0.f;
    float rmean = 0.f, rvar = 0.f;
    if (epi == EPI_BN_TRAIN && blockIdx.x == 0 && tid128 < ncr) {
      rmean = p.running_mean[tid128];
      rvar = p.running_var[tid128];
    }
    // ---------------- per-channel constants from the grid-wide sums
    const float cnt = static_cast<float>(p.num_pixels);
    float* s_scr = s_stats + (2 * EW + 5) * 64;  // [S][V] = 4 * ET floats, beside the per-warp partials still being read
    {
      // every CTA sums the per-CTA rows in the same fixed order: thread (slice, 4 values) takes rows slice, slice + S,
      // ... (16-byte loads, all independent), the S slice sums are combined in order below
      const int V = 2 * p.nc, Q = V / 4, S = ET / Q;  // 8/16/32 column groups -> 16/8/4 row slices
      const int cg = tid128 % Q, sl = tid128 / Q;
      *reinterpret_cast<float4*>(s_scr + sl * V + 4 * cg) =
          published_row_sum4<10>(rows_cur, V, cg, sl, S, static_cast<int>(gridDim.x));
    }
    if (tid128 == 0) WACIL_STAMP(8);  // every CTA's sums seen
    asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
    // check-out ticket: taken now (this CTA has read every row), looked at after phase 2 — the atomic's round trip
    // then lies neither on the path to the constants nor between the last store and the end of the kernel
    unsigned ticket = 0;
    if (tid128 == 0) ticket = atomicAdd(p.grid_bar + 1, 1u);
    if (tid128 < p.nc) {
      const int c = tid128;
      float t0 = 0.f, t1 = 0.f;
      const int V = 2 * p.nc, S = 4 * ET / V;
      for (int sl = 0; sl < S; ++sl) {
        t0 += s_scr[sl * V + c];
        t1 += s_scr[sl * V + p.nc + c];
      }
      if (epi == EPI_BN_TRAIN) {
        const float mean = t0 / cnt;
        float var = t1 / cnt - mean * mean;
        var = fmaxf(var, 0.f);
        const float rstd = rsqrtf(var + p.eps);
        const float sc = gam * rstd;
        s_c0[c] = sc;
        s_c1[c] = bet - mean * sc;
        if (blockIdx.x == 0 && c < ncr) {
          p.saved[c] = mean;
          p.saved[ncr + c] = rstd;
          p.running_mean[c] = (1.f - p.momentum) * rmean + p.momentum * mean;
          p.running_var[c] = (1.f - p.momentum) * rvar + p.momentum * var * (cnt / (cnt - 1.f));
        }
      } else {
        s_c0[c] = gam * s_rstd[c];
        s_c1[c] = t0 / cnt;
        s_c2[c] = t1 / cnt;
        if (blockIdx.x == 0 && c < ncr) {  // the parameter gradients of this BatchNorm
          p.dbeta[c] = t0;
          p.dgamma[c] = t1;
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
    if (tid128 == 0) WACIL_STAMP(13);  // per-channel constants ready
    // ---------------- phase 2: normalise (forward) / BN backward, from TMEM only (accumulator or gz, and the stash)
    as = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++as) {
      const int pix = tile * 128 + m;
      const bool valid = pix < p.num_pixels;
      const size_t base = static_cast<size_t>(valid ? pix : 0) * p.nc;
      size_t up_base = 0;
      if (epi == EPI_BNBWD && p.gy_up != nullptr) {
        const int w = pix % p.Wo, h = (pix / p.Wo) % p.Ho, n = pix / (p.Wo * p.Ho);
        up_base = ((static_cast<size_t>(n) * (2 * p.Ho) + 2 * h) * (2 * p.Wo) + 2 * w) * p.nc;
      }
      for (int c0 = 0; c0 < p.nc; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + lane_base + as * p.nc + c0, v);
        if (epi == EPI_BN_TRAIN) {
          {
            uint4 ya, yb;
            pack16_bf16(v, ya, yb);  // the stored (rounded) conv output is what gets normalised
            unpack16_bf16(ya, yb, v);
            const float4* k0 = reinterpret_cast<const float4*>(s_c0 + c0);
            const float4* k1 = reinterpret_cast<const float4*>(s_c1 + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 sc = k0[i], sh = k1[i];
              v[4 * i] = v[4 * i] * sc.x + sh.x;
              v[4 * i + 1] = v[4 * i + 1] * sc.y + sh.y;
              v[4 * i + 2] = v[4 * i + 2] * sc.z + sh.z;
              v[4 * i + 3] = v[4 * i + 3] * sc.w + sh.w;
            }
          }
          if (stash) {
            float r[16];
            tmem_ld16(tmem_base + lane_base + stash_col + as * p.nc + c0, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += r[i];
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          if (valid) store16_bf16(p.act_out + base + c0, v);
        } else {
          float xh[16], o[16];
          tmem_ld16(tmem_base + lane_base + stash_col + as * p.nc + c0, xh);
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = s_c0[c0 + i] * (v[i] - s_c1[c0 + i] - xh[i] * s_c2[c0 + i]);
          if (valid) {
            store16_bf16(p.gy + base + c0, o);
            if (p.gy_up != nullptr) store16_bf16(p.gy_up + up_base + c0, o);
            if (p.gz != nullptr) store16_bf16(p.gz + base + c0, v);
          }
        }
      }
    }
    if (tid128 == 0) {
      WACIL_STAMP(9);  // phase 2 done
      // launch counter: the last CTA to check out knows that every CTA has read every row of this launch's table
      if (ticket == gridDim.x - 1) {
        p.grid_bar[1] = 0u;
        __threadfence();
        p.grid_bar[0] = launch_no + 1u;
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&accfull[as], aphase);
      tc_fence_after();
      if (threadIdx.x == 64 && tile == static_cast<int>(blockIdx.x)) WACIL_STAMP(6);
      const int pix = tile * 128 + m;
      const bool valid = pix < p.num_pixels;
      const __nv_bfloat16* res_ptr = nullptr;
      if (epi == EPI_INFER && p.res != nullptr && valid) {
        if (p.res_stride == 1) {
          res_ptr = p.res + static_cast<size_t>(pix) * p.res_c;
        } else {
          const int w = pix % p.Wo;
          const int h = (pix / p.Wo) % p.Ho;
          const int n = pix / (p.Wo * p.Ho);
          res_ptr = p.res + ((static_cast<size_t>(n) * (2 * p.Ho) + 2 * h) * (2 * p.Wo) + 2 * w) * p.res_c;
        }
      }
      for (int c0 = 0; c0 < p.nc; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * p.nc + c0, v);
        if (epi == EPI_INFER) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + c0 + i);
          if (res_ptr != nullptr && c0 < p.res_c) {
            float r[16];
            load16_bf16(res_ptr + c0, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += r[i];
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        uint4 oa, ob;
        pack16_bf16(v, oa, ob);
        if (valid) {
          uint4* op = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(pix) * p.nc + c0);
          op[0] = oa;
          op[1] = ob;
        } else {
          oa = ob = make_uint4(0u, 0u, 0u, 0u);
        }
        if (epi == EPI_STATS) {
          float s1[16], s2[16];
          unpack16_bf16(oa, ob, s1);
#pragma unroll
          for (int i = 0; i < 16; ++i) s2[i] = s1[i] * s1[i];
          const float t1 = warp_colsum16(s1, lane);
          const float t2 = warp_colsum16(s2, lane);
          if ((lane & 1) == 0) {
            const int c = c0 + warp_colsum16_col(lane);
            s_stats[(quarter * 2 + 0) * 64 + c] += t1;
            s_stats[(quarter * 2 + 1) * 64 + c] += t2;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&accempty[as]);
      if (++as == nacc) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (epi == EPI_STATS) {
      // the four epilogue warps only: named barrier 1, 128 threads
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int e = threadIdx.x - 64;
      for (int c = e; c < p.nc; c += 128) {
        atomicAdd(&p.stats[c], s_stats[0 * 64 + c] + s_stats[2 * 64 + c] + s_stats[4 * 64 + c] + s_stats[6 * 64 + c]);
        atomicAdd(&p.stats[p.nc + c],
                  s_stats[1 * 64 + c] + s_stats[3 * 64 + c] + s_stats[5 * 64 + c] + s_stats[7 * 64 + c]);
      }
    }
    if (threadIdx.x == 64) WACIL_STAMP(9);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
  if (threadIdx.x == 0) WACIL_STAMP(10);  // teardown done
}

// Three entry points over one body.  Register caps measured on the step graph (profiles/r02_register_caps_ab.log): an
// SM holds the CTA of the running layer, the CTA of the NEXT layer (programmatic dependent launch: its prologue runs
// under this one's tail) and a CTA of the side branch (teacher / weight gradient), 192 threads each.  Above 128 the
// branches stop overlapping; the fused backward at 120 (a few spilled words) is 13 us per step faster than at 128;
// the forward at 112 / 120 is slower than at 128; the run-time flavour needs 106.
#ifndef WACIL_FWD_REGS
#define WACIL_FWD_REGS 128
#endif
#ifndef WACIL_BWD_REGS
#define WACIL_BWD_REGS 120
#endif
#ifndef WACIL_GEN_REGS
#define WACIL_GEN_REGS 128
#endif
__global__ void __maxnreg__(WACIL_GEN_REGS)
    conv_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                     const ConvGemmParams p) {  // RAW / STATS / INFER
  conv_gemm_body<-1>(mapA, mapW, p);
}
__global__ void __maxnreg__(WACIL_FWD_REGS)
    conv_gemm_kernel_bn_fwd(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                            const ConvGemmParams p) {  // conv + batch statistics + BatchNorm + shortcut + ReLU
  conv_gemm_body<EPI_BN_TRAIN>(mapA, mapW, p);
}
__global__ void __maxnreg__(WACIL_BWD_REGS)
    conv_gemm_kernel_bn_bwd(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                            const ConvGemmParams p) {  // data gradient + ReLU mask + BatchNorm backward
  conv_gemm_body<EPI_BNBWD>(mapA, mapW, p);
}
typedef void (*ConvKernelFn)(const CUtensorMap, const CUtensorMap, const ConvGemmParams);
inline ConvKernelFn conv_kernel_for(int epi) {
  return epi == EPI_BN_TRAIN ? conv_gemm_kernel_bn_fwd : (epi == EPI_BNBWD ? conv_gemm_kernel_bn_bwd : conv_gemm_kernel);
}

// ------------------------------------------------------------------------------------------------
struct WgradParams {
  int cin;            // channels of X (M rows per tap): 16/32/64
  int cout;           // channels of dY (N): 16/32/64
  int taps;           // 9 or 1
  int stride;         // conv stride (selects the X map / tap coordinates)
  int Ht, Nt;         // tile over OUTPUT pixels: Nt*Ht*Wo == 128
  int tiles_h;
  int num_tiles;      // pixel tiles in total
  int ring;           // group ring depth (32 KB each)
  float* part;        // [pixel chunks][taps*cin][cout] fp32: every CTA stores ITS partial sum with plain stores (no atomics);
                      // wgrad_reduce_kernel adds the slots in chunk order -> bit-reproducible gradients
  int gsplit;         // CTAs per pixel chunk (1 .. groups): CTA (chunk, j) accumulates the 128-row tap groups j, j + gsplit, ...
                      // of the chunk's tiles and writes those rows of the chunk's slot.  The 8x8 layers have 64 pixel tiles
                      // and five 128-row groups of 64 channels each: splitting the groups puts 160 CTAs to work instead of
                      // 32 and keeps a CTA at 64 TMEM columns instead of 512
  int pdl;            // launched with programmatic stream serialization
};
// Pixel chunks of a wgrad launch that own tiles (= partial slots written): contiguous chunks of `per` tiles
__host__ __device__ inline int wgrad_active_ctas(int num_tiles, int chunks) {
  const int per = (num_tiles + chunks - 1) / chunks;
  return (num_tiles + per - 1) / per;
}

__global__ void __launch_bounds__(kConvThreads, 2)
    wgrad_gemm_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDY,
                      const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int xrow = p.cin * 2;                 // bytes per pixel row of an X tap tile
  const int yrow = p.cout * 2;                // bytes per pixel row of the dY tile
  const uint32_t slot_bytes = 128u * xrow;    // one tap tile
  const int tpg = 128 / p.cin;                // taps stacked per 128-row M group
  const int groups = (p.taps + tpg - 1) / tpg;
  const uint32_t group_bytes = 128u * 128u * 2u;  // tpg * slot_bytes == 32 KB always
  const uint32_t y_bytes = 128u * yrow;

  uint8_t* smem_g = smem;
  uint8_t* smem_y = smem + p.ring * group_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_y + 2 * ((y_bytes + 1023) & ~1023u));
  uint64_t* gfull = bars;              // [ring]
  uint64_t* gempty = bars + p.ring;    // [ring]
  uint64_t* yfull = bars + 2 * p.ring; // [2]
  uint64_t* yempty = yfull + 2;        // [2]
  uint64_t* accfull = yempty + 2;      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);
  const uint32_t y_stride = (y_bytes + 1023) & ~1023u;

  // this CTA: pixel chunk `chunk` (contiguous tiles), tap groups gj, gj + gsplit, ...
  const int gsplit = p.gsplit > 0 ? p.gsplit : 1;
  const int chunk = blockIdx.x / gsplit, gj = blockIdx.x % gsplit;
  const int chunks = gridDim.x / gsplit;
  const int my_groups = (groups - gj + gsplit - 1) / gsplit;
  const uint32_t tmem_cols = next_pow2_cols(static_cast<uint32_t>(my_groups * p.cout));
  const int per = (p.num_tiles + chunks - 1) / chunks;
  const int t_begin = chunk * per;
  const int t_end = min(p.num_tiles, t_begin + per);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapDY);
    for (int i = 0; i < p.ring; ++i) {
      mbar_init(&gfull[i], 1);
      mbar_init(&gempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&yfull[i], 1);
      mbar_init(&yempty[i], 1);
    }
    mbar_init(accfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.pdl) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }

  // Producer and MMA warps run their loops with all 32 lanes (uniform control flow: the descriptor arithmetic stays in
  // uniform registers); the TMA / tcgen05 instructions are issued by the elect.sync leader (see conv_gemm_body).
  if (warp == 0) {
    // TMA loads issued by one thread run one after the other (~750 cycles each, tma_probe.cu); every box goes out from
    // its own lane and the lane window rotates (bx), so the up to ten boxes of a pixel tile overlap
    const bool leader = lane == 0;
    int gs = 0, ys = 0, bx = 0;
    uint32_t gphase = 0, yphase = 0;
    const int ring = p.ring, taps = p.taps;
    for (int tile = t_begin; tile < t_end; ++tile) {
      const int n0 = (tile / p.tiles_h) * p.Nt;
      const int h0 = (tile % p.tiles_h) * p.Ht;
      mbar_wait(&yempty[ys], yphase ^ 1);
      if (leader) mbar_expect_tx_relaxed(&yfull[ys], y_bytes);
      __syncwarp();
      if (((lane - bx) & 31) == 0) tma_load_2d(smem_y + ys * y_stride, &mapDY, &yfull[ys], 0, tile * 128);
      bx = (bx + 1) & 31;
      if (++ys == 2) {
        ys = 0;
        yphase ^= 1;
      }
      for (int g = gj; g < groups; g += gsplit) {
        const int tap_lo = g * tpg;
        const int ntap = min(tpg, taps - tap_lo);
        mbar_wait(&gempty[gs], gphase ^ 1);
        if (leader) mbar_expect_tx_relaxed(&gfull[gs], ntap * slot_bytes);
        __syncwarp();
        {
          const int j = (lane - bx) & 31;
          if (j < ntap) {
            int c0, c1, c2, c3, c4;
            tap_coords(taps, p.stride, tap_lo + j, p.cin, h0, n0, c0, c1, c2, c3, c4);
            tma_load_5d(smem_g + gs * group_bytes + j * slot_bytes, &mapX, &gfull[gs], c0, c1, c2, c3, c4);
          }
          bx = (bx + ntap) & 31;
        }
        if (++gs == ring) {
          gs = 0;
          gphase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (t_begin < t_end) {
      const bool leader = elect_one_sync();
      const uint32_t xlayout = umma_layout_for_rowbytes(xrow);
      const uint32_t ylayout = umma_layout_for_rowbytes(yrow);
      const uint32_t idesc = umma_idesc_bf16(128, p.cout, 1, 1);
      // MN-major: LBO = stride between MN atoms (tap slots), SBO = stride between 8-pixel K groups; the 8 K-steps of
      // 16 pixels of a tile differ in the start-address field only
      const uint64_t a_hi = umma_smem_desc(0, slot_bytes, 8u * xrow, xlayout);
      const uint64_t b_hi = umma_smem_desc(0, y_bytes, 8u * yrow, ylayout);
      const uint32_t ka16 = (16u * xrow) >> 4, kb16 = (16u * yrow) >> 4;
      const int ring = p.ring, cout = p.cout;
      int gs = 0, ys = 0;
      uint32_t gphase = 0, yphase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        mbar_wait(&yfull[ys], yphase);
        tc_fence_after();
        const uint64_t b0 = b_hi | (smem_u32(smem_y + ys * y_stride) >> 4);
        const bool first_tile = tile == t_begin;
        for (int g = gj, gi = 0; g < groups; g += gsplit, ++gi) {
          mbar_wait(&gfull[gs], gphase);
          tc_fence_after();
          const uint64_t a0 = a_hi | (smem_u32(smem_g + gs * group_bytes) >> 4);
          const uint32_t d_tmem = tmem_base + gi * cout;
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // 8 K-steps of 16 pixels
            if (j == 0) { if (leader) umma_bf16(d_tmem, a0, b0, idesc, first_tile ? 0u : 1u); }
            else if (leader) umma_bf16_acc(d_tmem, a0 + j * ka16, b0 + j * kb16, idesc);
          }
          if (leader) umma_commit(&gempty[gs]);
          if (++gs == ring) {
            gs = 0;
            gphase ^= 1;
          }
        }
        if (leader) umma_commit(&yempty[ys]);
        if (++ys == 2) {
          ys = 0;
          yphase ^= 1;
        }
      }
      if (leader) umma_commit(accfull);
    }
  } else if (t_begin < t_end) {
    const int quarter = warp & 3;
    mbar_wait(accfull, 0);
    tc_fence_after();
    const int rows = p.taps * p.cin;
    for (int g = gj, gi = 0; g < groups; g += gsplit, ++gi) {
      const int row = g * 128 + quarter * 32 + lane;
      for (int c0 = 0; c0 < p.cout; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + gi * p.cout + c0, v);
        if (row < rows) {
          float4* dst = reinterpret_cast<float4*>(p.part + (static_cast<size_t>(chunk) * rows + row) * p.cout + c0);
          dst[0] = make_float4(v[0], v[1], v[2], v[3]);
          dst[1] = make_float4(v[4], v[5], v[6], v[7]);
          dst[2] = make_float4(v[8], v[9], v[10], v[11]);
          dst[3] = make_float4(v[12], v[13], v[14], v[15]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// Ordered reduction of the per-chunk weight-gradient slots: dst[e] = part[0][e] + part[1][e] + ... in a fixed tree.
// A warp shares one float4 of the result: lane j adds slots j, j+32, ... in order (independent loads, issued together),
// then a fixed xor-shuffle tree combines the 32 sums.  blockIdx.y selects the layer.
struct WgradReduceDesc {
  const float* part;
  float* dst;
  int nparts;
  int elems;  // taps*cin*cout, a multiple of 4
};
__device__ __forceinline__ void wgrad_reduce_body(const WgradReduceDesc& d) {
  const int n4 = d.elems >> 2;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < n4; e += warps) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* src = reinterpret_cast<const float4*>(d.part) + e;
    for (int r0 = lane; r0 < d.nparts; r0 += 32 * 4) {  // four independent 16-byte loads in flight per lane
      float4 t[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = r0 + 32 * j;
        t[j] = r < d.nparts ? __ldcg(src + static_cast<size_t>(r) * n4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc.x += t[j].x; acc.y += t[j].y; acc.z += t[j].z; acc.w += t[j].w; }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (lane == 0) reinterpret_cast<float4*>(d.dst)[e] = acc;
  }
}
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const WgradReduceDesc* __restrict__ descs) {
  wgrad_reduce_body(descs[blockIdx.y]);
}
__global__ void __launch_bounds__(256) wgrad_reduce_one_kernel(const WgradReduceDesc d) { wgrad_reduce_body(d); }

// ------------------------------------------------------------------------------------------------
// Host side: tensor-map construction and launch helpers (no torch, no allocation).
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_tmap_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_tmapEncodeTiled>(ptr);
  return fn;
}

inline CUtensorMapSwizzle swizzle_for_rowbytes(int rowbytes) {
  return rowbytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                         : (rowbytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// Activation map for a conv whose INPUT is x[N][H][W][C] (bf16) and whose output tile is
// Nt images x Ht rows x (full output width).  stride 1: dims {C, W, H, N, 1}; stride 2: the parity-plane
// view {2C, W/2, 2, H/2, N} so that every tap is a unit-stride box.  Returns 0 on success.
inline int make_act_map(CUtensorMap* map, const void* base, int N, int H, int W, int C, int stride, int Ht,
                        int Nt, int box_rows = 0) {  // box_rows > 0 (stride 1): rows per box instead of Ht (row-merged loads)
  PFN_tmapEncodeTiled enc = get_tmap_encode();
  if (!enc) return -100;
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  const cuuint64_t e = 2;
  if (stride == 1) {
    dims[0] = C; dims[1] = W; dims[2] = H; dims[3] = N; dims[4] = 1;
    strides[0] = C * e; strides[1] = W * C * e; strides[2] = (cuuint64_t)H * W * C * e;
    strides[3] = (cuuint64_t)N * H * W * C * e;
    box[0] = C; box[1] = W; box[2] = box_rows > 0 ? box_rows : Ht; box[3] = Nt; box[4] = 1;
  } else {
    dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = N;
    strides[0] = 2 * C * e; strides[1] = W * C * e; strides[2] = 2 * (cuuint64_t)W * C * e;
    strides[3] = (cuuint64_t)H * W * C * e;
    box[0] = C; box[1] = W / 2; box[2] = 1; box[3] = Ht; box[4] = Nt;
  }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_rowbytes(C * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1000 - static_cast<int>(r);
}

// Row-major 2D bf16 matrix [rows][cols] with a {cols, box_rows} box (cols*2 bytes must be 32/64/128).
inline int make_mat_map(CUtensorMap* map, const void* base, long long rows, int cols, int box_rows) {
  PFN_tmapEncodeTiled enc = get_tmap_encode();
  if (!enc) return -100;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_rowbytes(cols * 2),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1000 - static_cast<int>(r);
}

inline void conv_tile_shape(int Ho, int Wo, int& Ht, int& Nt, int& tiles_h) {
  // 128 output pixels per tile, full-width rows.
  const int rows = 128 / Wo;  // rows of width Wo per tile
  if (rows <= Ho) {
    Ht = rows; Nt = 1; tiles_h = Ho / rows;
  } else {
    Ht = Ho; Nt = rows / Ho; tiles_h = 1;
  }
}

inline size_t conv_smem_bytes(const ConvGemmParams& p) {
  const size_t w = ((size_t)p.taps * p.nc * p.kc * 2 + 1023) & ~size_t(1023);
  const size_t stage = p.rmerge ? (size_t)3 * (p.Ht + 2) * p.Wo * p.kc * 2 : (size_t)p.tps * 128 * p.kc * 2;
  return 1024 + w + (size_t)p.stages * stage + (2 * p.stages + 1 + 2 * p.acc_stages) * 8 + 16 +
         21 * 64 * 4 + 64;  // per-channel scratch sized for 8 epilogue warps
}
// Row-merged loads apply to stride-1 3x3 convolutions whose 128-pixel tile lies inside one image (32x32 and 16x16
// maps) and whose image row is a whole number of 1 KB swizzle spans (so that the r-shifted windows stay atom-aligned).
inline bool conv_can_rmerge(int kc, int nc, int taps, int stride, int Ho, int Wo, size_t total_budget = 113 * 1024) {
  if (!(stride == 1 && taps == 9 && Ho * Wo >= 128 && (Wo * kc * 2) % 1024 == 0)) return false;
  const size_t w = ((size_t)9 * nc * kc * 2 + 1023) & ~size_t(1023);
  const size_t tile = (size_t)3 * (128 / Wo + 2) * Wo * kc * 2;  // three boxes of Ht+2 image rows
  return w + 6144 + 2 * tile <= total_budget;                    // at least a double buffer inside the footprint
}
inline void conv_pick_rmerge_pipeline(int kc, int nc, int Ht, int Wo, int& tps, int& stages,
                                      size_t total_budget = 113 * 1024) {
  const size_t w = ((size_t)9 * nc * kc * 2 + 1023) & ~size_t(1023);
  const size_t budget = total_budget - 6144 - w;
  const size_t tile = (size_t)3 * (Ht + 2) * Wo * kc * 2;  // three boxes
  tps = 9;  // taps per stage: the whole tile
  int st = (int)(budget / tile);
  stages = st > 4 ? 4 : (st < 2 ? 2 : st);
}
// Pipeline shape: a stage carries `tps` tap tiles so the producer / MMA threads pay one mbarrier round trip per
// stage instead of per tap (the MMAs themselves are only 8..32 cycles each at these channel counts).
// Per-CTA footprint is kept under half an SM in BOTH shared memory (<= 113 KB) and registers (192 threads x <= 160 regs =
// 30.7 K) so that CTAs of the two concurrently running graph branches (teacher || student forward, weight gradients ||
// data-gradient chain) can share an SM.  Measured on the overlapped step graph (profiles/r01_footprint_experiments.md):
// 1.48 ms with this footprint vs 1.78 ms with 194 KB pipelines and 1.71 ms with 320 threads x 168 regs, although
// those variants are as fast or faster in isolation: the kernels are latency-bound, a second resident CTA hides it.
// (The occupancy API reports 1 CTA/SM for any tcgen05 kernel — csrc/occ_probe.cu — so a cooperative grid is <= #SMs.)
constexpr size_t kConvSmemMax = 113 * 1024;  // 2 x (113 KB + 1 KB reserved) = 228 KB per SM
constexpr size_t kConvSmemWide = 208 * 1024;  // launches on <= half of the SMs (they do not share an SM with a side branch)
inline void conv_pick_pipeline(int kc, int nc, int taps, int& tps, int& stages, size_t total_budget = kConvSmemMax) {
  const size_t w = ((size_t)taps * nc * kc * 2 + 1023) & ~size_t(1023);
  const size_t budget = total_budget - 6144 - w;  // alignment slack, barriers, per-channel scratch
  const size_t tap_bytes = (size_t)128 * kc * 2;
  const int cand[3] = {9, 3, 1};
  tps = 1;
  stages = 2;
  for (int c : cand) {
    if (c > taps || taps % c) continue;
    const int st = (int)(budget / (c * tap_bytes));
    if (st >= 2) { tps = c; stages = st > 6 ? 6 : st; break; }
  }
}
inline size_t wgrad_smem_bytes(const WgradParams& p) {
  const size_t y = (((size_t)128 * p.cout * 2) + 1023) & ~size_t(1023);
  return 1024 + (size_t)p.ring * 32768 + 2 * y + (2 * p.ring + 5) * 8 + 64;
}

}  // namespace wacil
