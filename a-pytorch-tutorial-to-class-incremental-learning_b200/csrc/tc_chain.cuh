// tc_chain.cuh — several convolution layers of the CIFAR ResNet in ONE persistent cooperative launch
// (reference op sequence: resnet.py:40-53,107-116 forward; autograd of template.py:270 backward).
//
// conv_gemm_kernel (tc_conv.cuh) runs one layer per launch: at these sizes a launch is ~2.5 us of HBM time inside
// ~17 us of latency (setup, first operands, grid barrier, drain), and a training step chains 61 of them.  Here a CTA
// keeps its barriers, TMEM allocation and role warps alive across a whole run of layers; the only things that
// separate two layers are
//   * the grid barrier BatchNorm needs anyway (batch statistics / d-gamma, d-beta over ALL pixels), and
//   * per-tile "ready" flags in global memory: the TMA producer of layer l+1 waits only for the (at most a handful of)
//     producer tiles whose rows its halo'd box touches, so loads and MMAs of layer l+1 start while the other tiles
//     of layer l are still in their post-barrier phase.
// Three epilogue families, one kernel:
//   EPI_BN_TRAIN  conv -> batch statistics -> grid barrier -> BN + shortcut + ReLU          (student forward)
//   EPI_BNBWD     dgrad of layer l -> ReLU mask + shortcut gradient + d-gamma/d-beta of layer l-1 -> grid barrier
//                 -> BatchNorm backward                                                     (data-gradient chain)
//   EPI_INFER     conv + folded-BN bias + shortcut + ReLU, no grid barrier                  (teacher / eval / features)
//
// Warp roles (64 + 32*EW threads): warp 0 TMA producer, warp 1 TMEM alloc + MMA issue, warps 2..2+EW-1 epilogue
// (EW = 4 or 8; with 8, two groups of four warps take alternate tiles).  Every TMA box of a stage is issued by its own
// lane (loads of one lane queue behind each other, lanes overlap) after a relaxed expect_tx.
// Two instantiations: FUSED (BatchNorm modes, <= 168 registers) and inference only (<= 96 registers).  Every CTA stays
// under HALF an SM in registers (192 x 168 = 32 256) and shared memory (<= 113 KB) and under 256 TMEM columns: only then
// do CTAs of a second kernel share the SM (measured: a 43 K-register chain CTA and an 18 K-register weight-gradient CTA
// ran strictly one after the other) — the teacher's inference chain beside the student's forward, the weight-gradient
// kernels beside the data-gradient chain.
// Hand-off between layers: all layers of a launch share one tile geometry (a launch never crosses a resolution
// change), tile t belongs to CTA t mod G, and a 3x3 box of tile t touches tiles t-1, t, t+1 only — so a CTA needs the
// rows of CTAs b-1, b, b+1.  Every epilogue warp publishes "my rows of this layer are stored" ONCE per layer (one
// gpu-scope release per warp and layer; per-tile releases stalled the epilogue warps ~1 k cycles per tile on the
// store acknowledgements) and the producer polls three per-CTA counters before the first load of the next layer.
//
// Determinism: BatchNorm sums go through per-CTA partial rows added in CTA order by every CTA; nothing is atomic
// except the barrier counter itself.
#pragma once
#include "tc_conv.cuh"

namespace wacil {

constexpr int kChainMaxStages = 6;   // operand ring depth (mbarrier pairs)
constexpr int kChainMaxSlots = 8;    // TMEM accumulator slots = tiles of a CTA in a fused layer
// Operand pipelines per CTA (producer warp + MMA issuer warp each, tiles dealt round-robin).  Two were measured: same
// time per layer as one (16.9 vs 16.8 us for a 32x32 layer) — the main loop is not bound by either warp's own chain
// of barrier round trips — while the extra 64 threads push a FUSED CTA over half an SM's registers, which stops CTAs
// of a second kernel from sharing the SM (profiles/r02_chain_experiments.md).  One pipeline, 192 threads.
constexpr int kChainPipes = 1;
#ifndef WACIL_CHAIN_REGS
#define WACIL_CHAIN_REGS 160
#endif

struct ChainLayer {
  // ---- geometry of this layer's implicit GEMM (3x3, pad 1)
  int kc, nc;            // K per tap / N
  int stride;            // 1, or 2 (A map = parity-plane view)
  int Ht, Nt, tiles_h;   // tile = Nt images x Ht rows x full width (128 pixels)
  int num_tiles, num_pixels, Ho, Wo;
  int rmerge, tps, stages;  // operand pipeline shape (see ConvGemmParams)
  int epi;
  int tmem_sync;         // accumulator slot layout differs from the previous layer's: wait until all its slots are free
  int ring_sync;         // operand ring shape differs from the previous layer's: drain it first
  // ---- dependency on the previous layer of the chain
  int dep;               // 0: the input was complete before the launch; 1: wait for CTAs b-1, b, b+1 of that layer
  const unsigned* dep_flags;  // [gridDim.x] epilogue warps of CTA c that have stored all their rows of the previous layer
  unsigned* flags;       // [gridDim.x] the same for THIS layer's output (nullptr: nobody in this launch reads it)
  // ---- epilogue operands (a subset per mode, see ConvGemmParams)
  int res_stride, res_c, gres_mode;
  float eps, momentum;
  __nv_bfloat16* out;             // BN_TRAIN: raw conv output y;  INFER: the layer output
  __nv_bfloat16* act_out;         // BN_TRAIN: a = relu(bn(y) + shortcut)
  const __nv_bfloat16* res;       // shortcut source or nullptr
  const float* bias;              // INFER
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float* saved;                   // [2*nc] batch mean, rstd (written by BN_TRAIN, read by BNBWD)
  float* stats;                   // [gridDim.x][2*nc] per-CTA partial sums
  const __nv_bfloat16* g_res;
  const __nv_bfloat16* act_in;
  const __nv_bfloat16* y_in;
  float* dgamma;
  float* dbeta;
  __nv_bfloat16* gy;
  __nv_bfloat16* gy_up;
  __nv_bfloat16* gz;
};

struct ChainParams {
  const ChainLayer* layers;   // [n_layers] in global memory
  const CUtensorMap* maps;    // [2*n_layers] in global memory: activation map, weight map per layer (64-byte aligned)
  int n_layers;
  unsigned* grid_bar;         // monotonic arrival counter, zero before the launch
  int pdl;                    // launched with programmatic stream serialization
  uint32_t w_bytes_max;       // weight tile region (multiple of 1024)
  uint32_t ring_bytes;        // operand ring region (multiple of 1024)
  uint32_t tmem_cols;         // power of two >= the widest layer's accumulator + stash columns
  long long* dbg;             // optional [gridDim.x][n_layers][48] clock64 stamps: 0 phase 1 done, 1 grid barrier passed,
                              // 2 phase 2 stored, 3 published | 4 producer: dependencies met, 5 last load issued |
                              // 6 MMA warp: first operands landed, 7 last MMA issued
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// TMEM slot hand-back: the ordering that matters (tcgen05.ld before the next tcgen05.mma) comes from the tcgen05 fences
// around the synchronisation; a releasing arrive would additionally wait for the tile's global stores to be acknowledged.
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One epilogue warp has stored all its rows of a layer: count it in this CTA's flag (gpu-scope release, cumulative over
// the warp's stores after __syncwarp).
__device__ __forceinline__ void chain_publish(unsigned* flags, int lane) {
  __syncwarp();
  if (lane == 0 && flags != nullptr) {
    fence_proxy_async();  // consumers read these rows through TMA (async proxy)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(flags + blockIdx.x) : "memory");
  }
}

template <int ET>
__device__ __forceinline__ void chain_grid_barrier(unsigned* bar, unsigned target, int tid) {
  asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
  if (tid == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    const long long t0 = clock64();
    while (static_cast<int>(ld_acquire_gpu(bar) - target) < 0) {
      if (clock64() - t0 > 4000000000LL) {
        printf("wacil: chain grid barrier timeout (block %d, target %u)\n", blockIdx.x, target);
        __trap();
      }
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
}

template <int EW, bool FUSED>
__global__ void __maxnreg__(FUSED ? (EW == 4 ? WACIL_CHAIN_REGS : 96) : 128) conv_chain_kernel(const ChainParams cp) {
  constexpr int ET = EW * 32;  // epilogue threads
  constexpr int NG = EW / 4;   // epilogue groups (tiles alternate between them)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + cp.w_bytes_max;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + cp.ring_bytes);
  uint64_t* full = bars;                            // [kChainMaxStages]
  uint64_t* empty = full + kChainMaxStages;         // [kChainMaxStages]
  uint64_t* wfull = empty + kChainMaxStages;        // [1]
  uint64_t* wempty = wfull + 1;                     // [1]
  uint64_t* accfull = wempty + 1;                   // [kChainMaxSlots]
  uint64_t* accempty = accfull + kChainMaxSlots;    // [kChainMaxSlots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + kChainMaxSlots);
  float* s_stats = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kChainMaxStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(wfull, 1);
    mbar_init(wempty, kChainPipes);  // every MMA issuer has finished reading the weight tile
    for (int i = 0; i < kChainMaxSlots; ++i) {
      mbar_init(&accfull[i], 1);
      mbar_init(&accempty[i], 4);  // the four warps (one per TMEM lane quarter) that own a tile
    }
    fence_mbar_init();
  }
  if (warp == kChainPipes) tmem_alloc(tmem_slot, cp.tmem_cols);
  for (int i = threadIdx.x; i < (2 * EW + 5) * 64; i += blockDim.x) s_stats[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (cp.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");

  const int nl = cp.n_layers;
  const int G = gridDim.x;
  // register budget: 128 per thread at launch (two CTAs per SM); TMA / MMA warps need few, the epilogue many
  if (warp < kChainPipes) {
    // ================================================================================== TMA producers
    // Two independent operand pipelines: producer p feeds MMA issuer p through its own half of the ring and its own
    // full/empty barriers, and the pair handles the tiles whose running index is p modulo 2 (all tap groups of a tile
    // stay in one pipeline).  Sharing stage barriers between two consumers does not work: a consumer that waits for a
    // use two phases ahead of the barrier sees "parity complete" at once.  Within a stage every box goes out from its own
    // lane (rotating window `bx`): loads of one lane queue behind each other, lanes overlap.
    const bool leader = lane == 0;
    const int me = warp;      // pipeline index; producer 0 also loads the weight tiles
    uint64_t* myfull = full + me * (kChainMaxStages / kChainPipes);
    uint64_t* myempty = empty + me * (kChainMaxStages / kChainPipes);
    uint8_t* myring = smem_a + me * (cp.ring_bytes / kChainPipes);
    uint32_t ring_phase = 0;  // bit s: uses of my stage s so far, mod 2
    int stage = 0, wuse = 0, seq = 0;
    int bx = 0;               // rotating first lane of the next group of boxes
    for (int l = 0; l < nl; ++l) {
      const ChainLayer* L = cp.layers + l;
      const int ntiles = L->num_tiles;
      if (l == nl - 1 && cp.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
      if (static_cast<int>(blockIdx.x) >= ntiles) continue;  // this CTA owns no tile of the layer
      const CUtensorMap* mapA = cp.maps + 2 * l;
      const CUtensorMap* mapW = mapA + 1;
      const int kc = L->kc, nc = L->nc, stride = L->stride, Ht = L->Ht, Nt = L->Nt, tiles_h = L->tiles_h;
      const int tps = L->tps, nstages = L->stages >= kChainPipes ? L->stages / kChainPipes : 1;  // my share of the layer's stages
      const bool rm = L->rmerge != 0;
      const int rowbytes = kc * 2;
      const uint32_t a_bytes = 128u * rowbytes;
      const uint32_t box_bytes = static_cast<uint32_t>(Ht + 2) * L->Wo * rowbytes;
      const uint32_t stage_bytes = rm ? 3 * box_bytes : tps * a_bytes;
      const int groups = rm ? 1 : 9 / tps;
      const uint32_t w_tap_bytes = nc * rowbytes;
      if (leader) {
        tma_prefetch_desc(mapA);
        tma_prefetch_desc(mapW);
      }
      // weights of this layer: the tile is free once the previous layer's last MMAs (both issuers) have completed
      if (me == 0) {
        mbar_wait(wempty, (wuse & 1) ^ 1);
        // expect_tx strictly before the loads (bytes completing ahead of their expectation wedged the barrier on
        // hardware), relaxed: no release fence in front of it
        if (leader) mbar_expect_tx_relaxed(wfull, 9 * w_tap_bytes);
        __syncwarp();
        const int t = (lane - bx - 1) & 31;  // tap t of the weight tile from lane (bx + 1 + t) % 32
        if (t < 9) tma_load_2d(smem_w + t * w_tap_bytes, mapW, wfull, 0, t * nc);
        bx = (bx + 9) & 31;
      }
      ++wuse;
      if (L->ring_sync) {  // stages change size: every stage of the old shape must have been consumed
        for (int s = 0; s < kChainMaxStages / kChainPipes; ++s) mbar_wait(&myempty[s], ((ring_phase >> s) & 1) ^ 1);
        stage = 0;
      }
      if (L->dep != 0) {
        // rows of tiles t-1, t, t+1 of the previous layer = CTAs b-1, b, b+1 (mod G): three counters, once per layer
        const unsigned* dflags = L->dep_flags;
        if (lane < 3) {
          const int c = (static_cast<int>(blockIdx.x) + lane - 1 + G) % G;
          const long long t0 = clock64();
          while (ld_acquire_gpu(dflags + c) < static_cast<unsigned>(EW)) {
            if (clock64() - t0 > 4000000000LL) {
              printf("wacil: chain flag timeout (block %d layer %d waits for block %d)\n", blockIdx.x, l, c);
              __trap();
            }
          }
        }
        __syncwarp();
        fence_proxy_async();  // those rows were written through the generic proxy, TMA reads through the async one
      }
      if (cp.dbg != nullptr && leader && me == 0) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 4] = clock64();
      for (int tile = blockIdx.x; tile < ntiles; tile += G, ++seq) {
        if ((seq % kChainPipes) != me) continue;  // another pipeline's tile
        const int n0 = (tile / tiles_h) * Nt;
        const int h0 = (tile % tiles_h) * Ht;
        for (int g = 0; g < groups; ++g) {
          mbar_wait(&myempty[stage], ((ring_phase >> stage) & 1) ^ 1);
          if (cp.dbg != nullptr && leader && g == 0 && (tile - blockIdx.x) / G < 8) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 16 + (tile - blockIdx.x) / G] = clock64();
          uint8_t* dst = myring + stage * stage_bytes;
          const int j = (lane - bx - 1) & 31;  // box j of this stage from lane (bx + 1 + j) % 32 (lane 0: expect_tx)
          if (rm) {
            if (leader) mbar_expect_tx_relaxed(&myfull[stage], 3 * box_bytes);
            __syncwarp();
            if (j < 3) tma_load_5d(dst + j * box_bytes, mapA, &myfull[stage], 0, j - 1, h0 - 1, n0, 0);
            bx = (bx + 3) & 31;
          } else {
            const int tap_lo = g * tps;
            if (leader) mbar_expect_tx_relaxed(&myfull[stage], tps * a_bytes);
            __syncwarp();
            if (j < tps) {
              int c0, c1, c2, c3, c4;
              tap_coords(9, stride, tap_lo + j, kc, h0, n0, c0, c1, c2, c3, c4);
              tma_load_5d(dst + j * a_bytes, mapA, &myfull[stage], c0, c1, c2, c3, c4);
            }
            bx = (bx + tps) & 31;
          }
          ring_phase ^= 1u << stage;
          if (++stage == nstages) stage = 0;
        }
      }
      if (cp.dbg != nullptr && leader && me == 0) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 5] = clock64();
    }
  } else if (warp < 2 * kChainPipes) {
    // ================================================================================== MMA issuers
    // Issuer m consumes pipeline m (see the producers): the tiles whose running index is m modulo 2.
    const bool leader = elect_one_sync();
    const int me = warp - kChainPipes;
    uint64_t* myfull = full + me * (kChainMaxStages / kChainPipes);
    uint64_t* myempty = empty + me * (kChainMaxStages / kChainPipes);
    uint8_t* myring = smem_a + me * (cp.ring_bytes / kChainPipes);
    uint32_t ring_phase = 0, acc_phase = 0;
    int stage = 0, wuse = 0, seq = 0;
    for (int l = 0; l < nl; ++l) {
      const ChainLayer* L = cp.layers + l;
      const int ntiles = L->num_tiles;
      if (l == nl - 1 && cp.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
      if (static_cast<int>(blockIdx.x) >= ntiles) continue;
      const int kc = L->kc, nc = L->nc, tps = L->tps, nstages = L->stages >= kChainPipes ? L->stages / kChainPipes : 1;
      const bool rm = L->rmerge != 0;
      const bool fused = FUSED && L->epi >= EPI_BN_TRAIN;
      const int rowbytes = kc * 2;
      const uint32_t a_bytes = 128u * rowbytes;
      const uint32_t win_bytes = static_cast<uint32_t>(L->Wo) * rowbytes;
      const uint32_t box_bytes = static_cast<uint32_t>(L->Ht + 2) * win_bytes;
      const uint32_t stage_bytes = rm ? 3 * box_bytes : tps * a_bytes;
      const int groups = rm ? 1 : 9 / tps;
      const uint32_t w_tap_bytes = nc * rowbytes;
      const uint32_t layout = umma_layout_for_rowbytes(rowbytes);
      const uint32_t idesc = umma_idesc_bf16(128, nc, 0, 0);
      const int ksteps = kc / 16;
      const uint64_t desc_hi = umma_smem_desc(0, 16, 8u * rowbytes, layout);
      const uint64_t a_ring_desc = desc_hi | (smem_u32(myring) >> 4);
      const uint64_t w_desc = desc_hi | (smem_u32(smem_w) >> 4);
      const uint32_t stage16 = stage_bytes >> 4, a16 = a_bytes >> 4, win16 = win_bytes >> 4, wtap16 = w_tap_bytes >> 4;
      const uint32_t box16 = box_bytes >> 4;
      mbar_wait(wfull, wuse & 1);
      ++wuse;
      tc_fence_after();
      if (L->tmem_sync)  // the slots move: every accumulator of the previous layer must have been read
        for (int s = 0; s < kChainMaxSlots; ++s) mbar_wait(&accempty[s], ((acc_phase >> s) & 1) ^ 1);
      if (L->ring_sync) stage = 0;
      int k = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += G, ++k, ++seq) {
        const int slot = fused ? k : (seq & 1);
        if ((seq % kChainPipes) != me) {  // another issuer's tile
          acc_phase ^= 1u << slot;
          continue;
        }
        const uint32_t d_tmem = tmem_base + (fused ? slot * nc : slot * 64);
        // the two waits of a tile in parallel (half a warp each): their round trips overlap
        if (lane < 16) mbar_wait(&accempty[slot], ((acc_phase >> slot) & 1) ^ 1);
        else mbar_wait(&myfull[stage], (ring_phase >> stage) & 1);
        __syncwarp();
        for (int g = 0; g < groups; ++g) {
          if (g > 0) mbar_wait(&myfull[stage], (ring_phase >> stage) & 1);
          tc_fence_after();
          if (cp.dbg != nullptr && leader && k == 0 && g == 0) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 6] = clock64();
          if (cp.dbg != nullptr && leader && g == 0 && k < 8) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 24 + k] = clock64();
          const uint64_t adesc = a_ring_desc + stage * stage16;
          if (rm) {
            if (ksteps == 1) issue_rmerge_tile<1>(leader, d_tmem, adesc, w_desc, idesc, box16, win16, wtap16);
            else if (ksteps == 2) issue_rmerge_tile<2>(leader, d_tmem, adesc, w_desc, idesc, box16, win16, wtap16);
            else issue_rmerge_tile<4>(leader, d_tmem, adesc, w_desc, idesc, box16, win16, wtap16);
          } else {
            const uint64_t bdesc = w_desc + (g * tps) * wtap16;
            const bool first = g == 0;
            if (tps == 9) {
              if (ksteps == 1) issue_tap_group<1, 9>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
              else if (ksteps == 2) issue_tap_group<2, 9>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
              else issue_tap_group<4, 9>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
            } else if (tps == 3) {
              if (ksteps == 1) issue_tap_group<1, 3>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
              else if (ksteps == 2) issue_tap_group<2, 3>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
              else issue_tap_group<4, 3>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
            } else {
              if (ksteps == 1) issue_tap_group<1, 1>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
              else if (ksteps == 2) issue_tap_group<2, 1>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
              else issue_tap_group<4, 1>(leader, d_tmem, adesc, bdesc, idesc, a16, wtap16, first);
            }
          }
          if (leader) umma_commit(&myempty[stage]);
          ring_phase ^= 1u << stage;
          if (++stage == nstages) stage = 0;
        }
        if (leader) umma_commit(&accfull[slot]);
        if (cp.dbg != nullptr && leader && k < 8) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 32 + k] = clock64();
        acc_phase ^= 1u << slot;
      }
      if (leader) umma_commit(wempty);  // the weight tile may be overwritten once these MMAs have completed
      if (cp.dbg != nullptr && leader) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 7] = clock64();
    }
  } else {
    // ================================================================================== epilogue warps
    const int ew = warp - 2 * kChainPipes;
    const int quarter = warp & 3;
    const int group = ew >> 2;            // which of the NG groups of four warps
    const int m = quarter * 32 + lane;
    const int tid = threadIdx.x - 64 * kChainPipes;  // index among the ET epilogue threads
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    float* s_part = s_stats;                    // [EW warps][2][64] private per-warp partial sums (zero at rest)
    float* s_c0 = s_stats + EW * 128;           // [nc] scale    | k = gamma * rstd
    float* s_c1 = s_stats + EW * 128 + 64;      // [nc] shift    | d-beta / M
    float* s_c2 = s_stats + EW * 128 + 128;     // [nc]          | d-gamma / M
    float* s_mean = s_stats + EW * 128 + 192;   // [nc]
    float* s_rstd = s_stats + EW * 128 + 256;   // [nc]
    uint32_t acc_phase = 0;
    int seq = 0;          // tiles of this CTA so far (all layers): INFER slot parity
    unsigned nbar = 0;    // grid barriers passed
    for (int l = 0; l < nl; ++l) {
      const ChainLayer* L = cp.layers + l;
      const int ntiles = L->num_tiles, npix = L->num_pixels, nc = L->nc, epi = L->epi;
      const int Ho = L->Ho, Wo = L->Wo;
      if (l == nl - 1 && cp.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
      const int my_tiles = static_cast<int>(blockIdx.x) < ntiles ? (ntiles - static_cast<int>(blockIdx.x) + G - 1) / G : 0;
      if (!FUSED || epi == EPI_INFER) {
        // ------------------------------------------------------------------ bias + shortcut + ReLU, tile by tile
        const __nv_bfloat16* res = L->res;
        const int res_stride = L->res_stride, res_c = L->res_c;
        const float* bias = L->bias;
        __nv_bfloat16* outp = L->out;
        int k = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += G, ++k, ++seq) {
          const int slot = seq & 1;
          if ((k % NG) == group) {
            mbar_wait(&accfull[slot], (acc_phase >> slot) & 1);
            tc_fence_after();
            const int pix = tile * 128 + m;
            const bool valid = pix < npix;
            const __nv_bfloat16* res_ptr = nullptr;
            if (res != nullptr && valid) {
              if (res_stride == 1) {
                res_ptr = res + static_cast<size_t>(pix) * res_c;
              } else {
                const int w = pix % Wo, h = (pix / Wo) % Ho, n = pix / (Wo * Ho);
                res_ptr = res + ((static_cast<size_t>(n) * (2 * Ho) + 2 * h) * (2 * Wo) + 2 * w) * res_c;
              }
            }
            for (int c0 = 0; c0 < nc; c0 += 16) {
              float v[16];
              tmem_ld16(tmem_base + lane_base + slot * 64 + c0, v);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += __ldg(bias + c0 + i);
              if (res_ptr != nullptr && c0 < res_c) {
                float r[16];
                load16_bf16(res_ptr + c0, r);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += r[i];
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
              if (valid) store16_bf16(outp + static_cast<size_t>(pix) * nc + c0, v);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(&accempty[slot]);
          }
          acc_phase ^= 1u << slot;
        }
        chain_publish(L->flags, lane);
        continue;
      }
      if constexpr (FUSED) {
      // -------------------------------------------------------------------- fused BatchNorm modes
      // slots move (channel count / tiles per CTA change): the other epilogue group may still be reading the old ones
      if (L->tmem_sync) asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
      const bool has_stash = (epi == EPI_BNBWD) || (L->res != nullptr);
      const int nacc = (ntiles + G - 1) / G;  // slots of the layer (tiles of CTA 0)
      const uint32_t stash_col = static_cast<uint32_t>(nacc * nc);
      const float* saved = L->saved;
      if (epi == EPI_BNBWD) {
        for (int c = tid; c < nc; c += ET) {
          s_mean[c] = saved[c];
          s_rstd[c] = saved[nc + c];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
      }
      // ---------------- phase 1: per-channel reductions; stages what phase 2 needs in TMEM
      {
        const __nv_bfloat16* res = L->res;
        const int res_stride = L->res_stride, res_c = L->res_c, gres_mode = L->gres_mode;
        const __nv_bfloat16* g_res = L->g_res;
        const __nv_bfloat16* act_in = L->act_in;
        const __nv_bfloat16* y_in = L->y_in;
        int k = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += G, ++k) {
          if ((k % NG) != group) continue;
          const int pix = tile * 128 + m;
          const bool valid = pix < npix;
          const size_t base = static_cast<size_t>(valid ? pix : 0) * nc;
          const __nv_bfloat16* gres_ptr = nullptr;
          const __nv_bfloat16* res_ptr = nullptr;
          if (epi == EPI_BNBWD && g_res != nullptr && valid) {
            if (gres_mode == 1) {
              gres_ptr = g_res + base;
            } else {
              const int w = pix % Wo, h = (pix / Wo) % Ho, n = pix / (Wo * Ho);
              if (((w | h) & 1) == 0)
                gres_ptr = g_res + ((static_cast<size_t>(n) * (Ho / 2) + h / 2) * (Wo / 2) + w / 2) * (2 * nc);
            }
          }
          if (epi == EPI_BN_TRAIN && res != nullptr && valid) {
            if (res_stride == 1) {
              res_ptr = res + static_cast<size_t>(pix) * res_c;
            } else {
              const int w = pix % Wo, h = (pix / Wo) % Ho, n = pix / (Wo * Ho);
              res_ptr = res + ((static_cast<size_t>(n) * (2 * Ho) + 2 * h) * (2 * Wo) + 2 * w) * res_c;
            }
          }
          if (epi == EPI_BN_TRAIN && has_stash) {
            // the shortcut does not depend on this tile's accumulator: fetch it while the MMAs of the tile are in flight
            for (int c0 = 0; c0 < nc; c0 += 16) {
              float r[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = 0.f;
              if (res_ptr != nullptr && c0 < res_c) load16_bf16(res_ptr + c0, r);
              tmem_st16(tmem_base + lane_base + stash_col + k * nc + c0, r);
            }
          }
          mbar_wait(&accfull[k], (acc_phase >> k) & 1);
          tc_fence_after();
          if (cp.dbg != nullptr && tid == 0 && k == 0) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 11] = clock64();
          for (int c0 = 0; c0 < nc; c0 += 16) {
            float v[16], s2[16];
            tmem_ld16(tmem_base + lane_base + k * nc + c0, v);
            if (epi == EPI_BN_TRAIN) {
              uint4 oa, ob;
              pack16_bf16(v, oa, ob);  // the statistics are taken over the values as stored (bf16); phase 2 stores them
              if (!valid) oa = ob = make_uint4(0u, 0u, 0u, 0u);
              unpack16_bf16(oa, ob, v);
#pragma unroll
              for (int i = 0; i < 16; ++i) s2[i] = v[i] * v[i];
            } else {
              float a[16], yv[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) { a[i] = 0.f; yv[i] = 0.f; }
              if (valid) {
                load16_bf16(act_in + base + c0, a);
                load16_bf16(y_in + base + c0, yv);
                if (gres_ptr != nullptr) {
                  float r[16];
                  load16_bf16(gres_ptr + c0, r);
#pragma unroll
                  for (int i = 0; i < 16; ++i) v[i] += r[i];
                }
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                v[i] = (valid && a[i] > 0.f) ? v[i] : 0.f;                         // gz (masked gradient)
                yv[i] = valid ? (yv[i] - s_mean[c0 + i]) * s_rstd[c0 + i] : 0.f;   // x-hat
              }
              tmem_st16(tmem_base + lane_base + k * nc + c0, v);
              tmem_st16(tmem_base + lane_base + stash_col + k * nc + c0, yv);
#pragma unroll
              for (int i = 0; i < 16; ++i) s2[i] = v[i] * yv[i];
            }
            const float t1 = warp_colsum16(v, lane);
            const float t2 = warp_colsum16(s2, lane);
            if ((lane & 1) == 0) {
              const int c = c0 + warp_colsum16_col(lane);
              s_part[(ew * 2 + 0) * 64 + c] += t1;
              s_part[(ew * 2 + 1) * 64 + c] += t2;
            }
          }
        }
      }
      if (cp.dbg != nullptr && tid == 0) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 12] = clock64();
      asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
      if (cp.dbg != nullptr && tid == 0) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 0] = clock64();
      {
        float* g0 = L->stats + static_cast<size_t>(blockIdx.x) * (2 * nc);
        for (int c = tid; c < nc; c += ET) {
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int wv = 0; wv < EW; ++wv) {
            a0 += s_part[(2 * wv + 0) * 64 + c];
            a1 += s_part[(2 * wv + 1) * 64 + c];
          }
          g0[c] = a0;
          g0[nc + c] = a1;
        }
      }
      const int myc = tid < nc ? tid : 0;
      const float gam = L->gamma[myc];
      const float bet = epi == EPI_BN_TRAIN ? L->beta[myc] : 0.f;
      float rmean = 0.f, rvar = 0.f;
      if (epi == EPI_BN_TRAIN && blockIdx.x == 0 && tid < nc) {
        rmean = L->running_mean[tid];
        rvar = L->running_var[tid];
      }
      ++nbar;
      chain_grid_barrier<ET>(cp.grid_bar, nbar * static_cast<unsigned>(G), tid);
      if (cp.dbg != nullptr && tid == 0) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 1] = clock64();
      // ---------------- per-channel constants from the per-CTA partial rows, added in CTA order by every CTA
      const float cnt = static_cast<float>(npix);
      {
        const int V = 2 * nc, Q = V / 4, S = ET / Q;  // 8/16/32 column groups of 4 values -> ET/Q row slices
        const int cg = tid % Q, sl = tid / Q;
        // scratch [S][V] over the (already consumed) per-warp partials
        *reinterpret_cast<float4*>(s_part + sl * V + 4 * cg) = ordered_row_sum4(L->stats, V, cg, sl, S, G);
      }
      if (cp.dbg != nullptr && tid == 0) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 8] = clock64();
      asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
      if (tid < nc) {
        const int c = tid;
        const int V = 2 * nc, S = 4 * ET / V;
        float t0 = 0.f, t1 = 0.f;
        for (int sl = 0; sl < S; ++sl) {
          t0 += s_part[sl * V + c];
          t1 += s_part[sl * V + nc + c];
        }
        if (epi == EPI_BN_TRAIN) {
          const float mean = t0 / cnt;
          float var = t1 / cnt - mean * mean;
          var = fmaxf(var, 0.f);
          const float rstd = rsqrtf(var + L->eps);
          const float sc = gam * rstd;
          s_c0[c] = sc;
          s_c1[c] = bet - mean * sc;
          if (blockIdx.x == 0) {
            L->saved[c] = mean;
            L->saved[nc + c] = rstd;
            const float mo = L->momentum;
            L->running_mean[c] = (1.f - mo) * rmean + mo * mean;
            L->running_var[c] = (1.f - mo) * rvar + mo * var * (cnt / (cnt - 1.f));
          }
        } else {
          s_c0[c] = gam * s_rstd[c];
          s_c1[c] = t0 / cnt;
          s_c2[c] = t1 / cnt;
          if (blockIdx.x == 0) {  // the parameter gradients of this BatchNorm
            L->dbeta[c] = t0;
            L->dgamma[c] = t1;
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
      // this warp's private partial rows back to zero for the next layer (the scratch above overlapped them)
      for (int i = lane; i < 128; i += 32) s_part[ew * 128 + i] = 0.f;
      if (cp.dbg != nullptr && tid == 0) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 9] = clock64();
      // ---------------- phase 2: normalise (forward) / BatchNorm backward, from TMEM only
      {
        __nv_bfloat16* outp = L->out;
        __nv_bfloat16* act_out = L->act_out;
        __nv_bfloat16* gy = L->gy;
        __nv_bfloat16* gy_up = L->gy_up;
        __nv_bfloat16* gz = L->gz;
        int k = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += G, ++k) {
          if ((k % NG) != group) continue;
          const int pix = tile * 128 + m;
          const bool valid = pix < npix;
          const size_t base = static_cast<size_t>(valid ? pix : 0) * nc;
          size_t up_base = 0;
          if (epi == EPI_BNBWD && gy_up != nullptr) {
            const int w = pix % Wo, h = (pix / Wo) % Ho, n = pix / (Wo * Ho);
            up_base = ((static_cast<size_t>(n) * (2 * Ho) + 2 * h) * (2 * Wo) + 2 * w) * nc;
          }
          for (int c0 = 0; c0 < nc; c0 += 16) {
            float v[16];
            tmem_ld16(tmem_base + lane_base + k * nc + c0, v);
            if (epi == EPI_BN_TRAIN) {
              {
                uint4 ya, yb;
                pack16_bf16(v, ya, yb);  // the stored (rounded) conv output is what gets normalised
                if (valid) {
                  uint4* op = reinterpret_cast<uint4*>(outp + base + c0);
                  op[0] = ya;
                  op[1] = yb;
                }
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
              if (has_stash) {
                float r[16];
                tmem_ld16(tmem_base + lane_base + stash_col + k * nc + c0, r);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += r[i];
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
              if (valid) store16_bf16(act_out + base + c0, v);
            } else {
              float xh[16], o[16];
              tmem_ld16(tmem_base + lane_base + stash_col + k * nc + c0, xh);
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = s_c0[c0 + i] * (v[i] - s_c1[c0 + i] - xh[i] * s_c2[c0 + i]);
              if (valid) {
                store16_bf16(gy + base + c0, o);
                if (gy_up != nullptr) store16_bf16(gy_up + up_base + c0, o);
                if (gz != nullptr) store16_bf16(gz + base + c0, v);
              }
            }
          }
          // the tile's accumulator slot is free: the MMA warp may start the next layer's tile on it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_relaxed(&accempty[k]);
          if (cp.dbg != nullptr && tid == 0 && k == 0) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 10] = clock64();
        }
        if (cp.dbg != nullptr && tid == 0) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 2] = clock64();
        chain_publish(L->flags, lane);
      }
      if (cp.dbg != nullptr && tid == 0) cp.dbg[(static_cast<size_t>(blockIdx.x) * nl + l) * 48 + 3] = clock64();
      acc_phase ^= (1u << my_tiles) - 1u;
      seq += my_tiles;
      }  // FUSED
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kChainPipes) {
    tc_fence_after();
    tmem_dealloc(tmem_base, cp.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ host helpers
inline size_t chain_smem_bytes(uint32_t w_bytes_max, uint32_t ring_bytes, int ew) {
  return 1024 + static_cast<size_t>(w_bytes_max) + ring_bytes + (2 * kChainMaxStages + 2 + 2 * kChainMaxSlots) * 8 + 16 +
         16 + static_cast<size_t>(2 * ew + 5) * 64 * 4 + 64;
}

}  // namespace wacil
