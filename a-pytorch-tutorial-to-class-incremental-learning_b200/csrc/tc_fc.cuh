// tc_fc.cuh — the classifier head as tcgen05 tiles (template.py:99-101: logits = cat_i Linear_i(feats); autograd's
// backward of it).  One kernel, fc_gemm_kernel, computes
//
//     out[m][n] = sum_k A(m, k) * B(n, k)  (+ bias[n])          fp32 in, fp32 out, kind::tf32 (10-bit mantissa products,
//                                                                fp32 accumulation in TMEM)
//
// for the three products of the head, all straight from the row-major fp32 tensors the caller owns (no transposes,
// no casts): an operand is either K-major ([MN rows][K contiguous]) or MN-major ([K rows][MN contiguous]) and the
// major-ness goes into the UMMA instruction descriptor.
//
//     forward     logits [B][C] = feats  [B][D] . W[C][D]^T + b   A = feats   (K-major),  B = W     (K-major)
//     dfeats      dfeats [B][D] = dlogits[B][C] . W[C][D]         A = dlogits (K-major),  B = W     (MN-major)
//     dW          dW     [C][D] = dlogits[B][C]^T . feats[B][D]   A = dlogits (MN-major), B = feats (MN-major)
//
// Tiles: 128 rows of M x bn <= 256 columns of N per CTA; a pipeline stage holds kb x 32 floats of K (32 floats = one
// 128-byte swizzled row).  The producer's fixed cost per stage (~700 cycles of mbarrier round trip + TMA issue,
// profiles/r02_chain_experiments.md) sets the tile shape: bn = 256 x one K box when the grid still fills the SMs,
// otherwise narrower tiles with two K boxes per stage (fc_gemm_shape).
//   K-major operand : one TMA box {32 floats of K, rows} -> canonical K-major SW128 (SBO = 1 KB per 8 rows); the four
//                     K = 8 steps of a stage advance the descriptor start by 32 bytes.
//   MN-major operand: one TMA box {32 floats of MN, 32 rows of K} per 32-float chunk of the tile, chunks 4 KB apart
//                     (descriptor LBO).  4-byte MN-major operands have ONE legal shared-memory layout: the 128-byte
//                     swizzle with 32-byte atomicity (descriptor layout type 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B:
//                     32-byte chunks XORed with the row index mod 4), whose atoms are 4 K rows = 512 bytes (SBO);
//                     a K = 8 step is two of them = 1 KB of start address.
// Out-of-range rows/columns/K are zero-filled by TMA, so M, N, K are arbitrary; the row pitches must be multiples of
// 16 bytes (tensor-map rule).  Warp roles as in tc_conv.cuh: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue.
#pragma once
#include "tc_prims.cuh"
#include "tc_conv.cuh"  // get_tmap_encode

namespace wacil {

constexpr int kFcMaxStages = 4;
constexpr int kFcKTile = 32;                      // floats of K per box = one 128-byte row
constexpr int kFcABytes = 128 * kFcKTile * 4;     // 16 KB: the A operand of one K box

struct FcGemmParams {
  int M, N, K;
  int a_mn, b_mn;      // 0 = K-major, 1 = MN-major
  int bn;              // N tile: multiple of 32, <= 256
  int kb;              // K boxes (of 32 floats) per pipeline stage
  int stages;          // <= kFcMaxStages
  float* out;
  long long ldo;       // floats between rows of out
  const float* bias;   // [N] or nullptr
};

// kind::tf32 instruction descriptor: D fp32, A/B tf32 (format 2), major bits, N >> 3, M >> 4.
__host__ __device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 2u << 7;
  d |= 2u << 10;
  d |= static_cast<uint32_t>(a_mn_major & 1) << 15;
  d |= static_cast<uint32_t>(b_mn_major & 1) << 16;
  d |= static_cast<uint32_t>(N >> 3) << 17;
  d |= static_cast<uint32_t>(M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

inline size_t fc_gemm_smem_bytes(int bn, int kb, int stages) {
  return 1024 + static_cast<size_t>(stages) * kb * (kFcABytes + static_cast<size_t>(bn) * kFcKTile * 4) +
         (2 * kFcMaxStages + 1) * 8 + 16;
}
// Tile shape for an M x N x K product on `sms` SMs (see the header comment); every choice keeps the ring at 192 KB.
inline void fc_gemm_shape(int M, int N, int sms, int& bn, int& kb, int& stages) {
  const long long tm = (M + 127) / 128;
  if (tm * ((N + 255) / 256) * 5 >= 4LL * sms) { bn = 256; kb = 1; stages = 4; return; }
  if (tm * ((N + 127) / 128) * 5 >= 4LL * sms) { bn = 128; kb = 2; stages = 3; return; }
  bn = N >= 64 ? 64 : (N + 31) / 32 * 32;
  kb = 2;
  stages = 4;
}

__global__ void __launch_bounds__(192)
    fc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   const FcGemmParams p) {
  extern __shared__ uint8_t fc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fc_smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t b_bytes = static_cast<uint32_t>(p.bn) * kFcKTile * 4;  // the B operand of one K box
  const uint32_t a_stage = static_cast<uint32_t>(p.kb) * kFcABytes;
  const uint32_t stage_bytes = a_stage + static_cast<uint32_t>(p.kb) * b_bytes;
  const int nstages = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + nstages * stage_bytes);
  uint64_t* empty = full + kFcMaxStages;
  uint64_t* accfull = empty + kFcMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * p.bn;
  const int kstage = kFcKTile * p.kb;  // floats of K per stage
  const int nkt = (p.K + kstage - 1) / kstage;
  const uint32_t tmem_cols = next_pow2_cols(static_cast<uint32_t>(p.bn));

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&mapA);
      tma_prefetch_desc(&mapB);
      for (int s = 0; s < kFcMaxStages; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      mbar_init(accfull, 1);
      fence_mbar_init();
    }
  } else if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    const bool leader = elect_one_sync();
    int stage = 0;
    uint32_t phase = 0;
    for (int kt = 0; kt < nkt; ++kt) {
      mbar_wait(&empty[stage], phase ^ 1);
      if (leader) {
        uint8_t* a_s = smem + stage * stage_bytes;
        uint8_t* b_s = a_s + a_stage;
        const int k0 = kt * kstage;
        mbar_expect_tx(&full[stage], stage_bytes);
        // K-major: kb boxes {32 floats of K, rows};  MN-major: one box {32 floats of MN, kb * 32 rows of K} per chunk
        if (!p.a_mn) {
          for (int q = 0; q < p.kb; ++q) tma_load_2d(a_s + q * kFcABytes, &mapA, &full[stage], k0 + q * kFcKTile, m0);
        } else {
          for (int c = 0; c < 4; ++c) tma_load_2d(a_s + c * p.kb * 4096, &mapA, &full[stage], m0 + 32 * c, k0);
        }
        if (!p.b_mn) {
          for (int q = 0; q < p.kb; ++q) tma_load_2d(b_s + q * b_bytes, &mapB, &full[stage], k0 + q * kFcKTile, n0);
        } else {
          for (int c = 0; c < p.bn / 32; ++c) tma_load_2d(b_s + c * p.kb * 4096, &mapB, &full[stage], n0 + 32 * c, k0);
        }
      }
      __syncwarp();
      if (++stage == nstages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp in the loop, elected lane issues)
    const bool leader = elect_one_sync();
    const uint32_t idesc = umma_idesc_tf32(128, p.bn, p.a_mn, p.b_mn);
    // K-major (SW128, type 2): LBO unused (16), SBO = 8 rows x 128 B;  MN-major (SW128 with 32-byte atoms, type 1):
    // LBO = chunk stride (kb x 4 KB), SBO = next 4 K rows (512 B)
    const uint32_t lbo_mn = static_cast<uint32_t>(p.kb) * 4096u;
    const uint64_t a_hi = p.a_mn ? umma_smem_desc(0, lbo_mn, 512, 1) : umma_smem_desc(0, 16, 1024, 2);
    const uint64_t b_hi = p.b_mn ? umma_smem_desc(0, lbo_mn, 512, 1) : umma_smem_desc(0, 16, 1024, 2);
    // a K = 8 step: K-major +32 B inside the box (4 per box, then the next box); MN-major +8 rows = 1 KB throughout
    const uint32_t a_step16 = (p.a_mn ? 1024u : 32u) >> 4, b_step16 = (p.b_mn ? 1024u : 32u) >> 4;
    const uint32_t a_box16 = (p.a_mn ? 4096u : static_cast<uint32_t>(kFcABytes)) >> 4;
    const uint32_t b_box16 = (p.b_mn ? 4096u : b_bytes) >> 4;
    const int kb = p.kb;
    int stage = 0;
    uint32_t phase = 0;
    for (int kt = 0; kt < nkt; ++kt) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
      const uint64_t a0 = a_hi | ((a_addr >> 4) & 0x3FFF);
      const uint64_t b0 = b_hi | (((a_addr + a_stage) >> 4) & 0x3FFF);
      for (int q = 0; q < kb; ++q) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (leader)
            umma_tf32(tmem_base, a0 + q * a_box16 + j * a_step16, b0 + q * b_box16 + j * b_step16, idesc,
                      (kt | q | j) != 0 ? 1u : 0u);
        }
      }
      if (leader) umma_commit(&empty[stage]);
      __syncwarp();
      if (++stage == nstages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (leader) umma_commit(accfull);
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue: TMEM -> registers -> (+ bias) -> global
    const int quarter = warp & 3;
    mbar_wait(accfull, 0);
    tc_fence_after();
    const int m = m0 + quarter * 32 + lane;
    const bool vec = (p.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
    for (int c0 = 0; c0 < p.bn; c0 += 16) {
      const int n = n0 + c0;
      if (n >= p.N) break;  // warp-uniform
      float v[16];
      tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c0, v);
      if (p.bias != nullptr) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (n + i < p.N) v[i] += __ldg(p.bias + n + i);
      }
      if (m < p.M) {
        float* dst = p.out + static_cast<long long>(m) * p.ldo + n;
        if (vec && n + 16 <= p.N) {
          float4* d4 = reinterpret_cast<float4*>(dst);
          d4[0] = make_float4(v[0], v[1], v[2], v[3]);
          d4[1] = make_float4(v[4], v[5], v[6], v[7]);
          d4[2] = make_float4(v[8], v[9], v[10], v[11]);
          d4[3] = make_float4(v[12], v[13], v[14], v[15]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (n + i < p.N) dst[i] = v[i];
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

// db[c] = sum_b dlogits[b][c], fixed order: 32 columns x 8 row groups per block, rows g, g+8, ... then the 8 partials.
__global__ void __launch_bounds__(256) fc_bias_grad_kernel(const float* __restrict__ dlogits, long long ld, int B, int C,
                                                           float* __restrict__ db) {
  __shared__ float s_p[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), g = threadIdx.x >> 5;
  float acc = 0.f;
  if (col < C)
    for (int b = g; b < B; b += 8) acc += __ldg(dlogits + static_cast<long long>(b) * ld + col);
  s_p[g][threadIdx.x & 31] = acc;
  __syncthreads();
  if (g == 0 && col < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += s_p[i][threadIdx.x];
    db[col] = s;
  }
}

// Row-major fp32 matrix [rows][cols] with row pitch `ld` floats; box {32 floats, box_rows}, OOB = 0; 128-byte swizzle
// with 16-byte (K-major operands) or 32-byte (MN-major operands) atomicity.
inline int make_f32_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld,
                        int box_rows, bool atom32) {
  PFN_tmapEncodeTiled enc = get_tmap_encode();
  if (!enc) return -100;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1000 - static_cast<int>(r);
}

}  // namespace wacil
