// tma_probe2.cu — the chain producer's issue pattern in isolation: one warp, a ring of `depth` stages, three
// boxes per stage on ONE mbarrier.  mode 0: lane 0 issues all three; 1: lanes 0,1,2; 2: rotating lanes (bx).
#include <cstdio>
#include <cstdlib>
#include "tc_conv.cuh"
using namespace wacil;

__global__ void __launch_bounds__(128) probe2(const __grid_constant__ CUtensorMap map, int tiles, int depth, uint32_t box_bytes,
                                             int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
    tma_prefetch_desc(&map);
  }
  __syncthreads();
  const uint32_t stage_bytes = 3 * box_bytes;
  uint32_t phase_bits = 0;
  int bx = 0;
  const long long t0 = clock64();
  for (int it = 0; it < tiles + depth; ++it) {
    if (it >= depth) {  // consume the oldest stage (the warp that filled it)
      const int s = (it - depth) % depth;
      if ((it - depth) % nw == wid) mbar_wait(&full[s], (phase_bits >> s) & 1);
      phase_bits ^= 1u << s;
    }
    if (it < tiles && it % nw == wid) {
      const int s = it % depth;
      const int img = (blockIdx.x + it / 7) % 128, row = (it % 7) * 4;
      if (lane == 0) mbar_expect_tx(&full[s], stage_bytes);
      __syncwarp();
      uint8_t* dst = smem + s * stage_bytes;
      if (mode == 0) {
        if (lane == 0)
          for (int j = 0; j < 3; ++j) tma_load_5d(dst + j * box_bytes, &map, &full[s], 0, j - 1, row - 1, img, 0);
      } else {
        const int j = mode == 1 ? lane : ((lane - bx) & 31);
        if (j < 3) tma_load_5d(dst + j * box_bytes, &map, &full[s], 0, j - 1, row - 1, img, 0);
        bx = (bx + 3) & 31;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
  const int N = 128;
  void* buf;
  cudaMalloc(&buf, (size_t)N * 32 * 32 * 16 * 2);
  cudaMemset(buf, 0, (size_t)N * 32 * 32 * 16 * 2);
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  CUtensorMap map;
  if (make_act_map(&map, buf, N, 32, 32, 16, 1, 4, 1, 6)) { printf("map failed\n"); return 1; }
  const uint32_t box_bytes = 6 * 32 * 32;
  const char* names[3] = {"lane 0 issues 3 boxes", "lanes 0,1,2", "rotating lanes"};
  printf("stage-1 row-merged tile (3 boxes x 6 KB on one mbarrier), 148 CTAs x 1 warp\n");
  for (int nw : {1, 2, 4})
   for (int mode = 0; mode < 3; mode += 2)
    for (int depth : {2, 4, 8}) {
      if (depth < nw) continue;
      const int tiles = 56;
      for (int rep = 0; rep < 2; ++rep) probe2<<<148, 32 * nw, 200 * 1024>>>(map, tiles, depth, box_bytes, mode, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%d producer warps, %-24s depth %d: %8lld cycles, %7.1f cycles/tile, %6.1f B/clk/SM\n", nw, names[mode], depth, mx, (double)mx / tiles,
             (double)tiles * 3 * box_bytes / mx);
    }
  return 0;
}
