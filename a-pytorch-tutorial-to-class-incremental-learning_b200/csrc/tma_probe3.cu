// tma_probe3.cu — where do the cycles of a producer iteration go?  One thread: expect_tx, three 5-D box loads, wait for
// the stage issued `depth` iterations ago; clock64 around every instruction, averaged.
#include <cstdio>
#include "tc_conv.cuh"
using namespace wacil;

__global__ void __launch_bounds__(32) probe3(const __grid_constant__ CUtensorMap map, int tiles, int depth, uint32_t box_bytes,
                                             long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
    tma_prefetch_desc(&map);
    const uint32_t stage_bytes = 3 * box_bytes;
    uint32_t phase_bits = 0;
    long long acc[6] = {0, 0, 0, 0, 0, 0};
    for (int it = 0; it < tiles + depth; ++it) {
      const long long a = clock64();
      if (it >= depth) {
        const int s = (it - depth) % depth;
        mbar_wait(&full[s], (phase_bits >> s) & 1);
        phase_bits ^= 1u << s;
      }
      const long long b = clock64();
      acc[0] += b - a;
      if (it < tiles) {
        const int s = it % depth;
        const int img = (blockIdx.x + it / 7) % 128, row = (it % 7) * 4;
        uint8_t* dst = smem + s * stage_bytes;
        mbar_expect_tx(&full[s], stage_bytes);
        const long long c = clock64();
        tma_load_5d(dst, &map, &full[s], 0, -1, row - 1, img, 0);
        const long long d = clock64();
        tma_load_5d(dst + box_bytes, &map, &full[s], 0, 0, row - 1, img, 0);
        const long long e = clock64();
        tma_load_5d(dst + 2 * box_bytes, &map, &full[s], 0, 1, row - 1, img, 0);
        const long long f = clock64();
        acc[1] += c - b; acc[2] += d - c; acc[3] += e - d; acc[4] += f - e;
      }
      acc[5] += clock64() - a;
    }
    for (int i = 0; i < 6; ++i) out[blockIdx.x * 6 + i] = acc[i];
  }
}

int main() {
  const int N = 128;
  void* buf;
  cudaMalloc(&buf, (size_t)N * 32 * 32 * 16 * 2);
  cudaMemset(buf, 0, (size_t)N * 32 * 32 * 16 * 2);
  long long* d;
  cudaMalloc(&d, 148 * 6 * sizeof(long long));
  cudaFuncSetAttribute(probe3, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  CUtensorMap map;
  if (make_act_map(&map, buf, N, 32, 32, 16, 1, 4, 1, 6)) { printf("map failed\n"); return 1; }
  const uint32_t box_bytes = 6 * 32 * 32;
  for (int grid : {148, 1})
    for (int depth : {2, 4}) {
      const int tiles = 56;
      for (int rep = 0; rep < 2; ++rep) probe3<<<grid, 32, 200 * 1024>>>(map, tiles, depth, box_bytes, d);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
      long long h[6];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("grid %3d depth %d, cycles per iteration: wait %6.1f | expect_tx %6.1f | tma#1 %6.1f tma#2 %6.1f tma#3 %6.1f | total %7.1f\n",
             grid, depth, (double)h[0] / tiles, (double)h[1] / tiles, (double)h[2] / tiles, (double)h[3] / tiles, (double)h[4] / tiles,
             (double)h[5] / tiles);
    }
  return 0;
}
