// tma_probe.cu — per-SM throughput of cp.async.bulk.tensor box loads from an L2-resident activation tensor, by box
// shape (diagnostic).  One CTA per SM; one thread keeps `depth` boxes in flight through a shared-memory ring and
// reports bytes per SM cycle.  The shapes are the ones the conv kernels use ([N][H][W][C] bf16 maps, C*2-byte rows):
//   rows32 : C=16  (32-byte rows,  SWIZZLE_32B), box = 6 image rows x 32 pixels   (stage-1 row-merged box, 6 KB)
//   rows64 : C=32  (64-byte rows,  SWIZZLE_64B), box = 10 rows x 16 pixels        (stage-2 box, 10 KB)
//   rows128: C=64  (128-byte rows, SWIZZLE_128B), box = 128 pixels                (stage-3 tap tile, 16 KB)
//   and the same bytes with other row widths (views of one buffer).   Usage: bin/tma_probe
#include <cstdio>
#include <cstdlib>
#include "tc_conv.cuh"
using namespace wacil;

__global__ void __launch_bounds__(256) tma_probe_kernel(const __grid_constant__ CUtensorMap map, int boxes, int depth,
                                                       uint32_t box_bytes, int hrows, int himgs, int pitch_imgs,
                                                       long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_all[8][8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 64; ++i) mbar_init(&full_all[0][0] + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  int wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int lanes = pitch_imgs > 1 ? pitch_imgs : 0;  // lanes mode: `pitch_imgs` lanes of warp 0 issue
  bool issuer = (threadIdx.x & 31) == 0;
  if (lanes) { issuer = threadIdx.x < lanes; wid = threadIdx.x; nw = lanes; }
  if (issuer) {
    uint64_t* full = full_all[wid];
    tma_prefetch_desc(&map);
    const uint32_t slot = (box_bytes + 1023) & ~1023u;
    smem += wid * depth * slot;
    boxes /= nw;
    {
      const long long t0 = clock64();
      uint32_t phase_bits = 0u;
      int issued = 0, done = 0;
      while (done < boxes) {
        while (issued < boxes && issued - done < depth) {
          const int s = issued % depth;
          mbar_expect_tx(&full[s], box_bytes);
          // a different box every time: image (blockIdx.x * pitch + i) mod himgs, row block i mod hrows
          const int img = (blockIdx.x + wid * 17 + issued / hrows) % himgs;
          const int row = (issued % hrows) * 4;
          tma_load_5d(smem + s * slot, &map, &full[s], 0, 0, row - 1, img, 0);
          ++issued;
        }
        const int s = done % depth;
        mbar_wait(&full[s], (phase_bits >> s) & 1);
        phase_bits ^= 1u << s;
        ++done;
      }
      if (wid == 0) out[blockIdx.x] = clock64() - t0;
    }
  }
}

int main() {
  const int N = 128;  // images
  size_t bytes = (size_t)N * 32 * 32 * 16 * 2;  // 4 MiB: L2 resident
  void* buf;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 0, bytes);
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Shape { const char* name; int H, W, C, box_rows; };
  const Shape shapes[] = {
      {"32x32x16ch  32B rows, box 6 rows x 32 px ( 6 KB)", 32, 32, 16, 6},
      {"32x16x32ch  64B rows, box 6 rows x 16 px ( 6 KB)", 32, 16, 32, 6},
      {"32x8x64ch  128B rows, box 6 rows x  8 px ( 6 KB)", 32, 8, 64, 6},
      {"32x32x16ch  32B rows, box 18 rows x 32 px (18 KB)", 32, 32, 16, 18},
      {"32x8x64ch  128B rows, box 18 rows x  8 px (18 KB)", 32, 8, 64, 18},
  };
  printf("%-52s %6s %6s | %10s %10s\n", "shape (same 4 MiB buffer)", "grid", "depth", "cycles", "B/clk/SM");
  for (const Shape& sh : shapes) {
    CUtensorMap map;
    if (make_act_map(&map, buf, N, sh.H, sh.W, sh.C, 1, 4, 1, sh.box_rows)) { printf("map failed\n"); return 1; }
    const uint32_t box_bytes = (uint32_t)sh.box_rows * sh.W * sh.C * 2;
    for (int nw : {1, 2, 4, -2, -4, -8})
      for (int depth : {2}) {
        const int streams = nw < 0 ? -nw : nw;
        if ((size_t)streams * depth * ((box_bytes + 1023) & ~1023u) > 190 * 1024) continue;
        const int grid = 148, boxes = 64;
        for (int rep = 0; rep < 2; ++rep)
          tma_probe_kernel<<<grid, nw < 0 ? 32 : 32 * nw, 200 * 1024>>>(map, boxes, depth, box_bytes, 7, N, nw < 0 ? -nw : 1, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-52s %6d %6d | %10lld %10.1f  (%d issuing %s)\n", sh.name, grid, depth, mx, (double)boxes * box_bytes / mx, streams,
               nw < 0 ? "lanes of one warp" : "warps");
      }
  }
  return 0;
}
