// mma_probe.cu — how long does a chain of small tcgen05.mma instructions take?  (diagnostic, one CTA per SM)
// One thread issues `count` MMAs (M=128, N=n, K=16, bf16, operands resident in shared memory, SW32/64/128 rows) that
// accumulate round-robin into `chains` independent TMEM accumulators, commits, and waits for completion.
// Prints cycles per MMA for the issue loop alone and up to completion.  Usage: bin/mma_probe
#include <cstdio>
#include <cstdlib>
#include "tc_prims.cuh"
using namespace wacil;

__global__ void __launch_bounds__(128) probe_kernel(int n, int rowbytes, int count, int chains, int walk, int commit_every,
                                                      long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t layout = umma_layout_for_rowbytes(rowbytes);
    const uint64_t desc_hi = umma_smem_desc(0, 16, 8u * rowbytes, layout);
    const uint32_t idesc = umma_idesc_bf16(128, n, 0, 0);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32 * 1024);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int rep = 0; rep < 2; ++rep) {  // rep 0 warms up
      const long long t0 = clock64();
      const uint64_t adesc = desc_hi | (a_addr >> 4), bdesc = desc_hi | (b_addr >> 4);
      // no integer division in the timed loop: the chain index advances by compare-and-reset
      int c = 0, wk = 0, ce = 0;
      uint32_t d = tmem_base;
      uint64_t ad = adesc, bd = bdesc;
      for (int i = 0; i < count; ++i) {
        umma_bf16(d, ad, bd, idesc, i >= chains);
        d += n;
        if (++c == chains) { c = 0; d = tmem_base; }
        if (walk) {  // a different 4 KB activation tile and weight tile per MMA (7 tiles, then wrap)
          ad += 4096 >> 4;
          bd += 1024 >> 4;
          if (++wk == 7) { wk = 0; ad = adesc; bd = bdesc; }
        }
        if (commit_every > 0 && ++ce == commit_every) { ce = 0; umma_commit(&bar2); }
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      if (rep == 1) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 2 * sizeof(long long));
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("%5s %8s %6s %6s %5s %7s | %12s %12s\n", "N", "rowbytes", "count", "chains", "walk", "commit/", "issue cyc/mma", "total cyc/mma");
  const int ns[] = {16, 64};
  for (int rb : {32, 128})
    for (int n : ns)
      for (int walk : {0, 1})
        for (int ce : {0, 9, 3, 1}) {
          const int count = 63, chains = 1;
          probe_kernel<<<148, 128, 64 * 1024>>>(n, rb, count, chains, walk, ce, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[2];
          cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
          printf("%5d %8d %6d %6d %5d %7d | %12.1f %12.1f\n", n, rb, count, chains, walk, ce, (double)h[0] / count, (double)h[1] / count);
        }
  return 0;
}
