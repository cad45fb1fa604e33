// occ_probe.cu — prints what limits the residency of conv_gemm_kernel / wgrad_gemm_kernel (diagnostic).
#include <cstdio>
#include "tc_conv.cuh"
using namespace wacil;
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("SM smem %zu, per-block optin %zu, reserved %zu, regs/SM %d, max threads/SM %d\n", pr.sharedMemPerMultiprocessor,
         pr.sharedMemPerBlockOptin, pr.reservedSharedMemPerBlock, pr.regsPerMultiprocessor, pr.maxThreadsPerMultiProcessor);
  const void* ks[2] = {(const void*)conv_gemm_kernel, (const void*)wgrad_gemm_kernel};
  for (int k = 0; k < 2; ++k) {
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, ks[k]);
    printf("kernel %d: regs %d static smem %zu local %zu maxThreads %d maxDyn %d carveout %d\n", k, fa.numRegs,
           fa.sharedSizeBytes, fa.localSizeBytes, fa.maxThreadsPerBlock, fa.maxDynamicSharedSizeBytes, fa.preferredShmemCarveout);
    cudaFuncSetAttribute(ks[k], cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    for (int carve : {-1, 100}) {
      cudaFuncSetAttribute(ks[k], cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      for (size_t smem : {16384, 49152, 65536, 83352, 98304, 113 * 1024}) {
        int occ = -1;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ks[k], kConvThreads, smem);
        printf("  carveout %3d smem %6zu -> occ %d (%s)\n", carve, smem, occ, cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
