// test_tc.cu — standalone self-test of the tcgen05 conv / wgrad kernels against a double-precision CPU
// loop on the same bf16-rounded inputs.  Build: make -C csrc bin/test_tc ; run on a B200.
// Usage: test_tc [case-index|-1] [time [stages]]   (no argument = verify all cases in one process;
//        "time" skips the CPU check and reports back-to-back launch time, optional A-ring depth override)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <string>
#include "tc_conv.cuh"

using namespace wacil;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

static uint32_t rng_state = 12345u;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f * 2.f - 1.f;
}
static float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

struct Case {
  const char* name;
  int N, H, W, kc, nc, stride, taps;
  int epi;      // conv epilogue to test
  int res_mode; // 0 none, 1 same, 2 downsample-A (only with EPI_INFER)
};

static bool g_timing = false;
static int g_stages = 0;  // override of the A ring depth (experiments)

static int run_conv(const Case& c) {
  const int Ho = c.H / c.stride, Wo = c.W / c.stride;
  const size_t nx = (size_t)c.N * c.H * c.W * c.kc, nw = (size_t)c.taps * c.nc * c.kc;
  const size_t npix = (size_t)c.N * Ho * Wo, ny = npix * c.nc;
  std::vector<float> x(nx), w(nw), bias(c.nc), res;
  for (auto& v : x) v = bf16r(frand());
  for (auto& v : w) v = bf16r(frand() * 0.2f);
  for (auto& v : bias) v = frand() * 0.5f;
  int rH = Ho, rW = Wo, rC = c.nc, rs = 1;
  if (c.res_mode == 2) { rH = 2 * Ho; rW = 2 * Wo; rC = c.nc / 2; rs = 2; }
  if (c.res_mode) {
    res.resize((size_t)c.N * rH * rW * rC);
    for (auto& v : res) v = bf16r(frand());
  }
  std::vector<__nv_bfloat16> xb(nx), wb(nw), rb(res.size());
  for (size_t i = 0; i < nx; ++i) xb[i] = __float2bfloat16_rn(x[i]);
  for (size_t i = 0; i < nw; ++i) wb[i] = __float2bfloat16_rn(w[i]);
  for (size_t i = 0; i < res.size(); ++i) rb[i] = __float2bfloat16_rn(res[i]);

  // CPU reference
  std::vector<double> yref(ny, 0.0);
  for (int n = 0; n < (g_timing ? 0 : c.N); ++n)
    for (int ho = 0; ho < Ho; ++ho)
      for (int wo = 0; wo < Wo; ++wo)
        for (int t = 0; t < c.taps; ++t) {
          const int r = c.taps == 9 ? t / 3 : 1, s = c.taps == 9 ? t % 3 : 1;
          const int hi = ho * c.stride + r - 1, wi = wo * c.stride + s - 1;
          if (hi < 0 || hi >= c.H || wi < 0 || wi >= c.W) continue;
          const float* xp = &x[(((size_t)n * c.H + hi) * c.W + wi) * c.kc];
          for (int o = 0; o < c.nc; ++o) {
            const float* wp = &w[((size_t)t * c.nc + o) * c.kc];
            double acc = 0;
            for (int k = 0; k < c.kc; ++k) acc += (double)xp[k] * wp[k];
            yref[(((size_t)n * Ho + ho) * Wo + wo) * c.nc + o] += acc;
          }
        }
  std::vector<double> sref(2 * c.nc, 0.0);
  if (c.epi == EPI_INFER) {
    for (size_t pix = 0; pix < npix; ++pix) {
      const int wo = pix % Wo, ho = (pix / Wo) % Ho, n = pix / ((size_t)Wo * Ho);
      for (int o = 0; o < c.nc; ++o) {
        double v = yref[pix * c.nc + o] + bias[o];
        if (c.res_mode == 1) v += res[pix * c.nc + o];
        if (c.res_mode == 2 && o < rC) v += res[(((size_t)n * rH + 2 * ho) * rW + 2 * wo) * rC + o];
        yref[pix * c.nc + o] = v > 0 ? v : 0;
      }
    }
  }
  if (c.epi == EPI_STATS)
    for (size_t i = 0; i < ny; ++i) {
      const double r = bf16r((float)yref[i]);
      sref[i % c.nc] += r;
      sref[c.nc + i % c.nc] += r * r;
    }

  __nv_bfloat16 *dx, *dw, *dy, *dres = nullptr;
  float *dstats, *dbias;
  CK(cudaMalloc(&dx, nx * 2));
  CK(cudaMalloc(&dw, nw * 2));
  CK(cudaMalloc(&dy, ny * 2));
  CK(cudaMalloc(&dstats, 296 * 2 * c.nc * 8));  // room for the fused kernels' per-CTA partial rows
  CK(cudaMalloc(&dbias, c.nc * 4));
  CK(cudaMemcpy(dx, xb.data(), nx * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, wb.data(), nw * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, bias.data(), c.nc * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dy, 0xFF, ny * 2));
  CK(cudaMemset(dstats, 0, 296 * 2 * c.nc * 8));
  if (c.res_mode) {
    CK(cudaMalloc(&dres, rb.size() * 2));
    CK(cudaMemcpy(dres, rb.data(), rb.size() * 2, cudaMemcpyHostToDevice));
  }

  ConvGemmParams p{};
  p.kc = c.kc; p.nc = c.nc; p.taps = c.taps; p.stride = c.stride;
  conv_tile_shape(Ho, Wo, p.Ht, p.Nt, p.tiles_h);
  p.num_pixels = (int)npix;
  p.num_tiles = (int)((npix + 127) / 128);
  p.Ho = Ho; p.Wo = Wo;
  {
    const char* ev = getenv("WACIL_RMERGE");
    p.rmerge = (!(ev && ev[0] == '0') && conv_can_rmerge(c.kc, c.nc, c.taps, c.stride, Ho, Wo)) ? 1 : 0;
  }
  if (p.rmerge) conv_pick_rmerge_pipeline(c.kc, c.nc, p.Ht, Wo, p.tps, p.stages);
  else conv_pick_pipeline(c.kc, c.nc, c.taps, p.tps, p.stages);
  if (g_stages > 0 && !p.rmerge) { p.tps = g_stages / 100; p.stages = g_stages % 100; }  // e.g. 904 = 9 taps/stage, 4 stages
  p.epi = c.epi;
  p.acc_stages = 2;
  p.out = dy; p.stats = dstats; p.bias = dbias; p.res = dres; p.res_stride = rs; p.res_c = rC;
  CUtensorMap mapA, mapW;
  int e1 = make_act_map(&mapA, dx, c.N, c.H, c.W, c.kc, c.stride, p.Ht, p.Nt, p.rmerge ? p.Ht + 2 : 0);
  int e2 = make_mat_map(&mapW, dw, (long long)c.taps * c.nc, c.kc, c.nc);
  if (e1 || e2) { printf("  tensor map encode failed: %d %d\n", e1, e2); return 1; }
  int grid = p.num_tiles < 296 ? p.num_tiles : 296;
  __nv_bfloat16* dact = nullptr;
  float* dscratch = nullptr;
  unsigned* dbar = nullptr;
  long long* ddbg = nullptr;
  const bool fused = c.epi == EPI_BN_TRAIN;
  if (g_timing) {
    CK(cudaMalloc(&ddbg, 296 * 16 * sizeof(long long)));
    CK(cudaMemset(ddbg, 0, 296 * 16 * sizeof(long long)));
    p.dbg = ddbg;
  }
  if (fused) {
    CK(cudaMalloc(&dact, ny * 2));
    CK(cudaMalloc(&dscratch, 8 * 128 * 4));
    CK(cudaMemset(dscratch, 0, 8 * 128 * 4));
    CK(cudaMalloc(&dbar, 64));
    CK(cudaMemset(dbar, 0, 64));
    p.acc_stages = (p.num_tiles + grid - 1) / grid;
    p.grid_bar = dbar; p.act_out = dact; p.gamma = dbias; p.beta = dbias; p.running_mean = dscratch;
    p.running_var = dscratch + 128; p.saved = dscratch + 256; p.eps = 1e-5f; p.momentum = 0.1f;
    if (!c.res_mode) p.res = nullptr;
  }
  if (fused) p.acc_stages = (p.num_tiles + 147) / 148;  // upper bound for the shared-memory size
  const size_t smem = conv_smem_bytes(p);
  const int nthreads = kConvThreads;
  CK(cudaFuncSetAttribute((const void*)conv_kernel_for(p.epi), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute((const void*)conv_kernel_for(p.epi), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  {
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)conv_kernel_for(p.epi), kConvThreads, smem));
    if (g_timing) printf("   occupancy %d CTAs/SM at %zu B smem\n", occ, smem);
    if (grid > occ * 148) grid = occ * 148;
    if (fused) { p.acc_stages = (p.num_tiles + grid - 1) / grid; }
  }
  auto launch = [&]() {
    if (fused) {
      void* args[] = {&mapA, &mapW, &p};
      CK(cudaLaunchCooperativeKernel((void*)conv_kernel_for(p.epi), dim3(grid), dim3(nthreads), args, smem, 0));
    } else {
      conv_gemm_kernel<<<grid, kConvThreads, smem>>>(mapA, mapW, p);
    }
  };
  launch();
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  if (g_timing) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 50;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    {
      // per-CTA timeline of the last launch (SM cycles since kernel entry of that CTA), CTA 0 and the slowest CTA
      std::vector<long long> h(296 * 16);
      CK(cudaMemcpy(h.data(), ddbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      int slow = 0;
      for (int b = 0; b < grid; ++b)
        if (h[b * 16 + 10] - h[b * 16] > h[slow * 16 + 10] - h[slow * 16]) slow = b;
      const char* names[11] = {"entry", "setup", "tma-issued", "w-landed", "a-landed", "mma-issued", "acc0-done",
                               "phase1", "gridbar", "phase2", "exit"};
      for (int b : {0, slow}) {
        printf("   cta %3d:", b);
        for (int k = 1; k <= 10; ++k)
          if (h[b * 16 + k]) printf(" %s %lld", names[k], h[b * 16 + k] - h[b * 16]);
        if (h[b * 16 + 13]) printf(" [constants %lld]", h[b * 16 + 13] - h[b * 16]);
        printf("  | producer blocked on free slots %lld, MMA thread blocked on operands %lld  (cycles)\n", h[b * 16 + 11],
               h[b * 16 + 12]);
      }
    }
    printf("%-34s conv  tps %d stages %2d grid %3d smem %6zu : %7.2f us/launch (warm L2, back-to-back)\n", c.name, p.tps, p.stages, grid,
           smem, 1e3 * ms / reps);
    cudaFree(dx); cudaFree(dw); cudaFree(dy); cudaFree(dstats); cudaFree(dbias); if (dres) cudaFree(dres);
    return 0;
  }

  std::vector<__nv_bfloat16> yb(ny);
  std::vector<float> st(2 * c.nc);
  CK(cudaMemcpy(yb.data(), dy, ny * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(st.data(), dstats, 2 * c.nc * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0; size_t bad = 0, firstbad = (size_t)-1;
  for (size_t i = 0; i < ny; ++i) {
    const double g = __bfloat162float(yb[i]), r = yref[i];
    const double err = fabs(g - r), tol = 2e-2 + 1e-2 * fabs(r);
    if (!(err <= tol)) { if (!bad) firstbad = i; ++bad; }
    if (err > maxerr || std::isnan(g)) maxerr = std::isnan(g) ? 1e30 : err;
    if (fabs(r) > maxref) maxref = fabs(r);
  }
  double serr = 0;
  if (c.epi == EPI_STATS)
    for (int i = 0; i < 2 * c.nc; ++i) serr = fmax(serr, fabs(st[i] - sref[i]) / (1.0 + fabs(sref[i])));
  const bool ok = bad == 0 && serr < 2e-3;
  printf("%-34s conv  maxerr %.4g (max|ref| %.3g) bad %zu/%zu stats_relerr %.3g  %s\n", c.name, maxerr, maxref,
         bad, ny, serr, ok ? "PASS" : "FAIL");
  if (fused) { printf("%-34s conv  (fused mode: timing only, verified by the python tests)\n", c.name); return 0; }
  if (bad) {
    size_t i = firstbad;
    printf("   first bad idx %zu (pix %zu ch %zu): got %g ref %g\n", i, i / c.nc, i % c.nc,
           (double)__bfloat162float(yb[i]), yref[i]);
  }
  cudaFree(dx); cudaFree(dw); cudaFree(dy); cudaFree(dstats); cudaFree(dbias); if (dres) cudaFree(dres);
  return ok ? 0 : 1;
}

static int run_wgrad(const Case& c) {
  // X: [N][H][W][cin=kc], dY: [N][Ho][Wo][cout=nc]
  const int Ho = c.H / c.stride, Wo = c.W / c.stride;
  const int cin = c.kc, cout = c.nc;
  const size_t nx = (size_t)c.N * c.H * c.W * cin, npix = (size_t)c.N * Ho * Wo, ny = npix * cout;
  const size_t nw = (size_t)c.taps * cin * cout;
  std::vector<float> x(nx), dy(ny);
  for (auto& v : x) v = bf16r(frand());
  for (auto& v : dy) v = bf16r(frand() * 0.25f);
  std::vector<__nv_bfloat16> xb(nx), yb(ny);
  for (size_t i = 0; i < nx; ++i) xb[i] = __float2bfloat16_rn(x[i]);
  for (size_t i = 0; i < ny; ++i) yb[i] = __float2bfloat16_rn(dy[i]);
  std::vector<double> ref(nw, 0.0);
  for (int n = 0; n < (g_timing ? 0 : c.N); ++n)
    for (int ho = 0; ho < Ho; ++ho)
      for (int wo = 0; wo < Wo; ++wo)
        for (int t = 0; t < c.taps; ++t) {
          const int r = c.taps == 9 ? t / 3 : 1, s = c.taps == 9 ? t % 3 : 1;
          const int hi = ho * c.stride + r - 1, wi = wo * c.stride + s - 1;
          if (hi < 0 || hi >= c.H || wi < 0 || wi >= c.W) continue;
          const float* xp = &x[(((size_t)n * c.H + hi) * c.W + wi) * cin];
          const float* yp = &dy[(((size_t)n * Ho + ho) * Wo + wo) * cout];
          for (int ci = 0; ci < cin; ++ci)
            for (int co = 0; co < cout; ++co) ref[((size_t)t * cin + ci) * cout + co] += (double)xp[ci] * yp[co];
        }
  __nv_bfloat16 *dx, *ddy;
  float *ddw, *dpart;
  WgradReduceDesc* ddesc;
  CK(cudaMalloc(&dx, nx * 2));
  CK(cudaMalloc(&ddy, ny * 2));
  CK(cudaMalloc(&ddw, nw * 4));
  CK(cudaMalloc(&dpart, 296 * nw * 4));
  CK(cudaMalloc(&ddesc, sizeof(WgradReduceDesc)));
  CK(cudaMemcpy(dx, xb.data(), nx * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ddy, yb.data(), ny * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(ddw, 0, nw * 4));
  WgradParams p{};
  p.cin = cin; p.cout = cout; p.taps = c.taps; p.stride = c.stride;
  conv_tile_shape(Ho, Wo, p.Ht, p.Nt, p.tiles_h);
  p.num_tiles = (int)((npix + 127) / 128);
  p.ring = 2;
  p.part = dpart;
  CUtensorMap mapX, mapDY;
  int e1 = make_act_map(&mapX, dx, c.N, c.H, c.W, cin, c.stride, p.Ht, p.Nt);
  int e2 = make_mat_map(&mapDY, ddy, (long long)npix, cout, 128);
  if (e1 || e2) { printf("  tensor map encode failed: %d %d\n", e1, e2); return 1; }
  const size_t smem = wgrad_smem_bytes(p);
  CK(cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int grid = p.num_tiles < 148 ? p.num_tiles : 148;
  if (grid > 3 && !g_timing) grid = (grid + 2) / 3;  // force several tiles per CTA to exercise TMEM accumulation
  if (g_timing) { grid = (p.num_tiles + 1) / 2; if (grid > 296) grid = 296; }
  {
    WgradReduceDesc rd{dpart, ddw, wgrad_active_ctas(p.num_tiles, grid), (int)nw};
    p.gsplit = 1;
    CK(cudaMemcpy(ddesc, &rd, sizeof(rd), cudaMemcpyHostToDevice));
  }
  wgrad_gemm_kernel<<<grid, kConvThreads, smem>>>(mapX, mapDY, p);
  wgrad_reduce_kernel<<<dim3(148, 1), 256>>>(ddesc);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  if (g_timing) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 50;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) wgrad_gemm_kernel<<<grid, kConvThreads, smem>>>(mapX, mapDY, p);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-34s wgrad grid %3d : %7.2f us/launch (warm L2, back-to-back)\n", c.name, grid, 1e3 * ms / reps);
    cudaFree(dx); cudaFree(ddy); cudaFree(ddw); cudaFree(dpart); cudaFree(ddesc);
    return 0;
  }
  std::vector<float> got(nw);
  CK(cudaMemcpy(got.data(), ddw, nw * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0; size_t bad = 0, firstbad = 0;
  for (size_t i = 0; i < nw; ++i) {
    const double err = fabs(got[i] - ref[i]), tol = 1e-2 + 2e-3 * fabs(ref[i]);
    if (!(err <= tol)) { if (!bad) firstbad = i; ++bad; }
    if (err > maxerr) maxerr = err;
    if (std::isnan(got[i])) maxerr = 1e30;
    if (fabs(ref[i]) > maxref) maxref = fabs(ref[i]);
  }
  printf("%-34s wgrad maxerr %.4g (max|ref| %.3g) bad %zu/%zu  %s\n", c.name, maxerr, maxref, bad, nw,
         bad ? "FAIL" : "PASS");
  if (bad) {
    size_t i = firstbad;
    printf("   first bad idx %zu (tap %zu ci %zu co %zu): got %g ref %g\n", i, i / ((size_t)cin * cout),
           (i / cout) % cin, i % cout, (double)got[i], ref[i]);
  }
  cudaFree(dx); cudaFree(ddy); cudaFree(ddw); cudaFree(dpart); cudaFree(ddesc);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  const Case cases[] = {
      {"s1 32x32 16->16 raw", 4, 32, 32, 16, 16, 1, 9, EPI_RAW, 0},
      {"s1 32x32 16->16 stats", 4, 32, 32, 16, 16, 1, 9, EPI_STATS, 0},
      {"s1 16x16 32->32 stats", 6, 16, 16, 32, 32, 1, 9, EPI_STATS, 0},
      {"s1 8x8 64->64 stats", 8, 8, 8, 64, 64, 1, 9, EPI_STATS, 0},
      {"s1 8x8 64->64 odd-N guard", 5, 8, 8, 64, 64, 1, 9, EPI_STATS, 0},
      {"s2 32x32 16->32 stats", 4, 32, 32, 16, 32, 2, 9, EPI_STATS, 0},
      {"s2 16x16 32->64 stats", 6, 16, 16, 32, 64, 2, 9, EPI_STATS, 0},
      {"s1 32x32 32->16 raw (dgrad up)", 3, 32, 32, 32, 16, 1, 9, EPI_RAW, 0},
      {"s1 16x16 64->32 raw (dgrad up)", 4, 16, 16, 64, 32, 1, 9, EPI_RAW, 0},
      {"s1 16x16 32->32 infer+res", 4, 16, 16, 32, 32, 1, 9, EPI_INFER, 1},
      {"s2 32x32 16->32 infer", 4, 32, 32, 16, 32, 2, 9, EPI_INFER, 0},
      {"s1 16x16 32->32 infer+resA", 4, 16, 16, 32, 32, 1, 9, EPI_INFER, 2},
      {"1x1 8x8 64->128 raw", 8, 8, 8, 64, 128, 1, 1, EPI_RAW, 0},
      {"large s1 32x32 16->16 stats N=128", 128, 32, 32, 16, 16, 1, 9, EPI_STATS, 0},
      {"large s1 16x16 32->32 stats N=128", 128, 16, 16, 32, 32, 1, 9, EPI_STATS, 0},
      {"large s1 8x8 64->64 stats N=128", 128, 8, 8, 64, 64, 1, 9, EPI_STATS, 0},
      {"large s1 32x32 16->16 raw N=128", 128, 32, 32, 16, 16, 1, 9, EPI_RAW, 0},
      {"large s1 32x32 16->16 infer N=128", 128, 32, 32, 16, 16, 1, 9, EPI_INFER, 1},
      {"fused s1 32x32 16->16 bn N=128", 128, 32, 32, 16, 16, 1, 9, EPI_BN_TRAIN, 1},
      {"fused s1 16x16 32->32 bn N=128", 128, 16, 16, 32, 32, 1, 9, EPI_BN_TRAIN, 1},
      {"fused s1 8x8 64->64 bn N=128", 128, 8, 8, 64, 64, 1, 9, EPI_BN_TRAIN, 1},
  };
  const Case wcases[] = {
      {"s1 32x32 16x16", 4, 32, 32, 16, 16, 1, 9, 0, 0},
      {"s1 16x16 32x32", 6, 16, 16, 32, 32, 1, 9, 0, 0},
      {"s1 8x8 64x64", 8, 8, 8, 64, 64, 1, 9, 0, 0},
      {"s2 32x32 16x32", 4, 32, 32, 16, 32, 2, 9, 0, 0},
      {"s2 16x16 32x64", 6, 16, 16, 32, 64, 2, 9, 0, 0},
      {"s1 32x32 16x16 N=128", 128, 32, 32, 16, 16, 1, 9, 0, 0},
      {"s1 16x16 32x32 N=128", 128, 16, 16, 32, 32, 1, 9, 0, 0},
      {"s1 8x8 64x64 N=128", 128, 8, 8, 64, 64, 1, 9, 0, 0},
  };
  const int nc = sizeof(cases) / sizeof(cases[0]), nw = sizeof(wcases) / sizeof(wcases[0]);
  int only = argc > 1 ? atoi(argv[1]) : -1;
  if (argc > 2 && std::string(argv[2]) == "time") g_timing = true;
  if (argc > 3) g_stages = atoi(argv[3]);
  int fails = 0;
  for (int i = 0; i < nc + nw; ++i) {
    if (only >= 0 && i != only) continue;
    fails += i < nc ? run_conv(cases[i]) : run_wgrad(wcases[i - nc]);
    fflush(stdout);
  }
  printf("test_tc: %d failing case(s)\n", fails);
  return fails ? 1 : 0;
}
