// test_kd.cu — standalone check + timing sweep of the KD-loss kernels (loss_kernels.cuh) at the BASELINE cfg-5 shapes.
// Usage: bin/test_kd [B C]   (default 1024 8192).  Every variant is checked against a float64 CPU evaluation of
// utils.py:127-132 (loss, and the gradient of a few rows) and timed with CUDA events over rotating buffers
// (working set > L2), min over rounds of back-to-back launches.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "loss_kernels.cuh"

using namespace wacil;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

template <typename T> T to_t(float v);
template <> float to_t<float>(float v) { return v; }
template <> __nv_bfloat16 to_t<__nv_bfloat16>(float v) { return __float2bfloat16(v); }
template <typename T> float to_f(T v);
template <> float to_f<float>(float v) { return v; }
template <> float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

struct Ref {
  double loss;
  std::vector<std::vector<double>> grad_rows;  // rows 0, B/2, B-1
};

template <typename T>
Ref cpu_ref(const std::vector<T>& s, const std::vector<T>& t, int B, int C, double Tm) {
  Ref r;
  r.loss = 0;
  const int rows[3] = {0, B / 2, B - 1};
  r.grad_rows.resize(3);
  std::vector<double> ps(C), pt(C);
  for (int b = 0; b < B; ++b) {
    double ms = -1e300, mt = -1e300;
    for (int c = 0; c < C; ++c) {
      ms = fmax(ms, to_f(s[(size_t)b * C + c]) / Tm);
      mt = fmax(mt, to_f(t[(size_t)b * C + c]) / Tm);
    }
    double zs = 0, zt = 0;
    for (int c = 0; c < C; ++c) {
      ps[c] = exp(to_f(s[(size_t)b * C + c]) / Tm - ms);
      pt[c] = exp(to_f(t[(size_t)b * C + c]) / Tm - mt);
      zs += ps[c];
      zt += pt[c];
    }
    double kl = 0;
    for (int c = 0; c < C; ++c) {
      const double a = ps[c] / zs, q = pt[c] / zt;
      if (q > 0) kl += q * (log(q) - log(a));
    }
    r.loss += kl;
    for (int k = 0; k < 3; ++k)
      if (rows[k] == b) {
        r.grad_rows[k].resize(C);
        for (int c = 0; c < C; ++c) r.grad_rows[k][c] = Tm / B * (ps[c] / zs - pt[c] / zt);
      }
  }
  r.loss *= Tm * Tm / B;
  return r;
}

template <typename T> struct Bufs {
  static constexpr int NB = 6;
  T *s[NB], *t[NB], *g[NB];
  float* loss;
  float* part;
  unsigned* counter;
  int B, C;
};

template <typename T, typename F>
void run_variant(const char* name, Bufs<T>& bf, const Ref& ref, F launch) {
  const int B = bf.B, C = bf.C;
  CK(cudaMemset(bf.g[0], 0xff, sizeof(T) * (size_t)B * C));
  launch(0);
  CK(cudaDeviceSynchronize());
  float loss = 0;
  CK(cudaMemcpy(&loss, bf.loss, 4, cudaMemcpyDeviceToHost));
  const int rows[3] = {0, B / 2, B - 1};
  double gerr = 0, gmax = 0;
  std::vector<T> row(C);
  for (int k = 0; k < 3; ++k) {
    CK(cudaMemcpy(row.data(), bf.g[0] + (size_t)rows[k] * C, sizeof(T) * C, cudaMemcpyDeviceToHost));
    for (int c = 0; c < C; ++c) {
      gerr = fmax(gerr, fabs(to_f(row[c]) - ref.grad_rows[k][c]));
      gmax = fmax(gmax, fabs(ref.grad_rows[k][c]));
    }
  }
  const double lrel = fabs(loss - ref.loss) / fabs(ref.loss);
  const double gtol = sizeof(T) == 4 ? 1e-4 : 1e-2;
  const bool ok = lrel < 1e-5 && gerr <= gtol * gmax;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e9f;
  for (int r = 0; r < 7; ++r) {
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 40; ++i) launch(i % Bufs<T>::NB);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    best = fminf(best, ms * 1000.f / 40);
  }
  const double bytes = 3.0 * B * C * sizeof(T) + 4.0 * B;
  printf("%-44s %7.2f us %6.0f GB/s  loss rel.err %.2e  grad err %.2e (max %.2e)  %s\n", name, best,
         bytes / best / 1e3, lrel, gerr, gmax, ok ? "PASS" : "FAIL");
  fflush(stdout);
}

template <typename T, int THREADS, int TPR, int EPT, int NS>
void stream_variant(Bufs<T>& bf, const Ref& ref, float Tm, bool finish, bool pdl = false) {
  using Cfg = KdStreamCfg<T, THREADS, TPR, EPT, NS>;
  if (TPR * EPT < bf.C) return;
  if (Cfg::kSmemBytes > 227 * 1024) return;
  CK(cudaFuncSetAttribute(kd_stream_kernel<T, THREADS, TPR, EPT, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)Cfg::kSmemBytes));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int grid = (bf.B + Cfg::G - 1) / Cfg::G;
  if (grid > sms) grid = sms;
  char name[128];
  snprintf(name, sizeof name, "stream<%s thr%d tpr%d ept%d ns%d>%s%s grid %d", sizeof(T) == 4 ? "f32" : "bf16", THREADS, TPR,
           EPT, NS, finish ? "" : " nofinish", pdl ? " pdl" : "", grid);
  const float invT = 1.f / Tm, gs = Tm / bf.B, ls = Tm * Tm / bf.B;
  if (!finish) {  // loss is not produced: pre-store the reference so that the check only sees the gradient
    const float l = (float)ref.loss;
    CK(cudaMemcpy(bf.loss, &l, 4, cudaMemcpyHostToDevice));
  }
  run_variant<T>(name, bf, ref, [&](int k) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kd_stream_kernel<T, THREADS, TPR, EPT, NS>, (const T*)bf.s[k], (int64_t)bf.C,
                          (const T*)bf.t[k], (int64_t)bf.C, bf.B, bf.C, invT, gs, ls, bf.loss, bf.g[k], (int64_t)bf.C, 0,
                          bf.part, finish ? bf.counter : (unsigned*)nullptr));
  });
}

template <typename T> void sweep(int B, int C, float Tm) {
  printf("---- %s B=%d C=%d T=%g\n", sizeof(T) == 4 ? "fp32" : "bf16", B, C, Tm);
  std::vector<T> hs((size_t)B * C), ht((size_t)B * C);
  unsigned st = 12345u;
  auto rnd = [&]() {  // sum of 4 uniforms ~ N(0, 1/3) scaled to sigma 2
    float a = 0;
    for (int i = 0; i < 4; ++i) {
      st = st * 1664525u + 1013904223u;
      a += (st >> 8) * (1.f / 16777216.f) - 0.5f;
    }
    return a * 3.4641f;
  };
  for (size_t i = 0; i < hs.size(); ++i) {
    hs[i] = to_t<T>(rnd());
    ht[i] = to_t<T>(rnd());
  }
  Ref ref = cpu_ref(hs, ht, B, C, Tm);
  Bufs<T> bf;
  bf.B = B;
  bf.C = C;
  for (int k = 0; k < Bufs<T>::NB; ++k) {
    CK(cudaMalloc(&bf.s[k], sizeof(T) * hs.size()));
    CK(cudaMalloc(&bf.t[k], sizeof(T) * hs.size()));
    CK(cudaMalloc(&bf.g[k], sizeof(T) * hs.size()));
    CK(cudaMemcpy(bf.s[k], hs.data(), sizeof(T) * hs.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(bf.t[k], ht.data(), sizeof(T) * hs.size(), cudaMemcpyHostToDevice));
  }
  CK(cudaMalloc(&bf.loss, 4));
  CK(cudaMalloc(&bf.part, 4 * (size_t)B));
  CK(cudaMalloc(&bf.counter, 64));
  CK(cudaMemset(bf.counter, 0, 64));
  const float invT = 1.f / Tm, gs = Tm / B, ls = Tm * Tm / B;
  if (C <= 8192 && C > 4096) {
    run_variant<T>("rows<256,32> (log2 domain)", bf, ref, [&](int k) {
      kd_rows_kernel<T, 256, 32, true><<<B, 256>>>(bf.s[k], C, bf.t[k], C, B, C, invT, gs, ls, bf.loss, bf.g[k], C, 0,
                                                   bf.part, bf.counter);
    });
    run_variant<T>("rows<256,32> nofinish", bf, ref, [&](int k) {
      kd_rows_kernel<T, 256, 32, true><<<B, 256>>>(bf.s[k], C, bf.t[k], C, B, C, invT, gs, ls, bf.loss, bf.g[k], C, 0,
                                                   bf.part, nullptr);
    });
  }
  // the configurations the library dispatches (wacil_api.cu: kd_dispatch), with and without PDL / the loss finish
  if (sizeof(T) == 4) {
    if (C <= 1024) { stream_variant<T, 512, 64, 16, 16>(bf, ref, Tm, true); stream_variant<T, 512, 64, 16, 16>(bf, ref, Tm, true, true); }
    else if (C <= 2048) { stream_variant<T, 512, 64, 32, 12>(bf, ref, Tm, true); stream_variant<T, 512, 64, 32, 12>(bf, ref, Tm, true, true); }
    else if (C <= 4096) { stream_variant<T, 512, 128, 32, 6>(bf, ref, Tm, true); stream_variant<T, 512, 128, 32, 6>(bf, ref, Tm, true, true); }
    else {
      stream_variant<T, 512, 256, 32, 3>(bf, ref, Tm, true);
      stream_variant<T, 512, 256, 32, 3>(bf, ref, Tm, true, true);
      stream_variant<T, 512, 256, 32, 3>(bf, ref, Tm, false, true);
    }
  } else {
    if (C <= 1024) { stream_variant<T, 512, 64, 16, 16>(bf, ref, Tm, true); stream_variant<T, 512, 64, 16, 16>(bf, ref, Tm, true, true); }
    else if (C <= 2048) { stream_variant<T, 512, 64, 32, 16>(bf, ref, Tm, true); stream_variant<T, 512, 64, 32, 16>(bf, ref, Tm, true, true); }
    else if (C <= 4096) { stream_variant<T, 512, 128, 32, 12>(bf, ref, Tm, true); stream_variant<T, 512, 128, 32, 12>(bf, ref, Tm, true, true); }
    else {
      stream_variant<T, 512, 256, 32, 6>(bf, ref, Tm, true);
      stream_variant<T, 512, 256, 32, 6>(bf, ref, Tm, true, true);
      stream_variant<T, 512, 256, 32, 6>(bf, ref, Tm, false, true);
    }
  }
  for (int k = 0; k < Bufs<T>::NB; ++k) {
    cudaFree(bf.s[k]);
    cudaFree(bf.t[k]);
    cudaFree(bf.g[k]);
  }
  cudaFree(bf.loss);
  cudaFree(bf.part);
  cudaFree(bf.counter);
}

int main(int argc, char** argv) {
  int B = argc > 2 ? atoi(argv[1]) : 1024, C = argc > 2 ? atoi(argv[2]) : 8192;
  sweep<float>(B, C, 2.f);
  sweep<__nv_bfloat16>(B, C, 2.f);
  return 0;
}
