// wacil_api.cu — implementation of include/wacil.h: C ABI, launch plans, argument checking.
// Host code only sequences kernels; it never allocates device memory and never synchronises.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/wacil.h"
#include "ew_kernels.cuh"
#include "loss_kernels.cuh"
#include "tc_conv.cuh"
#include "tc_chain.cuh"
#include "tc_fc.cuh"
#include "head_kernels.cuh"

using namespace wacil;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return static_cast<int>(e) > 0 ? static_cast<int>(e) : 1;
}
#define WACIL_LAUNCH_CHECK(what)                           \
  do {                                                     \
    cudaError_t e_ = cudaGetLastError();                   \
    if (e_ != cudaSuccess) return cuda_fail(e_, what);     \
  } while (0)

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WACIL_PDL");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
inline int ew_grid(long long work_items, int threads = 256) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

constexpr float kBnEps = 1e-5f;
constexpr float kBnMomentum = 0.1f;
constexpr int kFeatDim = 64;

// ---------------------------------------------------------------------------------------- layer geometry
int build_layers(int depth, int max_classes, std::vector<LayerDesc>& L, wacil_net_layout_t* lay) {
  if (depth < 8 || (depth - 2) % 6 != 0) return fail(WACIL_E_SHAPE, "depth %d is not 6n+2", depth);
  if (max_classes < 0) return fail(WACIL_E_SHAPE, "max_classes %d", max_classes);
  const int n = (depth - 2) / 6;
  L.clear();
  auto add = [&](int cin, int cout, int stride, int H_in) {
    LayerDesc d{};
    d.cin = cin;
    d.cin_pad = cin < 16 ? 16 : cin;
    d.cout = cout;
    d.stride = stride;
    d.H_in = H_in;
    d.H_out = H_in / stride;
    L.push_back(d);
  };
  add(3, 16, 1, 32);
  int inpl = 16, H = 32;
  const int planes[3] = {16, 32, 64};
  for (int s = 0; s < 3; ++s)
    for (int b = 0; b < n; ++b) {
      const int stride = (s > 0 && b == 0) ? 2 : 1;
      add(inpl, planes[s], stride, H);
      H /= stride;
      add(planes[s], planes[s], 1, H);
      inpl = planes[s];
    }
  long long off = 0, boff = 0, woff = 0, bioff = 0;
  for (auto& d : L) {
    d.w_off = off;
    off += 9LL * d.cin_pad * d.cout;
  }
  for (auto& d : L) {
    d.gamma_off = off;
    off += d.cout;
    d.beta_off = off;
    off += d.cout;
    d.rmean_off = boff;
    boff += d.cout;
    d.rvar_off = boff;
    boff += d.cout;
    d.wf_off = woff;
    woff += 9LL * d.cin_pad * d.cout;
    d.wd_off = woff;
    woff += 9LL * d.cin_pad * d.cout;
    d.bias_off = bioff;
    bioff += d.cout;
  }
  if (lay) {
    lay->depth = depth;
    lay->n_layers = static_cast<int>(L.size());
    lay->feat_dim = kFeatDim;
    off = static_cast<long long>(align_up(off, 4));
    lay->fc_weight_off = off;
    off += static_cast<long long>(max_classes) * kFeatDim;
    lay->fc_bias_off = off;
    off += max_classes;
    lay->param_elems = static_cast<long long>(align_up(off, 4));
    lay->bn_elems = boff;
  }
  return 0;
}

struct Carve {
  size_t off = 0;
  size_t take(size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  }
};

struct Offsets {
  size_t table, wbuf, bias, stats, bstats, saved, a0, dfeats, gbar, sync_f, sync_b, gz[2], tmp[4], wred, ctab, cmaps;
  std::vector<size_t> y, a, gy, gyup, wpart;
  std::vector<size_t> flag_f, flag_b;  // per chain element: offset of its ready flags inside sync_f / sync_b
  size_t sync_f_bytes = 0, sync_b_bytes = 0;
  size_t stat_elems = 0, wbuf_elems = 0, total = 0;
};

constexpr int kSyncCounters = 64;  // grid-barrier counters (one per chain launch) at the head of a sync region

int env_int(const char* name, int dflt, int lo, int hi) {
  const char* e = getenv(name);
  if (!e) return dflt;
  int v = atoi(e);
  return v < lo ? lo : (v > hi ? hi : v);
}

// Grid of the weight-gradient kernel of a layer (pixel tiles are split into contiguous chunks, one per CTA) — also the
// number of partial-sum slots its workspace holds.
int wgrad_grid(int num_tiles) {
  int g = (num_tiles + 1) / 2;
  // at most one CTA per SM: a second resident weight-gradient CTA (75 KB of shared memory each) leaves no room for the
  // CTA of the cooperative data-gradient launch that has to fit on every SM at once (1.128 vs 1.185 ms per step with
  // 296 CTAs, profiles/r02_xp3.log)
  if (g > num_sms()) g = num_sms();
  return g < 1 ? 1 : g;
}

inline int tiles_of(int batch, int H) { return (batch * H * H + 127) / 128; }

void carve_workspace(const std::vector<LayerDesc>& L, int batch, int mode, Offsets& o) {
  Carve c;
  const size_t nl = L.size();
  o.table = c.take(nl * sizeof(LayerDesc));
  size_t wel = 0, sel = 0;
  for (auto& d : L) {
    wel += 2ull * 9 * d.cin_pad * d.cout;
    sel += 2ull * d.cout;
  }
  o.wbuf_elems = wel;
  o.stat_elems = sel;
  o.wbuf = c.take(wel * 2);
  o.bias = c.take(sel * 4);
  // per layer [CTAs][2*cout] partial BatchNorm sums of the fused kernels (forward: sum / sumsq; backward: d-beta / d-gamma)
  o.stats = c.take(sel * 8 * num_sms());   // per layer two buffers of [CTAs][2*cout] words (fused forward)
  o.bstats = c.take(sel * 8 * num_sms());  // the same for the fused backward
  o.saved = c.take(sel * 4);
  o.a0 = c.take(static_cast<size_t>(batch) * 32 * 32 * 16 * 2);
  o.dfeats = c.take(static_cast<size_t>(batch) * kFeatDim * 4);
  o.gbar = c.take(64 + 16 * nl);  // per fused launch {launch counter, check-out counter}: forward layer l at word 2 l, backward at 2 (nl + l)
  // chain tables: one ChainLayer + two tensor maps per chain element (forward: nl; backward: nl - 1)
  o.ctab = c.take(2 * nl * sizeof(ChainLayer));
  o.cmaps = c.take(2 * nl * 2 * sizeof(CUtensorMap));
  // sync regions (zeroed at the start of every forward / backward): barrier counters, then per-CTA ready counters
  o.flag_f.assign(nl, 0);
  o.flag_b.assign(nl, 0);
  size_t ff = kSyncCounters * 4, fb = kSyncCounters * 4;
  for (size_t l = 0; l < nl; ++l) {  // one counter per CTA and chain element
    o.flag_f[l] = ff;
    ff += static_cast<size_t>(num_sms()) * 4;
    if (l >= 1) {
      o.flag_b[l] = fb;
      fb += static_cast<size_t>(num_sms()) * 4;
    }
  }
  o.sync_f_bytes = ff;
  o.sync_f = c.take(ff);
  o.sync_b_bytes = mode == WACIL_NET_TRAIN ? fb : 0;
  o.sync_b = mode == WACIL_NET_TRAIN ? c.take(fb) : 0;
  size_t max_act = 0;
  for (auto& d : L) {
    const size_t in_b = static_cast<size_t>(batch) * d.H_in * d.H_in * d.cin_pad * 2;
    const size_t out_b = static_cast<size_t>(batch) * d.H_out * d.H_out * d.cout * 2;
    if (in_b > max_act) max_act = in_b;
    if (out_b > max_act) max_act = out_b;
  }
  o.y.assign(nl, 0);
  o.a.assign(nl, 0);
  o.gy.assign(nl, 0);
  o.gyup.assign(nl, 0);
  o.wpart.assign(nl, 0);
  o.wred = 0;
  o.gz[0] = o.gz[1] = 0;
  for (int i = 0; i < 4; ++i) o.tmp[i] = 0;
  if (mode == WACIL_NET_TRAIN) {
    o.wred = c.take(nl * sizeof(WgradReduceDesc));
    for (size_t l = 0; l < nl; ++l) {  // weight-gradient partial slots [CTAs][9*cin][cout] fp32
      const int tiles = tiles_of(batch, L[l].H_out);
      o.wpart[l] = c.take(static_cast<size_t>(wgrad_grid(tiles)) * 9 * L[l].cin_pad * L[l].cout * 4);
    }
    for (size_t l = 0; l < nl; ++l) {
      const size_t out_b = static_cast<size_t>(batch) * L[l].H_out * L[l].H_out * L[l].cout * 2;
      o.y[l] = c.take(out_b);
      o.a[l] = c.take(out_b);
      // every layer keeps its own gradient buffer: the data-gradient chain never waits for a weight-gradient kernel
      // (which reads GY_l) to release one
      o.gy[l] = c.take(out_b);
      if (L[l].stride == 2) o.gyup[l] = c.take(out_b * 4);
    }
    o.gz[0] = c.take(max_act);
    o.gz[1] = c.take(max_act);
  } else {
    // four rotating activation buffers: a CTA may run a layer ahead of a distant one (only neighbouring tiles are
    // ordered by the ready flags), the fourth buffer keeps a block input alive one block longer than its last reader
    for (int i = 0; i < 4; ++i) o.tmp[i] = c.take(max_act);
  }
  o.total = c.off;
}

struct ConvLaunch {
  CUtensorMap mapA, mapW;
  ConvGemmParams p;
  int grid;
  size_t smem;
  int dep_up = 0;  // chain element: the input tensor is a zero-upsampled copy of the previous element's output
};

struct WgradLaunch {
  CUtensorMap mapX, mapDY;
  WgradParams p;
  int grid;
  size_t smem;
};

int conv_max_resident(size_t smem);

// ring_budget > 0: the operand ring of a chain launch has this many bytes (the weight tile lives beside it);
// 0: the single-layer kernel's own half-SM budget.
int setup_conv(ConvLaunch& c, const void* in, int N, int H, int W, int kc, int nc, int stride, int taps,
               const void* wts, int epi, size_t ring_budget = 0) {
  ConvGemmParams& p = c.p;
  memset(&p, 0, sizeof(p));
  const int Ho = H / stride, Wo = W / stride;
  p.kc = kc;
  p.nc = nc;
  p.taps = taps;
  p.stride = stride;
  conv_tile_shape(Ho, Wo, p.Ht, p.Nt, p.tiles_h);
  p.num_pixels = N * Ho * Wo;
  p.num_tiles = (p.num_pixels + 127) / 128;
  p.Ho = Ho;
  p.Wo = Wo;
  size_t total = kConvSmemMax;
  if (ring_budget > 0) total = ring_budget + 6144 + (((size_t)taps * nc * kc * 2 + 1023) & ~size_t(1023));
  p.rmerge = conv_can_rmerge(kc, nc, taps, stride, Ho, Wo, total) ? 1 : 0;
  if (p.rmerge) conv_pick_rmerge_pipeline(kc, nc, p.Ht, Wo, p.tps, p.stages, total);
  else conv_pick_pipeline(kc, nc, taps, p.tps, p.stages, total);
  p.epi = epi;
  p.acc_stages = 2;
  p.res_stride = 1;
  p.res_c = nc;
  int e = make_act_map(&c.mapA, in, N, H, W, kc, stride, p.Ht, p.Nt, p.rmerge ? p.Ht + 2 : 0);
  if (e) return fail(WACIL_E_DRIVER, "activation tensor map (N=%d H=%d C=%d s=%d): %d", N, H, kc, stride, e);
  e = make_mat_map(&c.mapW, wts, static_cast<long long>(taps) * nc, kc, nc);
  if (e) return fail(WACIL_E_DRIVER, "weight tensor map (%dx%d): %d", nc, kc, e);
  c.smem = conv_smem_bytes(p);
  // one CTA per SM: oversubscribing the non-cooperative launches 2x was measured slower (1.43 vs 1.36 ms/step)
  const int cap = ring_budget > 0 ? num_sms() : conv_max_resident(c.smem);
  c.grid = p.num_tiles < cap ? p.num_tiles : cap;
  return 0;
}

constexpr size_t kChainSmemMax = 227 * 1024;

int conv_attrs() {
  static bool set = false;
  if (set) return 0;
  cudaError_t e = cudaSuccess;
  for (int epi : {static_cast<int>(EPI_RAW), static_cast<int>(EPI_BN_TRAIN), static_cast<int>(EPI_BNBWD)}) {
    const void* k = reinterpret_cast<const void*>(conv_kernel_for(epi));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kConvSmemWide));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(kConvSmemMax));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
  const void* chains[3] = {reinterpret_cast<const void*>(conv_chain_kernel<4, true>),
                           reinterpret_cast<const void*>(conv_chain_kernel<8, true>),
                           reinterpret_cast<const void*>(conv_chain_kernel<4, false>)};
  for (const void* k : chains) {
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kChainSmemMax));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv/wgrad/chain)");
  set = true;
  return 0;
}

// CTAs of conv_gemm_kernel that fit on the device at once with `smem` bytes of dynamic shared memory.
int conv_max_resident(size_t smem) {
  int occ = 0;
  if (conv_attrs()) return num_sms();
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_gemm_kernel_bn_fwd, kConvThreads, smem) != cudaSuccess || occ < 1)
    occ = 1;
  if (occ > 2) occ = 2;
  return occ * num_sms();
}

int launch_conv(const ConvLaunch& c, cudaStream_t st) {
  if (c.smem > kConvSmemWide) return fail(WACIL_E_SHAPE, "conv configuration needs %zu bytes of shared memory", c.smem);
  { int e0 = conv_attrs(); if (e0) return e0; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(c.grid);
  cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = c.smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (c.p.epi >= EPI_BN_TRAIN) {  // fused BatchNorm modes wait for each other's partial sums: all CTAs must be co-resident
    attrs[na].id = cudaLaunchAttributeCooperative;
    attrs[na].val.cooperative = 1;
    ++na;
  }
  if (c.p.pdl) {  // start under the tail of the previous kernel of this stream (the kernel waits before reading)
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_kernel_for(c.p.epi), c.mapA, c.mapW, c.p);
  if (e != cudaSuccess) return cuda_fail(e, "conv_gemm_kernel");
  return 0;
}

// Switch a prepared single-layer launch to a fused mode: one TMEM accumulator stage per tile of a CTA.
int make_fused(ConvLaunch& c, int epi, unsigned* grid_bar) {
  c.p.epi = epi;
  c.p.acc_stages = (c.p.num_tiles + c.grid - 1) / c.grid;
  if (c.p.acc_stages * c.p.nc > 512)
    return fail(WACIL_E_SHAPE, "batch too large for the fused BatchNorm path: %d tiles x %d channels per CTA > 512 TMEM "
                "columns", c.p.acc_stages, c.p.nc);
  c.p.grid_bar = grid_bar;
  c.p.eps = kBnEps;
  c.p.momentum = kBnMomentum;
  c.smem = conv_smem_bytes(c.p);
  const int cap = conv_max_resident(c.smem);  // cooperative launch: the whole grid must be resident
  if (c.grid > cap) {
    c.grid = cap;
    c.p.acc_stages = (c.p.num_tiles + c.grid - 1) / c.grid;
    c.smem = conv_smem_bytes(c.p);
  }
  if (c.p.acc_stages * c.p.nc > 256)
    return fail(WACIL_E_SHAPE, "batch too large for the fused BatchNorm path (%d tiles x %d channels per CTA)",
                c.p.acc_stages, c.p.nc);
  return 0;
}

int setup_wgrad(WgradLaunch& w, const void* x, const void* dy, float* part, int N, int H, int W, int cin, int cout,
                int stride, int taps) {
  WgradParams& p = w.p;
  memset(&p, 0, sizeof(p));
  const int Ho = H / stride, Wo = W / stride;
  p.cin = cin;
  p.cout = cout;
  p.taps = taps;
  p.stride = stride;
  conv_tile_shape(Ho, Wo, p.Ht, p.Nt, p.tiles_h);
  const long long npix = static_cast<long long>(N) * Ho * Wo;
  p.num_tiles = static_cast<int>((npix + 127) / 128);
  p.ring = 2;
  p.part = part;
  {
    // the 8x8 layers (64 channels): few pixel tiles, five 128-row tap groups -> one CTA per (pixel chunk, group)
    const int tpg = 128 / cin, groups = (taps + tpg - 1) / tpg;
    p.gsplit = (cin >= 64 && groups > 1) ? groups : 1;
  }
  int e = make_act_map(&w.mapX, x, N, H, W, cin, stride, p.Ht, p.Nt);
  if (e) return fail(WACIL_E_DRIVER, "wgrad activation tensor map: %d", e);
  e = make_mat_map(&w.mapDY, dy, npix, cout, 128);
  if (e) return fail(WACIL_E_DRIVER, "wgrad dY tensor map: %d", e);
  w.grid = wgrad_grid(p.num_tiles) * p.gsplit;
  w.smem = wgrad_smem_bytes(p);
  return 0;
}

int launch_wgrad(const WgradLaunch& w, cudaStream_t st) {
  { int e0 = conv_attrs(); if (e0) return e0; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(w.grid);
  cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = w.smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = w.p.pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, wgrad_gemm_kernel, w.mapX, w.mapDY, w.p);
  if (e != cudaSuccess) return cuda_fail(e, "wgrad_gemm_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------- chain launches
// Tuning knobs of the chain launches (host side only; the defaults are the measured best, profiles/r02_chain_experiments.md):
//   WACIL_CHAIN_EW     epilogue warps of the fused (BatchNorm) chains: 4 | 8
//   WACIL_CHAIN_FWD    layers per forward launch (a launch never crosses a resolution change)
//   WACIL_CHAIN_BWD    layers per backward launch (the weight gradients of a launch's layers start when it ends)
//   WACIL_CHAIN_SMEM   shared-memory budget of a chain CTA in KB (<= 227)
int chain_ew() { static int v = env_int("WACIL_CHAIN_EW", 4, 4, 8); return v == 8 ? 8 : 4; }
// WACIL_ENGINE = layer | chain: one cooperative conv_gemm_kernel launch per layer (programmatic dependent launches hide
// the prologues; kernels of the teacher / weight-gradient branches interleave between them), or several layers per
// persistent conv_chain_kernel launch.  Measured on the step graph (profiles/r02_chain_experiments.md): see kDefaultEngine.
enum { ENGINE_LAYER = 0, ENGINE_CHAIN = 1 };
constexpr int kDefaultEngine = ENGINE_LAYER;
int engine_kind() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WACIL_ENGINE");
    v = e == nullptr ? kDefaultEngine : (e[0] == 'c' ? ENGINE_CHAIN : ENGINE_LAYER);
  }
  return v;
}
int chain_fwd_len() { static int v = env_int("WACIL_CHAIN_FWD", 64, 1, 1000); return v; }
int chain_bwd_len() { static int v = env_int("WACIL_CHAIN_BWD", 4, 1, 1000); return v; }
size_t chain_smem_budget() { static int v = env_int("WACIL_CHAIN_SMEM", 113, 48, 227); return static_cast<size_t>(v) * 1024; }

// CTAs of a fused backward launch over `tiles` tiles (per-layer engine): no more than the SMs, and no more than the
// number of tile rounds needs — 256 tiles run as 128 CTAs x 2 rather than 108 x 2 + 40 x 1: the same critical path, 20 SMs
// left to the weight-gradient branch, 20 fewer rows in the statistics exchange.  Measured on the step graph
// (profiles/r02_balanced_grid_ab.log): backward 599 -> 593 us; the same rule on the forward launches costs 2 us.
int balanced_grid(int tiles, bool balance) {
  const int cap = std::min(tiles, num_sms());
  if (!balance || cap <= 0) return cap;
  const int rounds = (tiles + cap - 1) / cap;
  return (tiles + rounds - 1) / rounds;
}

struct ChainPlan {
  ChainParams cp;
  int grid = 0, ew = 4, first = 0, count = 0;
  bool fused = true;  // BatchNorm chain (train forward / data gradients) or inference-only instantiation
  size_t smem = 0;
};

inline uint32_t w_tile_bytes(int kc, int nc) { return static_cast<uint32_t>((9u * nc * kc * 2 + 1023) & ~1023u); }

// Operand-ring bytes a chain whose widest weight tile is w_max gets inside the shared-memory budget.
// `wide`: the launch uses at most half of the SMs (8x8 layers: 64 tiles), so its CTAs and those of the kernel running
// beside it land on different SMs and the half-SM footprint rule does not apply.
size_t chain_ring_bytes(uint32_t w_max, int ew, bool wide) {
  const size_t fixed = chain_smem_bytes(w_max, 0, ew);
  const size_t budget = wide ? std::max(chain_smem_budget(), static_cast<size_t>(200 * 1024)) : chain_smem_budget();
  if (budget <= fixed + 8192) return 8192;
  return (budget - fixed) & ~size_t(2047);  // two pipelines: each half of the ring starts on a 1 KB boundary
}

int launch_chain(const ChainPlan& c, cudaStream_t st) {
  { int e0 = conv_attrs(); if (e0) return e0; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(c.grid);
  cfg.blockDim = dim3(64 * kChainPipes + 32 * c.ew);
  cfg.dynamicSmemBytes = c.smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  attrs[na].id = cudaLaunchAttributeCooperative;  // grid barriers and cross-CTA ready flags: all CTAs co-resident
  attrs[na].val.cooperative = 1;
  ++na;
  if (c.cp.pdl) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  cudaError_t e = !c.fused ? cudaLaunchKernelEx(&cfg, conv_chain_kernel<4, false>, c.cp)
                  : (c.ew == 8 ? cudaLaunchKernelEx(&cfg, conv_chain_kernel<8, true>, c.cp)
                               : cudaLaunchKernelEx(&cfg, conv_chain_kernel<4, true>, c.cp));
  if (e != cudaSuccess) return cuda_fail(e, "conv_chain_kernel");
  return 0;
}

}  // namespace

// ============================================================================================ the engine
struct wacil_net {
  int depth, batch, max_classes, mode;
  std::vector<LayerDesc> L;
  wacil_net_layout_t lay;
  Offsets off;
  float *params, *grads, *bn;
  uint8_t* ws;
  std::vector<ConvLaunch> fprop, dgrad;   // chain elements: forward layer l; backward element of layer l (l >= 1)
  std::vector<WgradLaunch> wgrad;
  std::vector<ChainPlan> fchains, bchains;
  int engine = ENGINE_CHAIN;              // ENGINE_LAYER: every chain plan holds ONE element, launched as conv_gemm_kernel
  bool fwd_done = false;
  cudaStream_t side = nullptr;            // optional second stream: weight gradients run beside the data-gradient chain
  std::vector<cudaEvent_t> ev_chain;      // backward chain i finished (main stream)
  cudaEvent_t ev_join = nullptr;          // all weight gradients reduced
  // optional fused optimiser (wacil_net_set_optimizer): backward applies gradient mean + SGD itself, after the join
  struct {
    bool on = false;
    float* v = nullptr;
    const float* lr_dev = nullptr;
    float mom = 0.f, wd = 0.f;
    int world = 1, rank = 0;
    const float* grads[kMaxRanks] = {};
    unsigned* pads[kMaxRanks] = {};
    unsigned* epoch = nullptr;
  } opt;
  cudaStream_t red = nullptr;             // third branch (owned by the plan): the ordered reductions of the weight-gradient slots,
  std::vector<cudaEvent_t> ev_wg;         // so that a reduction never sits between two weight-gradient launches of the side stream
  const float* last_feats = nullptr;
  bool folded = false;
  int launches = 0;

  template <typename T> T* at(size_t o) const { return reinterpret_cast<T*>(ws + o); }
  __nv_bfloat16* act_in(int l) const { return l == 0 ? at<__nv_bfloat16>(off.a0) : at<__nv_bfloat16>(off.a[l - 1]); }
};

namespace {

// Fill one ChainLayer from a prepared ConvLaunch (geometry + epilogue operands) and its place in the launch.
void fill_chain_layer(ChainLayer& cl, const ConvLaunch& c, const ConvLaunch* prev, int grid, unsigned* flags,
                      const unsigned* dep_flags) {
  memset(&cl, 0, sizeof(cl));
  const ConvGemmParams& p = c.p;
  cl.kc = p.kc; cl.nc = p.nc; cl.stride = p.stride; cl.Ht = p.Ht; cl.Nt = p.Nt; cl.tiles_h = p.tiles_h;
  cl.num_tiles = p.num_tiles; cl.num_pixels = p.num_pixels; cl.Ho = p.Ho; cl.Wo = p.Wo;
  cl.rmerge = p.rmerge; cl.tps = p.tps; cl.stages = p.stages; cl.epi = p.epi;
  cl.flags = flags;
  if (prev != nullptr) {
    const ConvGemmParams& q = prev->p;
    const int nacc = (p.num_tiles + grid - 1) / grid, pacc = (q.num_tiles + grid - 1) / grid;
    cl.tmem_sync = (p.nc != q.nc || nacc != pacc) ? 1 : 0;
    const size_t st = p.rmerge ? (size_t)3 * (p.Ht + 2) * p.Wo * p.kc * 2 : (size_t)p.tps * 128 * p.kc * 2;
    const size_t sq = q.rmerge ? (size_t)3 * (q.Ht + 2) * q.Wo * q.kc * 2 : (size_t)q.tps * 128 * q.kc * 2;
    cl.ring_sync = (st != sq || p.stages != q.stages) ? 1 : 0;
    cl.dep = 1;  // same tile geometry as the previous element (split_chains): CTAs b-1, b, b+1 hold the rows a tile reads
    cl.dep_flags = dep_flags;
  }
  cl.res_stride = p.res_stride; cl.res_c = p.res_c; cl.gres_mode = p.gres_mode;
  cl.eps = kBnEps; cl.momentum = kBnMomentum;
  cl.out = p.out; cl.act_out = p.act_out; cl.res = p.res; cl.bias = p.bias;
  cl.gamma = p.gamma; cl.beta = p.beta; cl.running_mean = p.running_mean; cl.running_var = p.running_var;
  cl.saved = p.saved; cl.stats = p.stats;
  cl.g_res = p.g_res; cl.act_in = p.act_in; cl.y_in = p.y_in; cl.dgamma = p.dgamma; cl.dbeta = p.dbeta;
  cl.gy = p.gy; cl.gy_up = p.gy_up; cl.gz = p.gz;
}

// Group consecutive elements [0, n) into launches: a new launch starts where `starts_group(i)` says the resolution
// changes, or after `max_len` elements.
template <typename F>
std::vector<std::pair<int, int>> split_chains(int n, int max_len, F starts_group) {
  std::vector<std::pair<int, int>> out;
  int first = 0;
  for (int i = 1; i <= n; ++i) {
    if (i == n || starts_group(i) || i - first >= max_len) {
      out.emplace_back(first, i - first);
      first = i;
    }
  }
  return out;
}

}  // namespace

extern "C" {

int wacil_version(void) { return 100; }
const char* wacil_last_error(void) { return g_err.c_str(); }

int wacil_net_layout(int depth, int max_classes, wacil_net_layout_t* out) {
  if (!out) return fail(WACIL_E_NULL, "out is NULL");
  std::vector<LayerDesc> L;
  return build_layers(depth, max_classes, L, out);
}

int wacil_net_layer_layout(int depth, int max_classes, int layer, wacil_layer_layout_t* out) {
  if (!out) return fail(WACIL_E_NULL, "out is NULL");
  std::vector<LayerDesc> L;
  wacil_net_layout_t lay;
  int e = build_layers(depth, max_classes, L, &lay);
  if (e) return e;
  if (layer < 0 || layer >= static_cast<int>(L.size())) return fail(WACIL_E_SHAPE, "layer %d out of range", layer);
  const LayerDesc& d = L[layer];
  out->cin = d.cin; out->cin_pad = d.cin_pad; out->cout = d.cout; out->stride = d.stride;
  out->H_in = d.H_in; out->H_out = d.H_out;
  out->w_off = d.w_off; out->gamma_off = d.gamma_off; out->beta_off = d.beta_off;
  out->rmean_off = d.rmean_off; out->rvar_off = d.rvar_off;
  return 0;
}

size_t wacil_net_workspace_bytes(int depth, int batch, int max_classes, int mode) {
  std::vector<LayerDesc> L;
  wacil_net_layout_t lay;
  if (build_layers(depth, max_classes, L, &lay)) return 0;
  Offsets o;
  carve_workspace(L, batch, mode, o);
  return o.total;
}

int wacil_net_create(wacil_net_t** out, int depth, int batch, int max_classes, int mode, float* params,
                     float* grads, float* bn_running, void* workspace, size_t workspace_bytes) {
  if (!out || !params || !bn_running || !workspace) return fail(WACIL_E_NULL, "NULL argument");
  if (mode != WACIL_NET_TRAIN && mode != WACIL_NET_INFER) return fail(WACIL_E_SHAPE, "mode %d", mode);
  if (mode == WACIL_NET_TRAIN && !grads) return fail(WACIL_E_NULL, "grads is NULL in TRAIN mode");
  if (batch < 1) return fail(WACIL_E_SHAPE, "batch %d", batch);
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023) || (reinterpret_cast<uintptr_t>(params) & 15) ||
      (grads && (reinterpret_cast<uintptr_t>(grads) & 15)))
    return fail(WACIL_E_ALIGN, "workspace must be 1024-byte aligned, arenas 16-byte aligned");
  wacil_net* n = new wacil_net();
  n->depth = depth; n->batch = batch; n->max_classes = max_classes; n->mode = mode;
  int e = build_layers(depth, max_classes, n->L, &n->lay);
  if (e) { delete n; return e; }
  carve_workspace(n->L, batch, mode, n->off);
  if (workspace_bytes < n->off.total) {
    delete n;
    return fail(WACIL_E_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, n->off.total);
  }
  n->params = params; n->grads = grads; n->bn = bn_running; n->ws = static_cast<uint8_t*>(workspace);
  const int nl = static_cast<int>(n->L.size());
  cudaError_t ce = cudaMemcpy(n->ws + n->off.table, n->L.data(), nl * sizeof(LayerDesc), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) { delete n; return cuda_fail(ce, "layer table upload"); }
  __nv_bfloat16* wbuf = n->at<__nv_bfloat16>(n->off.wbuf);
  const int ew_fused = chain_ew();
  const bool pdl = pdl_enabled();
  n->engine = engine_kind();
  const bool per_layer = n->engine == ENGINE_LAYER;
  const int fwd_len = per_layer ? 1 : chain_fwd_len(), bwd_len = per_layer ? 1 : chain_bwd_len();
  ce = cudaMemset(n->ws + n->off.gbar, 0, 64 + 16 * nl);
  // the published statistics rows: a zero word = "not written" (a recycled workspace may hold an older plan's rows)
  if (ce == cudaSuccess) ce = cudaMemset(n->ws + n->off.stats, 0, n->off.stat_elems * 8 * num_sms());
  if (ce == cudaSuccess) ce = cudaMemset(n->ws + n->off.bstats, 0, n->off.stat_elems * 8 * num_sms());
  if (ce != cudaSuccess) { delete n; return cuda_fail(ce, "launch counter / statistics rows clear"); }
  // ENGINE_LAYER: finish a chain element as a stand-alone conv_gemm_kernel launch
  auto finish_layer_launch = [&](ConvLaunch& c, int grid, bool first_of_forward, int counter_slot) -> int {
    c.grid = grid;
    c.p.acc_stages = c.p.epi >= EPI_BN_TRAIN ? (c.p.num_tiles + grid - 1) / grid : 2;
    c.p.grid_bar = n->at<unsigned>(n->off.gbar) + 2 * counter_slot;
    c.p.eps = kBnEps;
    c.p.momentum = kBnMomentum;
    c.p.pdl = pdl ? 1 : 0;
    c.p.w_early = first_of_forward ? 0 : 1;  // the operand copies are older than the upstream kernel, except right after cast_weights
    c.smem = conv_smem_bytes(c.p);
    // a launch on at most half of the SMs (the 8x8 layers: 64 tiles) does not share its SMs with the side branches'
    // CTAs, so the half-SM footprint rule (kConvSmemMax) does not bind it
    const size_t cap = grid * 2 <= num_sms() ? kConvSmemWide : kConvSmemMax;
    while (c.smem > cap && c.p.stages > 2) {  // the ring was sized for the chain kernel's slightly larger budget
      --c.p.stages;
      c.smem = conv_smem_bytes(c.p);
    }
    if (c.smem > cap) return fail(WACIL_E_SHAPE, "layer launch needs %zu bytes of shared memory", c.smem);
    return 0;
  };
  std::vector<size_t> so(nl);
  {
    size_t o = 0;
    for (int l = 0; l < nl; ++l) { so[l] = o; o += 2 * n->L[l].cout; }
  }
  // a launch never spans a resolution change: the stride-2 layer opens the next group (its weight tile and operand
  // ring are sized with that group's)
  auto fwd_group_start = [&](int l) { return n->L[l].stride == 2; };
  n->fprop.resize(nl);
  std::vector<ChainLayer> tab(2 * nl);
  std::vector<CUtensorMap> maps(4 * nl);
  ChainLayer* dtab = n->at<ChainLayer>(n->off.ctab);
  CUtensorMap* dmaps = n->at<CUtensorMap>(n->off.cmaps);
  unsigned* sync_f = n->at<unsigned>(n->off.sync_f);

  // ------------------------------------------------------------------------------------------ forward chains
  const int fwd_ew = mode == WACIL_NET_TRAIN ? ew_fused : 4;
  const auto fsplit = split_chains(nl, fwd_len, fwd_group_start);
  float* bias = n->at<float>(n->off.bias);
  int cur = 0;  // INFER: rotating buffer holding the current block input
  for (size_t ci = 0; ci < fsplit.size(); ++ci) {
    const int first = fsplit[ci].first, count = fsplit[ci].second;
    uint32_t w_max = 0;
    for (int l = first; l < first + count; ++l) w_max = std::max(w_max, w_tile_bytes(n->L[l].cin_pad, n->L[l].cout));
    int group_tiles = 0;
    for (int l = first; l < first + count; ++l) group_tiles = std::max(group_tiles, tiles_of(batch, n->L[l].H_out));
    const size_t ring = chain_ring_bytes(w_max, fwd_ew, group_tiles * 2 <= num_sms());
    int max_tiles = 0;
    for (int l = first; l < first + count; ++l) {
      const LayerDesc& d = n->L[l];
      ConvLaunch& c = n->fprop[l];
      if (mode == WACIL_NET_TRAIN) {
        e = setup_conv(c, n->act_in(l), batch, d.H_in, d.H_in, d.cin_pad, d.cout, d.stride, 9, wbuf + d.wf_off,
                       EPI_BN_TRAIN, ring);
        if (e) { delete n; return e; }
        c.p.out = n->at<__nv_bfloat16>(n->off.y[l]);
        c.p.act_out = n->at<__nv_bfloat16>(n->off.a[l]);
        c.p.stats = n->at<float>(n->off.stats) + 2 * so[l] * num_sms();  // [grid <= SMs][2*cout] tagged partial sums
        c.p.saved = n->at<float>(n->off.saved) + so[l];
        c.p.gamma = n->params + d.gamma_off;
        c.p.beta = n->params + d.beta_off;
        c.p.running_mean = n->bn + d.rmean_off;
        c.p.running_var = n->bn + d.rvar_off;
        c.p.res = nullptr;
        if (l >= 2 && l % 2 == 0) {  // conv_b: add the block input through the shortcut
          const LayerDesc& da = n->L[l - 1];
          c.p.res = n->act_in(l - 1);
          const bool ds = (da.stride == 2 || da.cin != da.cout);
          c.p.res_stride = ds ? 2 : 1;
          c.p.res_c = ds ? da.cin : c.p.nc;
        }
      } else {
        // INFER: block input in tmp[cur], conv_a -> tmp[cur+1], conv_b (+ shortcut from tmp[cur]) -> tmp[cur+2]
        const void* in;
        __nv_bfloat16* outp;
        const __nv_bfloat16* res = nullptr;
        int res_stride = 1, res_c = d.cout;
        if (l == 0) {
          in = n->at<void>(n->off.a0);
          outp = n->at<__nv_bfloat16>(n->off.tmp[cur]);
        } else if (l % 2 == 1) {  // conv_a
          in = n->at<void>(n->off.tmp[cur]);
          outp = n->at<__nv_bfloat16>(n->off.tmp[(cur + 1) % 4]);
        } else {  // conv_b
          in = n->at<void>(n->off.tmp[(cur + 1) % 4]);
          outp = n->at<__nv_bfloat16>(n->off.tmp[(cur + 2) % 4]);
          res = n->at<__nv_bfloat16>(n->off.tmp[cur]);
          const LayerDesc& da = n->L[l - 1];
          if (da.stride == 2 || da.cin != da.cout) { res_stride = 2; res_c = da.cin; }
          cur = (cur + 2) % 4;
        }
        e = setup_conv(c, in, batch, d.H_in, d.H_in, d.cin_pad, d.cout, d.stride, 9, wbuf + d.wf_off, EPI_INFER, ring);
        if (e) { delete n; return e; }
        c.p.out = outp;
        c.p.bias = bias + d.bias_off;
        c.p.res = res;
        c.p.res_stride = res_stride;
        c.p.res_c = res_c;
      }
      max_tiles = std::max(max_tiles, c.p.num_tiles);
    }
    ChainPlan cpn;
    cpn.first = first; cpn.count = count; cpn.ew = fwd_ew;
    cpn.fused = mode == WACIL_NET_TRAIN;
    cpn.grid = balanced_grid(max_tiles, false);
    uint32_t cols = 0;
    for (int l = first; l < first + count; ++l) {
      const ConvLaunch& c = n->fprop[l];
      const int nacc = (c.p.num_tiles + cpn.grid - 1) / cpn.grid;
      uint32_t need = mode == WACIL_NET_TRAIN ? static_cast<uint32_t>(nacc * c.p.nc) * (c.p.res != nullptr ? 2u : 1u) : 128u;
      if (mode == WACIL_NET_TRAIN && nacc > kChainMaxSlots) need = 1024;
      cols = std::max(cols, need);
      const bool last = l == first + count - 1;
      unsigned* flags = last ? nullptr : sync_f + n->off.flag_f[l] / 4;
      const unsigned* dflags = l > first ? sync_f + n->off.flag_f[l - 1] / 4 : nullptr;
      fill_chain_layer(tab[l], c, l > first ? &n->fprop[l - 1] : nullptr, cpn.grid, flags, dflags);
      maps[2 * l] = c.mapA;
      maps[2 * l + 1] = c.mapW;
    }
    if (cols > 256) {
      delete n;
      return fail(WACIL_E_SHAPE, "batch %d too large for the fused BatchNorm path (%u TMEM columns per CTA > 256)", batch, cols);
    }
    cpn.cp.layers = dtab + first;
    cpn.cp.maps = dmaps + 2 * first;
    cpn.cp.n_layers = count;
    cpn.cp.grid_bar = sync_f + (per_layer ? 0 : ci);
    cpn.cp.pdl = pdl ? 1 : 0;
    cpn.cp.w_bytes_max = w_max;
    cpn.cp.ring_bytes = static_cast<uint32_t>(ring);
    cpn.cp.tmem_cols = 32;
    while (cpn.cp.tmem_cols < cols) cpn.cp.tmem_cols <<= 1;
    cpn.cp.dbg = nullptr;
    cpn.smem = chain_smem_bytes(w_max, static_cast<uint32_t>(ring), fwd_ew);
    if (!per_layer && static_cast<int>(ci) >= kSyncCounters) { delete n; return fail(WACIL_E_SHAPE, "too many chain launches"); }
    if (per_layer && (e = finish_layer_launch(n->fprop[first], cpn.grid, mode == WACIL_NET_TRAIN && first == 0, first))) { delete n; return e; }
    n->fchains.push_back(cpn);
  }

  if (mode == WACIL_NET_TRAIN) {
    // ---------------------------------------------------------------------------------------- backward
    n->dgrad.resize(nl);
    n->wgrad.resize(nl);
    std::vector<WgradReduceDesc> wred(nl);
    unsigned* sync_b = n->at<unsigned>(n->off.sync_b);
    for (int l = 0; l < nl; ++l) {
      const LayerDesc& d = n->L[l];
      // wgrad: X = layer input, dY = GY_l (its own buffer)
      e = setup_wgrad(n->wgrad[l], n->act_in(l), n->at<void>(n->off.gy[l]), n->at<float>(n->off.wpart[l]), batch, d.H_in,
                      d.H_in, d.cin_pad, d.cout, d.stride, 9);
      if (e) { delete n; return e; }
      n->wgrad[l].p.pdl = pdl ? 1 : 0;
      wred[l].part = n->at<float>(n->off.wpart[l]);
      wred[l].dst = n->grads + d.w_off;
      wred[l].nparts = wgrad_active_ctas(n->wgrad[l].p.num_tiles, n->wgrad[l].grid / n->wgrad[l].p.gsplit);
      wred[l].elems = 9 * d.cin_pad * d.cout;
      if (d.stride == 2) {
        ce = cudaMemset(n->ws + n->off.gyup[l], 0, static_cast<size_t>(batch) * d.H_in * d.H_in * d.cout * 2);
        if (ce != cudaSuccess) { delete n; return cuda_fail(ce, "gy_up clear"); }
      }
    }
    ce = cudaMemcpy(n->ws + n->off.wred, wred.data(), nl * sizeof(WgradReduceDesc), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { delete n; return cuda_fail(ce, "weight-gradient reduction table upload"); }
    // elements e = 0 .. nl-2 <-> layer l = nl-1-e: dgrad of layer l (stride-1 conv over GY_l or its zero-upsampled copy,
    // flipped [cin][cout] weights) fused with the ReLU mask + BatchNorm backward of layer l-1
    const int ne = nl - 1;
    auto layer_of = [&](int el) { return nl - 1 - el; };
    const auto bsplit = split_chains(ne, bwd_len, [&](int el) { return n->L[layer_of(el)].stride == 2; });
    for (size_t ci = 0; ci < bsplit.size(); ++ci) {
      const int first = bsplit[ci].first, count = bsplit[ci].second;
      uint32_t w_max = 0;
      for (int el = first; el < first + count; ++el) {
        const LayerDesc& d = n->L[layer_of(el)];
        w_max = std::max(w_max, w_tile_bytes(d.cout, d.cin_pad));
      }
      int group_tiles = 0;
      for (int el = first; el < first + count; ++el) group_tiles = std::max(group_tiles, tiles_of(batch, n->L[layer_of(el)].H_in));
      const size_t ring = chain_ring_bytes(w_max, ew_fused, group_tiles * 2 <= num_sms());
      int max_tiles = 0;
      for (int el = first; el < first + count; ++el) {
        const int l = layer_of(el), pl = l - 1;
        const LayerDesc& d = n->L[l];
        const LayerDesc& pd = n->L[pl];
        const void* src = d.stride == 2 ? n->at<void>(n->off.gyup[l]) : n->at<void>(n->off.gy[l]);
        ConvLaunch& g = n->dgrad[l];
        e = setup_conv(g, src, batch, d.H_in, d.H_in, d.cout, d.cin_pad, 1, 9, wbuf + d.wd_off, EPI_BNBWD, ring);
        if (e) { delete n; return e; }
        g.dep_up = d.stride == 2 ? 1 : 0;
        g.p.g_res = nullptr;
        const bool block_output = (pl == 0) || (pl % 2 == 0);
        if (block_output && pl + 2 < nl) {
          const LayerDesc& na = n->L[pl + 1];
          g.p.g_res = n->at<__nv_bfloat16>(n->off.gz[(pl / 2) % 2]);
          g.p.gres_mode = (na.stride == 2 || na.cin != na.cout) ? 2 : 1;
        }
        g.p.act_in = n->at<__nv_bfloat16>(n->off.a[pl]);
        g.p.y_in = n->at<__nv_bfloat16>(n->off.y[pl]);
        g.p.saved = n->at<float>(n->off.saved) + so[pl];
        g.p.gamma = n->params + pd.gamma_off;
        g.p.dgamma = n->grads + pd.gamma_off;
        g.p.dbeta = n->grads + pd.beta_off;
        g.p.stats = n->at<float>(n->off.bstats) + 2 * so[pl] * num_sms();
        g.p.gy = n->at<__nv_bfloat16>(n->off.gy[pl]);
        g.p.gy_up = pd.stride == 2 ? n->at<__nv_bfloat16>(n->off.gyup[pl]) : nullptr;
        g.p.gz = (pl >= 2 && pl % 2 == 0) ? n->at<__nv_bfloat16>(n->off.gz[((pl - 2) / 2) % 2]) : nullptr;
        max_tiles = std::max(max_tiles, g.p.num_tiles);
      }
      ChainPlan cpn;
      cpn.first = first; cpn.count = count; cpn.ew = ew_fused;
      cpn.grid = balanced_grid(max_tiles, per_layer);
      uint32_t cols = 0;
      for (int el = first; el < first + count; ++el) {
        const int l = layer_of(el);
        const ConvLaunch& c = n->dgrad[l];
        const int nacc = (c.p.num_tiles + cpn.grid - 1) / cpn.grid;
        uint32_t need = static_cast<uint32_t>(nacc * c.p.nc) * 2u;
        if (nacc > kChainMaxSlots) need = 1024;
        cols = std::max(cols, need);
        const bool last = el == first + count - 1;
        unsigned* flags = last ? nullptr : sync_b + n->off.flag_b[l] / 4;
        const unsigned* dflags = el > first ? sync_b + n->off.flag_b[l + 1] / 4 : nullptr;
        fill_chain_layer(tab[nl + el], c, el > first ? &n->dgrad[l + 1] : nullptr, cpn.grid, flags, dflags);
        maps[2 * (nl + el)] = c.mapA;
        maps[2 * (nl + el) + 1] = c.mapW;
      }
      if (cols > 256) {
        delete n;
        return fail(WACIL_E_SHAPE, "batch %d too large for the fused BatchNorm path (%u TMEM columns per CTA > 256)", batch, cols);
      }
      cpn.cp.layers = dtab + nl + first;
      cpn.cp.maps = dmaps + 2 * (nl + first);
      cpn.cp.n_layers = count;
      cpn.cp.grid_bar = sync_b + (per_layer ? 0 : ci);
      cpn.cp.pdl = pdl ? 1 : 0;
      cpn.cp.w_bytes_max = w_max;
      cpn.cp.ring_bytes = static_cast<uint32_t>(ring);
      cpn.cp.tmem_cols = 32;
      while (cpn.cp.tmem_cols < cols) cpn.cp.tmem_cols <<= 1;
      cpn.cp.dbg = nullptr;
      cpn.smem = chain_smem_bytes(w_max, static_cast<uint32_t>(ring), ew_fused);
      if (!per_layer && static_cast<int>(ci) >= kSyncCounters) { delete n; return fail(WACIL_E_SHAPE, "too many chain launches"); }
      if (per_layer && (e = finish_layer_launch(n->dgrad[layer_of(first)], cpn.grid, false, nl + layer_of(first)))) { delete n; return e; }
      n->bchains.push_back(cpn);
    }
  }
  ce = cudaMemcpy(dtab, tab.data(), tab.size() * sizeof(ChainLayer), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(dmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) { delete n; return cuda_fail(ce, "chain table upload"); }
  *out = n;
  return 0;
}

void wacil_net_destroy(wacil_net_t* net) {
  if (!net) return;
  for (cudaEvent_t e : net->ev_chain) cudaEventDestroy(e);
  for (cudaEvent_t e : net->ev_wg) cudaEventDestroy(e);
  if (net->ev_join) cudaEventDestroy(net->ev_join);
  if (net->red) cudaStreamDestroy(net->red);
  delete net;
}

int wacil_net_set_side_stream(wacil_net_t* n, void* side_stream) {
  if (!n) return fail(WACIL_E_NULL, "net is NULL");
  n->side = static_cast<cudaStream_t>(side_stream);
  if (n->side != nullptr && n->ev_join == nullptr) {
    n->ev_chain.resize(n->bchains.size() + 1);
    n->ev_wg.resize(n->L.size() / 4 + 2);
    cudaError_t e = cudaEventCreateWithFlags(&n->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&n->red, cudaStreamNonBlocking);
    for (size_t i = 0; i < n->ev_wg.size() && e == cudaSuccess; ++i)
      e = cudaEventCreateWithFlags(&n->ev_wg[i], cudaEventDisableTiming);
    for (size_t i = 0; i < n->ev_chain.size() && e == cudaSuccess; ++i)
      e = cudaEventCreateWithFlags(&n->ev_chain[i], cudaEventDisableTiming);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
  }
  return 0;
}

int wacil_net_last_launches(const wacil_net_t* net) { return net ? net->launches : 0; }

int wacil_net_set_optimizer(wacil_net_t* n, float* momentum_buf, const float* lr_dev, float momentum, float weight_decay,
                            int world, int rank, const void* const* grad_ptrs, void* const* pad_ptrs, void* epoch_ws) {
  if (!n) return fail(WACIL_E_NULL, "net is NULL");
  if (n->mode != WACIL_NET_TRAIN) return fail(WACIL_E_STATE, "the optimiser belongs to a TRAIN plan");
  if (momentum_buf == nullptr) {  // detach
    n->opt.on = false;
    return 0;
  }
  if (!lr_dev) return fail(WACIL_E_NULL, "lr_dev is NULL");
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return fail(WACIL_E_SHAPE, "world=%d rank=%d", world, rank);
  if (reinterpret_cast<uintptr_t>(momentum_buf) & 15) return fail(WACIL_E_ALIGN, "momentum arena must be 16-byte aligned");
  if (world > 1) {
    if (!grad_ptrs || !pad_ptrs || !epoch_ws) return fail(WACIL_E_NULL, "peer pointers / epoch workspace missing");
    for (int r = 0; r < world; ++r) {
      if (!grad_ptrs[r] || !pad_ptrs[r]) return fail(WACIL_E_NULL, "peer pointer %d is NULL", r);
      n->opt.grads[r] = static_cast<const float*>(grad_ptrs[r]);
      n->opt.pads[r] = static_cast<unsigned*>(pad_ptrs[r]);
    }
    if (n->opt.grads[rank] != n->grads) return fail(WACIL_E_STATE, "grad_ptrs[rank] is not this plan's gradient arena");
  }
  n->opt.v = momentum_buf; n->opt.lr_dev = lr_dev; n->opt.mom = momentum; n->opt.wd = weight_decay;
  n->opt.world = world; n->opt.rank = rank; n->opt.epoch = static_cast<unsigned*>(epoch_ws);
  n->opt.on = true;
  return 0;
}

int wacil_net_chain_count(const wacil_net_t* n, int backward) {
  if (!n) return 0;
  return static_cast<int>(backward ? n->bchains.size() : n->fchains.size());
}

int wacil_net_chain_info(const wacil_net_t* n, int backward, int index, int* first_layer, int* n_layers, int* grid,
                         int* epilogue_warps, size_t* smem_bytes) {
  if (!n) return fail(WACIL_E_NULL, "net is NULL");
  const std::vector<ChainPlan>& v = backward ? n->bchains : n->fchains;
  if (index < 0 || index >= static_cast<int>(v.size())) return fail(WACIL_E_SHAPE, "chain %d of %zu", index, v.size());
  const ChainPlan& c = v[index];
  const int nl = static_cast<int>(n->L.size());
  if (first_layer) *first_layer = backward ? nl - 1 - c.first : c.first;  // backward: the layer whose dgrad runs first
  if (n_layers) *n_layers = c.count;
  if (grid) *grid = c.grid;
  if (epilogue_warps) *epilogue_warps = c.ew;
  if (smem_bytes) *smem_bytes = c.smem;
  return 0;
}

int wacil_net_run_chain(wacil_net_t* n, int backward, int index, long long* dbg, void* stream) {
  if (!n) return fail(WACIL_E_NULL, "net is NULL");
  std::vector<ChainPlan>& v = backward ? n->bchains : n->fchains;
  if (index < 0 || index >= static_cast<int>(v.size())) return fail(WACIL_E_SHAPE, "chain %d of %zu", index, v.size());
  if (n->mode == WACIL_NET_INFER && !n->folded) return fail(WACIL_E_STATE, "INFER plan used before wacil_net_fold");
  cudaStream_t st = S(stream);
  cudaError_t ce = backward ? cudaMemsetAsync(n->ws + n->off.sync_b, 0, n->off.sync_b_bytes, st)
                            : cudaMemsetAsync(n->ws + n->off.sync_f, 0, n->off.sync_f_bytes, st);
  if (ce != cudaSuccess) return cuda_fail(ce, "sync clear");
  if (n->engine == ENGINE_LAYER) {  // one conv_gemm_kernel launch; dbg: [grid][16] clock64 stamps (WACIL_STAMP slots)
    ConvLaunch c = backward ? n->dgrad[static_cast<int>(n->L.size()) - 1 - v[index].first] : n->fprop[v[index].first];
    c.p.dbg = dbg;
    c.p.pdl = 0;
    c.p.w_early = 0;
    return launch_conv(c, st);
  }
  ChainPlan c = v[index];
  c.cp.dbg = dbg;
  c.cp.pdl = 0;
  return launch_chain(c, st);
}

int wacil_net_tensor(const wacil_net_t* n, int what, int layer, void** ptr, size_t* bytes) {
  if (!n || !ptr || !bytes) return fail(WACIL_E_NULL, "NULL argument");
  if (n->mode != WACIL_NET_TRAIN) return fail(WACIL_E_STATE, "only TRAIN plans keep their activations");
  const int nl = static_cast<int>(n->L.size());
  if (what != WACIL_T_INPUT && (layer < 0 || layer >= nl)) return fail(WACIL_E_SHAPE, "layer %d", layer);
  if (what == WACIL_T_INPUT) {
    *ptr = n->ws + n->off.a0;
    *bytes = static_cast<size_t>(n->batch) * 32 * 32 * 16 * 2;
    return 0;
  }
  const LayerDesc& d = n->L[layer];
  const size_t act = static_cast<size_t>(n->batch) * d.H_out * d.H_out * d.cout * 2;
  if (what == WACIL_T_CONV_OUT) { *ptr = n->ws + n->off.y[layer]; *bytes = act; return 0; }
  if (what == WACIL_T_ACT) { *ptr = n->ws + n->off.a[layer]; *bytes = act; return 0; }
  if (what == WACIL_T_BN_SAVED) {
    size_t o = 0;
    for (int l = 0; l < layer; ++l) o += 2 * n->L[l].cout;
    *ptr = n->ws + n->off.saved + o * 4;
    *bytes = 2 * d.cout * 4;
    return 0;
  }
  return fail(WACIL_E_SHAPE, "what %d", what);
}

int wacil_net_fold(wacil_net_t* n, void* stream) {
  if (!n) return fail(WACIL_E_NULL, "net is NULL");
  if (n->mode != WACIL_NET_INFER) return fail(WACIL_E_STATE, "fold is for INFER plans");
  dim3 grid(8, static_cast<unsigned>(n->L.size()));
  fold_bn_kernel<<<grid, 256, 0, S(stream)>>>(n->params, n->bn, n->at<__nv_bfloat16>(n->off.wbuf),
                                              n->at<float>(n->off.bias), n->at<LayerDesc>(n->off.table), kBnEps);
  WACIL_LAUNCH_CHECK("fold_bn_kernel");
  n->folded = true;
  return 0;
}

int wacil_net_forward(wacil_net_t* n, const float* x, int C_seen, float* feats, float* logits, void* stream) {
  if (!n || !x || !feats) return fail(WACIL_E_NULL, "NULL argument");
  if (C_seen < 0 || C_seen > n->max_classes) return fail(WACIL_E_SHAPE, "C_seen %d > max_classes %d", C_seen, n->max_classes);
  if (logits == nullptr && C_seen != 0) return fail(WACIL_E_NULL, "logits is NULL but C_seen = %d", C_seen);
  if (n->mode == WACIL_NET_INFER && !n->folded) return fail(WACIL_E_STATE, "INFER plan used before wacil_net_fold");
  cudaStream_t st = S(stream);
  const size_t nl = n->L.size();
  int launches = 0;
  int e;
  cudaError_t ce = cudaSuccess;
  if (n->engine == ENGINE_CHAIN)  // barrier counters + ready flags of the chain launches
    ce = cudaMemsetAsync(n->ws + n->off.sync_f, 0, n->off.sync_f_bytes, st);
  if (ce != cudaSuccess) return cuda_fail(ce, "sync clear");
  const long long npix0 = static_cast<long long>(n->batch) * 32 * 32;
  pack_input_kernel<<<static_cast<int>((npix0 + 255) / 256), 256, 0, st>>>(x, n->at<__nv_bfloat16>(n->off.a0),
                                                                         n->batch, 32 * 32);
  WACIL_LAUNCH_CHECK("pack_input_kernel");
  ++launches;
  if (n->mode == WACIL_NET_TRAIN) {
    dim3 grid(36, static_cast<unsigned>(nl));  // <= 4 elements per thread for the largest (64x64x9) layers
    cast_weights_kernel<<<grid, 256, 0, st>>>(n->params, n->at<__nv_bfloat16>(n->off.wbuf),
                                              n->at<LayerDesc>(n->off.table));
    WACIL_LAUNCH_CHECK("cast_weights_kernel");
    ++launches;
  }
  for (const ChainPlan& c : n->fchains) {
    if ((e = n->engine == ENGINE_LAYER ? launch_conv(n->fprop[c.first], st) : launch_chain(c, st))) return e;
    ++launches;
  }
  const __nv_bfloat16* last = n->mode == WACIL_NET_TRAIN ? n->at<__nv_bfloat16>(n->off.a[nl - 1]) : n->fprop[nl - 1].p.out;
  pool_fc_fwd_kernel<<<n->batch, 64, 0, st>>>(last, 64, kFeatDim, n->params + n->lay.fc_weight_off,
                                              n->params + n->lay.fc_bias_off, C_seen, feats, logits);
  WACIL_LAUNCH_CHECK("pool_fc_fwd_kernel");
  ++launches;
  n->launches = launches;
  n->fwd_done = true;
  n->last_feats = feats;  // the backward reads them (fc weight gradient)
  return 0;
}

int wacil_net_backward(wacil_net_t* n, const float* dlogits, int C_seen, void* stream) {
  if (!n || !dlogits) return fail(WACIL_E_NULL, "NULL argument");
  if (n->mode != WACIL_NET_TRAIN) return fail(WACIL_E_STATE, "backward needs a TRAIN plan");
  if (!n->fwd_done || !n->last_feats) return fail(WACIL_E_STATE, "backward before forward");
  if (C_seen <= 0 || C_seen > n->max_classes) return fail(WACIL_E_SHAPE, "C_seen %d", C_seen);
  cudaStream_t st = S(stream);
  const int nl = static_cast<int>(n->L.size());
  int launches = 0, e;
  cudaError_t ce = cudaMemsetAsync(n->grads, 0, static_cast<size_t>(n->lay.param_elems) * 4, st);
  if (ce != cudaSuccess) return cuda_fail(ce, "grad clear");
  if (n->engine == ENGINE_CHAIN) ce = cudaMemsetAsync(n->ws + n->off.sync_b, 0, n->off.sync_b_bytes, st);
  if (ce != cudaSuccess) return cuda_fail(ce, "sync clear");
  float* dfeats = n->at<float>(n->off.dfeats);
  const float* saved = n->at<float>(n->off.saved);
  std::vector<size_t> soff(nl);
  {
    size_t o = 0;
    for (int l = 0; l < nl; ++l) { soff[l] = o; o += 2 * n->L[l].cout; }
  }
  {
    // The last layer's gradient comes from the pooled features: ONE cooperative launch computes the classifier's data
    // gradient, the ReLU mask and the BatchNorm backward of that layer (head_kernels.cuh); the classifier's weight
    // gradient runs on the weight-gradient branch (below).
    const int l = nl - 1;
    const LayerDesc& d = n->L[l];
    HeadBwdParams hp{};
    hp.dlogits = dlogits;
    hp.W = n->params + n->lay.fc_weight_off;
    hp.B = n->batch;
    hp.C_seen = C_seen;
    hp.HW = d.H_out * d.H_out;
    hp.act = n->at<__nv_bfloat16>(n->off.a[l]);
    hp.y = n->at<__nv_bfloat16>(n->off.y[l]);
    hp.saved = saved + soff[l];
    hp.gamma = n->params + d.gamma_off;
    hp.dgamma = n->grads + d.gamma_off;
    hp.dbeta = n->grads + d.beta_off;
    hp.dfeats = dfeats;
    hp.gy = n->at<__nv_bfloat16>(n->off.gy[l]);
    hp.gz = (l >= 2 && l % 2 == 0) ? n->at<__nv_bfloat16>(n->off.gz[((l - 2) / 2) % 2]) : nullptr;
    hp.table = reinterpret_cast<unsigned*>(n->at<float>(n->off.bstats) + 2 * soff[l] * num_sms());  // this layer's rows
    hp.counter = n->at<unsigned>(n->off.gbar) + 4 * nl;  // past the per-launch counters of the fused conv launches
    hp.num_pixels = static_cast<long long>(n->batch) * hp.HW;
    const int grid = std::min(n->batch, num_sms());
    hp.pix_per_cta = static_cast<int>((hp.num_pixels + grid - 1) / grid);
    if (d.cout != kFeatDim || d.stride != 1 || (hp.pix_per_cta + hp.HW - 1) / hp.HW + 1 > kHeadMaxImgs)
      return fail(WACIL_E_SHAPE, "head backward: cout %d stride %d batch %d", d.cout, d.stride, n->batch);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kHeadThreads);
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeCooperative;  // the CTAs wait for each other's partial sums: all must be resident
    at[0].val.cooperative = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, head_bwd_kernel, hp);
    if (le != cudaSuccess) return cuda_fail(le, "head_bwd_kernel");
    ++launches;
  }
  // GY_{nl-1} is ready.  Every chain launch produces the gradients GY_{l-1} of its layers l; the weight gradient of
  // layer l needs (a_{l-1}, GY_l), so the weight gradients of a launch's OUTPUTS start when the launch ends and run
  // (on the side stream) beside the next launch.  Every layer has its own GY buffer: the chain never waits for them.
  cudaStream_t ws = n->side ? n->side : st;
  const bool par = n->side != nullptr;
#define WACIL_EV(call, what)                                   \
  do {                                                         \
    cudaError_t ce_ = (call);                                  \
    if (ce_ != cudaSuccess) return cuda_fail(ce_, what);       \
  } while (0)
  int pending = 0;  // weight gradients whose partial slots have not been reduced yet: layers l_lo .. l_lo + pending - 1
  size_t nred = 0;  // reductions launched so far
  auto wgrads = [&](int l_hi, int l_lo) -> int {  // weight gradients of layers l_hi .. l_lo (descending) + ordered reduction
    for (int l = l_hi; l >= l_lo; --l) {
      int e2 = launch_wgrad(n->wgrad[l], ws);
      if (e2) return e2;
      ++launches;
    }
    pending += l_hi - l_lo + 1;
    if (pending >= 4 || l_lo == 0) {  // a few layers per reduction launch, on the third branch
      cudaStream_t rs = par && n->red ? n->red : ws;
      if (rs != ws) {
        cudaError_t e_ = cudaEventRecord(n->ev_wg[nred], ws);
        if (e_ == cudaSuccess) e_ = cudaStreamWaitEvent(rs, n->ev_wg[nred], 0);
        if (e_ != cudaSuccess) return cuda_fail(e_, "event record / wait");
        ++nred;
      }
      wgrad_reduce_kernel<<<dim3(148, static_cast<unsigned>(pending)), 256, 0, rs>>>(n->at<WgradReduceDesc>(n->off.wred) + l_lo);
      cudaError_t e_ = cudaGetLastError();
      if (e_ != cudaSuccess) return cuda_fail(e_, "wgrad_reduce_kernel");
      ++launches;
      pending = 0;
    }
    return 0;
  };
  // Fused optimiser: gradient mean over the ranks (NVLink peer reads in rank order) + SGD, one launch after the join.
  // (An early instalment for the 8x8 layers' weights — 80 % of the arena, complete a third of the way into the backward
  // — was measured slower in every form: on the reduction branch, on its own branch, with 8 / 32 / all blocks, at one
  // and at two GPUs: 1.044 vs 1.032 ms and 1.066 vs 1.055 ms, profiles/r02_early_instalment_ab.log.  The step graph
  // is a chain of latency-bound launches; any extra kernel beside it costs more than the 20 us tail it saves.)
  auto optimise = [&](long long off0, long long n0, long long off1, long long n1, cudaStream_t os) -> int {
    if (n0 <= 0 && n1 <= 0) return 0;
    if (n0 <= 0) { off0 = off1; n0 = n1; n1 = 0; }
    if (n->opt.world == 1) {
      for (int sgm = 0; sgm < 2; ++sgm) {
        const long long o = sgm ? off1 : off0, len = sgm ? n1 : n0;
        if (len <= 0) continue;
        sgd_step_kernel<<<ew_grid(len / 4 + 1), 256, 0, os>>>(n->params + o, n->grads + o, n->opt.v + o, len, n->opt.lr_dev,
                                                             n->opt.mom, n->opt.wd, 1.f);
        cudaError_t e_ = cudaGetLastError();
        if (e_ != cudaSuccess) return cuda_fail(e_, "sgd_step_kernel");
        ++launches;
      }
      return 0;
    }
    AllreduceSgdParams q{};
    for (int r = 0; r < n->opt.world; ++r) { q.grads[r] = n->opt.grads[r]; q.pads[r] = n->opt.pads[r]; }
    q.world = n->opt.world; q.rank = n->opt.rank; q.p = n->params; q.v = n->opt.v; q.lr_dev = n->opt.lr_dev;
    q.mom = n->opt.mom; q.wd = n->opt.wd; q.epoch = n->opt.epoch;
    q.seg_off[0] = off0; q.n = n0; q.seg_off[1] = off1; q.seg_n1 = n1;
    int grid = ew_grid((n0 + n1) / 4);
    if (grid > num_sms() * 4) grid = num_sms() * 4;
    allreduce_sgd_kernel<<<grid, 256, 0, os>>>(q);
    cudaError_t e_ = cudaGetLastError();
    if (e_ != cudaSuccess) return cuda_fail(e_, "allreduce_sgd_kernel");
    ++launches;
    return 0;
  };
  if (par) {
    WACIL_EV(cudaEventRecord(n->ev_chain[0], st), "event record");
    WACIL_EV(cudaStreamWaitEvent(ws, n->ev_chain[0], 0), "event wait");
  }
  // the classifier's weight and bias gradients: first on the weight-gradient branch (head_bwd_kernel has done the data
  // gradient)
  fc_bwd_kernel<<<C_seen, 256, 0, ws>>>(dlogits, n->last_feats, n->batch, C_seen, kFeatDim,
                                        n->grads + n->lay.fc_weight_off, n->grads + n->lay.fc_bias_off);
  WACIL_LAUNCH_CHECK("fc_bwd_kernel");
  ++launches;
  if ((e = wgrads(nl - 1, nl - 1))) return e;
  for (size_t ci = 0; ci < n->bchains.size(); ++ci) {
    const ChainPlan& c = n->bchains[ci];
    if ((e = n->engine == ENGINE_LAYER ? launch_conv(n->dgrad[nl - 1 - c.first], st) : launch_chain(c, st))) return e;
    ++launches;
    if (par) {
      WACIL_EV(cudaEventRecord(n->ev_chain[ci + 1], st), "event record");
      WACIL_EV(cudaStreamWaitEvent(ws, n->ev_chain[ci + 1], 0), "event wait");
    }
    // elements first .. first+count-1 are layers l = nl-1-first .. ; they produced GY_{l-1}
    const int l_hi = nl - 1 - c.first - 1, l_lo = nl - 1 - (c.first + c.count - 1) - 1;
    if ((e = wgrads(l_hi, l_lo))) return e;
  }
  if (par) {  // join: the reductions are in order on their branch and each of them follows its weight gradients
    WACIL_EV(cudaEventRecord(n->ev_join, n->red ? n->red : ws), "event record");
    WACIL_EV(cudaStreamWaitEvent(st, n->ev_join, 0), "event wait");
  }
#undef WACIL_EV
  if (n->opt.on) {
    if ((e = optimise(0, n->lay.param_elems, 0, 0, st))) return e;
  }
  n->launches = launches;
  return 0;
}

// ============================================================================================ single ops
size_t wacil_kd_workspace_bytes(int B) { return 64 + static_cast<size_t>(B > 0 ? B : 0) * 3 * 4; }

extern "C++" {
namespace {
template <typename T>
int kd_dispatch(const void* s, int64_t s_ld, const void* t, int64_t t_ld, int B, int C, float T_, float scale,
                float* loss_out, void* g, int64_t g_ld, int accumulate, void* ws, cudaStream_t st) {
  const T* sp = static_cast<const T*>(s);
  const T* tp = static_cast<const T*>(t);
  T* gp = static_cast<T*>(g);
  unsigned* counter = static_cast<unsigned*>(ws);
  float* part = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + 64);
  const float invT = 1.f / T_;
  const float gscale = scale * T_ / static_cast<float>(B);
  const float lscale = scale * T_ * T_ / static_cast<float>(B);
  const size_t vb = 4 * sizeof(T);
  const bool vec = (C % 4 == 0) && (s_ld % 4 == 0) && (t_ld % 4 == 0) && (!g || g_ld % 4 == 0) &&
                   (reinterpret_cast<uintptr_t>(s) % vb == 0) && (reinterpret_cast<uintptr_t>(t) % vb == 0) &&
                   (!g || reinterpret_cast<uintptr_t>(g) % vb == 0);
  // 16-byte aligned rows wider than 512 columns: the persistent TMA-streamed kernel (WACIL_KD_STREAM=0 selects the
  // register-row kernels).  Configurations from the sweep in profiles/r01_kd_stream_sweep.log: as many row groups per
  // CTA as the row width allows (32 elements per thread), the ring as deep as ~190 KB of shared memory permits.
  {
    static int stream_env = -2;
    if (stream_env == -2) { const char* ev = getenv("WACIL_KD_STREAM"); stream_env = ev ? (ev[0] == '0' ? 0 : 1) : -1; }
    const size_t es = sizeof(T);
    const bool al16 = (C * es) % 16 == 0 && (s_ld * es) % 16 == 0 && (t_ld * es) % 16 == 0 && (!g || (g_ld * es) % 16 == 0) &&
                      reinterpret_cast<uintptr_t>(s) % 16 == 0 && reinterpret_cast<uintptr_t>(t) % 16 == 0 &&
                      (!g || reinterpret_cast<uintptr_t>(g) % 16 == 0);
    if (stream_env != 0 && al16 && C > 512 && C <= 8192) {
      int e = 0;
#define KD_STREAM(TPR, EPT, NS)                                                                                     \
  do {                                                                                                              \
    using Cfg = KdStreamCfg<T, 512, TPR, EPT, NS>;                                                                  \
    static_assert(Cfg::kSmemBytes <= 227 * 1024, "kd_stream_kernel ring exceeds the shared memory of an SM");       \
    static bool attr_set = false;                                                                                   \
    if (!attr_set) {                                                                                                \
      cudaError_t ce = cudaFuncSetAttribute(kd_stream_kernel<T, 512, TPR, EPT, NS>,                                 \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,                            \
                                            static_cast<int>(Cfg::kSmemBytes));                                     \
      if (ce != cudaSuccess) { e = cuda_fail(ce, "kd_stream_kernel smem attribute"); break; }                       \
      attr_set = true;                                                                                              \
    }                                                                                                               \
    int grid = (B + Cfg::G - 1) / Cfg::G;                                                                           \
    if (grid > num_sms()) grid = num_sms();                                                                         \
    cudaLaunchConfig_t cfg{};                                                                                       \
    cfg.gridDim = dim3(grid);                                                                                       \
    cfg.blockDim = dim3(512);                                                                                       \
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;                                                                         \
    cfg.stream = st;                                                                                                \
    cudaLaunchAttribute at[1];                                                                                      \
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                  \
    at[0].val.programmaticStreamSerializationAllowed = 1;                                                           \
    cfg.attrs = at;                                                                                                 \
    cfg.numAttrs = pdl_enabled() ? 1 : 0; /* setup runs under the previous kernel's tail; the kernel waits before reading */ \
    cudaError_t le = cudaLaunchKernelEx(&cfg, kd_stream_kernel<T, 512, TPR, EPT, NS>, sp, s_ld, tp, t_ld, B, C, invT, \
                                        gscale, lscale, loss_out, gp, g_ld, accumulate, part, counter);             \
    if (le != cudaSuccess) e = cuda_fail(le, "kd_stream_kernel");                                                   \
  } while (0)
      constexpr bool kF32 = sizeof(T) == 4;
      if (C <= 1024) KD_STREAM(64, 16, 16);
      else if (C <= 2048) KD_STREAM(64, 32, (kF32 ? 12 : 16));
      else if (C <= 4096) KD_STREAM(128, 32, (kF32 ? 6 : 12));
      else KD_STREAM(256, 32, (kF32 ? 3 : 6));
#undef KD_STREAM
      if (e) return e;
      return 0;
    }
  }
#define KD_LAUNCH(TPR, CACHE, VEC)                                                                              \
  kd_rows_kernel<T, TPR, CACHE, VEC>                                                                            \
      <<<(TPR >= 256 ? B : (B + (256 / TPR) - 1) / (256 / TPR)), (TPR >= 256 ? TPR : 256), 0, st>>>(            \
          sp, s_ld, tp, t_ld, B, C, invT, gscale, lscale, loss_out, gp, g_ld, accumulate, part, counter)
  if (C <= 128) {
    if (vec) KD_LAUNCH(32, 4, true); else KD_LAUNCH(32, 4, false);
  } else if (C <= 1024) {
    if (vec) KD_LAUNCH(32, 32, true); else KD_LAUNCH(32, 32, false);
  } else if (C <= 2048) {
    if (vec) KD_LAUNCH(256, 8, true); else KD_LAUNCH(256, 8, false);
  } else if (C <= 4096) {
    if (vec) KD_LAUNCH(512, 8, true); else KD_LAUNCH(512, 8, false);
  } else if (C <= 8192) {
    // unaligned / strided 4K..8K-column rows (the aligned ones take the streamed kernel above)
    if (vec) KD_LAUNCH(256, 32, true); else KD_LAUNCH(256, 32, false);  // 1024 x 8 measured 2x slower
  } else {
    return fail(WACIL_E_SHAPE, "C = %d > 8192 columns is not supported", C);
  }
#undef KD_LAUNCH
  WACIL_LAUNCH_CHECK("kd_rows_kernel");
  return 0;
}
}  // namespace
}  // extern "C++"

int wacil_kd_fwd_bwd(const void* s, int64_t s_ld, const void* t, int64_t t_ld, int B, int C, float T, float scale,
                     float* loss_out, void* grad_s, int64_t g_ld, int accumulate, int dtype, void* workspace,
                     void* stream) {
  if (!s || !t || !loss_out || !workspace) return fail(WACIL_E_NULL, "NULL argument");
  if (B <= 0 || C <= 0 || s_ld < C || t_ld < C || (grad_s && g_ld < C)) return fail(WACIL_E_SHAPE, "B=%d C=%d", B, C);
  if (!(T > 0.f)) return fail(WACIL_E_SHAPE, "temperature must be positive");
  if (dtype == WACIL_F32)
    return kd_dispatch<float>(s, s_ld, t, t_ld, B, C, T, scale, loss_out, grad_s, g_ld, accumulate, workspace, S(stream));
  if (dtype == WACIL_BF16)
    return kd_dispatch<__nv_bfloat16>(s, s_ld, t, t_ld, B, C, T, scale, loss_out, grad_s, g_ld, accumulate, workspace,
                                      S(stream));
  return fail(WACIL_E_DTYPE, "dtype %d", dtype);
}

int wacil_ce_kd_fwd_bwd(const float* logits, int64_t l_ld, const int64_t* targets, const float* t_logits,
                        int64_t t_ld, int B, int C_seen, int C_known, float T, float lambda_kd, float smooth,
                        float* losses_out, float* dlogits, int64_t d_ld, void* workspace, void* stream) {
  return wacil_ce_kd_fwd_bwd_w(logits, l_ld, targets, t_logits, t_ld, B, C_seen, C_known, T, 1.f, lambda_kd, smooth,
                               losses_out, dlogits, d_ld, workspace, stream);
}

int wacil_ce_kd_fwd_bwd_w(const float* logits, int64_t l_ld, const int64_t* targets, const float* t_logits,
                          int64_t t_ld, int B, int C_seen, int C_known, float T, float ce_weight, float lambda_kd,
                          float smooth, float* losses_out, float* dlogits, int64_t d_ld, void* workspace, void* stream) {
  if (!logits || !targets || !losses_out || !workspace) return fail(WACIL_E_NULL, "NULL argument");
  if (B <= 0 || C_seen <= 0 || l_ld < C_seen || (dlogits && d_ld < C_seen)) return fail(WACIL_E_SHAPE, "B=%d C=%d", B, C_seen);
  if (t_logits && (C_known <= 0 || C_known > C_seen || t_ld < C_known)) return fail(WACIL_E_SHAPE, "C_known=%d", C_known);
  if (!(T > 0.f)) return fail(WACIL_E_SHAPE, "temperature must be positive");
  if (!t_logits) C_known = 0;
  unsigned* counter = static_cast<unsigned*>(workspace);
  float* part = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + 64);
  const float invT = 1.f / T;
  const float gsc = lambda_kd * T / static_cast<float>(B);
  const float lsc = lambda_kd * T * T / static_cast<float>(B);
  cudaStream_t st = S(stream);
#define CK_LAUNCH(TPR, CACHE)                                                                                 \
  ce_kd_rows_kernel<TPR, CACHE><<<(B + (256 / TPR) - 1) / (256 / TPR), 256, 0, st>>>(                         \
      logits, l_ld, targets, t_logits, t_ld, B, C_seen, C_known, invT, gsc, lsc, smooth, ce_weight, losses_out,  \
      dlogits, d_ld, part, counter)
  if (C_seen <= 128) CK_LAUNCH(32, 4);
  else if (C_seen <= 1024) CK_LAUNCH(32, 32);
  else if (C_seen <= 8192) CK_LAUNCH(256, 32);
  else return fail(WACIL_E_SHAPE, "C_seen = %d > 8192", C_seen);
#undef CK_LAUNCH
  WACIL_LAUNCH_CHECK("ce_kd_rows_kernel");
  return 0;
}

size_t wacil_weight_align_workspace_bytes(int C_total) {
  return 64 + (static_cast<size_t>(C_total > 0 ? C_total : 0) + 2 * 1024) * 4;  // barrier counters + [CTAs][2] partial sums
}

extern "C++" {
namespace {
template <typename T>
int wa_launch(void* W, int C_total, int D, int n_new, float* gamma_out, void* ws, cudaStream_t st) {
  static bool attr_set = false;
  constexpr size_t kParkMax = 200 * 1024;  // shared memory for the parked new rows
  cudaError_t e = cudaSuccess;
  if (!attr_set) {
    e = cudaFuncSetAttribute(weight_align_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kParkMax));
    if (e != cudaSuccess) return cuda_fail(e, "weight_align_kernel smem attribute");
    attr_set = true;
  }
  int grid = (C_total + 31) / 32;  // a warp per row: no more CTAs than 32-row groups
  if (grid > num_sms()) grid = num_sms();
  if (grid > 1024) grid = 1024;
  const size_t rowbytes = static_cast<size_t>(D) * sizeof(T);
  int park_rows = (n_new + grid - 1) / grid;
  if (static_cast<size_t>(park_rows) * rowbytes > kParkMax) park_rows = static_cast<int>(kParkMax / rowbytes);
  if ((rowbytes & 15) != 0) park_rows = 0;  // rows must stay 16-byte aligned in shared memory
  const size_t smem = static_cast<size_t>(park_rows) * rowbytes;
  T* Wp = static_cast<T*>(W);
  unsigned* bar = static_cast<unsigned*>(ws);
  float* partial = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + 64);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kWaThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeCooperative;  // grid barrier between the norms and the scaling: all CTAs co-resident
  attrs[0].val.cooperative = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // prologue under the tail of the previous kernel
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  e = cudaLaunchKernelEx(&cfg, weight_align_kernel<T>, Wp, C_total, D, n_new, partial, bar, gamma_out, park_rows);
  if (e != cudaSuccess) return cuda_fail(e, "weight_align_kernel");
  return 0;
}
}  // namespace
}  // extern "C++"

int wacil_weight_align(void* W, int C_total, int D, int n_new, float* gamma_out, int dtype, void* workspace,
                       void* stream) {
  if (!W || !gamma_out || !workspace) return fail(WACIL_E_NULL, "NULL argument");
  if (C_total <= 1 || D <= 0 || n_new <= 0 || n_new >= C_total)
    return fail(WACIL_E_SHAPE, "C_total=%d D=%d n_new=%d (need 0 < n_new < C_total)", C_total, D, n_new);
  if (reinterpret_cast<uintptr_t>(W) & 15) return fail(WACIL_E_ALIGN, "W must be 16-byte aligned");
  if (dtype == WACIL_F32) return wa_launch<float>(W, C_total, D, n_new, gamma_out, workspace, S(stream));
  if (dtype == WACIL_BF16) return wa_launch<__nv_bfloat16>(W, C_total, D, n_new, gamma_out, workspace, S(stream));
  return fail(WACIL_E_DTYPE, "dtype %d", dtype);
}

// ---------------------------------------------------------------------------------------- classifier head (tcgen05)
namespace {
int fc_gemm_launch(const float* A, long long lda, int a_mn, const float* Bm, long long ldb, int b_mn, int M, int N,
                   int K, float* out, long long ldo, const float* bias, cudaStream_t st) {
  static bool attr_set = false;
  FcGemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.a_mn = a_mn; p.b_mn = b_mn;
  fc_gemm_shape(M, N, num_sms(), p.bn, p.kb, p.stages);
  p.out = out; p.ldo = ldo; p.bias = bias;
  CUtensorMap mapA, mapB;
  // K-major: rows = the M (N) extent, columns = K;  MN-major: rows = K, columns = the M (N) extent
  int e = a_mn ? make_f32_map(&mapA, A, K, M, lda, kFcKTile * p.kb, true) : make_f32_map(&mapA, A, M, K, lda, 128, false);
  if (e) return fail(WACIL_E_DRIVER, "classifier GEMM: tensor map of A (%dx%d, pitch %lld): %d", M, K, lda, e);
  e = b_mn ? make_f32_map(&mapB, Bm, K, N, ldb, kFcKTile * p.kb, true) : make_f32_map(&mapB, Bm, N, K, ldb, p.bn, false);
  if (e) return fail(WACIL_E_DRIVER, "classifier GEMM: tensor map of B (%dx%d, pitch %lld): %d", N, K, ldb, e);
  const size_t smem = fc_gemm_smem_bytes(256, 1, 4);  // the largest of the three shapes
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(fc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaFuncSetAttribute(fc_gemm_kernel)");
    attr_set = true;
  }
  const dim3 grid(static_cast<unsigned>((M + 127) / 128), static_cast<unsigned>((N + p.bn - 1) / p.bn));
  fc_gemm_kernel<<<grid, 192, fc_gemm_smem_bytes(p.bn, p.kb, p.stages), st>>>(mapA, mapB, p);
  WACIL_LAUNCH_CHECK("fc_gemm_kernel");
  return 0;
}
int fc_check(const void* ptr, long long ld, const char* what) {
  if (reinterpret_cast<uintptr_t>(ptr) & 15) return fail(WACIL_E_ALIGN, "%s must be 16-byte aligned", what);
  if (ld & 3) return fail(WACIL_E_ALIGN, "row pitch of %s (%lld floats) must be a multiple of 4", what, ld);
  return 0;
}
}  // namespace

int wacil_fc_forward(const float* feats, int64_t f_ld, const float* W, int64_t w_ld, const float* bias, float* logits,
                     int64_t l_ld, int B, int D, int C, void* stream) {
  if (!feats || !W || !logits) return fail(WACIL_E_NULL, "NULL argument");
  if (B <= 0 || D <= 0 || C <= 0 || f_ld < D || w_ld < D || l_ld < C)
    return fail(WACIL_E_SHAPE, "B=%d D=%d C=%d pitches %lld %lld %lld", B, D, C, (long long)f_ld, (long long)w_ld, (long long)l_ld);
  int e = fc_check(feats, f_ld, "feats");
  if (!e) e = fc_check(W, w_ld, "W");
  if (e) return e;
  return fc_gemm_launch(feats, f_ld, 0, W, w_ld, 0, B, C, D, logits, l_ld, bias, S(stream));
}

int wacil_fc_backward(const float* dlogits, int64_t g_ld, const float* feats, int64_t f_ld, const float* W,
                      int64_t w_ld, float* dfeats, int64_t df_ld, float* dW, int64_t dw_ld, float* db, int B, int D,
                      int C, void* stream) {
  if (!dlogits || !feats || !W) return fail(WACIL_E_NULL, "NULL argument");
  if (B <= 0 || D <= 0 || C <= 0 || g_ld < C || f_ld < D || w_ld < D || (dfeats && df_ld < D) || (dW && dw_ld < D))
    return fail(WACIL_E_SHAPE, "B=%d D=%d C=%d", B, D, C);
  int e = fc_check(dlogits, g_ld, "dlogits");
  if (!e) e = fc_check(feats, f_ld, "feats");
  if (!e) e = fc_check(W, w_ld, "W");
  if (e) return e;
  cudaStream_t st = S(stream);
  if (dfeats) {  // dfeats[b][d] = sum_c dlogits[b][c] W[c][d]
    e = fc_gemm_launch(dlogits, g_ld, 0, W, w_ld, 1, B, D, C, dfeats, df_ld, nullptr, st);
    if (e) return e;
  }
  if (dW) {  // dW[c][d] = sum_b dlogits[b][c] feats[b][d]
    e = fc_gemm_launch(dlogits, g_ld, 1, feats, f_ld, 1, C, D, B, dW, dw_ld, nullptr, st);
    if (e) return e;
  }
  if (db) {
    fc_bias_grad_kernel<<<(C + 31) / 32, 256, 0, st>>>(dlogits, g_ld, B, C, db);
    WACIL_LAUNCH_CHECK("fc_bias_grad_kernel");
  }
  return 0;
}

int wacil_sgd_step(float* p, const float* g, float* v, int64_t n, const float* lr_dev, float momentum,
                   float weight_decay, float grad_scale, void* stream) {
  if (!p || !g || !v || !lr_dev) return fail(WACIL_E_NULL, "NULL argument");
  if (n <= 0) return fail(WACIL_E_SHAPE, "n=%lld", static_cast<long long>(n));
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(v)) & 15)
    return fail(WACIL_E_ALIGN, "arenas must be 16-byte aligned");
  sgd_step_kernel<<<ew_grid(n / 4 + 1), 256, 0, S(stream)>>>(p, g, v, n, lr_dev, momentum, weight_decay, grad_scale);
  WACIL_LAUNCH_CHECK("sgd_step_kernel");
  return 0;
}

int wacil_allreduce_sgd(float* p, float* v, const void* const* grad_ptrs, void* const* pad_ptrs, int world, int rank,
                        int64_t n, const float* lr_dev, float momentum, float weight_decay, void* epoch_ws, void* stream) {
  if (!p || !v || !grad_ptrs || !pad_ptrs || !lr_dev || !epoch_ws) return fail(WACIL_E_NULL, "NULL argument");
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return fail(WACIL_E_SHAPE, "world=%d rank=%d", world, rank);
  if (n <= 0 || (n & 3)) return fail(WACIL_E_SHAPE, "n=%lld must be a positive multiple of 4", static_cast<long long>(n));
  AllreduceSgdParams q{};
  for (int r = 0; r < world; ++r) {
    if (!grad_ptrs[r] || !pad_ptrs[r]) return fail(WACIL_E_NULL, "peer pointer %d is NULL", r);
    if (reinterpret_cast<uintptr_t>(grad_ptrs[r]) & 15) return fail(WACIL_E_ALIGN, "peer gradient pointer %d", r);
    q.grads[r] = static_cast<const float*>(grad_ptrs[r]);
    q.pads[r] = static_cast<unsigned*>(pad_ptrs[r]);
  }
  q.world = world; q.rank = rank; q.p = p; q.v = v; q.n = n; q.lr_dev = lr_dev; q.mom = momentum; q.wd = weight_decay;
  q.seg_off[0] = 0; q.seg_off[1] = 0; q.seg_n1 = 0;
  q.epoch = static_cast<unsigned*>(epoch_ws);
  int grid = ew_grid(n / 4);
  if (grid > num_sms() * 4) grid = num_sms() * 4;
  allreduce_sgd_kernel<<<grid, 256, 0, S(stream)>>>(q);
  WACIL_LAUNCH_CHECK("allreduce_sgd_kernel");
  return 0;
}

size_t wacil_herd_workspace_bytes(int64_t n, int D) {
  if (n <= 0 || D <= 0) return 0;
  return align_up(static_cast<size_t>(n) * D * 4, 256) + static_cast<size_t>(n) * 4;
}

int wacil_herd_barycenter(const float* features, int64_t n, const int32_t* class_offsets, int n_classes, int D,
                          int quota, int32_t* out_idx, int32_t* out_count, void* workspace, void* stream) {
  if (!features || !class_offsets || !out_idx || !out_count || !workspace) return fail(WACIL_E_NULL, "NULL argument");
  if (n <= 0 || n_classes <= 0 || D <= 0 || D > 2048 || quota <= 0)
    return fail(WACIL_E_SHAPE, "n=%lld classes=%d D=%d quota=%d", static_cast<long long>(n), n_classes, D, quota);
  float* dnT = static_cast<float*>(workspace);
  int* rank = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + align_up(static_cast<size_t>(n) * D * 4, 256));
  herd_barycenter_kernel<<<n_classes, 256, 2 * D * sizeof(float), S(stream)>>>(features, class_offsets, D, quota, n, dnT,
                                                                              rank, out_idx, out_count);
  WACIL_LAUNCH_CHECK("herd_barycenter_kernel");
  return 0;
}

int wacil_herd_cluster(const float* features, int64_t n, const int32_t* class_offsets, int n_classes, int D, int quota,
                       int32_t* out_idx, int32_t* out_count, void* workspace, void* stream) {
  if (!features || !class_offsets || !out_idx || !out_count || !workspace) return fail(WACIL_E_NULL, "NULL argument");
  if (n <= 0 || n_classes <= 0 || D <= 0 || D > 2048 || quota <= 0)
    return fail(WACIL_E_SHAPE, "n=%lld classes=%d D=%d quota=%d", static_cast<long long>(n), n_classes, D, quota);
  herd_cluster_kernel<<<n_classes, 256, D * sizeof(float), S(stream)>>>(features, class_offsets, D, quota,
                                                                        static_cast<float*>(workspace), out_idx, out_count);
  WACIL_LAUNCH_CHECK("herd_cluster_kernel");
  return 0;
}

size_t wacil_eval_workspace_bytes(int B) { return 64 + static_cast<size_t>(B > 0 ? B : 0) * 3 * 4; }

int wacil_eval_rows(const float* logits, int64_t l_ld, const int64_t* targets, int B, int C, int topk, double* meter,
                    void* workspace, void* stream) {
  if (!logits || !targets || !meter || !workspace) return fail(WACIL_E_NULL, "NULL argument");
  if (B <= 0 || C <= 0 || l_ld < C || topk < 1) return fail(WACIL_E_SHAPE, "B=%d C=%d topk=%d", B, C, topk);
  unsigned* counter = static_cast<unsigned*>(workspace);
  float* part = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + 64);
  cudaStream_t st = S(stream);
#define EV_LAUNCH(TPR, CACHE)                                                                              \
  eval_rows_kernel<TPR, CACHE><<<(B + (256 / TPR) - 1) / (256 / TPR), 256, 0, st>>>(logits, l_ld, targets, B, C, topk, \
                                                                                     meter, part, counter)
  if (C <= 128) EV_LAUNCH(32, 4);
  else if (C <= 1024) EV_LAUNCH(32, 32);
  else if (C <= 8192) EV_LAUNCH(256, 32);
  else return fail(WACIL_E_SHAPE, "C = %d > 8192", C);
#undef EV_LAUNCH
  WACIL_LAUNCH_CHECK("eval_rows_kernel");
  return 0;
}

int wacil_augment_u8(const uint8_t* store, const int64_t* idx, const int32_t* aug, int B, int H, int W, int pad,
                     const float* mean3, const float* std3, float* out, void* stream) {
  if (!store || !mean3 || !std3 || !out) return fail(WACIL_E_NULL, "NULL argument");
  if (B <= 0 || H <= 0 || W <= 0 || pad < 0 || pad > 127) return fail(WACIL_E_SHAPE, "B=%d H=%d W=%d pad=%d", B, H, W, pad);
  for (int c = 0; c < 3; ++c)
    if (!(std3[c] > 0.f)) return fail(WACIL_E_SHAPE, "std[%d] must be positive", c);
  const long long npix = static_cast<long long>(B) * H * W;
  augment_u8_kernel<<<static_cast<int>((npix + 255) / 256), 256, 0, S(stream)>>>(
      store, idx, aug, B, H, W, pad, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], out);
  WACIL_LAUNCH_CHECK("augment_u8_kernel");
  return 0;
}

int wacil_conv3x3_fprop(const void* x, const void* w, void* y, float* stats, int N, int H, int W, int cin, int cout,
                        int stride, void* stream) {
  if (!x || !w || !y) return fail(WACIL_E_NULL, "NULL argument");
  if ((cin != 16 && cin != 32 && cin != 64) || cout % 16 || cout < 16 || cout > 128 || (stride != 1 && stride != 2) ||
      H != W || (H != 32 && H != 16 && H != 8) || (H / stride) < 8 || N < 1)
    return fail(WACIL_E_SHAPE, "conv3x3 N=%d H=%d W=%d cin=%d cout=%d stride=%d", N, H, W, cin, cout, stride);
  ConvLaunch c;
  int e = setup_conv(c, x, N, H, W, cin, cout, stride, 9, w, stats ? EPI_STATS : EPI_RAW);
  if (e) return e;
  c.p.out = static_cast<__nv_bfloat16*>(y);
  c.p.stats = stats;
  return launch_conv(c, S(stream));
}

size_t wacil_conv3x3_bn_relu_workspace_bytes(int cout) {
  return 64 + static_cast<size_t>(cout > 0 ? cout : 0) * 16 * num_sms();  // launch counter + [CTAs][2*cout] tagged sums
}

int wacil_conv3x3_bn_relu(const void* x, const void* w, const float* gamma, const float* beta, const void* res,
                          int res_mode, void* y, void* a, float* mean_rstd, float* running_mean, float* running_var,
                          int N, int H, int W, int cin, int cout, int stride, void* workspace, void* stream) {
  if (!x || !w || !gamma || !beta || !y || !a || !mean_rstd || !running_mean || !running_var || !workspace)
    return fail(WACIL_E_NULL, "NULL argument");
  if ((cin != 16 && cin != 32 && cin != 64) || (cout != 16 && cout != 32 && cout != 64) || (stride != 1 && stride != 2) ||
      H != W || (H != 32 && H != 16 && H != 8) || (H / stride) < 8 || N < 1 || res_mode < 0 || res_mode > 2 ||
      (res_mode != 0 && !res))
    return fail(WACIL_E_SHAPE, "conv3x3_bn_relu N=%d H=%d cin=%d cout=%d stride=%d res_mode=%d", N, H, cin, cout, stride,
                res_mode);
  unsigned* gbar = static_cast<unsigned*>(workspace);
  float* stats = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + 64);
  ConvLaunch c;
  int e = setup_conv(c, x, N, H, W, cin, cout, stride, 9, w, EPI_STATS);
  if (e) return e;
  if ((e = make_fused(c, EPI_BN_TRAIN, gbar))) return e;
  c.p.out = static_cast<__nv_bfloat16*>(y);
  c.p.act_out = static_cast<__nv_bfloat16*>(a);
  c.p.stats = stats;
  c.p.saved = mean_rstd;
  c.p.gamma = gamma;
  c.p.beta = beta;
  c.p.running_mean = running_mean;
  c.p.running_var = running_var;
  c.p.res = res_mode ? static_cast<const __nv_bfloat16*>(res) : nullptr;
  c.p.res_stride = res_mode == 2 ? 2 : 1;
  c.p.res_c = res_mode == 2 ? cout / 2 : cout;
  return launch_conv(c, S(stream));
}

size_t wacil_conv3x3_wgrad_workspace_bytes(int N, int H, int cin, int cout, int stride) {
  if (N < 1 || H < 1 || stride < 1 || cin < 1 || cout < 1) return 0;
  const int Ho = H / stride;
  return static_cast<size_t>(wgrad_grid((N * Ho * Ho + 127) / 128)) * 9 * cin * cout * 4;
}

int wacil_conv3x3_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int cin, int cout, int stride,
                        void* workspace, void* stream) {
  if (!x || !dy || !dw || !workspace) return fail(WACIL_E_NULL, "NULL argument");
  if ((cin != 16 && cin != 32 && cin != 64) || (cout != 16 && cout != 32 && cout != 64) ||
      (stride != 1 && stride != 2) || H != W || (H != 32 && H != 16 && H != 8) || (H / stride) < 8 || N < 1)
    return fail(WACIL_E_SHAPE, "wgrad N=%d H=%d cin=%d cout=%d stride=%d", N, H, cin, cout, stride);
  if (reinterpret_cast<uintptr_t>(workspace) & 15) return fail(WACIL_E_ALIGN, "workspace must be 16-byte aligned");
  WgradLaunch wl;
  int e = setup_wgrad(wl, x, dy, static_cast<float*>(workspace), N, H, W, cin, cout, stride, 9);
  if (e) return e;
  if ((e = launch_wgrad(wl, S(stream)))) return e;
  // per-CTA partial sums -> dw, added in CTA order (bit-reproducible)
  WgradReduceDesc rd{static_cast<const float*>(workspace), dw, wgrad_active_ctas(wl.p.num_tiles, wl.grid / wl.p.gsplit), 9 * cin * cout};
  wgrad_reduce_one_kernel<<<dim3(148, 1), 256, 0, S(stream)>>>(rd);
  WACIL_LAUNCH_CHECK("wgrad_reduce_one_kernel");
  return 0;
}

}  // extern "C"
