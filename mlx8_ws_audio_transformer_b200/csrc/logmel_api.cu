// C ABI of liblogmel_b200.so (declared in include/logmel.h).
//
// Build (see __graft_entry__.build / Makefile):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC \
//        -shared -o liblogmel_b200.so logmel_api.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/logmel.h"
#include "logmel_kernel.cuh"
#include "logmel_ws_kernel.cuh"
#include "logmel_tf_kernel.cuh"
#include "logmel_tables.h"

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

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
  return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
}

#define CUDA_TRY(expr)                                  \
  do {                                                  \
    cudaError_t e__ = (expr);                           \
    if (e__ != cudaSuccess) return cuda_fail(e__, #expr); \
  } while (0)

// Makes `device` current for the lifetime of the object and restores the caller's device on every
// exit path: the library never leaks a cudaSetDevice into the caller (torch reads the same state).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) {
      err = cudaSetDevice(device);
      if (err != cudaSuccess) prev = -1;
    } else if (err == cudaSuccess) {
      prev = -1;                                   // already current: nothing to restore
    }
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

constexpr int kMaxGroup = 64;
constexpr int LM_RETRY_PLAIN = -1000;   // internal: the warp-specialised kernel cannot take this filter bank

struct HostPipe {   // staging for lm_forward_host
  cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_h2d[2] = {}, ev_comp[2] = {}, ev_d2h[2] = {};
  float* d_wave[2] = {};
  float* d_out[2] = {};
  int* d_len[2] = {};
  void* d_scratch[2] = {};
  size_t wave_cap = 0, out_cap = 0, len_cap = 0, scratch_cap = 0;
  bool ready = false;
};

}  // namespace

// The thread-per-frame kernel (logmel_tf_kernel.cuh): attached to a Whisper handle when the filter
// bank has one of the two generated sparsity patterns; lm_forward picks it for batches that give
// every warp of the grid at least one clip.
struct TfLauncher {
  lm::TfTables tab;
  const void* kernel[2] = {nullptr, nullptr};        // frames per clip read from the arguments; [1]: clips split between pairs / CTAs
  const void* kernel3000[2] = {nullptr, nullptr};    // 30 s clips: the store offsets are immediates
  int n_mels = 0, n_sm = 0;
  int launch(const lm::KArgs& a, cudaStream_t st);
};

struct lm_handle {
  lm_config cfg{};
  int n_sm = 0, ctas_per_sm = 0, smem = 0, threads = 0, frames_per_tile = 0;
  std::mutex host_mu;
  HostPipe pipe;
  std::unique_ptr<TfLauncher> tf;
  long long tf_min_batch = 0;        // smallest batch routed to the thread-per-frame kernel
  int tf_pairs_forced = 0;           // tuning knob LM_TF_PAIRS_PER_CLIP: 1 or 4 warp pairs per clip, 0 = by cost
  int tf_slices_ok = 1;              // tuning knob LM_TF_SLICES=0: never spread a clip over several CTAs
  std::string tiled_name;            // the CTA-tiled kernel of this handle, as profilers print it
  int ctas_per_clip = 1;             // CTA-tiled kernels, steady state (tuning knob LM_CTAS_PER_CLIP)
  virtual ~lm_handle() {}
  virtual int launch(const lm::KArgs& a, int grid, cudaStream_t st) = 0;
};

int TfLauncher::launch(const lm::KArgs& a, cudaStream_t st) {
  void* args[] = {(void*)&tab, (void*)&a};
  const int split = a.group >= lm::TfGeo::PAIRS ? 1 : 0;
  const void* k = a.n_frames == 3000 ? kernel3000[split] : kernel[split];
  const int slices = a.group > lm::TfGeo::PAIRS ? a.group / lm::TfGeo::PAIRS : 1;
  cudaError_t e;
  if (slices > 1) {     // a clip over several CTAs that wait for each other: all of them must be resident
    e = cudaLaunchCooperativeKernel(k, dim3((unsigned)(a.batch * slices)), dim3(lm::TfGeo::THREADS), args, lm::TfGeo::SMEM_REQUEST, st);
  } else {              // clips spread over the SMs first (pair p of CTA c: clip p * grid + c)
    const int grid = (int)std::min<long long>(n_sm, a.batch);
    e = cudaLaunchKernel(k, dim3(grid), dim3(lm::TfGeo::THREADS), args, lm::TfGeo::SMEM_REQUEST, st);
  }
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernel(logmel_tf_kernel)");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

namespace {

template <class G>
constexpr size_t smem_bytes() {
  if constexpr (G::WS) return lm::LayWS<G>::BYTES;
  else return lm::Lay<G>::BYTES;
}

template <class G>
struct Impl : lm_handle {
  lm::Tables<G> tab;
  const void* kernel = nullptr;
  template <int K>
  static const void* kern() {
    if constexpr (G::WS) return (const void*)lm::logmel_ws_kernel<G, K>;
    else return (const void*)lm::logmel_kernel<G, K>;
  }
  static const void* pick(int log_mode) {
    switch (log_mode) {
      case LM_LOG_NONE: return kern<0>();
      case LM_LN_PLUS_EPS: return kern<2>();
      case LM_LOG10_CLAMP: return kern<1>();
      default: return kern<3>();
    }
  }
  int launch(const lm::KArgs& a, int grid, cudaStream_t st) override {
    void* args[] = {(void*)&tab, (void*)&a};
    cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(G::THREADS), args, smem_bytes<G>(), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchCooperativeKernel(logmel_kernel)");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
  }
};

template <class G>
int make(lm_handle** out, const lm_config* cfg, const float* window) {
  auto h = std::make_unique<Impl<G>>();
  h->cfg = *cfg;
  h->cfg.fbank = nullptr;
  h->cfg.window = nullptr;
  std::string err = lm::build_tables<G>(h->tab, window, cfg->fbank, cfg->n_mels);
  if (!err.empty()) return fail(LM_ERR_FBANK, "%s", err.c_str());
  if (G::WS && !h->tab.mel_scan) return LM_RETRY_PLAIN;   // in-place stage 2 needs the scan form of the bank
  const int smem = (int)smem_bytes<G>();
  h->kernel = Impl<G>::pick(cfg->log_mode);
  CUDA_TRY(cudaFuncSetAttribute(h->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, h->kernel, G::THREADS, smem));
  if (occ < 1) return fail(LM_ERR_NO_DEVICE, "forward kernel does not fit on this device (smem %d B)", smem);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  if (!prop.cooperativeLaunch) return fail(LM_ERR_NO_DEVICE, "device lacks cooperative launch");
  {
    char nm[96];
    snprintf(nm, sizeof(nm), "lm::%s<lm::Geo<%d, %d, %d, %d>, %d>", G::WS ? "logmel_ws_kernel" : "logmel_kernel", G::N, G::HOP,
             G::PK, G::WS, cfg->log_mode == LM_LOG_NONE ? 0 : cfg->log_mode == LM_LN_PLUS_EPS ? 2 : cfg->log_mode == LM_LOG10_CLAMP ? 1 : 3);
    h->tiled_name = nm;
  }
  if (const char* e = std::getenv("LM_CTAS_PER_CLIP")) h->ctas_per_clip = std::max(1, atoi(e));
  h->n_sm = prop.multiProcessorCount;
  h->ctas_per_sm = occ;
  h->smem = smem;
  h->threads = G::THREADS;
  h->frames_per_tile = G::F;
  *out = h.release();
  return 0;
}

template <int NM>
bool tf_fill(lm::TfTables& t, const float* window, const float* fbank) {
  using P = lm::TfMelPattern<NM>;
  int nnz = 0;
  for (int k = 0; k < 201; ++k)
    for (int m = 0; m < NM; ++m) nnz += fbank[(size_t)k * NM + m] != 0.0f;
  if (nnz != P::NNZ) return false;
  for (int i = 0; i < P::NNZ; ++i) {       // the kernel's mel weights are literals: the bank must be THAT bank
    unsigned bits;
    std::memcpy(&bits, &fbank[(size_t)P::bin[i] * NM + P::mel[i]], 4);
    if (bits != P::bits[i]) return false;
  }
  for (int cp = 0; cp < 10; ++cp) {
    for (int a = 0; a < 20; ++a) {
      const float w0 = window[20 * a + 2 * cp], w1 = window[20 * a + 2 * cp + 1];
      t.cp[cp].w[a] = lm::tf_pack2(w0, w1);
      t.cp[cp].nw[a] = lm::tf_pack2(-w0, -w1);
    }
    for (int k = 1; k <= 10; ++k) {
      float c[2], s[2];
      for (int h = 0; h < 2; ++h) {      // same rounding as build_tables(): cos / sin in double, rounded once
        const double ang = -2.0 * M_PI * (double)(2 * cp + h) * (double)k / 400.0;
        c[h] = (float)std::cos(ang);
        s[h] = (float)std::sin(ang);
      }
      t.cp[cp].twr[k - 1] = lm::tf_pack2(c[0], c[1]);
      t.cp[cp].ntwr[k - 1] = lm::tf_pack2(-c[0], -c[1]);
      t.cp[cp].twi[k - 1] = lm::tf_pack2(s[0], s[1]);
      t.cp[cp].ntwi[k - 1] = lm::tf_pack2(-s[0], -s[1]);
    }
  }
  return true;
}

// returns 0 with h->tf set when the bank qualifies, 0 with h->tf empty when it does not
int attach_tf(lm_handle* h, const lm_config* cfg, const float* window) {
  if (cfg->log_mode != LM_LOG10_CLAMP_WHISPER_NORM || (cfg->n_mels != 80 && cfg->n_mels != 128)) return 0;
  auto t = std::make_unique<TfLauncher>();
  std::memset(&t->tab, 0, sizeof(t->tab));
  const bool ok = cfg->n_mels == 80 ? tf_fill<80>(t->tab, window, cfg->fbank) : tf_fill<128>(t->tab, window, cfg->fbank);
  if (!ok) return 0;
  t->kernel[0] = cfg->n_mels == 80 ? (const void*)lm::logmel_tf_kernel<80, 0, false> : (const void*)lm::logmel_tf_kernel<128, 0, false>;
  t->kernel[1] = cfg->n_mels == 80 ? (const void*)lm::logmel_tf_kernel<80, 0, true> : (const void*)lm::logmel_tf_kernel<128, 0, true>;
  t->kernel3000[0] = cfg->n_mels == 80 ? (const void*)lm::logmel_tf_kernel<80, 3000, false> : (const void*)lm::logmel_tf_kernel<128, 3000, false>;
  t->kernel3000[1] = cfg->n_mels == 80 ? (const void*)lm::logmel_tf_kernel<80, 3000, true> : (const void*)lm::logmel_tf_kernel<128, 3000, true>;
  t->n_mels = cfg->n_mels;
  t->n_sm = h->n_sm;
  for (int i = 0; i < 2; ++i) {
    CUDA_TRY(cudaFuncSetAttribute(t->kernel[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lm::TfGeo::SMEM_REQUEST));
    CUDA_TRY(cudaFuncSetAttribute(t->kernel3000[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lm::TfGeo::SMEM_REQUEST));
  }
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, t->kernel[1], lm::TfGeo::THREADS, lm::TfGeo::SMEM_REQUEST));
  if (occ != 1) return fail(LM_ERR_NO_DEVICE, "thread-per-frame kernel: %d CTAs per SM (expected exactly 1: the CTA owns all of TMEM)", occ);
  h->tf = std::move(t);
  // Batches too small for a clip per CTA (fewer clips than half the SMs) spread a clip over several CTAs
  // (tf_pairs_per_clip); above that one round of the kernel costs a quarter clip.  The CTA-tiled kernel keeps the
  // launches this one cannot take: other banks / log modes, unaligned clips, a handful of very short clips
  // (profiles/r02_dispatch_sweep.txt: it loses at every batch size of 30 s clips).
  h->tf_min_batch = (long long)h->n_sm / 2 + 1;
  if (const char* e = std::getenv("LM_TF_MIN_BATCH")) h->tf_min_batch = std::max(1, atoi(e));   // tuning knob
  if (const char* e = std::getenv("LM_TF_PAIRS_PER_CLIP")) h->tf_pairs_forced = atoi(e) == 4 ? 4 : atoi(e) == 1 ? 1 : 0;
  if (const char* e = std::getenv("LM_TF_SLICES")) h->tf_slices_ok = atoi(e) != 0;
  if (h->tf_min_batch > h->n_sm) h->tf_slices_ok = 0;      // LM_TF_MIN_BATCH=<huge>: the CTA-tiled kernel for everything
  {
    int coop_ok = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, cfg->device));
    if (!coop_ok) h->tf_slices_ok = 0;
  }
  return 0;
}

int64_t frames_for(const lm_config& c, int64_t n_samples) {
  return 1 + n_samples / c.hop - (c.drop_last ? 1 : 0);
}

// the thread-per-frame kernel takes the launch when the handle has one, the clips are 16-byte
// aligned (cp.async), the batch gives every warp pair a clip and a clip has at most kTfMaxTiles tiles
bool use_tf(const lm_handle* h, int64_t batch, int64_t n_frames, bool aligned) {
  const int64_t tiles = (n_frames + lm::TfGeo::F - 1) / lm::TfGeo::F;
  if (!h->tf || !aligned || tiles > lm::kTfMaxTiles) return false;
  // enough clips for a quarter clip per warp pair -- or so few that a clip spreads over >= 2 CTAs (tf_pairs_per_clip)
  return batch >= h->tf_min_batch || (h->tf_slices_ok && batch >= 1 && 2 * batch <= h->n_sm && tiles >= 2 * lm::TfGeo::PAIRS);
}

// Thread-per-frame kernel: a clip per warp pair, or a clip per CTA with a quarter of its tiles per pair?  In tile
// periods: whole clips go round by round over 4 n_sm pairs, quarter clips over n_sm CTAs; the second pays one CTA
// barrier per clip and the slower clip-edge tiles inside one quarter (2 % here), so it is taken only when it saves
// a round: mid-size batches (8192 clips over 8 GPUs: 1024 = 1.73 rounds of pairs, but 6.92 of CTAs).
int tf_pairs_per_clip(const lm_handle* h, int64_t batch, int tiles) {
  if (h->tf_pairs_forced) return h->tf_pairs_forced;
  const int64_t P = lm::TfGeo::PAIRS, n = h->n_sm;
  const int64_t by_pair = ((batch + n * P - 1) / (n * P)) * tiles;
  const int64_t by_cta = ((batch + n - 1) / n) * ((tiles + P - 1) / P);
  // small batches: a clip per several CTAs (at least two, at most one tile period per pair, kMaxGroup scratch slots)
  const int64_t slices = std::min<int64_t>(std::min<int64_t>(n / std::max<int64_t>(batch, 1), kMaxGroup), (tiles + P - 1) / P);
  if (slices >= 2 && h->tf_slices_ok) return (int)(P * slices);
  return by_cta * 102 < by_pair * 100 ? (int)P : 1;
}

void choose_grid(const lm_handle* h, int64_t batch, int tiles, int* group, int* n_groups) {
  const int total = h->n_sm * h->ctas_per_sm;
  // Steady state (batch >= number of CTAs): ONE CTA per clip.  Measured on 4096 x 30 s clips,
  // CTAs per clip 8 / 4 / 2 / 1: 6.47 / 5.86 / 5.69 / 5.50 ms -- no inter-CTA agreement on the clip
  // maximum, no 11-vs-12-tile imbalance.  The price: 148 clips of un-normalised features (227 MB)
  // no longer fit the 126 MB L2, so the (pipelined) fix-up of tiles below max-8 reads DRAM; on an
  // input where EVERY tile needs it that costs 5 % (4.37 vs 4.16 ms per 2048 clips).
  // Small batches: spread every clip over as many CTAs as it has tiles.
  const int div = std::max(1, h->ctas_per_clip);  // CTAs per clip in steady state (LM_CTAS_PER_CLIP, read by lm_create)
  const int steady_groups = std::max(1, h->n_sm / div);
  int g = std::max(1, total / steady_groups);
  if (batch < steady_groups) g = (int)std::min<int64_t>(kMaxGroup, std::max<int64_t>(g, total / std::max<int64_t>(batch, 1)));
  g = std::max(1, std::min(std::min(g, kMaxGroup), tiles));
  int ng = std::max(1, total / g);
  if ((int64_t)ng > batch) ng = (int)batch;
  *group = g;
  *n_groups = ng;
}

}  // namespace

extern "C" {

int lm_version(void) { return LM_ABI_VERSION; }
const char* lm_last_error(void) { return g_err.c_str(); }
int64_t lm_launch_count(void) { return g_launches.load(); }

int lm_create(lm_handle** out, const lm_config* cfg) {
  if (!out || !cfg || !cfg->fbank) return fail(LM_ERR_NULL, "lm_create: out, cfg and cfg->fbank must be non-NULL");
  *out = nullptr;
  if (cfg->log_mode < LM_LOG_NONE || cfg->log_mode > LM_LOG10_CLAMP) return fail(LM_ERR_MODE, "unknown log_mode %d", cfg->log_mode);
  if (cfg->n_mels < 1 || cfg->n_mels > lm::kMaxMels) return fail(LM_ERR_FBANK, "n_mels=%d not in [1,%d]", cfg->n_mels, lm::kMaxMels);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(LM_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(LM_ERR_NO_DEVICE, "device %d out of range", cfg->device);
  DeviceGuard guard(cfg->device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err, "cudaSetDevice");
  std::vector<float> win;
  if (cfg->window) win.assign(cfg->window, cfg->window + cfg->n_fft);
  else if (cfg->n_fft > 0) win = lm::hann_periodic(cfg->n_fft);
  const int v = cfg->variant;
  if (cfg->n_fft == 400 && cfg->hop == 160) {
    if (v == 1) return make<lm::Geo<400, 160, 1>>(out, cfg, win.data());
    if (v == 2) return make<lm::Geo<400, 160, 2>>(out, cfg, win.data());
    int rc = make<lm::Geo<400, 160, 2, 1>>(out, cfg, win.data());   // CTA-tiled default: warp-specialised CTA
    if (rc == LM_RETRY_PLAIN) rc = make<lm::Geo<400, 160, 2>>(out, cfg, win.data());
    if (rc != 0) return rc;
    rc = attach_tf(*out, cfg, win.data());                            // + thread-per-frame kernel for large batches
    if (rc == 0 && v == 3) {
      if (!(*out)->tf) rc = fail(LM_ERR_FBANK, "variant 3 (thread-per-frame kernel) needs the Whisper normalisation and the 80- or 128-filter Slaney bank");
      else (*out)->tf_min_batch = 1;
    }
    if (rc != 0) {
      lm_destroy(*out);
      *out = nullptr;
    }
    return rc;
  }
  if (cfg->n_fft == 1024 && cfg->hop == 512) return make<lm::Geo<1024, 512, 1>>(out, cfg, win.data());
  if (cfg->n_fft == 1024 && cfg->hop == 128) return make<lm::Geo<1024, 128, 1>>(out, cfg, win.data());
  return fail(LM_ERR_GEOMETRY, "no kernel for n_fft=%d hop=%d (supported: 400/160, 1024/512, 1024/128)", cfg->n_fft, cfg->hop);
}

void lm_destroy(lm_handle* h) {
  if (!h) return;
  HostPipe& p = h->pipe;
  if (p.ready) {
    DeviceGuard guard(h->cfg.device);
    for (int i = 0; i < 2; ++i) {
      cudaFree(p.d_wave[i]);
      cudaFree(p.d_out[i]);
      cudaFree(p.d_len[i]);
      cudaFree(p.d_scratch[i]);
      cudaEventDestroy(p.ev_h2d[i]);
      cudaEventDestroy(p.ev_comp[i]);
      cudaEventDestroy(p.ev_d2h[i]);
    }
    cudaStreamDestroy(p.s_h2d);
    cudaStreamDestroy(p.s_comp);
    cudaStreamDestroy(p.s_d2h);
  }
  delete h;
}

int64_t lm_num_frames(const lm_handle* h, int64_t n_samples) {
  if (!h) return LM_ERR_NULL;
  return frames_for(h->cfg, n_samples);
}

size_t lm_scratch_bytes(const lm_handle* h, int64_t batch) {
  if (!h || batch < 0) return 0;
  return (size_t)batch * (kMaxGroup + 1) * 4 + 16;
}

int lm_kernel_info(const lm_handle* h, int32_t* n_sm, int32_t* ctas_per_sm, int32_t* smem_bytes,
                   int32_t* threads, int32_t* frames_per_tile) {
  if (!h) return fail(LM_ERR_NULL, "lm_kernel_info: NULL handle");
  if (n_sm) *n_sm = h->n_sm;
  if (ctas_per_sm) *ctas_per_sm = h->ctas_per_sm;
  if (smem_bytes) *smem_bytes = h->smem;
  if (threads) *threads = h->threads;
  if (frames_per_tile) *frames_per_tile = h->frames_per_tile;
  return 0;
}

const char* lm_kernel_name(const lm_handle* h, int64_t batch, int64_t n_samples) {
  if (!h) return "";
  const int64_t n_frames = frames_for(h->cfg, n_samples);
  if (use_tf(h, batch, n_frames, true)) {
    static const char* const names[2][2][2] = {
        {{"lm::logmel_tf_kernel<80, 0, 0>", "lm::logmel_tf_kernel<80, 0, 1>"},
         {"lm::logmel_tf_kernel<80, 3000, 0>", "lm::logmel_tf_kernel<80, 3000, 1>"}},
        {{"lm::logmel_tf_kernel<128, 0, 0>", "lm::logmel_tf_kernel<128, 0, 1>"},
         {"lm::logmel_tf_kernel<128, 3000, 0>", "lm::logmel_tf_kernel<128, 3000, 1>"}}};
    const int tiles = (int)((n_frames + lm::TfGeo::F - 1) / lm::TfGeo::F);
    const int split = tf_pairs_per_clip(h, batch, tiles) >= lm::TfGeo::PAIRS ? 1 : 0;   // third argument: clips are split
    return names[h->cfg.n_mels == 80 ? 0 : 1][n_frames == 3000 ? 1 : 0][split];
  }
  return h->tiled_name.c_str();
}

}  // extern "C"

namespace {
// d_wave (float32) or d_pcm (int16, `channels` interleaved channels): exactly one is non-NULL
int forward_impl(lm_handle* h, const float* d_wave, const int16_t* d_pcm, int channels, int64_t batch, int64_t clip_stride,
                 int64_t n_samples, const int32_t* d_lengths, float* d_out, float* d_clip_max, void* d_scratch,
                 size_t scratch_bytes, void* stream) {
  if (!h) return fail(LM_ERR_NULL, "lm_forward: NULL handle");
  if (batch == 0) return 0;
  if ((!d_wave && !d_pcm) || !d_out) return fail(LM_ERR_NULL, "lm_forward: the waveform and d_out must be non-NULL");
  if (d_pcm && channels != 1 && channels != 2) return fail(LM_ERR_SHAPE, "channels=%d: 16-bit PCM input is mono or stereo", channels);
  const lm_config& c = h->cfg;
  if (batch < 0 || batch > (1 << 24)) return fail(LM_ERR_SHAPE, "batch=%lld out of range", (long long)batch);
  if (n_samples <= c.n_fft / 2 || n_samples > (1LL << 30))
    return fail(LM_ERR_SHAPE, "n_samples=%lld: reflect padding needs more than n_fft/2=%d samples (as torch.stft)",
                (long long)n_samples, c.n_fft / 2);
  if (clip_stride < 0 || (!d_lengths && clip_stride < n_samples))
    return fail(LM_ERR_SHAPE, "clip_stride=%lld smaller than n_samples=%lld without d_lengths",
                (long long)clip_stride, (long long)n_samples);
  const int64_t n_frames = frames_for(c, n_samples);
  if (n_frames < 1) return fail(LM_ERR_SHAPE, "n_samples=%lld gives no frame", (long long)n_samples);
  const bool norm = c.log_mode == LM_LOG10_CLAMP_WHISPER_NORM;
  if (norm && (!d_scratch || scratch_bytes < lm_scratch_bytes(h, batch)))
    return fail(LM_ERR_SCRATCH, "scratch %zu B < required %zu B", scratch_bytes, lm_scratch_bytes(h, batch));
  if (norm && ((uintptr_t)d_scratch & 15)) return fail(LM_ERR_SCRATCH, "scratch must be 16-byte aligned");

  cudaStream_t st = (cudaStream_t)stream;
  DeviceGuard guard(c.device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err, "cudaSetDevice");

  lm::KArgs a{};
  a.wave = d_wave;
  a.pcm = d_pcm;
  a.pcm_channels = d_pcm ? channels : 0;
  a.clip_stride = clip_stride;
  a.lengths = d_lengths;
  a.out = d_out;
  a.clip_max = d_clip_max;
  a.batch = (int)batch;
  a.n_samples = (int)n_samples;
  a.n_frames = (int)n_frames;
  a.n_mels = c.n_mels;
  a.log_mode = c.log_mode;
  a.log_add = 0.0f;
  a.log_floor = 0.0f;
  a.log_scale = 1.0f;
  if (c.log_mode == LM_LOG10_CLAMP_WHISPER_NORM || c.log_mode == LM_LOG10_CLAMP) {
    a.log_floor = c.log_param;
    a.log_scale = 0.30102999566398120f;   // log10(2)
  } else if (c.log_mode == LM_LN_PLUS_EPS) {
    a.log_add = c.log_param;
    a.log_scale = 0.69314718055994531f;   // ln(2)
  }
  a.tiles_per_clip = (int)((n_frames + h->frames_per_tile - 1) / h->frames_per_tile);
  choose_grid(h, batch, a.tiles_per_clip, &a.group, &a.n_groups);
  a.vec_ok = (n_frames % 4 == 0) && (((uintptr_t)d_out & 15) == 0);
  // every clip starts 16-byte aligned (float32: TMA bulk copies / cp.async; PCM: 16-byte vector loads)
  // (a single clip has no stride to speak of)
  a.tma_ok = d_pcm ? ((((uintptr_t)d_pcm & 15) == 0) && ((clip_stride * channels) % 8 == 0 || batch == 1))
                   : ((((uintptr_t)d_wave & 15) == 0) && (clip_stride % 4 == 0 || batch == 1));
  const bool aligned = a.tma_ok != 0;
  a.timeline = nullptr;
#ifdef LM_TIMELINE
  a.timeline = reinterpret_cast<long long*>(d_clip_max);   // debug build: d_clip_max carries the stamp buffer
  a.clip_max = nullptr;
#endif
  if (norm) {
    a.gcnt = reinterpret_cast<int*>(d_scratch);
    a.gmax = reinterpret_cast<float*>(d_scratch) + ((batch + 3) / 4) * 4;
  }
  int rc;
  if (use_tf(h, batch, n_frames, aligned)) {
    a.tiles_per_clip = (int)((n_frames + lm::TfGeo::F - 1) / lm::TfGeo::F);
    a.group = tf_pairs_per_clip(h, batch, a.tiles_per_clip);
    a.n_groups = 0;
    if (a.group > lm::TfGeo::PAIRS) CUDA_TRY(cudaMemsetAsync(a.gcnt, 0, (size_t)batch * sizeof(int), st));
    rc = h->tf->launch(a, st);
  } else {
    if (d_pcm) a.tma_ok = 0;
    if (norm) CUDA_TRY(cudaMemsetAsync(a.gcnt, 0, (size_t)batch * sizeof(int), st));
    rc = h->launch(a, a.group * a.n_groups, st);
  }
  return rc;
}
}  // namespace

extern "C" {

int lm_forward(lm_handle* h, const float* d_wave, int64_t batch, int64_t clip_stride, int64_t n_samples,
               const int32_t* d_lengths, float* d_out, float* d_clip_max, void* d_scratch,
               size_t scratch_bytes, void* stream) {
  if (batch != 0 && !d_wave) return fail(LM_ERR_NULL, "lm_forward: d_wave and d_out must be non-NULL");
  return forward_impl(h, d_wave, nullptr, 0, batch, clip_stride, n_samples, d_lengths, d_out, d_clip_max, d_scratch,
                      scratch_bytes, stream);
}

int lm_forward_pcm16(lm_handle* h, const int16_t* d_pcm, int32_t channels, int64_t batch, int64_t clip_stride,
                     int64_t n_samples, const int32_t* d_lengths, float* d_out, float* d_clip_max, void* d_scratch,
                     size_t scratch_bytes, void* stream) {
  if (batch != 0 && !d_pcm) return fail(LM_ERR_NULL, "lm_forward_pcm16: d_pcm and d_out must be non-NULL");
  return forward_impl(h, nullptr, d_pcm, channels, batch, clip_stride, n_samples, d_lengths, d_out, d_clip_max, d_scratch,
                      scratch_bytes, stream);
}

int lm_host_register(void* p, size_t bytes) {
  if (!p) return fail(LM_ERR_NULL, "lm_host_register: NULL");
  CUDA_TRY(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
  return 0;
}
int lm_host_unregister(void* p) {
  if (!p) return fail(LM_ERR_NULL, "lm_host_unregister: NULL");
  CUDA_TRY(cudaHostUnregister(p));
  return 0;
}

}  // extern "C"

namespace {
// h_wave: float32 samples (channels == 0) or int16 PCM frames of `channels` interleaved channels
int forward_host_impl(lm_handle* h, const void* h_wave_v, int channels, int64_t batch, int64_t clip_stride, int64_t n_samples,
                      const int32_t* h_lengths, float* h_out) {
  const char* h_wave = static_cast<const char*>(h_wave_v);
  const size_t elem = channels ? (size_t)2 * channels : 4;         // bytes per sample / frame
  if (!h) return fail(LM_ERR_NULL, "lm_forward_host: NULL handle");
  if (channels != 0 && channels != 1 && channels != 2) return fail(LM_ERR_SHAPE, "channels=%d: 16-bit PCM input is mono or stereo", channels);
  if (batch == 0) return 0;
  if (!h_wave || !h_out) return fail(LM_ERR_NULL, "lm_forward_host: h_wave and h_out must be non-NULL");
  const lm_config& c = h->cfg;
  if (batch < 0) return fail(LM_ERR_SHAPE, "batch=%lld", (long long)batch);
  if (n_samples <= c.n_fft / 2) return fail(LM_ERR_SHAPE, "n_samples=%lld too short for reflect padding", (long long)n_samples);
  const int64_t n_frames = frames_for(c, n_samples);
  if (n_frames < 1) return fail(LM_ERR_SHAPE, "n_samples=%lld gives no frame", (long long)n_samples);
  std::lock_guard<std::mutex> lock(h->host_mu);
  DeviceGuard guard(c.device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err, "cudaSetDevice");

  // chunking: ~64 MiB of waveform per chunk, at least 1 clip, at most the whole batch
  // (a PCM row is padded to a multiple of 8 frames so that every staged clip starts 16-byte aligned)
  const int64_t dev_stride = channels ? (n_samples + 7) / 8 * 8 : n_samples;
  const size_t clip_in = (size_t)dev_stride * elem, clip_out = (size_t)c.n_mels * n_frames * 4;
  int64_t chunk = std::max<int64_t>(1, (int64_t)((64u << 20) / clip_in));
  chunk = std::min(chunk, batch);
  // copy width: clips shorter than n_samples on the host are copied at their stride
  const int64_t row = std::min(clip_stride, n_samples);

  HostPipe& p = h->pipe;
  if (!p.ready) {
    CUDA_TRY(cudaStreamCreateWithFlags(&p.s_h2d, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&p.s_comp, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&p.s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CUDA_TRY(cudaEventCreateWithFlags(&p.ev_h2d[i], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&p.ev_comp[i], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&p.ev_d2h[i], cudaEventDisableTiming));
    }
    p.ready = true;
  }
  const size_t need_w = (size_t)chunk * clip_in, need_o = (size_t)chunk * clip_out;
  const size_t need_l = (size_t)chunk * 4, need_s = lm_scratch_bytes(h, chunk);
  for (int i = 0; i < 2; ++i) {
    if (p.wave_cap < need_w) { cudaFree(p.d_wave[i]); p.d_wave[i] = nullptr; CUDA_TRY(cudaMalloc(&p.d_wave[i], need_w)); }
    if (p.out_cap < need_o) { cudaFree(p.d_out[i]); p.d_out[i] = nullptr; CUDA_TRY(cudaMalloc(&p.d_out[i], need_o)); }
    if (p.len_cap < need_l) { cudaFree(p.d_len[i]); p.d_len[i] = nullptr; CUDA_TRY(cudaMalloc(&p.d_len[i], need_l)); }
    if (p.scratch_cap < need_s) { cudaFree(p.d_scratch[i]); p.d_scratch[i] = nullptr; CUDA_TRY(cudaMalloc(&p.d_scratch[i], need_s)); }
  }
  p.wave_cap = std::max(p.wave_cap, need_w);
  p.out_cap = std::max(p.out_cap, need_o);
  p.len_cap = std::max(p.len_cap, need_l);
  p.scratch_cap = std::max(p.scratch_cap, need_s);

  int it = 0;
  for (int64_t c0 = 0; c0 < batch; c0 += chunk, ++it) {
    const int b = it & 1;
    const int64_t nb = std::min(chunk, batch - c0);
    // buffer b was last used by chunk it-2: its compute must be done before new input lands,
    // and its D2H before the kernel overwrites d_out[b]
    if (it >= 2) {
      CUDA_TRY(cudaStreamWaitEvent(p.s_h2d, p.ev_comp[b], 0));
      CUDA_TRY(cudaStreamWaitEvent(p.s_comp, p.ev_d2h[b], 0));
    }
    if (row == n_samples && clip_stride == n_samples && dev_stride == n_samples) {
      CUDA_TRY(cudaMemcpyAsync(p.d_wave[b], h_wave + (size_t)c0 * clip_stride * elem, (size_t)nb * clip_in,
                               cudaMemcpyHostToDevice, p.s_h2d));
    } else {
      if (row < dev_stride) CUDA_TRY(cudaMemsetAsync(p.d_wave[b], 0, (size_t)nb * clip_in, p.s_h2d));
      CUDA_TRY(cudaMemcpy2DAsync(p.d_wave[b], clip_in, h_wave + (size_t)c0 * clip_stride * elem, (size_t)clip_stride * elem,
                                 (size_t)row * elem, (size_t)nb, cudaMemcpyHostToDevice, p.s_h2d));
    }
    if (h_lengths)
      CUDA_TRY(cudaMemcpyAsync(p.d_len[b], h_lengths + c0, (size_t)nb * 4, cudaMemcpyHostToDevice, p.s_h2d));
    CUDA_TRY(cudaEventRecord(p.ev_h2d[b], p.s_h2d));
    CUDA_TRY(cudaStreamWaitEvent(p.s_comp, p.ev_h2d[b], 0));
    int rc = forward_impl(h, channels ? nullptr : p.d_wave[b], channels ? reinterpret_cast<const int16_t*>(p.d_wave[b]) : nullptr,
                          channels, nb, dev_stride, n_samples, h_lengths ? p.d_len[b] : nullptr, p.d_out[b], nullptr,
                          p.d_scratch[b], p.scratch_cap, p.s_comp);
    if (rc != 0) return rc;
    CUDA_TRY(cudaEventRecord(p.ev_comp[b], p.s_comp));
    CUDA_TRY(cudaStreamWaitEvent(p.s_d2h, p.ev_comp[b], 0));
    CUDA_TRY(cudaMemcpyAsync(h_out + c0 * (int64_t)c.n_mels * n_frames, p.d_out[b], (size_t)nb * clip_out,
                             cudaMemcpyDeviceToHost, p.s_d2h));
    CUDA_TRY(cudaEventRecord(p.ev_d2h[b], p.s_d2h));
  }
  CUDA_TRY(cudaStreamSynchronize(p.s_d2h));
  CUDA_TRY(cudaStreamSynchronize(p.s_comp));
  CUDA_TRY(cudaStreamSynchronize(p.s_h2d));
  return 0;
}
}  // namespace

extern "C" {

int lm_forward_host(lm_handle* h, const float* h_wave, int64_t batch, int64_t clip_stride, int64_t n_samples,
                    const int32_t* h_lengths, float* h_out) {
  return forward_host_impl(h, h_wave, 0, batch, clip_stride, n_samples, h_lengths, h_out);
}

int lm_forward_host_pcm16(lm_handle* h, const int16_t* h_pcm, int32_t channels, int64_t batch, int64_t clip_stride,
                          int64_t n_samples, const int32_t* h_lengths, float* h_out) {
  if (channels != 1 && channels != 2) return fail(LM_ERR_SHAPE, "channels=%d: 16-bit PCM input is mono or stereo", channels);
  return forward_host_impl(h, h_pcm, channels, batch, clip_stride, n_samples, h_lengths, h_out);
}

}  // extern "C"
