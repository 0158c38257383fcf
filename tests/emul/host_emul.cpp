// Host emulation of one CTA of the fused log-mel kernel, for CPU-only tests.
//
// It includes the SAME headers the CUDA kernel is built from (codelets, index maps, table
// builder) and walks warps and lanes sequentially, phase by phase, so the codelets, the
// shared-memory layouts, the bin folding and the banded mel tables are all exercised by
// `pytest -m "not gpu"` without a GPU.  It is a test harness, not a product path: the
// product library has no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../mlx8_ws_audio_transformer_b200/csrc/logmel_tables.h"

namespace {

using namespace lm;

inline float fin_log(float x, int mode, float param) {
  switch (mode) {
    case LOG10_CLAMP_WHISPER_NORM:
    case LOG10_CLAMP:
      return std::log2(std::max(x, param)) * 0.30102999566398120f;
    case LN_PLUS_EPS:
      return std::log2(x + param) * 0.69314718055994531f;
    default:
      return x;
  }
}

template <class G>
int run(const float* wave, long batch, long stride, const int32_t* lengths, int n_samples,
        int n_frames, const float* fbank, int n_mels, int log_mode, float log_param, float* out) {
  using T = typename ValT<G>::type;
  static Tables<G> tab;
  std::vector<float> win = hann_periodic(G::N);
  std::string err = build_tables<G>(tab, win.data(), fbank, n_mels);
  if (!err.empty()) {
    std::fprintf(stderr, "emul: %s\n", err.c_str());
    return -1;
  }
  std::vector<float> wave_s(G::WAVE_FLOATS);
  std::vector<T> Y(G::Y_ELEMS), P(G::P_ELEMS + 12 * 32);   // the pipelined scan reads (and ignores) a few bins past the end
  const int tiles = (n_frames + G::F - 1) / G::F;
  for (long c = 0; c < batch; ++c) {
    const float* clip = wave + c * stride;
    const int valid = lengths ? std::min<int>(lengths[c], n_samples) : n_samples;
    float* oc = out + c * (long)n_mels * n_frames;
    float cmax = -INFINITY;
    for (int t = 0; t < tiles; ++t) {
      const int f0 = t * G::F;
      const long s0 = (long)f0 * G::HOP - G::N / 2;
      for (int r = 0; r < G::SPAN; ++r)
        wave_s[wave_index<G>(r)] = load_sample(clip, s0 + r, n_samples, valid);
      // warp-specialised geometry: separate real / imaginary planes, stage 2 in place (its power
      // outputs overwrite the row's own slots of the real plane), mel reads them through scan_poff
      T* Yre = Y.data();
      T* Yim = Y.data() + G::YRE_ELEMS;
      for (int w = 0; w < G::NWK; ++w)
        for (int i = 0; i < G::S1_MAX && tab.s1_tasks[w][i] >= 0; ++i)
          for (int lane = 0; lane < 32; ++lane) {
            if (G::WS) stage1_task_ri<G, T>(wave_s.data(), Yre, Yim, tab.s1, tab.s1_tasks[w][i], lane);
            else stage1_task<G, T>(wave_s.data(), Y.data(), tab.s1, tab.s1_tasks[w][i], lane);
          }
      for (int w = 0; w < G::NWK; ++w)
        for (int i = 0; i < G::S2_MAX && tab.s2_rows[w][i] >= 0; ++i) {
          if (G::WS) {
            // a warp's loads all precede its stores: emulate with a snapshot of the row
            std::vector<T> snap(Y);
            for (int lane = 0; lane < 32; ++lane) {
              std::vector<T> work(snap);
              stage2_task_inplace<G, T>(work.data(), work.data() + G::YRE_ELEMS, tab.s2_rows[w][i], lane);
              const int k1 = tab.s2_rows[w][i];
              for (int j = 0; j < G::N2; ++j) Y[(k1 * G::N2 + j) * 32 + lane] = work[(k1 * G::N2 + j) * 32 + lane];
            }
          } else {
            for (int lane = 0; lane < 32; ++lane) stage2_task<G, T>(Y.data(), P.data(), tab.s2_rows[w][i], lane);
          }
        }
      const T* Pmel = G::WS ? Y.data() : P.data();
      for (int w = 0; w < G::NW_MEL; ++w)
        for (int lane = 0; lane < 32; ++lane) {
          int m = tab.mel_begin[w];
          auto emit = [&](T acc) {
            const float v[2] = {vlo(acc), vhi(acc)};
            for (int h = 0; h < G::PK; ++h) {
              const int f = f0 + lane + 32 * h;
              if (f >= n_frames) continue;
              const float s = fin_log(v[h], log_mode, log_param);
              oc[(long)m * n_frames + f] = s;
              cmax = std::max(cmax, s);
            }
            ++m;
          };
          if (G::WS && tab.mel_scan)   // the warp-specialised kernel's software-pipelined scan
            mel_task_pipe<G, T>(Pmel, tab, w, lane, [](T acc) { return acc; }, emit);
          else
            mel_task<G, T>(Pmel, tab, w, lane, emit);
        }
    }
    if (log_mode == LOG10_CLAMP_WHISPER_NORM) {
      const float thr = cmax - 8.0f;
      for (long i = 0; i < (long)n_mels * n_frames; ++i) oc[i] = (std::max(oc[i], thr) + 4.0f) * 0.25f;
    }
  }
  return 0;
}

}  // namespace

extern "C" int emul_logmel(int n_fft, int hop, int pk, const float* wave, long batch, long stride,
                           const int32_t* lengths, int n_samples, int n_frames, const float* fbank,
                           int n_mels, int log_mode, float log_param, float* out) {
#define CASE(N, H, K) \
  if (n_fft == N && hop == H && pk == K) \
    return run<lm::Geo<N, H, K>>(wave, batch, stride, lengths, n_samples, n_frames, fbank, n_mels, log_mode, log_param, out);
  CASE(400, 160, 1)
  CASE(400, 160, 2)
  if (n_fft == 400 && hop == 160 && pk == 3)   // pk 3: the warp-specialised geometry (two frames per lane)
    return run<lm::Geo<400, 160, 2, 1>>(wave, batch, stride, lengths, n_samples, n_frames, fbank, n_mels, log_mode, log_param, out);
  CASE(1024, 512, 1)
  CASE(1024, 128, 1)
#undef CASE
  return -2;
}
