// Host-side construction of the constant tables the kernel reads from its parameter bank.
// Plain C++ (no CUDA): shared by the C-ABI library and by the host emulation in tests/.
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "logmel_core.cuh"

namespace lm {

// Periodic Hann, as torch.hann_window(N) (feature_extraction_whisper.py:141;
// torchaudio Spectrogram's default window_fn) -- computed in double, rounded once.
inline std::vector<float> hann_periodic(int n) {
  std::vector<float> w(n);
  for (int i = 0; i < n; ++i) w[i] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)i / (double)n));
  return w;
}

// Fill Tables<G> from a window [N] and a dense filter bank [NBINS][n_mels] (row-major, the
// layout of WhisperFeatureExtractor.mel_filters and torchaudio's MelScale.fb).
// Returns an empty string on success, else the reason the bank is not supported.
template <class G>
std::string build_tables(Tables<G>& t, const float* window, const float* fbank, int n_mels) {
  std::memset(&t, 0, sizeof(t));
  if (n_mels < 1 || n_mels > kMaxMels) return "n_mels must be in [1, 128]";
  for (int b = 0; b < G::N2; ++b) {
    float* c = t.s1 + b * G::S1_STRIDE;
    for (int a = 0; a < G::N1; ++a) c[a] = window[G::N2 * a + b];
    for (int k = 1; k <= G::H1; ++k) {
      const double ang = -2.0 * M_PI * (double)b * (double)k / (double)G::N;
      c[G::N1 + 2 * (k - 1)] = (float)std::cos(ang);
      c[G::N1 + 2 * (k - 1) + 1] = (float)std::sin(ang);
    }
  }
  int off = 0;
  std::vector<int> cost(n_mels);
  for (int m = 0; m < n_mels; ++m) {
    int lo = -1, hi = -1;
    for (int k = 0; k < G::NBINS; ++k) {
      if (fbank[(size_t)k * n_mels + m] != 0.0f) {
        if (lo < 0) lo = k;
        hi = k + 1;
      }
    }
    if (lo < 0) lo = hi = 0;
    const int cnt = hi - lo;
    if (off + cnt > kMaxMelWeights) return "filter bank is not banded enough (more than 2048 weights)";
    t.mel_lo[m] = (unsigned short)lo;
    t.mel_cnt[m] = (unsigned short)cnt;
    t.mel_off[m] = (unsigned short)off;
    for (int j = 0; j < cnt; ++j) t.melw[off + j] = fbank[(size_t)(lo + j) * n_mels + m];
    off += cnt;
    cost[m] = cnt + 8;   // + clamp/log/max/store
  }
  // contiguous runs of filters per warp, balanced on cost
  long total = 0;
  for (int m = 0; m < n_mels; ++m) total += cost[m];
  int m = 0;
  long acc = 0;
  t.mel_begin[0] = 0;
  for (int w = 0; w < G::NW; ++w) {
    const long target = total * (w + 1) / G::NW;
    while (m < n_mels && (acc + cost[m] / 2 <= target || w == G::NW - 1)) acc += cost[m++];
    t.mel_begin[w + 1] = (unsigned short)m;
  }
  t.mel_begin[G::NW] = (unsigned short)n_mels;
  return std::string();
}

}  // namespace lm
