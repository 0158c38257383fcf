// Host-side construction of the constant tables the kernel reads from its parameter bank.
// Plain C++ (no CUDA): shared by the C-ABI library and by the host emulation in tests/.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "logmel_core.cuh"

// Cost of a finished filter relative to 7 per scanned bin, for the split of the filters over the mel
// warps.  The mel warps are co-critical in the warp-specialised CTA, so the split matters: measured
// on 4096 clips x 128 mels, cost 8 / 12 / 14 / 16 / 18 / 22 -> 5.49 / 5.42 / 5.44 / 5.50 / 5.67 / 5.89 ms.
#ifndef LM_MEL_EMIT_COST
#define LM_MEL_EMIT_COST 12
#endif

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
  // ---- stage-1 / stage-2 work tables, balanced per SM sub-partition (warp id mod 4)
  {
    std::memset(t.s1_tasks, -1, sizeof(t.s1_tasks));
    std::memset(t.s2_rows, -1, sizeof(t.s2_rows));
    std::vector<int> n1(G::NWK, 0), n2(G::NWK, 0);
    std::vector<long> load(4, 0);
    auto warp_of = [&](int part, const std::vector<int>& cnt, int cap) {   // least-busy warp of a partition
      int best = -1;
      for (int w = part; w < G::NW_FFT; w += 4)
        if (cnt[w] < cap && (best < 0 || cnt[w] < cnt[best])) best = w;
      return best;
    };
    if (G::S1_CONST_REGS) {
      // a warp's tasks share one column group (so its constants can stay in registers): the 4
      // frame octets of column group cg go to two warps, 2 + 2 ...
      int w = 0;
      for (int cg = 0; cg < G::CGROUPS; ++cg)
        for (int half = 0; half < 2; ++half, ++w) {
          if (w >= G::NWK) return "internal: stage-1 work table overflow";
          for (int i = 0; i < 2; ++i) t.s1_tasks[w][n1[w]++] = (signed char)((2 * half + i) * G::CGROUPS + cg);
        }
      // ... then warps that are still idle take one task from the busiest sub-partition until
      // every sub-partition (warp id mod 4) carries the same number of tasks
      for (int idle = w; idle < G::NWK; ++idle) {
        int load[4] = {0, 0, 0, 0};
        for (int q = 0; q < G::NWK; ++q) load[q % 4] += n1[q];
        int donor = -1;
        for (int q = 0; q < idle; ++q)
          if (n1[q] == 2 && q % 4 != idle % 4 && load[q % 4] > load[idle % 4] + 1 &&
              (donor < 0 || load[q % 4] > load[donor % 4] || (load[q % 4] == load[donor % 4] && q > donor)))
            donor = q;
        if (donor < 0) break;
        t.s1_tasks[idle][n1[idle]++] = t.s1_tasks[donor][--n1[donor]];
        t.s1_tasks[donor][n1[donor]] = -1;
      }
    } else {
      for (int task = 0; task < G::S1_TASKS; ++task) {     // equal-cost tasks: round robin over partitions
        int part = task % 4;
        int w = warp_of(part, n1, G::S1_MAX);
        if (w < 0) return "internal: stage-1 work table overflow";
        t.s1_tasks[w][n1[w]++] = (signed char)task;
      }
    }
    // rows by decreasing cost: general rows, then the half row k1 = H1, then the real row k1 = 0
    std::vector<std::pair<int, int>> rows;                  // (cost, k1)
    for (int k = 1; k < G::H1; ++k) rows.push_back({100, k});
    rows.push_back({77, G::H1});
    rows.push_back({44, 0});
    for (auto& r : rows) {
      int part = 0;
      for (int q = 1; q < 4; ++q)
        if (load[q] < load[part]) part = q;
      int w = warp_of(part, n2, G::S2_MAX);
      if (w < 0) return "internal: stage-2 work table overflow";
      t.s2_rows[w][n2[w]++] = (signed char)r.second;
      load[part] += r.first;
    }
    // A single issuing thread needs ~60 cycles per bulk copy (measured: 66 copies = 4 200 cycles,
    // longer than stage 2 itself), so the copies stay spread over all warps.
    t.loader_warp = -1;
    {
      int n = 0;
      for (int w = 0; w < G::NWK; ++w) t.tma_iss[w] = -1;
      if (G::WS)
        for (int w = 0; w < G::NW_FFT; ++w)
          if (n2[w] <= 1) t.tma_iss[w] = (signed char)n++;
      if (n < 3) {                      // too few light warps: everybody issues
        n = 0;
        for (int w = 0; w < G::NWK; ++w) t.tma_iss[w] = -1;
      }
      t.n_tma_iss = (signed char)n;
    }
  }
  // ---- banded supports
  std::vector<int> lo(n_mels, 0), cnt(n_mels, 0);
  for (int m = 0; m < n_mels; ++m) {
    int first = -1, last = -1;
    for (int k = 0; k < G::NBINS; ++k)
      if (fbank[(size_t)k * n_mels + m] != 0.0f) {
        if (first < 0) first = k;
        last = k;
      }
    if (first >= 0) {
      lo[m] = first;
      cnt[m] = last - first + 1;
    }
  }
  // ---- scan form: possible when every bin feeds at most two adjacent filters, in order
  std::vector<int> a_of(G::NBINS, 0);
  bool scan_ok = true;
  {
    int prev = 0;
    for (int k = 0; k < G::NBINS && scan_ok; ++k) {
      int first = -1, last = -1;
      for (int m = 0; m < n_mels; ++m)
        if (fbank[(size_t)k * n_mels + m] != 0.0f) {
          if (first < 0) first = m;
          last = m;
        }
      if (first < 0) {
        a_of[k] = prev;
      } else {
        if (last - first > 1 || first < prev) scan_ok = false;
        a_of[k] = first;
        prev = first;
      }
    }
  }
  std::vector<int> nbin_of(n_mels + 1, 0);           // bins with a(k) == m
  if (scan_ok)
    for (int k = 0; k < G::NBINS; ++k) nbin_of[a_of[k]]++;
  // ---- contiguous runs of filters per warp, balanced on their cost
  std::vector<long> cost(n_mels);
  long total = 0;
  for (int m = 0; m < n_mels; ++m)
    total += (cost[m] = scan_ok ? 7L * nbin_of[m] + LM_MEL_EMIT_COST : 9L * std::max(1, (cnt[m] + 3) / 4) + 14);
  {
    int m = 0;
    long acc = 0;
    t.mel_begin[0] = 0;
    for (int w = 0; w < G::NW_MEL; ++w) {
      const long target = total * (w + 1) / G::NW_MEL;
      while (m < n_mels && (acc + cost[m] / 2 <= target || w == G::NW_MEL - 1)) acc += cost[m++];
      t.mel_begin[w + 1] = (unsigned short)m;
    }
    for (int w = G::NW_MEL; w <= G::NWK; ++w) t.mel_begin[w] = (unsigned short)n_mels;
  }
  if (scan_ok) {
    // a(k) may advance by at most two between neighbouring bins (one filter without a bin of
    // its own); anything wider falls back to the gather form
    for (int k = 0; k + 1 < G::NBINS; ++k)
      if (a_of[k + 1] - a_of[k] > 2) scan_ok = false;
  }
  if (scan_ok) {
    Tables<G> keep = t;                      // the gather form below starts from this state
    t.mel_scan = 1;
    for (int k = 0; k < G::NBINS; ++k) {
      const int a = a_of[k];
      t.melw[2 * k] = fbank[(size_t)k * n_mels + a];
      t.melw[2 * k + 1] = (a + 1 < n_mels) ? fbank[(size_t)k * n_mels + a + 1] : 0.0f;
    }
    int soff = 0;
    bool ok = true;
    for (int w = 0; w < G::NW_MEL && ok; ++w) {
      const int m0 = t.mel_begin[w], m1 = t.mel_begin[w + 1];
      t.scan_soff[w] = (unsigned short)soff;
      if (m1 <= m0) continue;
      // bins with a(k) in [m0 - 1, m1 - 1] are contiguous because a(k) never decreases
      int k0 = 0;
      while (k0 < G::NBINS && a_of[k0] < m0 - 1) ++k0;
      int k1 = k0;
      while (k1 < G::NBINS && a_of[k1] <= m1 - 1) ++k1;
      if (k1 <= k0) { ok = false; break; }
      t.scan_bin0[w] = (unsigned short)k0;
      t.scan_nb[w] = (unsigned short)(k1 - k0);
      if (soff + (k1 - k0) + 4 > kMaxScanSteps) { ok = false; break; }
      int next_filter = m0;                   // the filter the next emitting shift produces
      for (int k = k0; k < k1; ++k) {
        const int cur = a_of[k];
        const int shifts = (k + 1 < k1) ? a_of[k + 1] - cur : (m1 - 1) - cur + 1;
        if (shifts < 0 || shifts > 2) { ok = false; break; }
        unsigned code = 0;
        for (int j = 0; j < shifts; ++j) {
          const int f = cur + j;               // filter leaving acc0
          code |= 1u << j;
          if (f < m0) code |= 4u << j;         // lead-in: belongs to the previous warp
          else if (f == next_filter) ++next_filter;
          else ok = false;
        }
        if (G::WS) t.scan_poff[soff] = (unsigned)(p_slot<G>(k) * 32 * (int)sizeof(typename ValT<G>::type));
        t.scan_code[soff++] = (unsigned char)code;
      }
      if (next_filter != m1) ok = false;
      soff = (soff + 3) & ~3;
    }
    if (ok) return std::string();
    t = keep;                                 // irregular bank: use the gather form
  }
  // ---- gather form: grouped, zero padded weights; a run uses the group count of its widest filter
  int off = 0;
  for (int w = 0; w < G::NW_MEL; ++w) {
    const int m0 = t.mel_begin[w], m1 = t.mel_begin[w + 1];
    int ng = 1;
    for (int m = m0; m < m1; ++m) ng = std::max(ng, (cnt[m] + 3) / 4);
    if (4 * ng > G::NBINS) return "filter bank is not banded (a support is wider than the spectrum)";
    if (off + (m1 - m0) * 4 * ng > kMaxMelWeights)
      return "filter bank is not banded enough (grouped weights exceed 3072)";
    t.mel_ng[w] = (unsigned short)ng;
    t.mel_woff[w] = (unsigned short)(off / 4);
    for (int m = m0; m < m1; ++m) {
      // keep the 4*ng bins that are read inside the spectrum: shift the window down if needed
      int start = lo[m];
      if (start + 4 * ng > G::NBINS) start = G::NBINS - 4 * ng;
      t.mel_lo[m] = (unsigned short)start;
      for (int j = 0; j < 4 * ng; ++j) {
        const int k = start + j;
        t.melw[off + j] = (k >= lo[m] && k < lo[m] + cnt[m]) ? fbank[(size_t)k * n_mels + m] : 0.0f;
      }
      off += 4 * ng;
    }
  }
  return std::string();
}

}  // namespace lm
