// Core of the fused log-mel kernel: geometry, shared-memory layout and the three per-tile
// phases (stage 1, stage 2, mel/log), written as host+device inline functions over one
// (warp, lane) pair so that the host emulation in tests/ runs exactly the device index maps.
//
// Operator: for frame t of a clip,
//     X[t,k] = sum_n w[n] x~[t*hop + n] exp(-2 pi i k n / N),   x~ = reflect_pad(zero_pad(x, L), N/2)
//     P = |X|^2,  M = F^T P,  then log / normalise per lm_log_mode.
// Replaces transformers/models/whisper/feature_extraction_whisper.py:135-164 (and its NumPy twin
// :105-133 + audio_utils.py:769-830) and torchaudio/functional/functional.py:123-144 +
// torchaudio/transforms/_transforms.py:407-419 + /root/reference/.charles/spectrogram.py:161-162.
//
// Work decomposition (DESIGN.md "kernels"): a CTA owns a tile of F = 32*PK consecutive frames
// of one clip; lane l of every warp owns frame l (and frame l+32 when PK == 2, packed in
// f32x2).  Warps own *tasks*:
//   stage 1  task b  in [0,N2):      window + real DFT-N1 over samples N2*a+b, twiddle -> Y[b][k1]
//   stage 2  task k1 in [0,N1/2]:    complex DFT-N2 over b, |.|^2               -> P[k1 + N1*k2]
//   mel      task = a run of filters: banded F^T P, log, running max, global store
// Because the lane index is always the frame, every shared-memory access below is conflict
// free by construction (consecutive lanes touch consecutive words).
#pragma once
#include "codelets_gen.cuh"

namespace lm {

enum : int { LOG_NONE = 0, LOG10_CLAMP_WHISPER_NORM = 1, LN_PLUS_EPS = 2, LOG10_CLAMP = 3 };

constexpr int kMaxMels = 128;
constexpr int kMaxMelWeights = 2048;

// ---------------------------------------------------------------------------------------
// compile-time geometry
// ---------------------------------------------------------------------------------------
template <int NFFT> struct Split;
template <> struct Split<400> { static constexpr int N1 = 20, N2 = 20; };
template <> struct Split<1024> { static constexpr int N1 = 32, N2 = 32; };

template <int NFFT, int HOP_, int PK_>
struct Geo {
  static constexpr int N = NFFT, HOP = HOP_, PK = PK_;
  static constexpr int N1 = Split<NFFT>::N1, N2 = Split<NFFT>::N2;
  static constexpr int H1 = N1 / 2;                 // stage-2 tasks are k1 = 0..H1
  static constexpr int NBINS = N / 2 + 1;
  static constexpr int NW = H1;                     // warps per CTA
  static constexpr int THREADS = NW * 32;
  static constexpr int F = 32 * PK;                 // frames per tile
  static constexpr int SPAN = (F - 1) * HOP + N;    // samples a tile touches
  static constexpr int PITCH = HOP + 1;             // odd pitch: lane stride = 1 bank
  static constexpr int ROWS = (SPAN + HOP - 1) / HOP;
  static constexpr int WAVE_FLOATS = ROWS * PITCH;
  static constexpr int S1_STRIDE = N1 + 2 * H1;     // per-column constants: w[N1], (twr,twi)[1..H1]
  // Y: k1 = 0 is real (one plane), k1 = 1..H1 complex (re, im interleaved per lane)
  static constexpr int Y0_ELEMS = N2 * 32;               // in units of T
  static constexpr int Y_ELEMS = Y0_ELEMS + H1 * N2 * 32 * 2;
  static constexpr int P_ELEMS = NBINS * 32;
  static_assert(HOP % N2 == 0, "a column must not straddle a hop row");
  static_assert(N2 % NW == 0, "stage-1 columns must divide evenly over the warps");
};

// kernel parameters that live in the constant bank (__grid_constant__)
template <class G>
struct Tables {
  float s1[G::N2 * G::S1_STRIDE];        // stage-1 constants per column b
  float melw[kMaxMelWeights];            // banded filter weights, filter after filter
  unsigned short mel_lo[kMaxMels];       // first bin of each filter's support
  unsigned short mel_cnt[kMaxMels];      // support length
  unsigned short mel_off[kMaxMels];      // offset of its weights in melw
  unsigned short mel_begin[G::NW + 1];   // filters [mel_begin[w], mel_begin[w+1]) belong to warp w
};

// ---------------------------------------------------------------------------------------
// waveform tile: sample index -> value, with the reference's padding rules
// ---------------------------------------------------------------------------------------
// s is an index into the clip padded to n_samples (L); indices outside [0, L) reflect about
// the ends without repeating the edge sample (torch.stft center=True, pad_mode="reflect";
// np.pad(mode="reflect") in audio_utils.py:769-771); samples at or past `valid` are the zero
// padding of feature_extraction_sequence_utils.py:276-277 / spectrogram.py:152-157.
LM_HD float load_sample(const float* __restrict__ clip, long s, int n_samples, int valid) {
  if (s < 0) s = -s;
  if (s >= n_samples) s = 2L * (n_samples - 1) - s;
  if (s < 0 || s >= valid) return 0.0f;
  return clip[s];
}

template <class G> LM_HD int wave_index(int r) { return r + r / G::HOP; }

// ---------------------------------------------------------------------------------------
// loads / stores of the value type
// ---------------------------------------------------------------------------------------
template <typename T> struct VT;
template <> struct VT<float> {
  static LM_HD float load_wave(const float* p, int /*hi_off*/) { return p[0]; }
};
template <> struct VT<f32x2> {
  static LM_HD f32x2 load_wave(const float* p, int hi_off) { return vpack(p[0], p[hi_off]); }
};

template <class G> struct ValT { using type = float; };
template <int N, int H> struct ValT<Geo<N, H, 2>> { using type = f32x2; };

// ---------------------------------------------------------------------------------------
// codelet dispatch
// ---------------------------------------------------------------------------------------
template <int NFFT> struct Codelets;
template <> struct Codelets<400> {
  template <typename T>
  static LM_HD void s1(const T (&x)[20], const float (&w)[20], const float (&tr)[11],
                       const float (&ti)[11], T (&yr)[11], T (&yi)[11]) {
    stage1_r20(x, w, tr, ti, yr, yi);
  }
  template <typename T> static LM_HD void s2(const T (&yr)[20], const T (&yi)[20], T (&p)[20]) {
    stage2_c20(yr, yi, p);
  }
  template <typename T> static LM_HD void s2_half(const T (&yr)[20], const T (&yi)[20], T (&p)[10]) {
    stage2_c20_half(yr, yi, p);
  }
  template <typename T> static LM_HD void s2_real(const T (&yr)[20], T (&p)[11]) {
    stage2_r20_half(yr, p);
  }
};
template <> struct Codelets<1024> {
  template <typename T>
  static LM_HD void s1(const T (&x)[32], const float (&w)[32], const float (&tr)[17],
                       const float (&ti)[17], T (&yr)[17], T (&yi)[17]) {
    stage1_r32(x, w, tr, ti, yr, yi);
  }
  template <typename T> static LM_HD void s2(const T (&yr)[32], const T (&yi)[32], T (&p)[32]) {
    stage2_c32(yr, yi, p);
  }
  template <typename T> static LM_HD void s2_half(const T (&yr)[32], const T (&yi)[32], T (&p)[16]) {
    stage2_c32_half(yr, yi, p);
  }
  template <typename T> static LM_HD void s2_real(const T (&yr)[32], T (&p)[17]) {
    stage2_r32_half(yr, p);
  }
};

// ---------------------------------------------------------------------------------------
// stage 1: column b of this lane's frame(s)
// ---------------------------------------------------------------------------------------
// wave_s: tile in shared memory (wave_index layout), Y: [k1][b][lane] as described in Geo.
template <class G, typename T>
LM_HD void stage1_task(const float* __restrict__ wave_s, T* __restrict__ Y,
                       const float* __restrict__ s1tab, int b, int lane) {
  constexpr int N1 = G::N1, N2 = G::N2, H1 = G::H1;
  const float* cst = s1tab + b * G::S1_STRIDE;
  const float* src = wave_s + lane * G::PITCH + b;
  T x[N1];
  float w[N1], tr[H1 + 1], ti[H1 + 1];
#pragma unroll
  for (int a = 0; a < N1; ++a) {
    // sample n = N2*a + b of frame `lane`: r = HOP*lane + n, and n / HOP == (N2*a) / HOP
    x[a] = VT<T>::load_wave(src + (N2 * a + (N2 * a) / G::HOP), 32 * G::PITCH);
    w[a] = cst[a];
  }
  tr[0] = 1.0f;
  ti[0] = 0.0f;
#pragma unroll
  for (int k = 1; k <= H1; ++k) {
    tr[k] = cst[N1 + 2 * (k - 1)];
    ti[k] = cst[N1 + 2 * (k - 1) + 1];
  }
  T yr[H1 + 1], yi[H1 + 1];
  Codelets<G::N>::template s1<T>(x, w, tr, ti, yr, yi);
  Y[b * 32 + lane] = yr[0];
  T* yc = Y + G::Y0_ELEMS;
#pragma unroll
  for (int k = 1; k <= H1; ++k) {
    T* d = yc + (((k - 1) * N2 + b) * 32 + lane) * 2;
    d[0] = yr[k];
    d[1] = yi[k];
  }
}

// ---------------------------------------------------------------------------------------
// stage 2: row k1, all columns; writes |X|^2 for the bins k1 + N1*k2 (folded to <= N/2)
// ---------------------------------------------------------------------------------------
template <class G, typename T>
LM_HD void stage2_task(const T* __restrict__ Y, T* __restrict__ P, int k1, int lane) {
  constexpr int N1 = G::N1, N2 = G::N2, H1 = G::H1, N = G::N;
  if (k1 == 0) {
    T yr[N2], p[N2 / 2 + 1];
#pragma unroll
    for (int b = 0; b < N2; ++b) yr[b] = Y[b * 32 + lane];
    Codelets<N>::template s2_real<T>(yr, p);
#pragma unroll
    for (int j = 0; j <= N2 / 2; ++j) P[(N1 * j) * 32 + lane] = p[j];
    return;
  }
  T yr[N2], yi[N2];
  const T* yc = Y + G::Y0_ELEMS + ((k1 - 1) * N2 * 32 + lane) * 2;
#pragma unroll
  for (int b = 0; b < N2; ++b) {
    yr[b] = yc[b * 64];
    yi[b] = yc[b * 64 + 1];
  }
  if (k1 == H1) {
    T p[N2 / 2];
    Codelets<N>::template s2_half<T>(yr, yi, p);
#pragma unroll
    for (int j = 0; j < N2 / 2; ++j) P[(H1 + N1 * j) * 32 + lane] = p[j];
  } else {
    T p[N2];
    Codelets<N>::template s2<T>(yr, yi, p);
#pragma unroll
    for (int j = 0; j < N2; ++j) {
      // k = k1 + N1*j with 0 < k1 < N1/2: k > N/2 exactly when j >= N2/2, and then the bin is
      // the conjugate N - k (same power)
      const int bin = (j >= N2 / 2) ? (N - N1 * j - k1) : (N1 * j + k1);
      P[bin * 32 + lane] = p[j];
    }
  }
}

// ---------------------------------------------------------------------------------------
// mel projection for the filters owned by warp `w`: banded gather, one filter at a time
// ---------------------------------------------------------------------------------------
template <class G, typename T, class Emit>
LM_HD void mel_task(const T* __restrict__ P, const Tables<G>& tab, int w, int lane, Emit&& emit) {
  const int m0 = tab.mel_begin[w], m1 = tab.mel_begin[w + 1];
  for (int m = m0; m < m1; ++m) {
    const int lo = tab.mel_lo[m], cnt = tab.mel_cnt[m];
    const float* wp = tab.melw + tab.mel_off[m];
    const T* src = P + lo * 32 + lane;
    T acc = vzero<T>();
    for (int j = 0; j < cnt; ++j) acc = vfmas(src[j * 32], wp[j], acc);
    emit(m, acc);
  }
}

}  // namespace lm
