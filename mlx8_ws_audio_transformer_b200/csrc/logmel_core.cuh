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
constexpr int kMaxMelWeights = 3072;
constexpr int kMaxScanSteps = 1280;

// ---------------------------------------------------------------------------------------
// compile-time geometry
// ---------------------------------------------------------------------------------------
template <int NFFT> struct Split;
template <> struct Split<400> { static constexpr int N1 = 20, N2 = 20; };
template <> struct Split<1024> { static constexpr int N1 = 32, N2 = 32; };

// WS_ = 1: warp-specialised CTA (logmel_ws_kernel.cuh): warps [0, NW_FFT) run the two DFT stages,
// warps [NW_FFT, NWK) run the mel projection / log / stores.
template <int NFFT, int HOP_, int PK_, int WS_ = 0>
struct Geo {
  static constexpr int N = NFFT, HOP = HOP_, PK = PK_, WS = WS_;
  static constexpr int N1 = Split<NFFT>::N1, N2 = Split<NFFT>::N2;
  static constexpr int H1 = N1 / 2;                 // stage-2 tasks are k1 = 0..H1
  static constexpr int NBINS = N / 2 + 1;
  // Warps per CTA: a multiple of 4 so that every SM sub-partition (warp id mod 4) holds the
  // same number of warps; the work tables below balance tasks per sub-partition.
  static constexpr int NWK = (NFFT == 400) ? 12 : 16;
  static constexpr int NW_FFT = WS_ ? 8 : NWK;      // warps that take stage-1 tasks and stage-2 rows
  static constexpr int NW_MEL = WS_ ? NWK - 8 : NWK;  // warps that take runs of mel filters
  static constexpr int MEL_W0 = WS_ ? 8 : 0;        // first of them
  // N = 400: a warp's stage-1 tasks all use the same four columns, so their window samples and
  // twiddles (S1_STRIDE floats per lane) are loaded once per kernel and stay in registers; the
  // alternative -- fetching them from shared memory per task -- costs as many LSU wavefronts as
  // the waveform samples themselves.  N = 1024 has too many constants (64) for that, and the
  // warp-specialised CTA spreads 20 tasks over 8 warps, which does not split by column group.
  static constexpr bool S1_CONST_REGS = (NFFT == 400) && !WS_;
  static constexpr int THREADS = NWK * 32;
  static constexpr int F = 32 * PK;                 // frames per tile
  static constexpr int SPAN = (F - 1) * HOP + N;    // samples a tile touches
  // Waveform rows of HOP samples, 4 pad words each: every row starts 16-byte aligned (one TMA
  // bulk copy per row) and the row stride is 4 banks, which the stage-1 lane map (8 frames x 4
  // adjacent columns per warp) turns into 32 distinct banks.
  static constexpr int PITCH = HOP + 4;
  static constexpr int ROWS = (SPAN + HOP - 1) / HOP;
  static constexpr int WAVE_FLOATS = ROWS * PITCH;
  static constexpr int S1_STRIDE = N1 + 2 * H1;     // per-column constants: w[N1], (twr,twi)[1..H1]
  // Y (units of T): one real plane [k1 = 0..H1][b][lane], then one imaginary plane
  // [k1 = 1..H1][b][lane] (row k1 = 0 is purely real).  Separate planes let stage 1 store each
  // 64-bit packed value straight from the register pair it was computed in.
  static constexpr int YRE_ELEMS = (H1 + 1) * N2 * 32;
  static constexpr int Y_ELEMS = YRE_ELEMS + H1 * N2 * 32;
  static constexpr int P_ELEMS = NBINS * 32;
  static_assert(HOP % N2 == 0, "a column must not straddle a hop row");
  static_assert(S1_STRIDE % 4 == 0, "stage-1 constants are fetched as float4");
  static_assert(HOP % 32 == 0, "the row pitch must be 4 banks past a multiple of 32");
  static constexpr int CGROUPS = N2 / 4;            // stage-1 tasks: (4 frame octets) x (N2/4 column quads)
  static constexpr int S1_TASKS = 4 * CGROUPS;
  static constexpr int S1_MAX = 3, S2_MAX = 2;      // most tasks / rows one warp can be handed
  static_assert(N2 % 4 == 0, "columns are handled four at a time");
  static_assert(S1_TASKS <= NW_FFT * S1_MAX && H1 + 1 <= NW_FFT * S2_MAX, "work tables too small");
};

// kernel parameters that live in the constant bank (__grid_constant__)
//
// Mel projection layout: warp w owns filters [mel_begin[w], mel_begin[w+1]).  Every filter of
// that run is stored as mel_ng[w] groups of 4 consecutive-bin weights (zero padded), starting
// at bin mel_lo[m]; its weights sit at melw[4 * (mel_woff[w] + (m - mel_begin[w]) * mel_ng[w])].
// A fixed group count per warp keeps the inner loop free of per-filter control flow, and one
// 128-bit constant load fetches the four weights of a group.
template <class G>
struct alignas(16) Tables {
  float s1[G::N2 * G::S1_STRIDE];        // stage-1 constants per column b (S1_STRIDE % 4 == 0)
  float melw[kMaxMelWeights];            // grouped filter weights
  unsigned short mel_lo[kMaxMels];       // first bin read for each filter
  unsigned short mel_begin[G::NWK + 1];
  unsigned short mel_ng[G::NWK];
  unsigned short mel_woff[G::NWK];       // in units of 4 weights
  // Scan form of the mel projection (mel_scan != 0; triangular banks, where every bin feeds at
  // most two ADJACENT filters a(k), a(k)+1 and a(k) never decreases): melw then holds one weight
  // pair per bin, melw[2k] for filter a(k) and melw[2k+1] for a(k)+1.  Warp w walks scan_nb[w]
  // consecutive bins from scan_bin0[w] on with two running sums (acc0: filter a, acc1: filter
  // a+1); scan_code (from scan_soff[w] on, one byte per bin) says what happens AFTER a bin:
  //   bit 0  first shift:  filter a is complete -> emit acc0 (unless bit 2), acc0 = acc1, acc1 = 0
  //   bit 1  second shift: a(k) jumps by two, filter a+1 has no bin of its own -> emit acc0
  //          (unless bit 3), acc0 = 0
  //   bit 2 / bit 3  the shifted-out filter belongs to the previous warp (lead-in): no emit
  // Every power value is read from shared memory once per warp instead of once per filter, and
  // the walk is one flat loop (no per-filter inner loops).
  unsigned short scan_bin0[G::NWK];
  unsigned short scan_soff[G::NWK];         // multiple of 4
  unsigned short scan_nb[G::NWK];
  alignas(4) unsigned char scan_code[kMaxScanSteps];   // read four codes at a time
  // Warp-specialised CTA only: where the power of each scanned bin lives.  Stage 2 overwrites
  // its own row of the real plane of Y with its |X|^2 outputs (row k1, slot j), so bin k is at
  // row-slot p_slot(k); scan_poff holds that position in BYTES, one entry per scan step.
  alignas(16) unsigned scan_poff[G::WS ? kMaxScanSteps : 4];
  unsigned char mel_scan;
  signed char s1_tasks[G::NWK][G::S1_MAX];   // stage-1 tasks of each warp (-1: none)
  signed char s2_rows[G::NWK][G::S2_MAX];    // stage-2 rows k1 of each warp (-1: none)
  signed char loader_warp;                   // a warp without stage-2 rows issues the tile's TMA copies (-1: none)
  // warp-specialised CTA: the FFT warps with ONE stage-2 row issue the TMA copies of the next tile
  // (tma_iss: their number 0..n_tma_iss-1, -1 for the others) while the two-row warps start stage 2
  signed char tma_iss[G::NWK];
  signed char n_tma_iss;
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

// The same for 16-bit PCM input (fused ingest): `pcm` holds interleaved frames of
// `ch` channels (1 or 2), and the sample is what the reference's loaders make of them --
// torchaudio.load's s / 32768 (/root/reference/AB/wavToWhisper.py:52, AB/memoToWav.py:19 writes s16 mono)
// and the mono mix waveform.mean(dim=0) of /root/reference/.charles/spectrogram.py:147-148.  Sums of
// two int16 and the scale by a power of two are exact in float32, so this equals the float path bit
// for bit.
LM_HD float pcm_to_float(int sum, int ch) { return (float)sum * (ch == 2 ? (1.0f / 65536.0f) : (1.0f / 32768.0f)); }
LM_HD float load_sample_pcm(const short* __restrict__ pcm, int ch, long s, int n_samples, int valid) {
  if (s < 0) s = -s;
  if (s >= n_samples) s = 2L * (n_samples - 1) - s;
  if (s < 0 || s >= valid) return 0.0f;
  return ch == 2 ? pcm_to_float((int)pcm[2 * s] + (int)pcm[2 * s + 1], 2) : pcm_to_float((int)pcm[s], 1);
}
LM_HD float load_sample_any(const float* __restrict__ clip, const short* __restrict__ pcm, int ch, long s, int n_samples,
                            int valid) {
  return pcm ? load_sample_pcm(pcm, ch, s, n_samples, valid) : load_sample(clip, s, n_samples, valid);
}

template <class G> LM_HD int wave_index(int r) { return r + 4 * (r / G::HOP); }

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
template <int N, int H, int W> struct ValT<Geo<N, H, 2, W>> { using type = f32x2; };

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
// stage 1: one task = 8 frame slots x 4 adjacent columns
// ---------------------------------------------------------------------------------------
// Lane l of the warp handles frame slot p = 8*fg + (l & 7) (frames p and p+32 when packed) and
// column b = 4*cg + (l >> 3).  With the 4-bank row stride the 32 lanes read 32 distinct banks.
// Y[k1][b][slot]: the slot is XOR-swizzled with the column (slot = p ^ 8*(b & 3)) so that these
// writes (4 columns x 8 slots) and the stage-2 reads (one column, 32 slots) are both conflict
// free without padding.  s1tab may live in shared memory (device) or anywhere (host emulation).
template <class G> LM_HD int y_slot(int p, int b) { return p ^ (8 * (b & 3)); }

template <class G>
LM_HD void stage1_consts(const float* __restrict__ s1tab, int task, int lane, float (&cst)[G::S1_STRIDE]) {
  // window samples and twiddles of this lane's column: S1_STRIDE floats, fetched 16 bytes at a time
  const int b = 4 * (task % G::CGROUPS) + (lane >> 3);
  const float4* c4 = reinterpret_cast<const float4*>(s1tab + b * G::S1_STRIDE);
#pragma unroll
  for (int i = 0; i < G::S1_STRIDE / 4; ++i) {
    const float4 v = c4[i];
    cst[4 * i] = v.x; cst[4 * i + 1] = v.y; cst[4 * i + 2] = v.z; cst[4 * i + 3] = v.w;
  }
}

template <class G, typename T>
LM_HD void stage1_load(const float* __restrict__ wave_s, int task, int lane, T (&x)[G::N1]) {
  const int b = 4 * (task % G::CGROUPS) + (lane >> 3);
  const int p = 8 * (task / G::CGROUPS) + (lane & 7);
  const float* src = wave_s + p * G::PITCH + b;
#pragma unroll
  for (int a = 0; a < G::N1; ++a)   // sample n = N2*a + b of frame p: r = HOP*p + n, and n / HOP == (N2*a) / HOP
    x[a] = VT<T>::load_wave(src + (G::N2 * a + 4 * ((G::N2 * a) / G::HOP)), 32 * G::PITCH);
}

template <class G, typename T>
LM_HD void stage1_compute(const T (&x)[G::N1], T* __restrict__ Y, const float (&cst)[G::S1_STRIDE], int task, int lane) {
  constexpr int N1 = G::N1, N2 = G::N2, H1 = G::H1;
  const int b = 4 * (task % G::CGROUPS) + (lane >> 3);
  const int p = 8 * (task / G::CGROUPS) + (lane & 7);
  float w[N1], tr[H1 + 1], ti[H1 + 1];
#pragma unroll
  for (int a = 0; a < N1; ++a) w[a] = cst[a];
  tr[0] = 1.0f;
  ti[0] = 0.0f;
#pragma unroll
  for (int k = 1; k <= H1; ++k) {
    tr[k] = cst[N1 + 2 * (k - 1)];
    ti[k] = cst[N1 + 2 * (k - 1) + 1];
  }
  T yr[H1 + 1], yi[H1 + 1];
  Codelets<G::N>::template s1<T>(x, w, tr, ti, yr, yi);
  T* dre = Y + b * 32 + y_slot<G>(p, b);
  T* dim = dre + G::YRE_ELEMS;
#pragma unroll
  for (int k = 0; k <= H1; ++k) dre[k * N2 * 32] = yr[k];
#pragma unroll
  for (int k = 1; k <= H1; ++k) dim[(k - 1) * N2 * 32] = yi[k];
}

template <class G, typename T>
LM_HD void stage1_task_c(const float* __restrict__ wave_s, T* __restrict__ Y,
                         const float (&cst)[G::S1_STRIDE], int task, int lane) {
  T x[G::N1];
  stage1_load<G, T>(wave_s, task, lane, x);
  stage1_compute<G, T>(x, Y, cst, task, lane);
}

// two tasks of one warp: the second task's samples are fetched before the first task's
// arithmetic starts, so the shared-memory latency and the FP32 work overlap
template <class G, typename T>
LM_HD void stage1_task_pair(const float* __restrict__ wave_s, T* __restrict__ Y,
                            const float (&cst)[G::S1_STRIDE], int ta, int tb, int lane) {
  T xa[G::N1], xb[G::N1];
  stage1_load<G, T>(wave_s, ta, lane, xa);
  stage1_load<G, T>(wave_s, tb, lane, xb);
  stage1_compute<G, T>(xa, Y, cst, ta, lane);
  stage1_compute<G, T>(xb, Y, cst, tb, lane);
}

template <class G, typename T>
LM_HD void stage1_task(const float* __restrict__ wave_s, T* __restrict__ Y,
                       const float* __restrict__ s1tab, int task, int lane) {
  float cst[G::S1_STRIDE];
  stage1_consts<G>(s1tab, task, lane, cst);
  stage1_task_c<G, T>(wave_s, Y, cst, task, lane);
}

// Same, with the real and imaginary planes of Y given separately (warp-specialised CTA: the
// real plane is double buffered, see logmel_ws_kernel.cuh).
template <class G, typename T>
LM_HD void stage1_compute_ri(const T (&x)[G::N1], T* __restrict__ Yre, T* __restrict__ Yim,
                             const float (&cst)[G::S1_STRIDE], int task, int lane) {
  constexpr int N1 = G::N1, N2 = G::N2, H1 = G::H1;
  const int b = 4 * (task % G::CGROUPS) + (lane >> 3);
  const int p = 8 * (task / G::CGROUPS) + (lane & 7);
  float w[N1], tr[H1 + 1], ti[H1 + 1];
#pragma unroll
  for (int a = 0; a < N1; ++a) w[a] = cst[a];
  tr[0] = 1.0f;
  ti[0] = 0.0f;
#pragma unroll
  for (int k = 1; k <= H1; ++k) {
    tr[k] = cst[N1 + 2 * (k - 1)];
    ti[k] = cst[N1 + 2 * (k - 1) + 1];
  }
  T yr[H1 + 1], yi[H1 + 1];
  Codelets<G::N>::template s1<T>(x, w, tr, ti, yr, yi);
  T* dre = Yre + b * 32 + y_slot<G>(p, b);
  T* dim = Yim + b * 32 + y_slot<G>(p, b);
#pragma unroll
  for (int k = 0; k <= H1; ++k) dre[k * N2 * 32] = yr[k];
#pragma unroll
  for (int k = 1; k <= H1; ++k) dim[(k - 1) * N2 * 32] = yi[k];
}

template <class G, typename T>
LM_HD void stage1_task_ri(const float* __restrict__ wave_s, T* __restrict__ Yre, T* __restrict__ Yim,
                          const float* __restrict__ s1tab, int task, int lane) {
  float cst[G::S1_STRIDE];
  stage1_consts<G>(s1tab, task, lane, cst);
  T x[G::N1];
  stage1_load<G, T>(wave_s, task, lane, x);
  stage1_compute_ri<G, T>(x, Yre, Yim, cst, task, lane);
}

// row-slot (units of 32 lanes of T, from the start of the real plane) at which the in-place
// stage 2 leaves the power of bin k: its own row k1, output index j
template <class G> LM_HD int p_slot(int k) {
  constexpr int N1 = G::N1, N2 = G::N2, H1 = G::H1, N = G::N;
  const int c = k % N1, q = k / N1;
  if (c == 0) return q;                                   // row 0: bins N1 * j
  if (c < H1) return c * N2 + q;                          // row c, j = q < N2 / 2
  if (c == H1) return H1 * N2 + q;                        // row H1: bins H1 + N1 * j
  return (N1 - c) * N2 + ((N - k - (N1 - c)) / N1);       // conjugate side of row N1 - c: j >= N2 / 2
}

// Stage 2 of row k1 IN PLACE: the row's inputs are read into registers, then its power outputs
// overwrite the row's own slots of the real plane (output j at slot j).  No other warp touches
// this row, and a warp's shared-memory accesses are performed in program order.
template <class G, typename T>
LM_HD void stage2_task_inplace(T* __restrict__ Yre, const T* __restrict__ Yim, int k1, int lane) {
  constexpr int N2 = G::N2, H1 = G::H1, N = G::N;
  T* row = Yre + k1 * N2 * 32;
  if (k1 == 0) {
    T yr[N2], p[N2 / 2 + 1];
#pragma unroll
    for (int b = 0; b < N2; ++b) yr[b] = row[b * 32 + y_slot<G>(lane, b)];
    Codelets<N>::template s2_real<T>(yr, p);
#pragma unroll
    for (int j = 0; j <= N2 / 2; ++j) row[j * 32 + lane] = p[j];
    return;
  }
  T yr[N2], yi[N2];
  const T* sim = Yim + (k1 - 1) * N2 * 32;
#pragma unroll
  for (int b = 0; b < N2; ++b) {
    yr[b] = row[b * 32 + y_slot<G>(lane, b)];
    yi[b] = sim[b * 32 + y_slot<G>(lane, b)];
  }
  if (k1 == H1) {
    T p[N2 / 2];
    Codelets<N>::template s2_half<T>(yr, yi, p);
#pragma unroll
    for (int j = 0; j < N2 / 2; ++j) row[j * 32 + lane] = p[j];
  } else {
    T p[N2];
    Codelets<N>::template s2<T>(yr, yi, p);
#pragma unroll
    for (int j = 0; j < N2; ++j) row[j * 32 + lane] = p[j];
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
    for (int b = 0; b < N2; ++b) yr[b] = Y[b * 32 + y_slot<G>(lane, b)];
    Codelets<N>::template s2_real<T>(yr, p);
#pragma unroll
    for (int j = 0; j <= N2 / 2; ++j) P[(N1 * j) * 32 + lane] = p[j];
    return;
  }
  T yr[N2], yi[N2];
  const T* sre = Y + k1 * N2 * 32;
  const T* sim = Y + G::YRE_ELEMS + (k1 - 1) * N2 * 32;
#pragma unroll
  for (int b = 0; b < N2; ++b) {
    yr[b] = sre[b * 32 + y_slot<G>(lane, b)];
    yi[b] = sim[b * 32 + y_slot<G>(lane, b)];
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
// mel projection for the filters owned by warp `w`: grouped banded gather, two filters in flight
// ---------------------------------------------------------------------------------------
// NG > 0: group count known at compile time (straight-line code, the two filters of a pair
// interleaved for ILP); NG == 0: run-time group count.  w4 points at the pair's weights:
// ng float4 for the first filter, then ng float4 for the second.
template <int NG, typename T>
LM_HD void mel_dot2(const T* __restrict__ s0, const T* __restrict__ s1, const float4* __restrict__ w4,
                    int ng, T& a0, T& a1) {
  a0 = vzero<T>();
  a1 = vzero<T>();
  if (NG > 0) {
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const float4 u = w4[g], v = w4[NG + g];
      a0 = vfmas(s0[g * 128], u.x, a0);
      a1 = vfmas(s1[g * 128], v.x, a1);
      a0 = vfmas(s0[g * 128 + 32], u.y, a0);
      a1 = vfmas(s1[g * 128 + 32], v.y, a1);
      a0 = vfmas(s0[g * 128 + 64], u.z, a0);
      a1 = vfmas(s1[g * 128 + 64], v.z, a1);
      a0 = vfmas(s0[g * 128 + 96], u.w, a0);
      a1 = vfmas(s1[g * 128 + 96], v.w, a1);
    }
  } else {
#pragma unroll 1
    for (int g = 0; g < ng; ++g) {
      const float4 u = w4[g], v = w4[ng + g];
      a0 = vfmas(s0[g * 128], u.x, a0);
      a1 = vfmas(s1[g * 128], v.x, a1);
      a0 = vfmas(s0[g * 128 + 32], u.y, a0);
      a1 = vfmas(s1[g * 128 + 32], v.y, a1);
      a0 = vfmas(s0[g * 128 + 64], u.z, a0);
      a1 = vfmas(s1[g * 128 + 64], v.z, a1);
      a0 = vfmas(s0[g * 128 + 96], u.w, a0);
      a1 = vfmas(s1[g * 128 + 96], v.w, a1);
    }
  }
}

// emit(acc) is called once per filter, in filter order
template <int NG, class G, typename T, class Emit>
LM_HD void mel_run(const T* __restrict__ P, const Tables<G>& tab, int m0, int m1, int ng, int woff,
                   int lane, Emit&& emit) {
  const T* Pl = P + lane;
  const float4* w4 = reinterpret_cast<const float4*>(tab.melw) + woff;
  int m = m0;
#pragma unroll 1
  for (; m + 1 < m1; m += 2, w4 += 2 * ng) {
    T a0, a1;
    mel_dot2<NG, T>(Pl + tab.mel_lo[m] * 32, Pl + tab.mel_lo[m + 1] * 32, w4, ng, a0, a1);
    emit(a0);
    emit(a1);
  }
  if (m < m1) {     // odd run length: the last filter alone
    T a0 = vzero<T>();
    const T* s = Pl + tab.mel_lo[m] * 32;
#pragma unroll 1
    for (int g = 0; g < ng; ++g) {
      const float4 u = w4[g];
      a0 = vfmas(s[g * 128], u.x, a0);
      a0 = vfmas(s[g * 128 + 32], u.y, a0);
      a0 = vfmas(s[g * 128 + 64], u.z, a0);
      a0 = vfmas(s[g * 128 + 96], u.w, a0);
    }
    emit(a0);
  }
}

template <class G, typename T, class Emit>
LM_HD void mel_task_scan(const T* __restrict__ P, const Tables<G>& tab, int w, int lane, Emit&& emit) {
  const int nb = tab.scan_nb[w];
  const unsigned char* code = tab.scan_code + tab.scan_soff[w];
  const T* src = P + tab.scan_bin0[w] * 32 + lane;
  const float* wp = tab.melw + 2 * tab.scan_bin0[w];
  T acc0 = vzero<T>(), acc1 = vzero<T>();
  // what follows a bin whose code is non-zero (warp-uniform branches)
  auto shift = [&](unsigned c) {
    if (!(c & 4u)) emit(acc0);
    acc0 = acc1;
    acc1 = vzero<T>();
    if (c & 2u) {
      if (!(c & 8u)) emit(acc0);
      acc0 = vzero<T>();
    }
  };
  int i = 0;
#pragma unroll 1
  for (; i + 4 <= nb; i += 4, src += 4 * 32, wp += 8) {
    const T p0 = src[0], p1 = src[32], p2 = src[64], p3 = src[96];
    const unsigned c4 = *reinterpret_cast<const unsigned*>(code + i);
    acc0 = vfmas(p0, wp[0], acc0);
    acc1 = vfmas(p0, wp[1], acc1);
    if (c4 & 0x000000ffu) shift(c4);
    acc0 = vfmas(p1, wp[2], acc0);
    acc1 = vfmas(p1, wp[3], acc1);
    if (c4 & 0x0000ff00u) shift(c4 >> 8);
    acc0 = vfmas(p2, wp[4], acc0);
    acc1 = vfmas(p2, wp[5], acc1);
    if (c4 & 0x00ff0000u) shift(c4 >> 16);
    acc0 = vfmas(p3, wp[6], acc0);
    acc1 = vfmas(p3, wp[7], acc1);
    if (c4 & 0xff000000u) shift(c4 >> 24);
  }
#pragma unroll 1
  for (; i < nb; ++i, src += 32, wp += 2) {
    const T p0 = src[0];
    const unsigned c = code[i];
    acc0 = vfmas(p0, wp[0], acc0);
    acc1 = vfmas(p0, wp[1], acc1);
    if (c) shift(c);
  }
}

// Software-pipelined scan for the warp-specialised CTA, where one warp per sub-partition runs the
// whole mel phase and nothing else hides its latencies:
//  * bins are fetched in groups of four, one group AHEAD of the arithmetic (power values from
//    shared memory, weights and codes from the constant bank);
//  * a finished filter is split into start(acc) -- the part with a long latency (MUFU.LG2) -- and
//    finish(value), which runs one filter later, when the result has long arrived.
// Reads up to two groups past the warp's last bin (never used: they follow the last emit); the
// tables are zero padded and the caller's P is followed by other shared memory.
template <class G, typename T, class Start, class Finish>
LM_HD void mel_task_pipe(const T* __restrict__ P, const Tables<G>& tab, int w, int lane, Start&& start,
                         Finish&& finish) {
  const int ng = (tab.scan_nb[w] + 3) >> 2;
  const unsigned* code4 = reinterpret_cast<const unsigned*>(tab.scan_code + tab.scan_soff[w]);
  const T* src = P + tab.scan_bin0[w] * 32 + lane;
  const float* wp = tab.melw + 2 * tab.scan_bin0[w];
  struct Grp {
    T p[4];
    float w[8];
    unsigned c;
  };
  // G::WS: P is the in-place output of stage 2 (p_slot layout), addressed through scan_poff
  const unsigned* poff = tab.scan_poff + (G::WS ? tab.scan_soff[w] : 0);
  const char* pl = reinterpret_cast<const char*>(P + lane);
  auto load = [&](Grp& g, int gi) {
    if (G::WS) {
#pragma unroll
      for (int j = 0; j < 4; ++j) g.p[j] = *reinterpret_cast<const T*>(pl + poff[4 * gi + j]);
    } else {
      const T* s = src + gi * 128;
      g.p[0] = s[0]; g.p[1] = s[32]; g.p[2] = s[64]; g.p[3] = s[96];
    }
    const float* q = wp + gi * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) g.w[j] = q[j];
    g.c = code4[gi];
  };
  T acc0 = vzero<T>(), acc1 = vzero<T>(), pend = vzero<T>();
  bool have = false;
  auto emit = [&](T acc) {
    const T cur = start(acc);
    if (have) finish(pend);
    pend = cur;
    have = true;
  };
  auto shift = [&](unsigned c) {
    if (!(c & 4u)) emit(acc0);
    acc0 = acc1;
    acc1 = vzero<T>();
    if (c & 2u) {
      if (!(c & 8u)) emit(acc0);
      acc0 = vzero<T>();
    }
  };
  auto run = [&](const Grp& g) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc0 = vfmas(g.p[j], g.w[2 * j], acc0);
      acc1 = vfmas(g.p[j], g.w[2 * j + 1], acc1);
      const unsigned c = (g.c >> (8 * j)) & 0xffu;
      if (c) shift(c);
    }
  };
  Grp ga, gb;
  if (ng > 0) load(ga, 0);
  int gi = 0;
#pragma unroll 1
  for (; gi + 2 <= ng; gi += 2) {
    load(gb, gi + 1);
    run(ga);
    load(ga, gi + 2);
    run(gb);
  }
  if (gi < ng) run(ga);
  if (have) finish(pend);
}

template <class G, typename T, class Emit>
LM_HD void mel_task(const T* __restrict__ P, const Tables<G>& tab, int w, int lane, Emit&& emit) {
  if (tab.mel_scan) {
    mel_task_scan<G, T>(P, tab, w, lane, emit);
    return;
  }
  const int m0 = tab.mel_begin[w], m1 = tab.mel_begin[w + 1];
  const int ng = tab.mel_ng[w], woff = tab.mel_woff[w];
  switch (ng) {      // warp-uniform
    case 1: mel_run<1, G, T>(P, tab, m0, m1, 1, woff, lane, emit); break;
    case 2: mel_run<2, G, T>(P, tab, m0, m1, 2, woff, lane, emit); break;
    case 3: mel_run<3, G, T>(P, tab, m0, m1, 3, woff, lane, emit); break;
    default: mel_run<0, G, T>(P, tab, m0, m1, ng, woff, lane, emit); break;
  }
}

}  // namespace lm
