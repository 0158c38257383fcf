// The fused log-mel kernel for sm_100a: waveform tile -> window -> two-stage real FFT ->
// |X|^2 -> banded mel -> log -> (Whisper) per-clip max and normalisation, one launch.
//
// Grid: persistent, cooperative.  CTAs are organised in *clip groups* of `group` CTAs; group g
// walks clips g, g + n_groups, ...; CTA `rank` of the group owns a contiguous run of frame
// tiles of that clip.  The un-normalised log-mel is written once, stays in L2 (the grid keeps
// only n_groups ~ 37 clips in flight: ~57 MB of the 126 MB L2 for 128 mels), the group agrees
// on the clip maximum through a release/acquire counter in global memory, and every CTA then
// rewrites its own slab with max(S, M - 8), (S + 4) / 4 while the lines are still L2 resident.
// The cooperative launch guarantees the co-residency the spin wait relies on.
#pragma once
#include <cuda_runtime.h>

#include "logmel_core.cuh"

namespace lm {

struct KArgs {
  const float* wave;         // float32 clips, or NULL when pcm is set
  const short* pcm;          // 16-bit PCM clips: interleaved frames of pcm_channels channels (fused ingest)
  int pcm_channels;          // 1 or 2 (the kernel averages the channels)
  long long clip_stride;     // samples (PCM: frames) between clip starts
  const int* lengths;
  float* out;
  float* clip_max;   // optional user-visible per-clip max
  float* gmax;       // scratch [batch * group]
  int* gcnt;         // scratch [batch], zeroed before launch
  int batch, n_samples, n_frames, n_mels;
  int log_mode;
  float log_add, log_floor, log_scale;   // y = log2(max(x + add, floor)) * scale
  int group, n_groups, tiles_per_clip;
  int vec_ok;        // slabs are float4-addressable
  int tma_ok;        // every clip starts 16-byte aligned: tiles can be fetched with bulk copies
  long long* timeline;   // debug builds (-DLM_TIMELINE): per-tile clock stamps of CTA 0, else NULL
};

#ifdef LM_TIMELINE
#ifndef LM_TL_CTA
#define LM_TL_CTA 0
#endif
#define LM_STAMP(slot)                                                                         \
  if (a.timeline && blockIdx.x == LM_TL_CTA && lane == 0 && tl_tile < 48)                              \
    a.timeline[(tl_tile * 16 + warp) * 8 + (slot)] = clock64();
#define LM_CSTAMP(slot)                                                                        \
  if (a.timeline && blockIdx.x == LM_TL_CTA && threadIdx.x == 0 && tl_clip < 16)                       \
    a.timeline[48 * 16 * 8 + tl_clip * 8 + (slot)] = clock64();
#else
#define LM_STAMP(slot)
#define LM_CSTAMP(slot)
#endif

template <class G>
struct Lay {   // dynamic shared-memory budget of one CTA: Y | P | waveform tile | stage-1 constants
  using T = typename ValT<G>::type;
  static constexpr size_t Y = (size_t)G::Y_ELEMS * sizeof(T);
  static constexpr size_t P = (size_t)G::P_ELEMS * sizeof(T);
  static constexpr size_t W = (size_t)G::WAVE_FLOATS * 4;
  static constexpr size_t S1 = G::S1_CONST_REGS ? 0 : (size_t)G::N2 * G::S1_STRIDE * 4;
  // when all of it does not fit, P reuses the waveform buffer (dead after stage 1) and the next
  // tile is fetched after the mel phase instead of behind stage 2
  static constexpr bool ALIAS = Y + P + W + S1 > 222 * 1024;
  static constexpr size_t PW = ALIAS ? (P > W ? P : W) : P + W;
  static constexpr size_t BYTES = Y + PW + S1;
  static constexpr int MIN_CTAS = BYTES <= 110 * 1024 ? 2 : 1;
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- TMA (bulk async copy) + mbarrier primitives -----------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// one lane of a converged warp (lets ptxas keep the bulk-copy operands in uniform registers)
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void bulk_g2s(float* dst_smem, const float* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Stage the SPAN samples of a tile into shared memory (wave_index layout), row by row.
//
// A row (HOP samples, the last one shorter) that lies inside the clip's audio is fetched with
// ONE TMA bulk copy (both addresses 16-byte aligned) onto the mbarrier.  In the common case --
// every row interior -- and when a warp without stage-2 work exists (`loader`), lane 0 of that
// warp issues all copies from a short uniform-register loop (about four instructions per copy)
// while the other warps run stage 2.  Otherwise lane 0 of warp w issues the copies of rows
// w, w + NWK, ..., and rows that touch a clip edge or the zero padding (at most a handful per
// clip), and every row of an unaligned input, are filled by their warp through load_sample
// (reflection / zero fill).
// Returns bit 0: some rows arrive through the mbarrier, bit 1: some rows were stored by threads.
// CTA-uniform plan of a tile: rows [r_lo, r_hi) are interior (one bulk copy each), the others are
// written by threads.  Returns bit 0: some rows arrive through the mbarrier, bit 1: some rows
// are stored by threads.
template <class G>
__device__ __forceinline__ int tile_plan(long long s0, int valid, bool tma_ok, int* r_lo_out, int* r_hi_out) {
  constexpr int FULL_ROWS = G::SPAN / G::HOP;
  constexpr int REM = G::SPAN - FULL_ROWS * G::HOP;
  constexpr int ROWS = FULL_ROWS + (REM > 0 ? 1 : 0);
  int r_lo = 0, r_hi = 0;
  if (tma_ok) {
    r_lo = s0 >= 0 ? 0 : (int)((-s0 + G::HOP - 1) / G::HOP);
    const long long room = (long long)valid - s0;              // samples of audio from s0 on
    r_hi = room <= 0 ? 0 : (int)min((long long)FULL_ROWS, room / G::HOP);
    if (REM > 0 && r_hi == FULL_ROWS && room >= G::SPAN) r_hi = ROWS;
    if (r_lo > r_hi) r_lo = r_hi;
  }
  *r_lo_out = r_lo;
  *r_hi_out = r_hi;
  const int n_tma = r_hi - r_lo;
  return (n_tma > 0 ? 1 : 0) | (n_tma < ROWS ? 2 : 0);
}

// NW warps (indices 0..NW-1 in `warp`) share the work.
// n_iss > 0: in the common all-interior case only n_iss of them issue copies (this warp is number
// `iss` among those, -1 if it is not one) -- the warps with the lighter stage-2 load.
template <class G, int NW = G::NWK>
__device__ __forceinline__ int load_tile(float* wave_s, const float* __restrict__ clip, long long s0, int n_samples,
                                         int valid, bool tma_ok, unsigned long long* bar, int warp, int lane,
                                         int loader, int iss = -1, int n_iss = 0, const short* pcm = nullptr,
                                         int pcm_ch = 0) {
  constexpr int FULL_ROWS = G::SPAN / G::HOP;
  constexpr int REM = G::SPAN - FULL_ROWS * G::HOP;
  constexpr int ROWS = FULL_ROWS + (REM > 0 ? 1 : 0);
  static_assert((G::HOP * 4) % 16 == 0 && (REM * 4) % 16 == 0 && (G::PITCH * 4) % 16 == 0, "bulk copies move 16-byte units");
  // interior rows form the contiguous range [r_lo, r_hi)   (CTA-uniform)
  int r_lo = 0, r_hi = 0;
  if (tma_ok) {
    r_lo = s0 >= 0 ? 0 : (int)((-s0 + G::HOP - 1) / G::HOP);
    const long long room = (long long)valid - s0;              // samples of audio from s0 on
    r_hi = room <= 0 ? 0 : (int)min((long long)FULL_ROWS, room / G::HOP);
    if (REM > 0 && r_hi == FULL_ROWS && room >= G::SPAN) r_hi = ROWS;
    if (r_lo > r_hi) r_lo = r_hi;
  }
  const int n_tma = r_hi - r_lo;
  if (n_tma == ROWS && loader >= 0) {
    if (warp == loader && lane == 0) {
      fence_proxy_async();    // earlier generic-proxy reads of this buffer are ordered before the async writes
      mbar_expect_tx(bar, (unsigned)(G::SPAN * 4));
      unsigned dst = smem_u32(wave_s);
      const float* src = clip + s0;
      const unsigned b = smem_u32(bar);
#pragma unroll 4
      for (int row = 0; row < FULL_ROWS; ++row, dst += G::PITCH * 4, src += G::HOP)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(src), "n"(G::HOP * 4), "r"(b)
                     : "memory");
      if (REM > 0)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(src), "n"(REM > 0 ? REM * 4 : 16), "r"(b)
                     : "memory");
    }
    __syncwarp();
    return 1;
  }
  if (n_tma > 0 && lane == 0) {
    fence_proxy_async();
    if (warp == 0) {
      const unsigned bytes = (unsigned)(min(r_hi, FULL_ROWS) - r_lo) * (G::HOP * 4u) + (r_hi == ROWS && REM > 0 ? REM * 4u : 0u);
      mbar_expect_tx(bar, bytes);
    }
  }
  if (n_tma == ROWS) {
    // every row is interior: one lane walks this warp's rows with two running pointers
    const int me = n_iss > 0 ? iss : warp, step = n_iss > 0 ? n_iss : NW;
    if (me >= 0 && elect_one()) {
      float* dst = wave_s + me * G::PITCH;
      const float* src = clip + s0 + me * G::HOP;
#pragma unroll 1
      for (int row = me; row < FULL_ROWS; row += step, dst += step * G::PITCH, src += step * G::HOP)
        bulk_g2s(dst, src, G::HOP * 4, bar);
      if (REM > 0 && me == FULL_ROWS % step)
        bulk_g2s(wave_s + FULL_ROWS * G::PITCH, clip + s0 + FULL_ROWS * G::HOP, REM * 4, bar);
    }
  } else {
#pragma unroll 1
    for (int row = warp; row < ROWS; row += NW) {
      const int len = (row < FULL_ROWS) ? G::HOP : REM;
      if (row >= r_lo && row < r_hi) {
        if (lane == 0) bulk_g2s(wave_s + row * G::PITCH, clip + s0 + (long long)row * G::HOP, len * 4, bar);
      } else {
        for (int i = lane; i < len; i += 32)
          wave_s[row * G::PITCH + i] =
              load_sample_any(clip, pcm, pcm_ch, (long)(s0 + (long long)row * G::HOP + i), n_samples, valid);
      }
    }
  }
  __syncwarp();
  return (n_tma > 0 ? 1 : 0) | (n_tma < ROWS ? 2 : 0);
}

// every sample the tile touches (after reflection about the padded length) is zero padding
template <class G>
__device__ __forceinline__ bool tile_is_silent(long long s0, int n_samples, int valid) {
  if (s0 < valid) return false;
  const long long last = s0 + G::SPAN - 1;
  if (last < n_samples) return true;
  return 2LL * (n_samples - 1) - last >= valid;
}

constexpr int kMaxLocalTiles = 64;

// KIND: 0 = mel power, 1 = log10(max(x, floor)), 2 = ln(x + eps), 3 = Whisper: kind 1 followed by
// max(S, clipmax - 8), (S + 4) / 4.
//
// Whisper normalisation without a second pass over the data: g(S) = (S + 4) * 0.25 is
// monotone, so max(g(S), g(c)) == g(max(S, c)) bit for bit, and
//     max(max(S', f), M - 8) == max(S', max(f, M - 8))        (S' = log10 of the UNclamped power,
//                                                              f = log10(floor))
// The first pass therefore stores g(S') -- already final wherever S' >= c := max(f, M - 8) --
// and records each tile's minimum; once the clip group has agreed on M, only tiles whose
// minimum lies below c are revisited with out = max(out, g(c)), while their lines are still in
// L2.  The agreement (a release/acquire counter in global memory) is not waited for at the end
// of the clip: the CTA goes on with the first tile of its next clip and resolves the previous
// clip afterwards, when the other CTAs of the group have long published their maxima.
// Silent tiles (all zero padding) are not computed at all: their constant is written when the
// clip is resolved.
template <class G, int KIND>
__global__ void __launch_bounds__(G::THREADS, Lay<G>::MIN_CTAS)
logmel_kernel(const __grid_constant__ Tables<G> tab, const KArgs a) {
  using T = typename ValT<G>::type;
  constexpr bool ALIAS = Lay<G>::ALIAS;
  constexpr bool NORM = KIND == 3;
  constexpr int NT = G::THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Y = reinterpret_cast<T*>(smem_raw);
  T* P = Y + G::Y_ELEMS;
  float* wave_s = ALIAS ? reinterpret_cast<float*>(P) : reinterpret_cast<float*>(P + G::P_ELEMS);
  float* s1_s = reinterpret_cast<float*>(smem_raw + Lay<G>::Y + Lay<G>::PW);
  __shared__ __align__(8) unsigned long long s_mbar;
  __shared__ float s_red[G::NWK];
  __shared__ float s_max;
  __shared__ float s_cta_max[2];
  __shared__ float s_tmin[NORM ? 2 : 1][NORM ? kMaxLocalTiles : 1][G::NWK];
  __shared__ unsigned char s_silent[NORM ? 2 : 1][NORM ? kMaxLocalTiles : 1];

  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int group_id = blockIdx.x / a.group;
  const int rank = blockIdx.x - group_id * a.group;
  const int loader = tab.loader_warp;
  const float log_floor = a.log_floor, log_add = a.log_add, log_scale = a.log_scale;
  // what a bin-wise silent frame produces, through exactly the arithmetic of the mel phase
  float silent_val = 0.0f;
  if (KIND == 1 || KIND == 3) silent_val = vlog2_clamp(0.0f, log_floor) * log_scale;
  if (KIND == 2) silent_val = vlog2_add(0.0f, log_add) * log_scale;

  // one-time setup: stage-1 constants into shared memory (the four column groups of a warp read
  // four different rows of the table at once), mbarrier for the TMA tile copies
  if (!G::S1_CONST_REGS)
    for (int i = threadIdx.x; i < G::N2 * G::S1_STRIDE; i += NT) s1_s[i] = tab.s1[i];
  if (threadIdx.x == 0) {
    mbar_init(&s_mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned parity = 0;
  // N = 400: this warp's stage-1 constants, resident in registers for the whole kernel
  float s1c[G::S1_CONST_REGS ? G::S1_STRIDE : 1];
  if (G::S1_CONST_REGS && tab.s1_tasks[warp][0] >= 0)
    stage1_consts<G>(tab.s1, tab.s1_tasks[warp][0], lane, reinterpret_cast<float(&)[G::S1_STRIDE]>(s1c));
#ifdef LM_TIMELINE
  int tl_tile = 0, tl_clip = 0;
#endif

  // this CTA's share of every clip
  const int t0 = (int)((long long)rank * a.tiles_per_clip / a.group);
  const int t1 = (int)((long long)(rank + 1) * a.tiles_per_clip / a.group);
  const bool track = NORM && (t1 - t0 <= kMaxLocalTiles);
  auto tile_s0 = [&](int t) { return (long long)t * G::F * G::HOP - G::N / 2; };

  // ---- deferred normalisation of a finished clip (NORM only) -------------------------------
  int pend_clip = -1, pend_valid = 0, pend_par = 0;
  auto resolve = [&]() {      // CTA-uniform; one CTA barrier inside
    if (threadIdx.x == 0) {
      float m;
      if (a.group > 1) {
        const float* slots = a.gmax + (long long)pend_clip * a.group;
        while (ld_acquire(a.gcnt + pend_clip) < a.group) __nanosleep(64);
        m = __ldcg(slots);
        for (int r = 1; r < a.group; ++r) m = fmaxf(m, __ldcg(slots + r));
      } else {
        m = s_cta_max[pend_par];
      }
      m = fmaxf(m, silent_val);                   // the clamp at `floor`, applied to the maximum
      if (rank == 0 && a.clip_max) a.clip_max[pend_clip] = m;
      s_max = m;
    }
    __syncthreads();
    // revisit only what the clamp actually touches
    const float thr = fmaxf(s_max - 8.0f, silent_val);
    const float cval = vaffine(thr, 0.25f, 1.0f);
    float* oc = a.out + (long long)pend_clip * a.n_mels * a.n_frames;
    for (int t = t0; t < t1; ++t) {
      const int fa = t * G::F;
      const int len = min(fa + G::F, a.n_frames) - fa;
      bool silent, fix;
      if (track) {
        silent = s_silent[pend_par][t - t0] != 0;
        float tm = INFINITY;
        if (!silent)
          for (int w = 0; w < G::NWK; ++w) tm = fminf(tm, s_tmin[pend_par][t - t0][w]);
        fix = !(tm >= thr);
      } else {
        silent = tile_is_silent<G>(tile_s0(t), a.n_samples, pend_valid);
        fix = true;
      }
      if (silent) {
        for (int m = warp; m < a.n_mels; m += G::NWK) {
          float* row = oc + (long long)m * a.n_frames + fa;
          for (int j = lane; j < len; j += 32) __stcs(row + j, cval);
        }
      } else if (fix) {
        if (a.vec_ok && (len & 3) == 0) {
          const int q = len >> 2;          // <= 16 float4 per row: two rows per warp pass
          for (int m = 2 * warp + (lane >> 4); m < a.n_mels; m += 2 * G::NWK) {
            float4* row = reinterpret_cast<float4*>(oc + (long long)m * a.n_frames + fa);
            const int j = lane & 15;
            if (j < q) {
              float4 v = __ldcg(row + j);
              v.x = fmaxf(v.x, cval); v.y = fmaxf(v.y, cval); v.z = fmaxf(v.z, cval); v.w = fmaxf(v.w, cval);
              __stcs(row + j, v);
            }
          }
        } else {
          for (int m = warp; m < a.n_mels; m += G::NWK) {
            float* row = oc + (long long)m * a.n_frames + fa;
            for (int j = lane; j < len; j += 32) row[j] = fmaxf(__ldcg(row + j), cval);
          }
        }
      }
    }
    pend_clip = -1;
  };

  int par = 0;
  for (int clip = group_id; clip < a.batch; clip += a.n_groups, par ^= 1) {
    const float* cptr = a.wave + (long long)clip * a.clip_stride;
    const short* pptr = a.pcm ? a.pcm + (long long)clip * a.clip_stride * a.pcm_channels : nullptr;
    int valid = a.n_samples;
    if (a.lengths) valid = min(max(a.lengths[clip], 0), a.n_samples);
    float* oc = a.out + (long long)clip * a.n_mels * a.n_frames;
    float rmax = -INFINITY;
    LM_CSTAMP(0)
    bool staged_tma = false;  // the tile being staged into wave_s arrives through the mbarrier

    // CTA-uniform: start filling wave_s with the tile that begins at sample s; returns whether
    // some rows were written by threads (then a CTA barrier must precede their use)
    auto fetch = [&](long long s) -> bool {
      const int how = load_tile<G>(wave_s, cptr, s, a.n_samples, valid, a.tma_ok != 0, &s_mbar, warp, lane, loader, -1, 0,
                                   pptr, a.pcm_channels);
      staged_tma = (how & 1) != 0;
      return (how & 2) != 0;
    };
    auto next_loud = [&](int t) {            // first tile >= t that has to be computed
      while (t < t1 && tile_is_silent<G>(tile_s0(t), a.n_samples, valid)) ++t;
      return t;
    };

    // ---- silent tiles (all zero padding) are never computed
    for (int t = t0; t < t1; ++t) {
      const bool silent = tile_is_silent<G>(tile_s0(t), a.n_samples, valid);       // CTA-uniform
      if (NORM && track && threadIdx.x == 0) s_silent[par][t - t0] = silent ? 1 : 0;
      if (!silent) continue;
      rmax = fmaxf(rmax, silent_val);
      if (!NORM) {                            // no normalisation: the constant can be written right away
        const int f0 = t * G::F, fend = min(f0 + G::F, a.n_frames);
        for (int m = warp; m < a.n_mels; m += G::NWK)
          for (int f = f0 + lane; f < fend; f += 32) oc[(long long)m * a.n_frames + f] = silent_val;
      }
    }

    auto wait_wave = [&]() __attribute__((always_inline)) {                  // the bulk copies of the staged tile have landed
      if (staged_tma) {
        mbar_wait(&s_mbar, parity);
        parity ^= 1u;
      }
    };
    auto do_s1 = [&]() __attribute__((always_inline)) {                      // stage 1: this warp's tasks of (8 frame slots x 4 columns)
      if (G::S1_CONST_REGS) {
        // 0, 1 or 2 tasks per warp; a pair is one straight-line block in which the second task's
        // samples are fetched before the first task's arithmetic starts
        const int ta = tab.s1_tasks[warp][0], tb = tab.s1_tasks[warp][1];
        if (tb >= 0)
          stage1_task_pair<G, T>(wave_s, Y, reinterpret_cast<const float(&)[G::S1_STRIDE]>(s1c), ta, tb, lane);
        else if (ta >= 0)
          stage1_task_c<G, T>(wave_s, Y, reinterpret_cast<const float(&)[G::S1_STRIDE]>(s1c), ta, lane);
      } else {
#pragma unroll 1
        for (int i = 0; i < G::S1_MAX; ++i) {
          const int task = tab.s1_tasks[warp][i];
          if (task < 0) break;
          stage1_task<G, T>(wave_s, Y, s1_s, task, lane);
        }
      }
    };
    // power -> log domain value S (KIND 3: the clamp at `floor` is applied when the clip is resolved)
    auto to_log = [&](T v) __attribute__((always_inline)) -> T {
      if (KIND == 1) return vmuls(vlog2_clamp(v, log_floor), log_scale);
      if (KIND == 2) return vmuls(vlog2_add(v, log_add), log_scale);
      if (KIND == 3) return vmuls(vlog2_raw(v), log_scale);
      return v;
    };
    auto do_mel = [&](int t) __attribute__((always_inline)) {                // mel projection, log, store, running max / tile min
      const int f0 = t * G::F;
      const int f = f0 + lane;
      const bool full = f0 + G::F <= a.n_frames;            // CTA-uniform: only a clip's last tile is not
      float* op = oc + (long long)tab.mel_begin[warp] * a.n_frames + f;
      const long long ostep = a.n_frames;
      float tmin = INFINITY;
      if (full) {
        mel_task<G, T>(P, tab, warp, lane, [&](T acc) {
          T v = to_log(acc);
          if (NORM) {
            rmax = vhmax(rmax, v);
            tmin = vhmin(tmin, v);
            v = vaffine(v, 0.25f, 1.0f);
          }
          op[0] = vlo(v);
          if (G::PK == 2) op[32] = vhi(v);
          op += ostep;
        });
      } else {
        const bool ok0 = f < a.n_frames, ok1 = (G::PK == 2) && (f + 32 < a.n_frames);
        mel_task<G, T>(P, tab, warp, lane, [&](T acc) {
          T v = to_log(acc);
          if (NORM) {
            if (ok0) { rmax = fmaxf(rmax, vlo(v)); tmin = fminf(tmin, vlo(v)); }
            if (ok1) { rmax = fmaxf(rmax, vhi(v)); tmin = fminf(tmin, vhi(v)); }
            v = vaffine(v, 0.25f, 1.0f);
          }
          if (ok0) op[0] = vlo(v);
          if (ok1) op[32] = vhi(v);
          op += ostep;
        });
      }
      if (track) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
        if (lane == 0) s_tmin[par][t - t0][warp] = tmin;
      }
    };

    // ---- computed tiles, software pipelined:
    //   S1(t) | barrier | TMA prefetch(t') + S2(t) | barrier | { M(t), S1(t') in either order } | ...
    // M(t) (shared-memory / MUFU bound) and S1(t') (FP32-pipe bound) share an interval; one third
    // of the warps runs them in the opposite order, so the two kinds of work overlap instead of
    // all warps hitting the same pipe at the same time.
    const bool s1_first = ((warp >> 2) & 1) != 0;
    int t = next_loud(t0);
    if (t < t1) {
      if (fetch(tile_s0(t))) __syncthreads();
      wait_wave();
      LM_STAMP(1)
      do_s1();
    }
    bool first = true;
    while (t < t1) {
      LM_STAMP(2)
      __syncthreads();                                        // Y(t) complete, waveform tile dead
      const int tn = next_loud(t + 1);
      const bool pre = !ALIAS && tn < t1;
      if (pre) fetch(tile_s0(tn));                            // behind stage 2; the barrier below covers thread-written rows
#pragma unroll 1
      for (int i = 0; i < G::S2_MAX; ++i) {                   // stage 2: this warp's rows k1
        const int k1 = tab.s2_rows[warp][i];
        if (k1 < 0) break;
        stage2_task<G, T>(Y, P, k1, lane);
      }
      LM_STAMP(4)
      __syncthreads();                                        // P(t) complete
      if (!ALIAS) {
        // one copy of each phase in the instruction stream; the order depends on the warp
#pragma unroll 1
        for (int ph = 0; ph < 2; ++ph) {
          if ((ph == 0) == s1_first) {
            if (pre) {
              wait_wave();
              do_s1();
            }
          } else {
            do_mel(t);
            LM_STAMP(6)
          }
        }
      } else {
        do_mel(t);
        __syncthreads();   // P shares the waveform buffer: the next tile is fetched only now
        if (tn < t1) {
          if (fetch(tile_s0(tn))) __syncthreads();
          wait_wave();
          do_s1();
        }
      }
      if (NORM && first && pend_clip >= 0) resolve();         // the previous clip, one tile into this one
      first = false;
#ifdef LM_TIMELINE
      ++tl_tile;
#endif
      t = tn;
    }

    LM_CSTAMP(1)
    if (NORM) {
      if (pend_clip >= 0) resolve();                          // this clip had no computed tile
      // ---- publish this CTA's maximum: warp shuffle -> CTA -> slot of the clip group
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
      if (lane == 0) s_red[warp] = rmax;
      __syncthreads();
      if (threadIdx.x == 0) {
        float m = s_red[0];
        for (int w = 1; w < G::NWK; ++w) m = fmaxf(m, s_red[w]);
        if (a.group > 1) {
          __stcg(a.gmax + (long long)clip * a.group + rank, m);
          __threadfence();
          atomicAdd(a.gcnt + clip, 1);
        } else {
          s_cta_max[par] = m;
        }
      }
      pend_clip = clip;
      pend_valid = valid;
      pend_par = par;
      LM_CSTAMP(2)
    }
#ifdef LM_TIMELINE
    ++tl_clip;
#endif
  }
  if (NORM && pend_clip >= 0) {
    __syncthreads();          // s_cta_max / s_tmin of the last clip are complete
    resolve();
  }
}

}  // namespace lm
