// Warp-specialised form of the fused log-mel kernel (Geo<..., WS = 1>): the hot Whisper path.
//
// The phase-synchronous kernel (logmel_kernel.cuh) runs every phase on all warps, so between two
// CTA barriers all warps first load, then compute: shared-memory time and FP32 time add up
// instead of overlapping, and the latency-bound mel/log/store phase sits on the critical path
// of every warp (ncu: FMA pipe 32 %, 30 % of warp time at barriers).  Here the CTA is split by
// role:
//
//   warps 0..7   (two per SM sub-partition)  stage 1 and stage 2 -- the FP32-pipe-bound work;
//                the codelets saturate the pipe from a single warp (profiles/r01_codelet_rate_*)
//   warps 8..11  (one per sub-partition)     mel projection, log, global stores, running max /
//                tile min -- the latency-bound work, software pipelined (mel_task_pipe), issued
//                into the slots the FFT warps leave.  (The TMA tile copies of the next tile are
//                issued by the FFT warps that own a single stage-2 row, see tab.tma_iss.)
//
// The two roles are decoupled by a full tile period.  Stage 2 writes the power spectrum IN PLACE
// over its own row of the real plane of Y (stage2_task_inplace), and that plane is double
// buffered (+56 KB instead of a separate 51 KB P buffer):
//
//   shared memory:  Yre[0] | Yre[1] | Yim | waveform tile | stage-1 constants      (210 KB)
//
// Per computed tile i (counted over the whole CTA), b = i & 1:
//   FFT warps:  sync Y_DONE | issue TMA(i+1) | S2(i): Yre[b], Yim -> P(i) in Yre[b]
//               | arrive P_FULL[b] | sync S2_DONE | wait TMA(i+1) | sync P_EMPTY[b^1] (tile i-1)
//               | S1(i+1): wave -> Yre[b^1], Yim
//   mel warps:  sync P_FULL[b] | mel(i): Yre[b] -> HBM | arrive P_EMPTY[b]
// so mel(i) may run through S1(i+1) AND S2(i+1); the FFT warps never execute a MUFU or a global
// store and never wait for the mel warps unless those fall a whole tile behind.  The per-clip
// bookkeeping of the Whisper normalisation (publish the maximum, resolve the previous clip)
// is done by the mel warps alone.
#pragma once
#include "logmel_kernel.cuh"

namespace lm {

enum : int { BAR_PFULL0 = 1, BAR_PFULL1 = 2, BAR_PEMPTY0 = 3, BAR_PEMPTY1 = 4, BAR_YDONE = 5, BAR_S2DONE = 6,
             BAR_WAVE = 7, BAR_MEL = 8 };

__device__ __forceinline__ void nbar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void nbar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <class G>
struct LayWS {   // dynamic shared memory of the warp-specialised CTA
  using T = typename ValT<G>::type;
  static constexpr size_t YRE = (size_t)G::YRE_ELEMS * sizeof(T);
  static constexpr size_t YIM = (size_t)G::H1 * G::N2 * 32 * sizeof(T);
  static constexpr size_t W = (size_t)G::WAVE_FLOATS * 4;
  static constexpr size_t S1 = (size_t)G::N2 * G::S1_STRIDE * 4;
  static constexpr size_t BYTES = 2 * YRE + YIM + W + S1;
  static_assert(BYTES <= 227 * 1024, "does not fit in shared memory");
};

template <class G, int KIND>
__global__ void __launch_bounds__(G::THREADS, 1)
logmel_ws_kernel(const __grid_constant__ Tables<G> tab, const KArgs a) {
  using T = typename ValT<G>::type;
  static_assert(G::WS == 1 && !G::S1_CONST_REGS, "warp-specialised geometry expected");
  constexpr bool NORM = KIND == 3;
  constexpr int NT = G::THREADS;
  constexpr int NTF = G::NW_FFT * 32;
  constexpr int NTM = G::NW_MEL * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Yre0 = reinterpret_cast<T*>(smem_raw);
  T* Yim = Yre0 + 2 * G::YRE_ELEMS;
  float* wave_s = reinterpret_cast<float*>(smem_raw + 2 * LayWS<G>::YRE + LayWS<G>::YIM);
  float* s1_s = wave_s + G::WAVE_FLOATS;
  __shared__ __align__(8) unsigned long long s_mbar;
  __shared__ float s_red[G::NW_MEL];
  __shared__ float s_max;
  __shared__ float s_cta_max[2];
  __shared__ float s_tmin[NORM ? 2 : 1][NORM ? kMaxLocalTiles : 1][G::NW_MEL];
  __shared__ unsigned char s_silent[NORM ? 2 : 1][NORM ? kMaxLocalTiles : 1];

  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const bool is_mel = warp >= G::NW_FFT;
  const int mw = warp - G::NW_FFT;                  // index among the mel warps
  const int group_id = blockIdx.x / a.group;
  const int rank = blockIdx.x - group_id * a.group;

  for (int i = threadIdx.x; i < G::N2 * G::S1_STRIDE; i += NT) s1_s[i] = tab.s1[i];
  if (threadIdx.x == 0) {
    mbar_init(&s_mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
#ifdef LM_TIMELINE
  int tl_tile = 0, tl_clip = 0;
#endif

  const int t0 = (int)((long long)rank * a.tiles_per_clip / a.group);
  const int t1 = (int)((long long)(rank + 1) * a.tiles_per_clip / a.group);
  auto tile_s0 = [&](int t) { return (long long)t * G::F * G::HOP - G::N / 2; };
  int it = 0;                                       // computed tiles so far (same sequence in both roles)

  if (!is_mel) {
    // =====================================================================================
    // FFT warps
    // =====================================================================================
    unsigned parity = 0;
    for (int clip = group_id; clip < a.batch; clip += a.n_groups) {
      const float* cptr = a.wave + (long long)clip * a.clip_stride;
      const short* pptr = a.pcm ? a.pcm + (long long)clip * a.clip_stride * a.pcm_channels : nullptr;
      int valid = a.n_samples;
      if (a.lengths) valid = min(max(a.lengths[clip], 0), a.n_samples);
      auto next_loud = [&](int t) {
        while (t < t1 && tile_is_silent<G>(tile_s0(t), a.n_samples, valid)) ++t;
        return t;
      };
      int how = 0;
      auto fetch = [&](long long s) {            // start filling wave_s (dead at this point)
        how = load_tile<G, G::NW_FFT>(wave_s, cptr, s, a.n_samples, valid, a.tma_ok != 0, &s_mbar, warp, lane, -1,
                                          tab.tma_iss[warp], tab.n_tma_iss, pptr, a.pcm_channels);
      };
      auto do_s1 = [&](int i) __attribute__((always_inline)) {   // tile i: wave -> Yre[i & 1], Yim
        if (how & 2) nbar_sync(BAR_WAVE, NTF);                   // rows written by threads
        if (how & 1) {
          mbar_wait(&s_mbar, parity);
          parity ^= 1u;
        }
        LM_STAMP(1)
        if (i >= 2) nbar_sync((i & 1) ? BAR_PEMPTY1 : BAR_PEMPTY0, NT);   // mel(i - 2) has left this plane
        T* Yre = Yre0 + (i & 1) * G::YRE_ELEMS;
#pragma unroll 1
        for (int k = 0; k < G::S1_MAX; ++k) {
          const int task = tab.s1_tasks[warp][k];
          if (task < 0) break;
          stage1_task_ri<G, T>(wave_s, Yre, Yim, s1_s, task, lane);
        }
      };

      // (one call site of do_s1 -- one copy of the stage-1 codelet in the instruction stream)
      int t = next_loud(t0);
      if (t < t1) fetch(tile_s0(t));
      while (t < t1) {
        do_s1(it);
        LM_STAMP(2)
        nbar_sync(BAR_YDONE, NTF);                 // Y(it) complete, waveform tile dead
        const int tn = next_loud(t + 1);
        const bool pre = tn < t1;
        if (pre) fetch(tile_s0(tn));
        LM_STAMP(3)
        T* Yre = Yre0 + (it & 1) * G::YRE_ELEMS;
#pragma unroll 1
        for (int k = 0; k < G::S2_MAX; ++k) {
          const int k1 = tab.s2_rows[warp][k];
          if (k1 < 0) break;
          stage2_task_inplace<G, T>(Yre, Yim, k1, lane);
        }
        LM_STAMP(4)
        nbar_arrive((it & 1) ? BAR_PFULL1 : BAR_PFULL0, NT);      // P(it) is in Yre[it & 1]
        nbar_sync(BAR_S2DONE, NTF);                // every FFT warp is done reading Yim
        ++it;
#ifdef LM_TIMELINE
        ++tl_tile;
#endif
        t = tn;
      }
    }
    return;
  }

  // =======================================================================================
  // mel warps
  // =======================================================================================
  const float log_floor = a.log_floor, log_add = a.log_add, log_scale = a.log_scale;
  float silent_val = 0.0f;
  if (KIND == 1 || KIND == 3) silent_val = vlog2_clamp(0.0f, log_floor) * log_scale;
  if (KIND == 2) silent_val = vlog2_add(0.0f, log_add) * log_scale;
  const bool track = NORM && (t1 - t0 <= kMaxLocalTiles);
  const bool leader = threadIdx.x == NTF;           // lane 0 of the first mel warp

  // ---- deferred normalisation of a finished clip (see logmel_kernel.cuh), mel warps only ----
  int pend_clip = -1, pend_valid = 0, pend_par = 0;
  auto resolve = [&]() {      // uniform over the mel warps; one mel-warp barrier inside
    if (leader) {
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
    nbar_sync(BAR_MEL, NTM);
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
          for (int w = 0; w < G::NW_MEL; ++w) tm = fminf(tm, s_tmin[pend_par][t - t0][w]);
        fix = !(tm >= thr);
      } else {
        silent = tile_is_silent<G>(tile_s0(t), a.n_samples, pend_valid);
        fix = true;
      }
      if (silent) {
        for (int m = mw; m < a.n_mels; m += G::NW_MEL) {
          float* row = oc + (long long)m * a.n_frames + fa;
          for (int j = lane; j < len; j += 32) __stcs(row + j, cval);
        }
      } else if (fix) {
        if (a.vec_ok && (len & 3) == 0) {
          // <= 16 float4 per row: two rows per warp pass, eight passes in flight -- the lines come
          // from L2 (~700 cycles), so the loads of a batch are all issued before the first use
          const int q = len >> 2;
          const int j = lane & 15;
          constexpr int UNR = 8;
          constexpr int STEP = 2 * G::NW_MEL;
          for (int m0 = 2 * mw + (lane >> 4); m0 < a.n_mels; m0 += UNR * STEP) {
            float4 v[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
              const int m = m0 + u * STEP;
              if (m < a.n_mels && j < q)
                v[u] = __ldcg(reinterpret_cast<const float4*>(oc + (long long)m * a.n_frames + fa) + j);
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
              const int m = m0 + u * STEP;
              if (m < a.n_mels && j < q) {
                float4 w = v[u];
                w.x = fmaxf(w.x, cval); w.y = fmaxf(w.y, cval); w.z = fmaxf(w.z, cval); w.w = fmaxf(w.w, cval);
                __stcs(reinterpret_cast<float4*>(oc + (long long)m * a.n_frames + fa) + j, w);
              }
            }
          }
        } else {
          for (int m = mw; m < a.n_mels; m += G::NW_MEL) {
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
    int valid = a.n_samples;
    if (a.lengths) valid = min(max(a.lengths[clip], 0), a.n_samples);
    float* oc = a.out + (long long)clip * a.n_mels * a.n_frames;
    float rmax = -INFINITY;
    LM_CSTAMP(0)
    auto next_loud = [&](int t) {
      while (t < t1 && tile_is_silent<G>(tile_s0(t), a.n_samples, valid)) ++t;
      return t;
    };

    // ---- silent tiles (all zero padding) are never computed
    for (int t = t0; t < t1; ++t) {
      const bool silent = tile_is_silent<G>(tile_s0(t), a.n_samples, valid);
      if (NORM && track && leader) s_silent[par][t - t0] = silent ? 1 : 0;
      if (!silent) continue;
      rmax = fmaxf(rmax, silent_val);
      if (!NORM) {                            // no normalisation: the constant can be written right away
        const int f0 = t * G::F, fend = min(f0 + G::F, a.n_frames);
        for (int m = mw; m < a.n_mels; m += G::NW_MEL)
          for (int f = f0 + lane; f < fend; f += 32) oc[(long long)m * a.n_frames + f] = silent_val;
      }
    }

    auto do_mel = [&](int t, const T* P) __attribute__((always_inline)) {
      const int f0 = t * G::F;
      const int f = f0 + lane;
      const bool full = f0 + G::F <= a.n_frames;
      float* op = oc + (long long)tab.mel_begin[mw] * a.n_frames + f;
      const long long ostep = a.n_frames;
      float tmin = INFINITY;
      // the long-latency part of a finished filter (MUFU.LG2); its consumer runs one filter later
      auto start = [&](T acc) __attribute__((always_inline)) -> T {
        if (KIND == 1) return vlog2_clamp(acc, log_floor);
        if (KIND == 2) return vlog2_add(acc, log_add);
        if (KIND == 3) return vlog2_raw(acc);
        return acc;
      };
      const bool ok0 = f < a.n_frames, ok1 = (G::PK == 2) && (f + 32 < a.n_frames);
      auto finish_full = [&](T lg) __attribute__((always_inline)) {
        T v = KIND == 0 ? lg : vmuls(lg, log_scale);
        if (NORM) {
          rmax = vhmax(rmax, v);
          tmin = vhmin(tmin, v);
          v = vaffine(v, 0.25f, 1.0f);
        }
        op[0] = vlo(v);
        if (G::PK == 2) op[32] = vhi(v);
        op += ostep;
      };
      auto finish_part = [&](T lg) __attribute__((always_inline)) {
        T v = KIND == 0 ? lg : vmuls(lg, log_scale);
        if (NORM) {
          if (ok0) { rmax = fmaxf(rmax, vlo(v)); tmin = fminf(tmin, vlo(v)); }
          if (ok1) { rmax = fmaxf(rmax, vhi(v)); tmin = fminf(tmin, vhi(v)); }
          v = vaffine(v, 0.25f, 1.0f);
        }
        if (ok0) op[0] = vlo(v);
        if (ok1) op[32] = vhi(v);
        op += ostep;
      };
      if (full) mel_task_pipe<G, T>(P, tab, mw, lane, start, finish_full);
      else mel_task_pipe<G, T>(P, tab, mw, lane, start, finish_part);
      if (track) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
        if (lane == 0) s_tmin[par][t - t0][mw] = tmin;
      }
    };

    bool first = true;
    for (int t = next_loud(t0); t < t1; t = next_loud(t + 1), ++it) {
      LM_STAMP(2)
      nbar_sync((it & 1) ? BAR_PFULL1 : BAR_PFULL0, NT);        // P(it) complete in Yre[it & 1]
      LM_STAMP(5)
      do_mel(t, Yre0 + (it & 1) * G::YRE_ELEMS);
      LM_STAMP(6)
      nbar_arrive((it & 1) ? BAR_PEMPTY1 : BAR_PEMPTY0, NT);    // the plane may be overwritten
      if (NORM && first && pend_clip >= 0) resolve();           // the previous clip, one tile into this one
      first = false;
#ifdef LM_TIMELINE
      ++tl_tile;
#endif
    }

    LM_CSTAMP(1)
    if (NORM) {
      if (pend_clip >= 0) resolve();                            // this clip had no computed tile
      // ---- publish this CTA's maximum: warp shuffle -> mel warps -> slot of the clip group
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
      if (lane == 0) s_red[mw] = rmax;
      nbar_sync(BAR_MEL, NTM);
      if (leader) {
        float m = s_red[0];
        for (int w = 1; w < G::NW_MEL; ++w) m = fmaxf(m, s_red[w]);
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
    nbar_sync(BAR_MEL, NTM);      // s_cta_max / s_tmin of the last clip are complete
    resolve();
  }
}

}  // namespace lm
