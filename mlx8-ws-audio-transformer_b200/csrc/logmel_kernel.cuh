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
  const float* wave;
  long long clip_stride;
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
};

template <class G>
struct Lay {   // shared-memory budget of one CTA
  using T = typename ValT<G>::type;
  static constexpr size_t Y = (size_t)G::Y_ELEMS * sizeof(T);
  static constexpr size_t P = (size_t)G::P_ELEMS * sizeof(T);
  static constexpr size_t W = (size_t)G::WAVE_FLOATS * 4;
  // when all three do not fit, P reuses the waveform buffer (dead after stage 1) and the next
  // tile is fetched after the mel phase instead of behind it
  static constexpr bool ALIAS = Y + P + W > 225 * 1024;
  static constexpr size_t BYTES = ALIAS ? Y + (P > W ? P : W) : Y + P + W;
  static constexpr int MIN_CTAS = BYTES <= 112 * 1024 ? 2 : 1;
};

__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <class G>
__device__ __forceinline__ void load_tile(float* wave_s, const float* __restrict__ clip, long long s0,
                                          int n_samples, int valid) {
  const bool fast = (s0 >= 0) && (s0 + G::SPAN <= (long long)valid);
  if (fast) {
    const float* src = clip + s0;
    for (int r = threadIdx.x; r < G::SPAN; r += G::THREADS) cp_async4(wave_s + wave_index<G>(r), src + r);
  } else {
    for (int r = threadIdx.x; r < G::SPAN; r += G::THREADS)
      wave_s[wave_index<G>(r)] = load_sample(clip, (long)(s0 + r), n_samples, valid);
  }
  cp_async_commit();
}

template <class G>
__global__ void __launch_bounds__(G::THREADS, Lay<G>::MIN_CTAS)
logmel_kernel(const __grid_constant__ Tables<G> tab, const KArgs a) {
  using T = typename ValT<G>::type;
  constexpr bool ALIAS = Lay<G>::ALIAS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Y = reinterpret_cast<T*>(smem_raw);
  T* P = Y + G::Y_ELEMS;
  float* wave_s = ALIAS ? reinterpret_cast<float*>(P) : reinterpret_cast<float*>(P + G::P_ELEMS);
  __shared__ float s_red[G::NW];
  __shared__ float s_max;

  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int group_id = blockIdx.x / a.group;
  const int rank = blockIdx.x - group_id * a.group;
  const bool use_log = a.log_mode != LOG_NONE;

  for (int clip = group_id; clip < a.batch; clip += a.n_groups) {
    const float* cptr = a.wave + (long long)clip * a.clip_stride;
    int valid = a.n_samples;
    if (a.lengths) valid = min(max(a.lengths[clip], 0), a.n_samples);
    float* oc = a.out + (long long)clip * a.n_mels * a.n_frames;
    const int t0 = (int)((long long)rank * a.tiles_per_clip / a.group);
    const int t1 = (int)((long long)(rank + 1) * a.tiles_per_clip / a.group);
    float rmax = -INFINITY;

    if (t0 < t1) {
      load_tile<G>(wave_s, cptr, (long long)t0 * G::F * G::HOP - G::N / 2, a.n_samples, valid);
      cp_async_wait_all();
    }
    __syncthreads();

    for (int t = t0; t < t1; ++t) {
      const int f0 = t * G::F;
      // ---- stage 1: columns b = warp, warp + NW, ...
#pragma unroll 1
      for (int b = warp; b < G::N2; b += G::NW) stage1_task<G, T>(wave_s, Y, tab.s1, b, lane);
      __syncthreads();
      // the waveform tile is dead: prefetch the next one behind stage 2 and the mel phase
      if (!ALIAS && t + 1 < t1)
        load_tile<G>(wave_s, cptr, (long long)(t + 1) * G::F * G::HOP - G::N / 2, a.n_samples, valid);
      // ---- stage 2: rows k1 = 1..H1-1 on warps 0..NW-2, the two half-size rows on the last warp
      if (warp < G::NW - 1) {
        stage2_task<G, T>(Y, P, warp + 1, lane);
      } else {
        stage2_task<G, T>(Y, P, 0, lane);
        stage2_task<G, T>(Y, P, G::H1, lane);
      }
      __syncthreads();
      // ---- mel projection, log, store of the un-normalised value, running max
      mel_task<G, T>(P, tab, warp, lane, [&](int m, T acc) {
        float v0 = vlo(acc), v1 = vhi(acc);
        if (use_log) {
          v0 = __log2f(fmaxf(v0 + a.log_add, a.log_floor)) * a.log_scale;
          if (G::PK == 2) v1 = __log2f(fmaxf(v1 + a.log_add, a.log_floor)) * a.log_scale;
        }
        const int f = f0 + lane;
        float* o = oc + (long long)m * a.n_frames + f;
        if (f < a.n_frames) {
          o[0] = v0;
          rmax = fmaxf(rmax, v0);
        }
        if (G::PK == 2 && f + 32 < a.n_frames) {
          o[32] = v1;
          rmax = fmaxf(rmax, v1);
        }
      });
      if (ALIAS) {
        __syncthreads();   // P shares the waveform buffer: reload only after the mel phase
        if (t + 1 < t1)
          load_tile<G>(wave_s, cptr, (long long)(t + 1) * G::F * G::HOP - G::N / 2, a.n_samples, valid);
      }
      cp_async_wait_all();
      __syncthreads();
    }

    if (a.log_mode == LOG10_CLAMP_WHISPER_NORM) {
      // ---- clip maximum: warp shuffle -> CTA -> clip group (release/acquire counter)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
      if (lane == 0) s_red[warp] = rmax;
      __syncthreads();
      if (threadIdx.x == 0) {
        float m = s_red[0];
        for (int w = 1; w < G::NW; ++w) m = fmaxf(m, s_red[w]);
        if (a.group > 1) {
          float* slots = a.gmax + (long long)clip * a.group;
          __stcg(slots + rank, m);
          __threadfence();
          atomicAdd(a.gcnt + clip, 1);
          while (ld_acquire(a.gcnt + clip) < a.group) __nanosleep(64);
          for (int r = 0; r < a.group; ++r) m = fmaxf(m, __ldcg(slots + r));
        }
        if (rank == 0 && a.clip_max) a.clip_max[clip] = m;
        s_max = m;
      }
      __syncthreads();
      // ---- normalise this CTA's slab in place (still L2 resident)
      const float thr = s_max - 8.0f;
      const int fa = t0 * G::F;
      const int fb = min(t1 * G::F, a.n_frames);
      const int len = fb - fa;
      if (len > 0) {
        if (a.vec_ok) {
          const int q = len >> 2;
          for (int m = warp; m < a.n_mels; m += G::NW) {
            float4* row = reinterpret_cast<float4*>(oc + (long long)m * a.n_frames + fa);
            for (int j = lane; j < q; j += 32) {
              float4 v = __ldcg(row + j);
              v.x = (fmaxf(v.x, thr) + 4.0f) * 0.25f;
              v.y = (fmaxf(v.y, thr) + 4.0f) * 0.25f;
              v.z = (fmaxf(v.z, thr) + 4.0f) * 0.25f;
              v.w = (fmaxf(v.w, thr) + 4.0f) * 0.25f;
              __stcs(row + j, v);
            }
          }
        } else {
          for (int m = warp; m < a.n_mels; m += G::NW) {
            float* row = oc + (long long)m * a.n_frames + fa;
            for (int j = lane; j < len; j += 32) row[j] = (fmaxf(__ldcg(row + j), thr) + 4.0f) * 0.25f;
          }
        }
      }
      __syncthreads();   // s_red / s_max are reused by the next clip
    }
  }
}

}  // namespace lm
