// Microbenchmark: cycles of each per-tile phase (stage 1, stage 2, mel) in isolation, with the
// real core functions and tables but no global-memory traffic.  One CTA per SM.
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../mlx8_ws_audio_transformer_b200/csrc/logmel_tables.h"
#include "../../mlx8_ws_audio_transformer_b200/csrc/logmel_kernel.cuh"
using namespace lm;

template <class G>
__global__ void __launch_bounds__(G::THREADS, 1) phases(const __grid_constant__ Tables<G> tab, float* sink, long long* cyc, int iters, int mode) {
  using T = typename ValT<G>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Y = reinterpret_cast<T*>(smem_raw);
  T* P = Y + G::Y_ELEMS;
  float* wave_s = reinterpret_cast<float*>(P + G::P_ELEMS);
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  for (int i = threadIdx.x; i < G::WAVE_FLOATS; i += G::THREADS) wave_s[i] = 0.001f * (i % 977) - 0.4f;
  for (int i = threadIdx.x; i < G::Y_ELEMS; i += G::THREADS) Y[i] = vpack(0.01f * (i % 113), 0.02f * (i % 71));
  for (int i = threadIdx.x; i < G::P_ELEMS; i += G::THREADS) P[i] = vpack(1.0f + (i % 13), 2.0f + (i % 7));
  float s1c[G::S1_STRIDE];
  if (tab.s1_tasks[warp][0] >= 0) stage1_consts<G>(tab.s1, tab.s1_tasks[warp][0], lane, s1c);
  __syncthreads();
  float acc = 0.f;
  long long t[4] = {0, 0, 0, 0};
  for (int it = 0; it < iters; ++it) {
    long long c0 = clock64();
    if (mode < 8 && (mode & 1)) {
      const int ta = tab.s1_tasks[warp][0], tb = tab.s1_tasks[warp][1];
      if (ta >= 0) stage1_task_pair<G, T>(wave_s, Y, s1c, ta, tb, lane);
    }
    __syncthreads();
    long long c1 = clock64();
    if (mode < 8 && (mode & 2)) {
#pragma unroll 1
      for (int i = 0; i < G::S2_MAX; ++i) {
        const int k1 = tab.s2_rows[warp][i];
        if (k1 < 0) break;
        stage2_task<G, T>(Y, P, k1, lane);
      }
    }
    __syncthreads();
    long long c2 = clock64();
    if (mode < 8 && (mode & 4)) {
      mel_task<G, T>(P, tab, warp, lane, [&](T a) {
        T v = vmuls(vlog2_clamp(a, 1e-10f), 0.30103f);
        acc = vhmax(acc, v);
        v = vmulc(vadds(v, 4.0f), 0.25f);
        sink[(blockIdx.x * G::THREADS + threadIdx.x)] = vlo(v) + vhi(v);
      });
    }
    if (mode == 8) {
      if (warp < 6) {
        for (int task = warp; task < G::S1_TASKS; task += 6) {
          float c[G::S1_STRIDE];
          stage1_consts<G>(tab.s1, task, lane, c);
          stage1_task_c<G, T>(wave_s, Y + (G::Y_ELEMS / 2) * 0, c, task, lane);
        }
      } else {
        for (int k1 = warp - 6; k1 <= G::H1; k1 += 5) stage2_task<G, T>(Y, P, k1, lane);
      }
    }
    if (mode == 10 || mode == 11 || mode == 12) {   // S1 decomposition: 10 = no loads, 11 = no stores, 12 = neither
      const int ta = tab.s1_tasks[warp][0], tb = tab.s1_tasks[warp][1];
      if (ta >= 0) {
        for (int rep = 0; rep < 2; ++rep) {
          const int task = rep ? tb : ta;
          T x[G::N1];
          if (mode == 11) stage1_load<G, T>(wave_s, task, lane, x);
          else
            for (int i = 0; i < G::N1; ++i) x[i] = vpack(acc + i, acc - i + it);
          if (mode == 10) stage1_compute<G, T>(x, Y, s1c, task, lane);
          else {
            float w[G::N1], tr[G::H1 + 1], ti[G::H1 + 1];
            for (int i = 0; i < G::N1; ++i) w[i] = s1c[i];
            tr[0] = 1.f; ti[0] = 0.f;
            for (int k = 1; k <= G::H1; ++k) { tr[k] = s1c[G::N1 + 2 * (k - 1)]; ti[k] = s1c[G::N1 + 2 * (k - 1) + 1]; }
            T yr[G::H1 + 1], yi[G::H1 + 1];
            Codelets<G::N>::template s1<T>(x, w, tr, ti, yr, yi);
            T sacc = vzero<T>();
            for (int k = 0; k <= G::H1; ++k) sacc = vadd(sacc, vadd(yr[k], yi[k]));
            acc += vlo(sacc) * 1e-9f;
          }
        }
      }
    }
    if (mode >= 13 && mode <= 15) {   // S2 decomposition on the general rows: 13 = no loads, 14 = no stores, 15 = neither
      const int k1 = tab.s2_rows[warp][0];
      if (k1 > 0 && k1 < G::H1) {
        T yr[G::N2], yi[G::N2], p[G::N2];
        if (mode == 14) {
          const T* sre = Y + k1 * G::N2 * 32;
          const T* sim = Y + G::YRE_ELEMS + (k1 - 1) * G::N2 * 32;
#pragma unroll
          for (int b = 0; b < G::N2; ++b) { yr[b] = sre[b * 32 + y_slot<G>(lane, b)]; yi[b] = sim[b * 32 + y_slot<G>(lane, b)]; }
        } else {
#pragma unroll
          for (int b = 0; b < G::N2; ++b) { yr[b] = vpack(acc + b, 1.f); yi[b] = vpack(acc - b, 2.f); }
        }
        Codelets<G::N>::template s2<T>(yr, yi, p);
        if (mode == 13) {
#pragma unroll
          for (int j = 0; j < G::N2; ++j) P[((j >= G::N2 / 2) ? (G::N - G::N1 * j - k1) : (G::N1 * j + k1)) * 32 + lane] = p[j];
        } else {
          T sacc = vzero<T>();
#pragma unroll
          for (int j = 0; j < G::N2; ++j) sacc = vadd(sacc, p[j]);
          acc += vlo(sacc) * 1e-9f;
        }
      }
    }
    if (mode == 9) {   // same split of warps, but the two kinds of work one after the other
      if (warp < 6) {
        for (int task = warp; task < G::S1_TASKS; task += 6) {
          float c[G::S1_STRIDE];
          stage1_consts<G>(tab.s1, task, lane, c);
          stage1_task_c<G, T>(wave_s, Y, c, task, lane);
        }
      }
      __syncthreads();
      if (warp >= 6) {
        for (int k1 = warp - 6; k1 <= G::H1; k1 += 5) stage2_task<G, T>(Y, P, k1, lane);
      }
    }
    __syncthreads();
    long long c3 = clock64();
    t[0] += c1 - c0; t[1] += c2 - c1; t[2] += c3 - c2;
  }
  sink[blockIdx.x * G::THREADS + threadIdx.x] += acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) { cyc[0] = t[0]; cyc[1] = t[1]; cyc[2] = t[2]; }
}

int main() {
  using G = Geo<400, 160, 2>;
  static Tables<G> tab;
  std::vector<float> win = hann_periodic(400);
  // synthetic triangular bank, 128 filters over 201 bins
  std::vector<float> fb(201 * 128, 0.f);
  for (int m = 0; m < 128; ++m) {
    float c = 1.f + m * 198.f / 128.f, wdt = 1.2f + m * 0.03f;
    for (int k = 1; k < 200; ++k) { float d = 1.f - fabsf(k - c) / wdt; if (d > 0) fb[k * 128 + m] = d * 0.02f; }
  }
  std::string err = build_tables<G>(tab, win.data(), fb.data(), 128);
  printf("tables: %s scan=%d\n", err.empty() ? "ok" : err.c_str(), (int)tab.mel_scan);
  float* sink; long long* cyc;
  cudaMalloc(&sink, 148 * G::THREADS * 4); cudaMalloc(&cyc, 32);
  size_t smem = Lay<G>::BYTES;
  cudaFuncSetAttribute(phases<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 50;
  for (int mode : {2, 13, 14, 15}) {
    phases<G><<<148, G::THREADS, smem>>>(tab, sink, cyc, iters, mode);
    phases<G><<<148, G::THREADS, smem>>>(tab, sink, cyc, iters, mode);
    long long h[3]; cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d (%s%s%s): S1 %7.0f  S2 %7.0f  mel %7.0f  cycles/tile   [%s]\n", mode, mode & 1 ? "S1 " : "", mode & 2 ? "S2 " : "",
           mode & 4 ? "mel" : "", (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters, cudaGetErrorString(e));
  }
  return 0;
}
