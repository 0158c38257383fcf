// Microbenchmark: how fast do the real stage-1 / stage-2 codelets run per SM sub-partition as a
// function of resident warps (packed f32x2 vs scalar float)?  Data comes from registers only, so
// this isolates the FP32 pipe + scheduling from shared memory.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../mlx8_ws_audio_transformer_b200/csrc/codelets_gen.cuh"
using namespace lm;

template <typename T> __device__ T mk(float a, float b);
template <> __device__ float mk<float>(float a, float) { return a; }
template <> __device__ f32x2 mk<f32x2>(float a, float b) { return vpack(a, b); }

template <typename T, int WHICH>
__global__ void k(float* out, long long* cyc, int iters, float s) {
  T x[20];
  float w[20], tr[11], ti[11];
  for (int i = 0; i < 20; ++i) { x[i] = mk<T>(threadIdx.x * 0.01f + i * s, i + s); w[i] = 0.5f + 0.01f * i * s; }
  for (int i = 0; i < 11; ++i) { tr[i] = 0.9f + 0.001f * i * s; ti[i] = 0.1f * s + 0.002f * i; }
  T yr[11], yi[11], acc = mk<T>(0.f, 0.f);
  T p[20], zr[20], zi[20];
  for (int i = 0; i < 20; ++i) { zr[i] = x[i]; zi[i] = mk<T>(i * s, 1.0f); }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (WHICH == 1) {
      stage1_r20(x, w, tr, ti, yr, yi);
#pragma unroll
      for (int i = 0; i < 11; ++i) { x[i] = vadd(x[i], yr[i]); x[i + 9] = vadd(x[i + 9], yi[i]); }
    } else {
      stage2_c20(zr, zi, p);
#pragma unroll
      for (int i = 0; i < 20; ++i) { zr[i] = vfmac(p[i], 1e-3f, zr[i]); }
    }
  }
  long long t1 = clock64();
  for (int i = 0; i < 20; ++i) acc = vadd(acc, vadd(x[i], zr[i]));
  out[blockIdx.x * blockDim.x + threadIdx.x] = vlo(acc) + vhi(acc);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <typename T, int WHICH>
void run(const char* name, int warps, int ops) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 200;
  k<T, WHICH><<<148, warps * 32>>>(out, cyc, iters, 1.0001f);
  k<T, WHICH><<<148, warps * 32>>>(out, cyc, iters, 1.0001f);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per_task = (double)h / iters;
  double per_smsp = per_task / (warps / 4.0);
  printf("%-26s warps/SM=%2d cycles/task/warp=%7.1f  cycles/task/SMSP=%7.1f  (math ops %d -> pipe bound %d)\n", name, warps,
         per_task, per_smsp, ops, (int)(ops * (sizeof(T) == 8 ? 2 : 1)));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 12, 16, 24}) {
    run<f32x2, 1>("stage1_r20 packed", w, 144 + 20);
    run<float, 1>("stage1_r20 scalar", w, 144 + 20);
    run<f32x2, 2>("stage2_c20 packed", w, 264 + 20);
    run<float, 2>("stage2_c20 scalar", w, 264 + 20);
  }
  return 0;
}
