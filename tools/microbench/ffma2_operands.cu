// Microbenchmark: does an FFMA2 with three DISTINCT 64-bit register operands cost more FMA-pipe / register-file
// time than one whose multiplier is an immediate or a uniform register?  (stage 1 of logmel_tf_kernel feeds its
// window / twiddle constants from registers loaded with LDS; stage 2 uses immediates.)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_operands ffma2_operands.cu && ./ffma2_operands
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
struct Tab { u64 c[64]; };

template <int MODE>
__global__ void k(const __grid_constant__ Tab tab, float* out, long long* cyc, int iters, const u64* gc) {
  u64 a[8], b[8], w[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float f = threadIdx.x * 0.001f + i;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"(f), "f"(f + 1.0f));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b[i]) : "f"(f * 0.5f), "f"(f * 0.25f));
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = gc[i + (threadIdx.x & 1)];     // constants in (vector) registers, all distinct
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(b[i]), "l"(w[(i + 4 * r) & 15]));        // 3 distinct register pairs
        if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(b[i]), "l"(tab.c[(i + 8 * r) & 63]));    // uniform-register multiplier
        if (MODE == 2) {                                                                                                      // immediate multiplier (broadcast)
          u64 imm;
          asm("mov.b64 %0, {%1, %1};" : "=l"(imm) : "f"(0.30901699f));
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(b[i]), "l"(imm));
        }
        if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b[i]));                                      // FADD2, 2 register pairs
        if (MODE == 4) asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(a[i]) : "l"(b[i]));                                  // 2 distinct pairs
      }
    }
  }
  long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += __uint_as_float((unsigned)a[i]) + __uint_as_float((unsigned)(a[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps, const Tab& tab, const u64* gc) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  const int iters = 1000;
  k<MODE><<<148, warps * 32>>>(tab, out, cyc, iters, gc);
  k<MODE><<<148, warps * 32>>>(tab, out, cyc, iters, gc);
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / (iters * 32.0);
  printf("%-44s warps/SM=%2d  cycles/instr/warp=%5.2f  FMA-pipe cycles per instr per SMSP=%5.2f\n", name, warps, per, per / (warps / 4.0));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  Tab tab;
  for (int i = 0; i < 64; ++i) tab.c[i] = 0x3f8000003f800000ull + i;
  u64* gc;
  cudaMalloc(&gc, 64 * 8);
  cudaMemcpy(gc, tab.c, 64 * 8, cudaMemcpyHostToDevice);
  for (int w : {4, 8}) {
    run<0>("FFMA2 a += b * w   (3 distinct register pairs)", w, tab, gc);
    run<1>("FFMA2 a += b * UR  (uniform-register multiplier)", w, tab, gc);
    run<2>("FFMA2 a += b * imm (immediate multiplier)", w, tab, gc);
    run<4>("FFMA2 a += b * b   (2 distinct register pairs)", w, tab, gc);
    run<3>("FADD2 a += b       (2 register pairs)", w, tab, gc);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
