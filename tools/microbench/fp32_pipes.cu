// Microbenchmark: issue rate and dependent latency of FFMA / FFMA2 / FADD2 on sm_100a.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32_pipes fp32_pipes.cu && ./fp32_pipes
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int MODE>
__global__ void k(float* out, long long* cyc, int iters, float s) {
  // MODE 0: scalar FFMA (register operands), 1: FFMA2 (register operands), 2: FFMA2 with scalar-broadcast b,
  // 3: FADD2, 4: scalar FFMA alternating with FFMA2
  unsigned long long v[ILP];
  float f[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    f[i] = threadIdx.x * 0.001f + i;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "f"(f[i]), "f"(f[i] + 1.0f));
  }
  unsigned long long b2, c2;
  asm("mov.b64 %0, {%1, %2};" : "=l"(b2) : "f"(s), "f"(s * 1.0001f));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c2) : "f"(s * 0.5f), "f"(s * 0.25f));
  float bs = s, cs = s * 0.5f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) f[i] = fmaf(f[i], bs, cs);
      if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(b2), "l"(c2));
      if (MODE == 2) {
        unsigned long long bb;
        asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(bs));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(bb), "l"(c2));
      }
      if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(c2));
      if (MODE == 4) {
        if (i & 1) f[i] = fmaf(f[i], bs, cs);
        else asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(b2), "l"(c2));
      }
    }
  }
  long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i]));
    acc += lo + hi + f[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP, int MODE>
void run(const char* name, int warps) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<ILP, MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0001f);
  k<ILP, MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0001f);
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per_instr = (double)h / (iters * ILP);                       // cycles per instruction per warp
  double smsp_rate = (warps / 4.0) / per_instr;                       // warp-instr per cycle per SMSP
  printf("%-22s ILP=%2d warps/SM=%2d  cycles/instr/warp=%6.2f  instr/cycle/SMSP=%5.3f\n", name, ILP, warps, per_instr, smsp_rate);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<1, 0>("FFMA dependent", 4);
  run<1, 1>("FFMA2 dependent", 4);
  run<1, 3>("FADD2 dependent", 4);
  for (int w : {4, 8, 12, 16, 32}) {
    run<8, 0>("FFMA", w);
    run<8, 1>("FFMA2 (3 reg)", w);
    run<8, 2>("FFMA2 (scalar b)", w);
    run<8, 3>("FADD2", w);
    run<8, 4>("FFMA + FFMA2 mix", w);
  }
  run<2, 1>("FFMA2 (3 reg)", 12);
  run<4, 1>("FFMA2 (3 reg)", 12);
  run<2, 1>("FFMA2 (3 reg)", 4);
  run<4, 1>("FFMA2 (3 reg)", 4);
  return 0;
}
