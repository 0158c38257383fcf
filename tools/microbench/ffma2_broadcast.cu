// Does FFMA2 take a scalar operand that both halves share, without a MOV to duplicate it?  (No GPU needed.)
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -cubin -o /tmp/bc.cubin tools/microbench/ffma2_broadcast.cu
//   cuobjdump -sass /tmp/bc.cubin | grep FFMA2
//
// nvcc 12.9 answers yes -- `mov.b64 {s, s}` + `fma.rn.f32x2` becomes ONE instruction with a 32-bit register
// broadcast to both lanes:
//
//   FFMA2 R8, R8.F32x2.HI_LO, R0.F32, R10.F32x2.HI_LO ;
//   FFMA2 R8, R10.F32x2.HI_LO, R12.F32, R8.F32x2.HI_LO ;
//
// i.e. constants that are the same for the two packed values (two FRAMES per thread: window samples, twiddles)
// cost one register and no packing instruction; constants that differ per half (two COLUMNS per thread, the
// thread-per-frame kernel's stage 1) need a 64-bit pair each.  DESIGN.md 4.3.
#include <cstdint>
__global__ void k(const float2* __restrict__ x, const float* __restrict__ s, float2* out) {
  float2 a = x[threadIdx.x], c = x[threadIdx.x + 32];
  float sv = s[0], sw = s[1];
  unsigned long long xa, xc, ss, sd, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(xa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(xc) : "f"(c.x), "f"(c.y));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(sv));
  asm("mov.b64 %0, {%1, %1};" : "=l"(sd) : "f"(sw));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(xa), "l"(ss), "l"(xc));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(r) : "l"(xc), "l"(sd));
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
  out[threadIdx.x] = make_float2(lo, hi);
}
