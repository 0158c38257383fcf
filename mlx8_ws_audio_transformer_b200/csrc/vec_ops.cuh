// Value types for the STFT codelets.
//
//   float      one frame per lane (scalar FFMA path, also the host emulation type)
//   lm::f32x2  two frames per lane, packed in one 64-bit register pair; on sm_100a every
//              operation below is a single FADD2 / FMUL2 / FFMA2 (Blackwell's packed fp32
//              pipe instructions), and a scalar second operand uses their broadcast form,
//              so window samples, twiddles and filter weights stay one register wide.
//
// The same header compiles on the host (g++ -x c++) where f32x2 is a plain pair of floats;
// tests/ uses that build to check the index maps and codelets without a GPU.
#pragma once

#if defined(__CUDACC__)
#define LM_HD __host__ __device__ __forceinline__
#define LM_D __device__ __forceinline__
#else
#define LM_HD inline
#define LM_D inline
#endif

#if !defined(__CUDACC__)
struct alignas(16) float4 { float x, y, z, w; };   // host emulation build only
#endif

namespace lm {

struct f32x2 {
#if defined(__CUDA_ARCH__)
  unsigned long long v;
#else
  float lo, hi;
#endif
};

// ---- float -----------------------------------------------------------------------------
template <typename T> LM_HD T vzero();
template <> LM_HD float vzero<float>() { return 0.0f; }
LM_HD float vadd(float a, float b) { return a + b; }
LM_HD float vsub(float a, float b) { return a - b; }
LM_HD float vneg(float a) { return -a; }
LM_HD float vmulc(float a, float k) { return a * k; }
LM_HD float vfmac(float a, float k, float c) { return __builtin_fmaf(a, k, c); }
LM_HD float vmul(float a, float b) { return a * b; }
LM_HD float vfma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
LM_HD float vfnma(float a, float b, float c) { return __builtin_fmaf(-a, b, c); }
LM_HD float vmuls(float a, float s) { return a * s; }
LM_HD float vfmas(float a, float s, float c) { return __builtin_fmaf(a, s, c); }
LM_HD float vfnmas(float a, float s, float c) { return __builtin_fmaf(-a, s, c); }
LM_HD float vlo(float a) { return a; }
LM_HD float vhi(float a) { return a; }

// ---- f32x2 -----------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
LM_D f32x2 vpack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
LM_D float vlo(f32x2 a) { return __uint_as_float((unsigned)(a.v & 0xffffffffull)); }
LM_D float vhi(f32x2 a) { return __uint_as_float((unsigned)(a.v >> 32)); }
LM_D f32x2 vadd(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
LM_D f32x2 vsub(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
LM_D f32x2 vmul(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
LM_D f32x2 vfma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
#else
LM_HD f32x2 vpack(float lo, float hi) {
  f32x2 r;
  r.lo = lo;
  r.hi = hi;
  return r;
}
LM_HD float vlo(f32x2 a) { return a.lo; }
LM_HD float vhi(f32x2 a) { return a.hi; }
LM_HD f32x2 vadd(f32x2 a, f32x2 b) { return vpack(a.lo + b.lo, a.hi + b.hi); }
LM_HD f32x2 vsub(f32x2 a, f32x2 b) { return vpack(a.lo - b.lo, a.hi - b.hi); }
LM_HD f32x2 vmul(f32x2 a, f32x2 b) { return vpack(a.lo * b.lo, a.hi * b.hi); }
LM_HD f32x2 vfma(f32x2 a, f32x2 b, f32x2 c) {
  return vpack(__builtin_fmaf(a.lo, b.lo, c.lo), __builtin_fmaf(a.hi, b.hi, c.hi));
}
#endif

template <> LM_HD f32x2 vzero<f32x2>() { return vpack(0.0f, 0.0f); }
// the 64-bit image of a pair (low word = first element), as stored in constant tables
LM_HD f32x2 vfrombits(unsigned long long b) {
  f32x2 r;
#if defined(__CUDA_ARCH__)
  r.v = b;
#else
  const unsigned lo = (unsigned)b, hi = (unsigned)(b >> 32);
  __builtin_memcpy(&r.lo, &lo, 4);
  __builtin_memcpy(&r.hi, &hi, 4);
#endif
  return r;
}

// ---- epilogue helpers: log2 of a clamped / offset value, horizontal max / min ------------
LM_HD float lm_log2(float x) {   // x is a normal positive number here
#if defined(__CUDA_ARCH__)
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return __builtin_log2f(x);
#endif
}
LM_HD float vlog2_clamp(float v, float floor) { return lm_log2(__builtin_fmaxf(v, floor)); }
LM_HD f32x2 vlog2_clamp(f32x2 v, float floor) {
  return vpack(lm_log2(__builtin_fmaxf(vlo(v), floor)), lm_log2(__builtin_fmaxf(vhi(v), floor)));
}
LM_HD float vlog2_raw(float v) { return lm_log2(v); }                      // log2(0) = -inf: clamped later
LM_HD f32x2 vlog2_raw(f32x2 v) { return vpack(lm_log2(vlo(v)), lm_log2(vhi(v))); }
LM_HD float vlog2_add(float v, float add) { return lm_log2(v + add); }
LM_HD f32x2 vlog2_add(f32x2 v, float add) { return vpack(lm_log2(vlo(v) + add), lm_log2(vhi(v) + add)); }
LM_HD float vhmax(float r, float v) { return __builtin_fmaxf(r, v); }
LM_HD float vhmax(float r, f32x2 v) { return __builtin_fmaxf(r, __builtin_fmaxf(vlo(v), vhi(v))); }
LM_HD float vhmin(float r, float v) { return __builtin_fminf(r, v); }
LM_HD float vhmin(float r, f32x2 v) { return __builtin_fminf(r, __builtin_fminf(vlo(v), vhi(v))); }
LM_HD float vadds(float a, float s) { return a + s; }
LM_HD f32x2 vneg(f32x2 a) { return vmul(a, vpack(-1.0f, -1.0f)); }
LM_HD f32x2 vmulc(f32x2 a, float k) { return vmul(a, vpack(k, k)); }
LM_HD f32x2 vfmac(f32x2 a, float k, f32x2 c) { return vfma(a, vpack(k, k), c); }
LM_HD f32x2 vfnma(f32x2 a, f32x2 b, f32x2 c) { return vfma(vneg(a), b, c); }
LM_HD f32x2 vmuls(f32x2 a, float s) { return vmul(a, vpack(s, s)); }
LM_HD f32x2 vfmas(f32x2 a, float s, f32x2 c) { return vfma(a, vpack(s, s), c); }
LM_HD f32x2 vfnmas(f32x2 a, float s, f32x2 c) { return vfma(a, vpack(-s, -s), c); }
LM_HD f32x2 vadds(f32x2 a, float s) { return vadd(a, vpack(s, s)); }
// a * k + c with scalar k, c: one FFMA / FFMA2.  With k = 0.25, c = 1 this is bit-identical to
// (a + 4) / 4 (scaling by a power of two commutes with rounding).
LM_HD float vaffine(float a, float k, float c) { return __builtin_fmaf(a, k, c); }
LM_HD f32x2 vaffine(f32x2 a, float k, float c) { return vfma(a, vpack(k, k), vpack(c, c)); }

}  // namespace lm
