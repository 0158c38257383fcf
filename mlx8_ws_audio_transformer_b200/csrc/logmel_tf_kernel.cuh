// Thread-per-frame form of the fused Whisper log-mel kernel: the steady-state hot path.
//
// The CTA-tiled kernels (logmel_kernel.cuh, logmel_ws_kernel.cuh) spread one frame over many
// threads, so the two transposes of the two-stage FFT (samples -> columns, columns -> rows) and
// the bin order of the mel projection all go through shared memory with index tables: ncu r01 shows
// them bound by shared-memory wavefronts (52 %) and FP32 issue (48 %) TOGETHER, and 60 % of their
// instructions are not floating point.  Here ONE THREAD OWNS ONE FRAME from the waveform to the
// stored features, and a warp owns 32 consecutive frames of one clip:
//
//   stage 1   the thread reads its 400 samples with LDS.128 (row pitch 164 floats: 41 quad-words,
//             odd, so the 8 lanes of a quarter-warp phase hit 8 different 16-byte banks) and runs
//             the windowed real DFT-20 of two adjacent columns at a time, packed in f32x2
//             (FFMA2/FADD2; the LDS.128 result registers ARE the packed pairs).  Window samples
//             and twiddles are the same for every lane: warp-uniform LDS.128 from a shared-memory
//             copy of the table (uniform-register loads, LDCU, were tried first: ptxas places them
//             a few instructions before their use and every FFMA2 waits on the short scoreboard,
//             5.65 vs 4.40 ms per 4096 clips).
//   transpose the 420 inter-stage values of the frame are parked in TENSOR MEMORY, used as
//             lane-private scratch: the warps of TMEM lane quadrant q own lanes 32q..32q+31, lane =
//             frame, column = (row pair, column b, re/im).  tcgen05.st / tcgen05.ld with the 32x32b
//             shape are exactly "each thread writes / reads N consecutive words of its own lane", so
//             the column -> row transpose costs no shared-memory bandwidth, no swizzle and no CTA
//             barrier (tools/microbench/tmem_scratch.cu: no MMA is needed to use TMEM).
//   stage 2   per ROW PAIR (k, k'): 80 values back from TMEM, two complex DFT-20 + |X|^2 as one
//             packed codelet (the TMEM column order makes (Y_k[b], Y_k'[b]) adjacent registers),
//             powers written over the pair in place; row 0 (real input) alone in scalar FP32.
//   mel       straight-line banded projection generated for the two Whisper banks
//             (tf_mel_gen.cuh): weights are immediates of the FFMAs, the sums of the frame live in
//             registers; then log2, running clip max / tile min (FMNMX3), (S+4)/4 as one FFMA and a
//             store that is coalesced across the warp (lane = consecutive frame).
//
// Two warps share a tile of 32 frames (and a TMEM lane quadrant, which one frame's 420 values nearly
// fill): both have lane = frame, role A takes column pairs 0-4 / row 0 and row pairs 0-1 / the low
// filters, role B column pairs 5-9 / row pairs 2-4 / the high filters, and they meet at three
// 64-thread named barriers per tile.  The two warps sit on the same SM sub-partition, so whenever
// one waits for shared or tensor memory the other issues.  A pair walks a whole clip, so the
// Whisper max - 8 rule needs no inter-CTA agreement (the per-tile minimum decides which tiles are
// revisited, as in the other kernels).  The waveform tile of the next 32 frames is fetched with
// cp.async (16-byte chunks, coalesced) in groups between the stage-2 codelets of the current one.
// Grid: one CTA of 4 pairs per SM (the CTA owns all 512 TMEM columns), no cooperative launch.
#pragma once
#include <type_traits>
#include <cstdint>
#include <cstring>

#include "logmel_kernel.cuh"
#include "tf_mel_gen.cuh"

namespace lm {

struct TfGeo {
  static constexpr int N = 400, HOP = 160, N1 = 20, N2 = 20, H1 = 10;
  static constexpr int F = 32;                          // frames per warp tile: lane = frame
  static constexpr int SPAN = (F - 1) * HOP + N;        // 5360 samples
  static constexpr int PITCH = HOP + 4;                 // 164 floats = 41 quad-words (odd)
  static constexpr int ROWS = (SPAN + HOP - 1) / HOP;   // 34
  static constexpr int TILE_FLOATS = ROWS * PITCH;
  static constexpr int PAIRS = 4;                       // warp pairs per CTA = TMEM lane quadrants
  static constexpr int WARPS = 2 * PAIRS;
  static constexpr int THREADS = WARPS * 32;
  static constexpr int Y_COLS = N2 + 2 * N2 * H1;       // 420 TMEM columns per frame
  static constexpr int PCM_STAGE_BYTES = SPAN * 4;       // raw 16-bit frames of a tile: 2 channels x 2 bytes at most
  static constexpr size_t SMEM = (size_t)PAIRS * (TILE_FLOATS * 4 + PCM_STAGE_BYTES);
  // ask for more than half of the SM's shared memory so that two CTAs (each allocating all of
  // TMEM) can never be co-resident
  static constexpr size_t SMEM_REQUEST = SMEM > 120 * 1024 ? SMEM : 120 * 1024;
};
constexpr int kTfMaxTiles = 128;

// stage-1 constants of column pair c = b / 2, packed (b, b + 1); twiddles of k1 = 1..10 at [k1 - 1]
// (each entry is the 64-bit image of an f32x2: low word = column b, high word = column b + 1, so a
// constant reaches the FFMA2 as one aligned uniform-register pair without any packing move)
struct alignas(16) TfPairConsts {      // one column pair: 640 bytes, read with warp-uniform LDS.128
  unsigned long long w[20], nw[20];
  unsigned long long twr[10], ntwr[10], twi[10], ntwi[10];
};
struct alignas(16) TfTables {
  TfPairConsts cp[10];
};
__host__ __device__ inline unsigned long long tf_pack2(float lo, float hi) {
  unsigned a, b;
  memcpy(&a, &lo, 4);
  memcpy(&b, &hi, 4);
  return (unsigned long long)a | ((unsigned long long)b << 32);
}

// ---- tensor memory as lane-private scratch -------------------------------------------------
#define LM_TM_R4(r, o) "f"(r[o]), "f"(r[o + 1]), "f"(r[o + 2]), "f"(r[o + 3])
#define LM_TM_W4(r, o) "=f"(r[o]), "=f"(r[o + 1]), "=f"(r[o + 2]), "=f"(r[o + 3])
__device__ __forceinline__ void tm_st1(uint32_t addr, float a) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ void tm_st2(uint32_t addr, float a, float b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void tm_st4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void tm_st8v(uint32_t addr, float a, float b, float c, float d, float e, float f, float g, float h) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "f"(a), "f"(b),
               "f"(c), "f"(d), "f"(e), "f"(f), "f"(g), "f"(h)
               : "memory");
}
__device__ __forceinline__ void tm_st32(uint32_t addr, const float* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(addr),
      LM_TM_R4(r, 0), LM_TM_R4(r, 4), LM_TM_R4(r, 8), LM_TM_R4(r, 12), LM_TM_R4(r, 16), LM_TM_R4(r, 20), LM_TM_R4(r, 24),
      LM_TM_R4(r, 28)
      : "memory");
}
__device__ __forceinline__ void tm_st8(uint32_t addr, const float* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr),
               LM_TM_R4(r, 0), LM_TM_R4(r, 4)
               : "memory");
}
__device__ __forceinline__ void tm_st16(uint32_t addr, const float* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16};" ::"r"(addr),
      LM_TM_R4(r, 0), LM_TM_R4(r, 4), LM_TM_R4(r, 8), LM_TM_R4(r, 12)
      : "memory");
}
__device__ __forceinline__ void tm_ld4(uint32_t addr, float* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : LM_TM_W4(r, 0) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tm_ld8(uint32_t addr, float* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : LM_TM_W4(r, 0), LM_TM_W4(r, 4)
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t addr, float* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : LM_TM_W4(r, 0), LM_TM_W4(r, 4), LM_TM_W4(r, 8), LM_TM_W4(r, 12)
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tm_ld32(uint32_t addr, float* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : LM_TM_W4(r, 0), LM_TM_W4(r, 4), LM_TM_W4(r, 8), LM_TM_W4(r, 12), LM_TM_W4(r, 16), LM_TM_W4(r, 20),
        LM_TM_W4(r, 24), LM_TM_W4(r, 28)
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tm_ld64(uint32_t addr, float* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, "
      "%38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, "
      "%60, %61, %62, %63}, [%64];"
      : LM_TM_W4(r, 0), LM_TM_W4(r, 4), LM_TM_W4(r, 8), LM_TM_W4(r, 12), LM_TM_W4(r, 16), LM_TM_W4(r, 20),
        LM_TM_W4(r, 24), LM_TM_W4(r, 28), LM_TM_W4(r, 32), LM_TM_W4(r, 36), LM_TM_W4(r, 40), LM_TM_W4(r, 44),
        LM_TM_W4(r, 48), LM_TM_W4(r, 52), LM_TM_W4(r, 56), LM_TM_W4(r, 60)
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// TMEM columns of a frame (420 of the 512).  Row 0 of Y is real: 20 columns, column b at [b].  Rows
// k1 = 1..10 are kept as the five ROW PAIRS q = (2q+1, 2q+2), 80 columns each, interleaved so that
// stage 2 can run both rows of a pair in one packed (f32x2) codelet:
//   [pair_base(q) + 4 b + {0, 1, 2, 3}] = re_k(b), re_k'(b), im_k(b), im_k'(b)        k = 2q+1, k' = 2q+2
// A load of the 80 columns returns (re_k(b), re_k'(b)) and (im_k(b), im_k'(b)) as adjacent registers
// = the packed operands.  A stage-1 codelet holds (column b, column b+1) packed instead; it writes the
// eight columns of (b, b+1) x (k, k') x (re, im) with eight single-column stores, so the 2 x 2
// transposes cost no register moves (the tensor-memory pipe is ~1 % busy).
// Stage 2 writes the powers back in place: [pair_base(q) + 2 j + {0, 1}] = P_k(j), P_k'(j).
__device__ __forceinline__ constexpr int tf_pair_base(int q) { return 20 + 80 * q; }

// one stage-1 codelet: the column pair cp = (2 cp, 2 cp + 1) of this lane's frame, taken from the .xy
// (HALF 0) or .zw (HALF 1) halves of the group's LDS.128 results.  The column pair enters only
// through the table pointer and two tensor-memory base addresses, so ONE copy of this code serves
// all ten column pairs (the instruction stream, not the arithmetic, was the cost of unrolling them:
// 102 KB of code, 8 % of warp time waiting for instructions).
__device__ __forceinline__ void tf_stage1_pair(const TfPairConsts* __restrict__ c, const f32x2 (&x)[20], uint32_t tm2,
                                               uint32_t tm8) {
  f32x2 w[20], nw[20], twr[11], ntwr[11], twi[11], ntwi[11], yr[11], yi[11];
#pragma unroll
  for (int i = 0; i < 20; ++i) {
    w[i] = vfrombits(c->w[i]);
    nw[i] = vfrombits(c->nw[i]);
  }
  twr[0] = ntwr[0] = twi[0] = ntwi[0] = vzero<f32x2>();     // unused by the codelet
#pragma unroll
  for (int k = 1; k <= 10; ++k) {
    twr[k] = vfrombits(c->twr[k - 1]);
    ntwr[k] = vfrombits(c->ntwr[k - 1]);
    twi[k] = vfrombits(c->twi[k - 1]);
    ntwi[k] = vfrombits(c->ntwi[k - 1]);
  }
  stage1_r20p<f32x2>(x, w, twr, twi, ntwi, nw, ntwr, yr, yi);
  tm_st2(tm2, vlo(yr[0]), vhi(yr[0]));                       // row 0: tm + 2 cp
#pragma unroll
  for (int q = 0; q < 5; ++q) {                              // row pairs: tm + pair_base(q) + 8 cp
    const int k = 2 * q + 1, k2 = 2 * q + 2;
    // eight single-column stores straight from the registers the values were computed in: one 8-column
    // store needs them as a register vector in (b, re/im, k) order, i.e. ~8 MOVs per store (3.93 vs 3.85 ms)
    const float v8[8] = {vlo(yr[k]), vlo(yr[k2]), vlo(yi[k]), vlo(yi[k2]), vhi(yr[k]), vhi(yr[k2]), vhi(yi[k]), vhi(yi[k2])};
#pragma unroll
    for (int e = 0; e < 8; ++e) tm_st1(tm8 + tf_pair_base(q) + e, v8[e]);
  }
}

// the samples of columns 4 g4 .. 4 g4 + 3 (two column pairs) of this lane's frame: 20 conflict-free LDS.128
__device__ __forceinline__ void tf_load_group(const float* mine, int g4, f32x2 (&xa)[20], f32x2 (&xb)[20]) {
#pragma unroll
  for (int i = 0; i < 20; ++i) {
    const int n = 20 * i;       // sample n = 20 i + 4 g4 .. + 3 of the frame; a hop row holds 160 samples + 4 pad
    const float4 v = *reinterpret_cast<const float4*>(mine + n + 4 * (n / TfGeo::HOP) + 4 * g4);
    xa[i] = vpack(v.x, v.y);
    xb[i] = vpack(v.z, v.w);
  }
}
// one column pair alone (columns 2 cp, 2 cp + 1): 20 LDS.64 -- two-way bank conflicts, i.e. the LSU time
// of the LDS.128 they replace, for one copy of the code instead of one per half
__device__ __forceinline__ void tf_load_pair(const float* mine, int cp, f32x2 (&x)[20]) {
#pragma unroll
  for (int i = 0; i < 20; ++i) {
    const int n = 20 * i;
    const float2 v = *reinterpret_cast<const float2*>(mine + n + 4 * (n / TfGeo::HOP) + 2 * cp);
    x[i] = vpack(v.x, v.y);
  }
}

// one row pair of stage 2: 80 columns in, two complex DFT-20 + |X|^2 as one packed codelet, 40 columns
// out (in place).  For the pair (9, 10) row 10 runs through the general codelet as well; its outputs
// j >= 10 repeat bins it already has and are not used.
// `between` runs while the tensor-memory loads are in flight (the cp.async issue of the next tile).
template <class Between>
__device__ __forceinline__ void tf_stage2_pair(uint32_t base, Between between) {
  float y[80];
  tm_ld64(base, y);
  tm_ld16(base + 64, y + 64);
  between();
  tm_wait_ld();
  f32x2 yr[20], yi[20], p[20];
#pragma unroll
  for (int b = 0; b < 20; ++b) {
    yr[b] = vpack(y[4 * b], y[4 * b + 1]);
    yi[b] = vpack(y[4 * b + 2], y[4 * b + 3]);
  }
  stage2_c20<f32x2>(yr, yi, p);
  float o[40];
#pragma unroll
  for (int j = 0; j < 20; ++j) {
    o[2 * j] = vlo(p[j]);
    o[2 * j + 1] = vhi(p[j]);
  }
  tm_st32(base, o);
  tm_st8(base + 32, o + 32);
}

__device__ __forceinline__ void pair_sync(int pair) {     // the two warps of a pair: named barrier 1 + pair
  asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
}

// mel projection of role R over the 11 rows of P, then log / clip statistics / store
template <int NM, int R>
__device__ __forceinline__ void tf_mel_store(uint32_t tm, float* op, long long ostep, float q_scale, float& hi_out,
                                             float& lo_out) {
  using P = TfMelPattern<NM>;
  constexpr int M0 = R == 0 ? 0 : P::M0, M1 = R == 0 ? P::M0 : NM;
  float acc[P::MAXHALF];
#pragma unroll
  for (int m = 0; m < P::MAXHALF; ++m) acc[m] = 0.0f;
  float p0[12];
  tm_ld8(tm, p0);
  tm_ld4(tm + 8, p0 + 8);
  float pa[40], pb[40];
  tm_ld32(tm + tf_pair_base(0), pa);
  tm_ld8(tm + tf_pair_base(0) + 32, pa + 32);
  tm_wait_ld();
  tf_mel_row0<NM, R>(p0, acc);
  // row pairs alternate between two register sets: the next pair is on its way while this one is used
#define LM_TF_PAIR(Q, CUR, NXT)                                          \
  if (Q < 4) {                                                           \
    tm_ld32(tm + tf_pair_base(Q + 1), NXT);                              \
    tm_ld8(tm + tf_pair_base(Q + 1) + 32, NXT + 32);                     \
  }                                                                      \
  tf_mel_pair<NM, Q, R>(CUR, acc);                                       \
  if (Q < 4) tm_wait_ld();
  LM_TF_PAIR(0, pa, pb)
  LM_TF_PAIR(1, pb, pa)
  LM_TF_PAIR(2, pa, pb)
  LM_TF_PAIR(3, pb, pa)
  LM_TF_PAIR(4, pa, pb)
#undef LM_TF_PAIR
  // running maximum / tile minimum of log2(mel): four independent chains of 3-input min / max
  float hi[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, lo[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
  op += M0 * ostep;
#pragma unroll
  for (int m = 0; m < M1 - M0; m += 2) {
    const float l0 = vlog2_raw(acc[m]), l1 = vlog2_raw(acc[m + 1]);
    hi[(m >> 1) & 3] = fmaxf(fmaxf(hi[(m >> 1) & 3], l0), l1);
    lo[(m >> 1) & 3] = fminf(fminf(lo[(m >> 1) & 3], l0), l1);
    op[m * ostep] = vaffine(l0, q_scale, 1.0f);
    op[(m + 1) * ostep] = vaffine(l1, q_scale, 1.0f);
  }
  hi_out = fmaxf(fmaxf(hi[0], hi[1]), fmaxf(hi[2], hi[3]));
  lo_out = fminf(fminf(lo[0], lo[1]), fminf(lo[2], lo[3]));
}

// NM: 80 or 128 (the two generated banks).  NF: frames per clip when known at compile time (3000:
// the store offsets m * NF become immediates), 0: read from the arguments.
// Whisper normalisation: log10, clip max - 8, (S + 4) / 4.
// SPLIT: the instantiation that can split a clip between warp pairs / CTAs (a.group >= 4).  The one-clip-per-pair
// steady state is its own instantiation so that its tile loop keeps compile-time bounds: the general loop measured
// 1.6 % slower on 4096 clips (3.89 vs 3.83 ms, A/B on one box; 3.85 with the split out).
template <int NM, int NF, bool SPLIT>
__global__ void __launch_bounds__(TfGeo::THREADS, 1)
logmel_tf_kernel(const __grid_constant__ TfTables ctab, const KArgs a) {
  using G = TfGeo;
  using MP = TfMelPattern<NM>;
  // Stage-1 constants as shared-memory operands (warp-uniform LDS.128, ~2 LSU cycles each) instead of
  // uniform-register loads: ptxas places an LDCU only a few instructions before its first use and
  // every FFMA2 then waits ~30 cycles on the short scoreboard (measured: 5.65 vs 4.40 ms per 4096 clips).
  __shared__ TfTables s_tab;
  for (int i = threadIdx.x; i < (int)(sizeof(TfTables) / 8); i += blockDim.x)
    reinterpret_cast<unsigned long long*>(&s_tab)[i] = reinterpret_cast<const unsigned long long*>(&ctab)[i];
  const TfTables& tab = s_tab;
  static_assert(MP::M0 % 2 == 0 && NM % 2 == 0, "the epilogue walks filters two at a time");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_tmem;
  __shared__ float s_tmin[G::PAIRS][kTfMaxTiles][2];
  __shared__ float s_pmax[2][G::PAIRS][2];

  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int pair = warp & 3;                 // = TMEM lane quadrant = SM sub-partition of both warps
  const int role = warp >> 2;                // 0: A, 1: B
  float* tile = reinterpret_cast<float*>(smem_raw) + pair * G::TILE_FLOATS;
  // raw 16-bit PCM of the next tile (fused ingest): lands here by cp.async, converted into `tile` at tile start
  const uint4* pstage = reinterpret_cast<const uint4*>(smem_raw + G::PAIRS * G::TILE_FLOATS * 4 + pair * G::PCM_STAGE_BYTES);

  // ---- the CTA takes the whole tensor memory of its SM: 512 columns x 128 lanes
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = __shfl_sync(0xffffffffu, s_tmem + ((uint32_t)(pair * 32) << 16), 0);   // the pair's lane quadrant

  const float log_floor = a.log_floor, log_scale = a.log_scale;
  const float silent_val = vlog2_clamp(0.0f, log_floor) * log_scale;
  const float q_scale = 0.25f * log_scale;          // (S + 4) / 4 with S = log2 * scale, one FFMA
  const int n_frames = NF ? NF : a.n_frames;
  const int T = a.tiles_per_clip;
  auto tile_s0 = [&](int t) { return (long long)t * G::F * G::HOP - G::N / 2; };

  // cp.async plan of four hop rows (160 16-byte chunks, five per lane): byte offsets in the tile
  unsigned dst_off[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const int i = lane + 32 * j;
    dst_off[j] = smem_u32(tile) + (unsigned)((i / 40) * G::PITCH * 4 + (i % 40) * 16);
  }

  // Two ways to hand out the work (a.group, chosen by the host from the batch size): a clip per warp pair (1: the
  // steady state of a large batch; pairs of an SM take clips 148 apart), or a clip per CTA (4: its four pairs take a
  // quarter of the tiles each and agree on the clip maximum through shared memory) -- the second wastes at most one
  // quarter-clip per SM where the first can waste a whole clip per pair, which is what mid-size batches need.  The
  // quarters are contiguous runs of tiles (measured 2.5 % faster than interleaving) unless per-clip lengths are given:
  // then every fourth tile, so that a short clip's few loud tiles spread over all four pairs (config 4: 1.02 -> 0.44 ms).
  // Small batches (fewer clips than half the SMs) go one step further: a clip per `slices` CTAs (a.group = 4 * slices,
  // grid = batch * slices, cooperative launch), every pair of them takes every (4 * slices)-th tile, and the CTAs
  // agree on the clip maximum through the scratch buffer (gmax / gcnt, one release-acquire round per launch).
  const bool coop = SPLIT && a.group >= G::PAIRS;
  const int slices = SPLIT && a.group > G::PAIRS ? a.group / G::PAIRS : 1;
  const int slice = slices > 1 ? (int)blockIdx.x % slices : 0;
  const int gp = slices > 1 ? (int)blockIdx.x / slices : coop ? (int)blockIdx.x : pair * (int)gridDim.x + (int)blockIdx.x;
  const int gn = slices > 1 ? a.batch : coop ? (int)gridDim.x : (int)gridDim.x * G::PAIRS;
  const bool inter = coop && (a.lengths != nullptr || slices > 1);
  const int tq = (T + G::PAIRS - 1) / G::PAIRS;
  const int t0 = !coop ? 0 : inter ? slice * G::PAIRS + pair : min(pair * tq, T);   // this pair's tiles: t0, t0 + tstep, ... < t1
  const int tstep = inter ? G::PAIRS * slices : 1;
  const int t1 = (!coop || inter) ? T : min((pair + 1) * tq, T);
  int par = 0;                                                       // clip parity: s_pmax is double-buffered when CTAs share clips
  for (int clip = gp; clip < a.batch; clip += gn, par ^= (SPLIT ? 1 : 0)) {
    const float* cptr = a.wave + (long long)clip * a.clip_stride;
    const short* pptr = a.pcm ? a.pcm + (long long)clip * a.clip_stride * a.pcm_channels : nullptr;
    int valid = a.n_samples;
    if (a.lengths) valid = min(max(a.lengths[clip], 0), a.n_samples);
    float* oc = a.out + (long long)clip * NM * n_frames;
    float rmax = -INFINITY;                   // of log2(mel), this warp's filters
    auto next_loud = [&](int t) {
      while (t < t1 && tile_is_silent<G>(tile_s0(t), a.n_samples, valid)) t += tstep;
      return t;
    };
    // Filling the pair's waveform tile (dead at that point) with tile t, half of it per warp.  Hop rows
    // inside the audio arrive as coalesced 16-byte cp.async chunks: four rows = 160 chunks = five per
    // lane per call of fetch_rows4 -- a burst of all twenty stalls the warp on the LSU queue (ncu:
    // lg_throttle 8 % of the kernel), so the steady state issues one group of five behind each
    // stage-2 row.  Tiles that touch a clip edge (reflection) or the zero padding are written by
    // the lanes, row by row, at once.
    const char* fsrc = nullptr;        // interior tile being fetched: this warp's first chunk; else nullptr
    bool pcm_staged = false;           // the tile being fetched is raw PCM in `pstage` and still has to be converted
    bool pcm_edge = false;             // ... and touches a clip edge or the padding: only the chunks inside are copied
    long long pcm_s0 = 0;              // ... its first sample
    const int spc = a.pcm_channels == 2 ? 4 : 8;               // PCM: samples per 16-byte chunk
    // PCM: the warps split the tile like the float32 copies do -- role A hop rows 0-15, role B rows 16-33
    const int pcm_c0 = role ? 16 * G::HOP / spc : 0;                                        // first chunk of this warp
    const int pcm_cn = role ? (G::SPAN - 16 * G::HOP) / spc : 16 * G::HOP / spc;            // and how many
    auto fetch_rows4 = [&](int blk) {  // rows 4 blk .. 4 blk + 3 of this warp's half (blk 4, role B: rows 32, 33)
      if (!fsrc) return;
      if (pcm_staged) {                // a quarter of this warp's raw chunks per call (blk 0..3)
        if (blk >= 4) return;
        const int per = spc == 8 ? 3 : 6;                       // 32-lane rounds per call: 12 (mono) / 24 (stereo) cover 350 / 700
        const unsigned dst = smem_u32(pstage) + 16 * (pcm_c0 + lane);
        if (!pcm_edge) {               // fsrc already points at this lane's first chunk
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const int k = blk * per + i;
            if (i < per && lane + 32 * k < pcm_cn)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512 * k), "l"(fsrc + 512 * k) : "memory");
          }
          return;
        }
#pragma unroll 1
        for (int k = blk * per; k < (blk + 1) * per; ++k) {     // a tile at a clip edge: only the chunks inside the audio
          const int c = lane + 32 * k;
          const long long n0 = pcm_s0 + (long long)(pcm_c0 + c) * spc;
          if (c < pcm_cn && n0 >= 0 && n0 + spc <= valid)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512 * k), "l"(fsrc + 512 * k) : "memory");
        }
        return;
      }
      const unsigned rb = (role * 4 + blk) * (4 * G::PITCH * 4);
      const char* src = fsrc + blk * (4 * G::HOP * 4);
      if (blk < 4) {
#pragma unroll
        for (int j = 0; j < 5; ++j)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_off[j] + rb), "l"(src + 512 * j) : "memory");
      } else if (role == 1) {          // the last one is half a row: 60 chunks
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_off[0] + rb), "l"(src) : "memory");
        if (lane < 28) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_off[1] + rb), "l"(src + 512) : "memory");
      }
    };
    // Raw PCM in the staging buffer -> float32 samples in the tile, this warp's rows: the sum of the channels times
    // 2^-15 / channels, exact in float32.  A group is four samples: 8 (mono) or 16 (stereo) staged bytes in, one
    // 16-byte store out, both conflict-free, at the byte offsets of the float32 copy plan (dst_off).  No I2F (the
    // quarter-rate conversion pipe): the integer is placed in the mantissa of 1.5 * 2^23, which gives the float
    // 12582912 + s, and one FFMA scales and removes the offset: (12582912 + s) * 2^-15 - 384 = s / 32768.  Stereo:
    // PRMT sign-extends the halves, one three-input add sums them onto the bit pattern.
    auto sext_lo = [](unsigned w) { int r; asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(r) : "r"(w)); return r; };
    auto sext_hi = [](unsigned w) { int r; asm("prmt.b32 %0, %1, 0, 0xBB32;" : "=r"(r) : "r"(w)); return r; };
    auto pcm_group = [&](auto stereo, unsigned src, float (&f)[4]) {       // staged bytes of one group -> four samples
      if constexpr (decltype(stereo)::value) {
        unsigned w[4];
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(src));
#pragma unroll
        for (int j = 0; j < 4; ++j)
          f[j] = fmaf(__int_as_float(sext_lo(w[j]) + sext_hi(w[j]) + 0x4B400000), 1.0f / 65536.0f, -192.0f);
      } else {
        unsigned w[2];
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(src));
        // mono, three ALU operations per two samples: flipping bit 15 of both halves turns s into the unsigned
        // s + 32768, which is then spliced under the exponent of 1.5 * 2^23 (LOP3 / PRMT) -- offset 385 instead of 384
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const unsigned u = w[j] ^ 0x80008000u;
          f[2 * j] = fmaf(__uint_as_float((u & 0xffffu) | 0x4B400000u), 1.0f / 32768.0f, -385.0f);
          f[2 * j + 1] = fmaf(__uint_as_float(__byte_perm(u, 0x4B400000u, 0x7632)), 1.0f / 32768.0f, -385.0f);
        }
      }
    };
    auto pcm_convert_tile = [&](auto stereo) {
      constexpr unsigned GB = decltype(stereo)::value ? 16 : 8;            // staged bytes per group
      const unsigned src0 = smem_u32(pstage) + (role * 16 * (G::HOP / 4) + lane) * GB;
      if (!pcm_edge) {
#pragma unroll 1
        for (int blk = 0; blk < 4; ++blk) {                                 // four hop rows = 160 groups, five per lane
          float f[5][4];
#pragma unroll
          for (int j = 0; j < 5; ++j) pcm_group(stereo, src0 + (blk * 160 + 32 * j) * GB, f[j]);
          const unsigned rb = (role * 4 + blk) * (4 * G::PITCH * 4);
#pragma unroll
          for (int j = 0; j < 5; ++j)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst_off[j] + rb), "f"(f[j][0]), "f"(f[j][1]),
                         "f"(f[j][2]), "f"(f[j][3]) : "memory");
        }
        if (role) {                                                         // rows 32 and 33: 60 groups
          const unsigned rb = 8 * (4 * G::PITCH * 4);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (j == 0 || lane < 28) {
              float f[4];
              pcm_group(stereo, src0 + (4 * 160 + 32 * j) * GB, f);
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst_off[j] + rb), "f"(f[0]), "f"(f[1]), "f"(f[2]),
                           "f"(f[3]) : "memory");
            }
          }
        }
        return;
      }
      // a tile that touches a clip edge or the padding: groups inside the audio were copied, the others are
      // reflected / zero samples fetched one by one
      const int g0 = role * 16 * (G::HOP / 4), g1 = role ? G::SPAN / 4 : 16 * (G::HOP / 4);
#pragma unroll 1
      for (int g = g0 + lane; g < g1; g += 32) {
        const int n0 = 4 * g;
        const long long c0 = pcm_s0 + (n0 / spc) * spc;                     // the 16-byte chunk the group lies in
        float f[4];
        if (c0 >= 0 && c0 + spc <= valid) {
          pcm_group(stereo, smem_u32(pstage) + g * GB, f);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) f[j] = load_sample_pcm(pptr, a.pcm_channels, (long)(pcm_s0 + n0 + j), a.n_samples, valid);
        }
        *reinterpret_cast<float4*>(tile + n0 + 4 * (n0 / G::HOP)) = make_float4(f[0], f[1], f[2], f[3]);
      }
    };
    auto pcm_convert = [&]() {
      if (spc == 4) pcm_convert_tile(std::true_type{});
      else pcm_convert_tile(std::false_type{});
    };
    auto fetch_begin = [&](int t) {
      const long long s0 = tile_s0(t);
      fsrc = nullptr;
      pcm_staged = false;
      const bool interior = s0 >= 0 && s0 + G::SPAN <= valid;
      if (interior && !pptr) {
        fsrc = reinterpret_cast<const char*>(cptr + s0) + 16 * lane + role * (4 * 4 * G::HOP * 4);
        return;
      }
      if (pptr && a.tma_ok) {          // 16-bit PCM (fused ingest): the raw frames go to the staging buffer
        fsrc = reinterpret_cast<const char*>(pptr + s0 * a.pcm_channels) + 16 * (pcm_c0 + lane);
        pcm_staged = true;
        pcm_edge = !interior;
        pcm_s0 = s0;
        return;
      }
#pragma unroll 1
      for (int r = role; r < G::ROWS; r += 2) {
        const long long sr = s0 + (long long)r * G::HOP;
        const int len = r == G::ROWS - 1 ? G::SPAN - (G::ROWS - 1) * G::HOP : G::HOP;
        float* drow = tile + r * G::PITCH;
        if (sr >= 0 && sr + len <= valid && !pptr) {
          for (int c = lane; c < len / 4; c += 32) cp_async16(drow + 4 * c, cptr + sr + 4 * c);
        } else {
#pragma unroll
          for (int k = 0; k < G::HOP / 32; ++k) {            // the row's loads go out back to back
            const int i = lane + 32 * k;
            if (i < len) drow[i] = load_sample_any(cptr, pptr, a.pcm_channels, (long)(sr + i), a.n_samples, valid);
          }
        }
      }
    };

    int t = next_loud(t0);
    if (t < t1) {
      fetch_begin(t);
#pragma unroll 1
      for (int blk = 0; blk < 5; ++blk) fetch_rows4(blk);
      cp_async_commit();
    }
    while (t < t1) {
      cp_async_wait_all();
      if (pcm_staged) {          // (16-bit PCM input only) the raw frames this warp copied -> its rows of the float tile,
        __syncwarp();            //  which has been dead since the stage-1 barrier of the previous tile
        pcm_convert();
      }
      pair_sync(pair);           // the tile is complete; the partner has finished reading P of the previous tile
      // Lanes past the clip's last frame (only in its last tile) redo the last valid frame: same
      // samples, same arithmetic, the same value stored to the same address -- no predicates anywhere.
      const int fl = min(lane, n_frames - 1 - t * G::F);
      // ================= stage 1: wave tile -> Y in tensor memory =================
      {
        const float* mine = tile + fl * G::PITCH;
        // role A: column groups 0, 1 and the first pair of group 2; role B: the second pair of group 2
        // and groups 3, 4.  One rolled loop = one copy of the two-codelet group body for both roles.
#pragma unroll 1
        for (int u = 0; u < 2; ++u) {
          const int g = role * 3 + u;
          f32x2 xa[20], xb[20];
          tf_load_group(mine, g, xa, xb);
          tf_stage1_pair(&tab.cp[2 * g], xa, tm + 4 * g, tm + 16 * g);
          tf_stage1_pair(&tab.cp[2 * g + 1], xb, tm + 4 * g + 2, tm + 16 * g + 8);
        }
        {
          const int cp = 4 + role;             // the column pairs of group 2, one per role
          f32x2 x[20];
          tf_load_pair(mine, cp, x);
          tf_stage1_pair(&tab.cp[cp], x, tm + 2 * cp, tm + 8 * cp);
        }
      }
      tm_wait_st();
      pair_sync(pair);           // Y is complete, the waveform tile is dead
      const int tn = next_loud(t + tstep);
      fsrc = nullptr;
      if (tn < t1) fetch_begin(tn);  // the copies themselves go out between the stage-2 rows

      // ================= stage 2: rows of Y -> |X|^2, in place =================
      // (role A: row 0 and row pairs 0, 1; role B: row pairs 2, 3, 4; the cp.async groups of the next
      //  tile go out in between)
      if (role == 0) {
        {
          float yr[20], p[12], pr[11];
          tm_ld16(tm, yr);
          tm_ld4(tm + 16, yr + 16);
          tm_wait_ld();
          fetch_rows4(0);
          stage2_r20_half<float>(yr, pr);
#pragma unroll
          for (int j = 0; j < 11; ++j) p[j] = pr[j];
          p[11] = 0.0f;
          tm_st8(tm, p);
          tm_st4(tm + 8, p[8], p[9], p[10], p[11]);
        }
      }
      {
        const int q0 = role ? 2 : 0, q1 = role ? 5 : 2, f0 = role ? 0 : 1;
#pragma unroll 1
        for (int q = q0; q < q1; ++q) {        // one copy of the pair codelet for both roles
          tf_stage2_pair(tm + (uint32_t)tf_pair_base(q), [&]() { fetch_rows4(f0 + q - q0); });
        }
        fetch_rows4(3);
        if (role) fetch_rows4(4);
      }
      cp_async_commit();
      tm_wait_st();
      pair_sync(pair);           // P is complete

      // ================= mel projection, log, store =================
      {
        float* op = oc + (t * G::F + fl);
        float hi, lo;
        if (role == 0) tf_mel_store<NM, 0>(tm, op, (long long)n_frames, q_scale, hi, lo);
        else tf_mel_store<NM, 1>(tm, op, (long long)n_frames, q_scale, hi, lo);
        rmax = fmaxf(rmax, hi);
        float wlo;                 // minimum over the warp in one instruction (NaN if any lane has one)
        asm volatile("redux.sync.min.NaN.f32 %0, %1, 0xffffffff;" : "=f"(wlo) : "f"(lo));
        if (lane == 0) s_tmin[pair][t][role] = wlo * log_scale;
      }
      t = tn;
    }

    // ================= the clip is complete: max - 8 clamp where it bites =================
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    if (lane == 0) s_pmax[par][pair][role] = rmax * log_scale;
    float cmax = silent_val;                                          // the clamp at `floor`, applied to the maximum
    if (coop) {                  // all eight warps walk the same clips: a CTA barrier is safe here
      __syncthreads();
#pragma unroll
      for (int q = 0; q < G::PAIRS; ++q) cmax = fmaxf(cmax, fmaxf(s_pmax[par][q][0], s_pmax[par][q][1]));
      if (slices > 1) {          // ... and the CTAs of the clip meet in global memory (all resident: cooperative launch)
        if (threadIdx.x == 0) {
          __stcg(a.gmax + (long long)clip * slices + slice, cmax);
          __threadfence();
          atomicAdd(a.gcnt + clip, 1);
          long long spins = 0;
          while (ld_acquire(a.gcnt + clip) < slices) {
            __nanosleep(32);
            if (++spins > (1ll << 24)) __trap();      // > 10 s: the launch was not cooperative / the counter not zeroed
          }
        }
        __syncthreads();
        for (int i = 0; i < slices; ++i) cmax = fmaxf(cmax, __ldcg(a.gmax + (long long)clip * slices + i));
      }
    } else {
      pair_sync(pair);
      cmax = fmaxf(cmax, fmaxf(s_pmax[par][pair][0], s_pmax[par][pair][1]));
    }
    if (role == 0 && lane == 0 && a.clip_max && (!coop || (pair == 0 && slice == 0))) a.clip_max[clip] = cmax;
    const float thr = fmaxf(cmax - 8.0f, silent_val);
    const float cval = vaffine(thr, 0.25f, 1.0f);
    const int mA = role == 0 ? 0 : MP::M0, mB = role == 0 ? MP::M0 : NM;            // this warp's filters
    for (int tt = t0; tt < t1; tt += tstep) {
      const int fa = tt * G::F;
      const int len = min(fa + G::F, n_frames) - fa;
      if (tile_is_silent<G>(tile_s0(tt), a.n_samples, valid)) {
        for (int m = mA; m < mB; ++m) {
          float* row = oc + (long long)m * n_frames + fa;
          if (lane < len) __stcs(row + lane, cval);
        }
        continue;
      }
      // the stored values are fma(log2, scale / 4, 1), the minimum was taken on log2 * scale: leave
      // a margin of a few ulps so that rounding can never skip a tile that needs the clamp
      if (s_tmin[pair][tt][role] > thr + 1e-5f) continue;   // (a NaN minimum takes the fix-up path)
      if (a.vec_ok && (len & 3) == 0) {
        const int q = len >> 2, j = lane & 7;         // <= 8 float4 per row: four rows per warp pass
#ifndef LM_TF_FIX_UNR
#define LM_TF_FIX_UNR 16
#endif
        constexpr int UNR = LM_TF_FIX_UNR;
        for (int m0 = mA + (lane >> 3); m0 < mB; m0 += 4 * UNR) {
          float4 v[UNR];
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int m = m0 + 4 * u;
            if (m < mB && j < q) v[u] = __ldcg(reinterpret_cast<const float4*>(oc + (long long)m * n_frames + fa) + j);
          }
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int m = m0 + 4 * u;
            if (m < mB && j < q) {
              float4 w = v[u];
              w.x = fmaxf(w.x, cval); w.y = fmaxf(w.y, cval); w.z = fmaxf(w.z, cval); w.w = fmaxf(w.w, cval);
              __stcs(reinterpret_cast<float4*>(oc + (long long)m * n_frames + fa) + j, w);
            }
          }
        }
      } else {
        for (int m = mA; m < mB; ++m) {
          float* row = oc + (long long)m * n_frames + fa;
          if (lane < len) row[lane] = fmaxf(__ldcg(row + lane), cval);
        }
      }
    }
    pair_sync(pair);             // s_pmax / s_tmin may be rewritten by the next clip
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512) : "memory");
}

}  // namespace lm
