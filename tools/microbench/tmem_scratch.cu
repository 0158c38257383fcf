// Microbenchmark: tensor memory (TMEM) as thread-private scratch on sm_100a -- no MMA involved.
//
// Question behind logmel_tf_kernel: can a warp park the 420 inter-stage values of "its" 32 frames
// in its TMEM lane quadrant (lane = frame, column = value) instead of shared memory?  Measures
//   * tcgen05.st / tcgen05.ld round trip correctness (32x32b.x32 and .x16 shapes, all 512 columns)
//   * store / load throughput per SM with 4 warps (one per quadrant) and 8 warps (two per quadrant)
//   * the same with a packed-FP32 stream in between (does LDTM/STTM steal FMA issue slots?)
//   * LDS.128 of a thread-per-frame waveform tile (pitch 164 floats): bank conflicts or not
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmem_scratch tmem_scratch.cu && ./tmem_scratch
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st32(uint32_t addr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(addr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t tmem_alloc_all(uint32_t* slot) {   // whole CTA; returns the base address
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  return *slot;
}
__device__ __forceinline__ void tmem_free_all(uint32_t base) {
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512));
}

// MODE 0: correctness -- write every column of the warp's range with a (lane, column, round) pattern, read back
// MODE 1: store throughput  2: load throughput  3: store+load with FFMA2 work in between  4: FFMA2 work alone
template <int MODE>
__global__ void k_tmem(int* errors, long long* cyc, float* sink, int rounds) {
  __shared__ uint32_t slot;
  const uint32_t base = tmem_alloc_all(&slot);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwq = blockDim.x / 128;                           // warps per quadrant
  const int ncol = 512 / nwq;                                 // columns of this warp
  const uint32_t wbase = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * ncol);
  uint32_t r[32];
  unsigned long long acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = (unsigned long long)(threadIdx.x + i) << 20;
  unsigned long long kk;
  asm("mov.b64 %0, {%1, %2};" : "=l"(kk) : "f"(1.0001f), "f"(0.9999f));
  int bad = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < rounds; ++it) {
    if (MODE == 0 || MODE == 1 || MODE == 3) {
      for (int c = 0; c < ncol; c += 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = (uint32_t)(it * 0x9e3779b9u) ^ (uint32_t)((threadIdx.x << 16) | (c + j));
        tmem_st32(wbase + c, r);
      }
      tmem_wait_st();
    }
    if (MODE == 3 || MODE == 4) {
#pragma unroll 1
      for (int q = 0; q < 64; ++q) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(acc[i]) : "l"(kk));
      }
    }
    if (MODE == 0 || MODE == 2 || MODE == 3) {
      for (int c = 0; c < ncol; c += 32) {
        tmem_ld32(wbase + c, r);
        tmem_wait_ld();
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            bad += r[j] != ((uint32_t)(it * 0x9e3779b9u) ^ (uint32_t)((threadIdx.x << 16) | (c + j)));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j & 7] ^= r[j];
        }
      }
    }
  }
  const long long t1 = clock64();
  if (MODE == 0 && bad) atomicAdd(errors, bad);
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += (float)(acc[i] & 0xffff);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  tmem_free_all(base);
}

// thread-per-frame waveform fetch: lane f reads 16 bytes at word PITCH * f + 4 * j
template <int PITCH>
__global__ void k_lds(long long* cyc, float* sink, int rounds) {
  extern __shared__ __align__(16) float tile[];
  for (int i = threadIdx.x; i < 36 * PITCH * (int)(blockDim.x / 32); i += blockDim.x) tile[i] = (float)i;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* mine = tile + warp * 36 * PITCH + lane * PITCH;
  float4 acc = {0, 0, 0, 0};
  const long long t0 = clock64();
  for (int it = 0; it < rounds; ++it) {
#pragma unroll 20
    for (int j = 0; j < 100; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(mine + 4 * j);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  const long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// warp-uniform (broadcast) shared-memory loads: how many LSU cycles does an LDS.128 / LDS.64 / LDS.32 cost when
// all 32 lanes read the SAME address?  (constants of the stage-1 codelets as shared-memory operands)
template <int WIDTH>
__global__ void k_lds_bcast(long long* cyc, float* sink, int rounds) {
  __shared__ __align__(16) float tab[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) tab[i] = (float)i;
  __syncthreads();
  float acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < rounds; ++it) {
#pragma unroll 32
    for (int j = 0; j < 128; ++j) {
      const float* p = tab + ((j * 4 + it) & 2044);
      const unsigned sp = (unsigned)__cvta_generic_to_shared(p);
      if (WIDTH == 4) {
        float x, y, z, w;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(sp));
        acc += x + w;
      } else if (WIDTH == 2) {
        float x, y;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(sp));
        acc += x + y;
      } else {
        float x;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(sp));
        acc += x;
      }
    }
  }
  const long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  int* err;
  long long* cyc;
  float* sink;
  cudaMalloc(&err, 4);
  cudaMalloc(&cyc, 8);
  cudaMalloc(&sink, 148 * 256 * 4);
  cudaMemset(err, 0, 4);
  const int rounds = 200;
  for (int threads : {128, 256}) {
    long long h[5];
    int herr = 0;
    k_tmem<0><<<148, threads>>>(err, cyc, sink, 8);
    cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
    printf("threads/CTA=%d  round-trip mismatches: %d  (%s)\n", threads, herr, cudaGetErrorString(cudaGetLastError()));
    k_tmem<1><<<148, threads>>>(err, cyc, sink, rounds); cudaMemcpy(&h[1], cyc, 8, cudaMemcpyDeviceToHost);
    k_tmem<2><<<148, threads>>>(err, cyc, sink, rounds); cudaMemcpy(&h[2], cyc, 8, cudaMemcpyDeviceToHost);
    k_tmem<3><<<148, threads>>>(err, cyc, sink, rounds); cudaMemcpy(&h[3], cyc, 8, cudaMemcpyDeviceToHost);
    k_tmem<4><<<148, threads>>>(err, cyc, sink, rounds); cudaMemcpy(&h[4], cyc, 8, cudaMemcpyDeviceToHost);
    const double bytes = 128.0 * 512 * 4;          // the whole TMEM of an SM per round
    printf("  store: %.1f cycles/round -> %.1f B/cycle/SM\n", (double)h[1] / rounds, bytes * rounds / h[1]);
    printf("  load : %.1f cycles/round -> %.1f B/cycle/SM\n", (double)h[2] / rounds, bytes * rounds / h[2]);
    printf("  FFMA2 alone (512 per warp per round): %.1f cycles/round; with store+load of all columns: %.1f "
           "(sum would be %.1f)\n",
           (double)h[4] / rounds, (double)h[3] / rounds, (double)(h[1] + h[2] + h[4]) / rounds);
    printf("  %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  }
  {
    long long h;
    cudaFuncSetAttribute(k_lds<164>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 36 * 164 * 4);
    k_lds<164><<<148, 128, 4 * 36 * 164 * 4>>>(cyc, sink, rounds);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS.128 pitch 164, 4 warps: %.2f cycles per LDS.128 per warp (4 wavefronts each => 4.0 is conflict free at one warp per SMSP... "
           "SM-wide 4 warps share 1 LSU: 16.0 = bandwidth bound)\n", (double)h / (rounds * 100));
    cudaFuncSetAttribute(k_lds<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 36 * 164 * 4);
    k_lds<160><<<148, 128, 4 * 36 * 164 * 4>>>(cyc, sink, rounds);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS.128 pitch 160, 4 warps: %.2f cycles per LDS.128 per warp\n", (double)h / (rounds * 100));
    printf("  %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  }
  for (int threads : {128, 256, 512}) {
    long long h4, h2, h1;
    k_lds_bcast<4><<<148, threads>>>(cyc, sink, rounds); cudaMemcpy(&h4, cyc, 8, cudaMemcpyDeviceToHost);
    k_lds_bcast<2><<<148, threads>>>(cyc, sink, rounds); cudaMemcpy(&h2, cyc, 8, cudaMemcpyDeviceToHost);
    k_lds_bcast<1><<<148, threads>>>(cyc, sink, rounds); cudaMemcpy(&h1, cyc, 8, cudaMemcpyDeviceToHost);
    const double n = (double)rounds * 128 * (threads / 32);
    printf("broadcast LDS, %2d warps/SM: SM cycles per warp-instruction  .128: %.2f   .64: %.2f   .32: %.2f\n", threads / 32,
           h4 / n, h2 / n, h1 / n);
  }
  printf("  %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
