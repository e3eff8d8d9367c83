// Micro-benchmark (sm_100a): tensor memory (TMEM) as a per-lane table store.
//  1. correctness of tcgen05.st / tcgen05.ld .32x32b round trips from all 8 warps of a CTA (warps w and
//     w+4 share TMEM lanes 32*(w%4)..+31), with 2 CTAs resident per SM (each allocates 128 columns);
//  2. throughput of tcgen05.ld.32x32b.x8 alone, and next to conflict-free LDS.64 traffic — does the
//     TMEM read path run in parallel with the shared-memory crossbar?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mb_tmem tools/microbench_tmem.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_ld8(unsigned taddr, unsigned (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const unsigned (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

constexpr int kCols = 128;

// MODE 0: 16 x tcgen05.ld.x8 per iteration; MODE 1: 16 x LDS.64; MODE 2: both
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float* out, int iters, long long* cycles, int* errors) {
  __shared__ float2 sm[2048];
  __shared__ unsigned s_taddr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float2(i, -i);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_taddr)),
                 "n"(kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const unsigned tbase = s_taddr;
  const unsigned tq = tbase + ((unsigned)(32 * (warp & 3)) << 16);  // this warp's TMEM lane quadrant
  // warps 0..3 fill their quadrant: lane l, column c holds 1000 * (quadrant lane) + c
  if (warp < 4) {
    for (int c0 = 0; c0 < kCols; c0 += 8) {
      unsigned v[8];
      for (int j = 0; j < 8; ++j) v[j] = 1000u * (unsigned)(32 * warp + lane) + (unsigned)(c0 + j);
      tmem_st8(tq + c0, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  // every warp (including 4..7, which did not write) checks its quadrant
  int bad = 0;
  for (int c0 = 0; c0 < kCols; c0 += 8) {
    unsigned v[8];
    tmem_ld8(tq + c0, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 8; ++j) bad += v[j] != 1000u * (unsigned)(32 * (warp & 3) + lane) + (unsigned)(c0 + j);
  }
  if (bad) atomicAdd(errors, bad);

  float acc0 = 0.f, acc1 = 0.f;
  unsigned accu = 0;
  int base = warp * 64;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (MODE == 0 || MODE == 2) {
        unsigned v[8];
        tmem_ld8(tq + 8 * (j & 15), v);
        if ((j & 3) == 3) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        accu += v[0] ^ v[7];
      }
      if (MODE == 1 || MODE == 2) {
        const float2 v = sm[(base + 32 * j + lane) & 2047];
        acc0 += v.x; acc1 += v.y;
      }
    }
    base = (base + 7) & 1023;
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + (float)accu;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(kCols));
}

template <int MODE>
void run(const char* name, float* out, long long* cyc, int* err) {
  const int iters = 2000;
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  cudaMemset(err, 0, 4);
  for (int rep = 0; rep < 2; ++rep) {
    k<MODE><<<2 * n_sm, 256>>>(out, iters, cyc, err);
    cudaDeviceSynchronize();
  }
  long long h = 0;
  int e = 0;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  cudaMemcpy(&e, err, sizeof(e), cudaMemcpyDeviceToHost);
  printf("%-44s %8.2f SM-cycles per 16-op group of one warp (16 warps/SM)   round-trip errors %d   %s\n", name,
         (double)h / iters / 16.0, e, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* out;
  long long* cyc;
  int* err;
  cudaMalloc(&out, 4 * 1024 * 1024);
  cudaMalloc(&cyc, 8);
  cudaMalloc(&err, 4);
  run<0>("16 x tcgen05.ld.32x32b.x8 (1 KB per op)", out, cyc, err);
  run<1>("16 x LDS.64 (256 B per op)", out, cyc, err);
  run<2>("16 x tcgen05.ld.x8 + 16 x LDS.64", out, cyc, err);
  return 0;
}
