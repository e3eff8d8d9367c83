// Micro-benchmark (sm_100a): do warp shuffles compete with shared-memory loads for the same
// 128 B/clk/SM data path?  Decides whether table loads that both half-warps duplicate (window,
// twiddles) are cheaper as "each half loads every other entry + SHFL.xor 16".
//   nvcc -arch=sm_100a -O3 -o mb_lds_shfl tools/microbench_lds_shfl.cu && ./mb_lds_shfl
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: 16 x LDS.64 (conflict-free, 2 wavefronts each)      per iteration
// MODE 1: 16 x SHFL.IDX                                        per iteration
// MODE 2: 16 x LDS.64 + 16 x SHFL interleaved                  per iteration
// MODE 3: 16 x LDS.64 where both half-warps read the SAME 128 bytes (duplicate table read)
// MODE 4: 16 x LDS.32 where both half-warps read the same 64 bytes
// MODE 5: 32 x LDS.64 (reference for MODE 2: same instruction count, all loads)
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, long long* cycles) {
  __shared__ float2 sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float2(i, -i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float acc0 = 0.f, acc1 = 0.f;
  float sh = lane;
  int base = (threadIdx.x >> 5) * 64;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (MODE == 0 || MODE == 2 || MODE == 5) {
        const float2 v = sm[(base + 32 * j + lane) & 2047];
        acc0 += v.x; acc1 += v.y;
      }
      if (MODE == 5) {
        const float2 v = sm[(base + 32 * j + 512 + lane) & 2047];
        acc0 += v.x; acc1 += v.y;
      }
      if (MODE == 3) {
        const float2 v = sm[(base + 32 * j + (lane & 15)) & 2047];
        acc0 += v.x; acc1 += v.y;
      }
      if (MODE == 4) {
        const float v = reinterpret_cast<const float*>(sm)[(base + 32 * j + (lane & 15)) & 4095];
        acc0 += v;
      }
      if (MODE == 1 || MODE == 2) {
        // independent shuffles (throughput, not latency): the source does not depend on earlier results
        const float g = __shfl_xor_sync(0xffffffffu, sh + (float)j, 16);
        acc1 += g;
      }
      if (MODE == 6 && (j & 1) == 0) {
        // each half-warp loads every other table entry, then the halves swap both components
        const float2 v = sm[(base + 32 * j + (lane & 15) + (lane & 16)) & 2047];
        const float gx = __shfl_xor_sync(0xffffffffu, v.x, 16);
        const float gy = __shfl_xor_sync(0xffffffffu, v.y, 16);
        acc0 += v.x + gx; acc1 += v.y + gy;
      }
    }
    sh += acc1;
    base = (base + 7) & 1023;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + sh;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, long long* cyc) {
  const int iters = 2000;
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  k<MODE><<<2 * n_sm, 256>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  k<MODE><<<2 * n_sm, 256>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  // 16 warps per SM, each runs `iters` iterations
  printf("%-52s %8.2f SM-cycles per (16-op group of one warp)\n", name, (double)h / iters / 16.0);
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 4 * 1024 * 1024);
  cudaMalloc(&cyc, 8);
  run<0>("16 LDS.64 distinct (2 wavefronts each)", out, cyc);
  run<1>("16 SHFL", out, cyc);
  run<2>("16 LDS.64 + 16 SHFL", out, cyc);
  run<5>("32 LDS.64", out, cyc);
  run<3>("16 LDS.64, halves read the same 128 B", out, cyc);
  run<4>("16 LDS.32, halves read the same 64 B", out, cyc);
  run<6>("8 LDS.64 (halves: different entries) + 16 SHFL.xor16", out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
