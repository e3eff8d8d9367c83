// Micro-benchmark: scalar FP32 (FFMA/FADD) vs packed f32x2 (fma.rn.f32x2 / add.f32x2) issue throughput
// on sm_100a, alone and interleaved with shared-memory loads.  Decides whether the FFT butterflies
// should be written with packed ops ("FP32 pipe").   nvcc -arch=sm_100a -O3 -o mb mb.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define NACC 8

__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float s, int iters) {
  __shared__ float sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = s * i;
  __syncthreads();
  const float a = s * 1.0001f, b = s * 0.5f;
  if (MODE == 0) {  // scalar FFMA
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  } else if (MODE == 1) {  // scalar FADD
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = acc[i] + a;
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  } else if (MODE == 2 || MODE == 3 || MODE == 4) {  // packed fma / add / mul
    unsigned long long acc[NACC];
    const unsigned long long pa = pk(a, a), pb = pk(b, b);
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = pk(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        if (MODE == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(pa), "l"(pb));
        if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc[i]) : "l"(pa));
        if (MODE == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(acc[i]) : "l"(pa));
      }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      float x, y;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i]));
      r += x + y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  } else if (MODE == 5) {  // scalar FFMA + 1 LDS per 4 FFMA
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
      const float l0 = sm[(idx + it) & 1023], l1 = sm[(idx + 2 * it) & 1023];
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, (i == 0) ? l0 : ((i == 4) ? l1 : b));
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  } else if (MODE == 6) {  // packed FFMA2 x4 (= 8 flops-lanes) + 2 LDS : same math as MODE 5
    unsigned long long acc[NACC / 2];
    const unsigned long long pa = pk(a, a);
#pragma unroll
    for (int i = 0; i < NACC / 2; ++i) acc[i] = pk(threadIdx.x + i, threadIdx.x - i);
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
      const float l0 = sm[(idx + it) & 1023], l1 = sm[(idx + 2 * it) & 1023];
      const unsigned long long pl = pk(l0, l1), pb = pk(b, b);
#pragma unroll
      for (int i = 0; i < NACC / 2; ++i)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(pa), "l"(i == 0 ? pl : pb));
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < NACC / 2; ++i) {
      float x, y;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i]));
      r += x + y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  }
}

template <int MODE>
void run(const char* name, double flop_lanes_per_iter, float* d) {
  int dev_sms = 148;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int blocks = dev_sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(d, 1.0f, 64);
  cudaDeviceSynchronize();
  float best = 1e9;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, 1.0f, ITERS);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double lane_ops = (double)blocks * threads * ITERS * flop_lanes_per_iter;
  const double per_sm_per_s = lane_ops / (best * 1e-3) / dev_sms;
  printf("%-34s %8.3f ms  %7.1f lane-ops/clk/SM @%d MHz (nominal max clock)  %7.2f Glane-ops/s/SM\n", name, best,
         per_sm_per_s / (clk_khz * 1e3), clk_khz / 1000, per_sm_per_s / 1e9);
}

int main() {
  float* d;
  cudaMalloc(&d, 148 * 8 * 256 * sizeof(float) * 2);
  run<0>("scalar FFMA", NACC, d);
  run<1>("scalar FADD", NACC, d);
  run<2>("packed fma.rn.f32x2", 2 * NACC, d);
  run<3>("packed add.rn.f32x2", 2 * NACC, d);
  run<4>("packed mul.rn.f32x2", 2 * NACC, d);
  run<5>("scalar 8 FFMA + 2 LDS", NACC, d);
  run<6>("packed 4 FFMA2 + 2 LDS (same math)", NACC, d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
