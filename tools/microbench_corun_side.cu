// Tuning aid: a resident "side" kernel with a configurable footprint and behaviour, launched next to the
// persistent fbank kernel from Python (tools/corun_sleep.py) to find out what slows the fbank kernel when
// another kernel shares its SMs.   nvcc -arch=sm_100a -O3 --shared -Xcompiler -fPIC -o build/libcorun_side.so
//   mode 0: sleep (nanosleep loop, no issue pressure)   mode 1: spin on the clock (ALU issue pressure)
//   mode 2: stream global memory (LDG/STG float4)        mode 3: shared-memory loads
//   mode 4: 8 loads in flight per thread + stores          mode 5: the same, loads only
#include <cuda_runtime.h>
extern "C" __global__ void side_kernel(int mode, long long ns, float4* buf, long long n4, long long* counter) {
  extern __shared__ float4 sm[];
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  float4 acc = make_float4(0, 0, 0, 0);
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  do {
    if (mode == 0) {
      __nanosleep(1000);
    } else if (mode == 1) {
      for (int k = 0; k < 64; ++k) acc.x = acc.x * 1.0001f + 1.f;
    } else if (mode == 2) {
      float4 v = buf[i & (n4 - 1)];
      v.x += 1.f;
      buf[i & (n4 - 1)] = v;
      i += (long long)gridDim.x * blockDim.x;
    } else if (mode == 4 || mode == 5) {  // 8 independent 16-byte loads in flight per thread, then (mode 4) 8 stores
      float4 v[8];
      const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = buf[(i + k * stride) & (n4 - 1)];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k].x += 1.f;
        if (mode == 4) buf[(i + k * stride) & (n4 - 1)] = v[k];
        else acc.x += v[k].x;
      }
      i += 8 * stride;
    } else {
      for (int k = 0; k < 16; ++k) {
        float4 v = sm[(threadIdx.x + k * 32) & 255];
        acc.x += v.x;
      }
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while ((long long)(t - t0) < ns);
  if (acc.x == 12345.f) buf[0] = acc;
  // bytes this thread moved (modes 2, 4, 5): i advanced by gridDim.x * blockDim.x per 16-byte element
  if (counter != nullptr && threadIdx.x == 0 && blockIdx.x == 0)
    *counter = (i - (long long)blockIdx.x * blockDim.x - threadIdx.x) * 16;
}
extern "C" int launch_side(int grid, int threads, int smem, int mode, long long ns, void* buf, long long n4, void* stream) {
  cudaFuncSetAttribute(side_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  side_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(mode, ns, (float4*)buf, n4, (long long*)buf + n4 * 2);
  return (int)cudaGetLastError();
}
