// Decision gate for a tensor-core DFT — the THROUGHPUT half, on the GPU:
// how fast do the 5th-gen tensor cores run the GEMM shapes a radix-16 x radix-16 DFT of the fbank path
// needs?  (The numerics half is tools/sim_tc_dft.py.)
//
// The DFT-as-GEMM formulation (frames are the M dimension, 128 per tile; per stage 16 small GEMMs
// [128 x 32] x [32 x 32], window / DC-mean / twiddles folded into the constant matrices) issues
// tcgen05.mma with M = 128, N = 32, K = 32 bytes per instruction (8 tf32 or 16 bf16 elements).  This
// program measures, for kind::tf32 and kind::f16 (bf16), N = 32 and N = 64, A from shared memory (SS) and
// from tensor memory (TS):
//   * cycles per tcgen05.mma, one CTA per SM (the instruction stream of one elected thread),
//   * and checks the result of one accumulation chain against integer arithmetic (descriptors are right).
// From that: MMA-bound frames/s/SM = 128 frames / (32 GEMMs x K-steps x products x cycles per MMA).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mb_umma_dft tools/microbench_umma_dft.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// K-major, no swizzle: element (row, k) at (row / 8) * SBO + (k_chunk) * LBO + (row % 8) * 16 + byte-in-chunk
constexpr int kLBO = 128;         // between the 16-byte K chunks of one 8-row group
constexpr int kKBytes = 128;      // bytes of K per row of a stage operand (32 tf32 or 64 bf16)
constexpr int kSBO = (kKBytes / 16) * kLBO;  // between 8-row groups

__host__ __device__ constexpr uint64_t smem_desc(unsigned addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(kLBO >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) |
         (1ull << 46);  // version = 1 (Blackwell), base_offset 0, layout_type 0 (no swizzle)
}
// instruction descriptor: c_format F32 (1) [4,6), a/b format [7,10)/[10,13), K-major both, N >> 3 [17,23), M >> 4 [24,29)
__host__ __device__ constexpr uint32_t instr_desc(int fmt, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

template <bool kTF32>
__device__ __forceinline__ void mma_ss(unsigned d, uint64_t a, uint64_t b, uint32_t idesc, unsigned acc) {
  if (kTF32)
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
template <bool kTF32>
__device__ __forceinline__ void mma_ts(unsigned d, unsigned a_tmem, uint64_t b, uint32_t idesc, unsigned acc) {
  if (kTF32)
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(d),
                 "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(d),
                 "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred P1;\nW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ int a_val(int r, int k) { return ((r + 3 * k) % 7) - 3; }
__device__ __forceinline__ int b_val(int n, int k) { return ((5 * n + k) % 5) - 2; }

constexpr int kCols = 512;  // D: up to 4 independent accumulators of N columns at 0, A operand (TS): 32 columns at 256
constexpr int kAccs = 4;    // independent accumulator tiles the throughput loop rotates over (a DFT stage has 16)

template <bool kTF32, int N, bool kTS>
__global__ void __launch_bounds__(128, 1) probe(int n_mma, int n_issuers, long long* cycles, int* errors) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                    // 128 rows x 128 bytes
  unsigned char* sB = smem + 128 * kKBytes;    // N rows x 128 bytes
  __shared__ unsigned s_taddr;
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ __align__(8) unsigned long long s_bar2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kEl = kTF32 ? 4 : 2;                  // bytes per element
  constexpr int kK = kKBytes / kEl;                   // K elements of the stage (32 tf32 / 64 bf16)
  auto put = [&](unsigned char* base, int row, int k, int v) {
    const int off = (row / 8) * kSBO + ((k * kEl) / 16) * kLBO + (row % 8) * 16 + (k * kEl) % 16;
    if (kTF32) *reinterpret_cast<float*>(base + off) = (float)v;
    else *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16((float)v);
  };
  for (int i = tid; i < 128 * kK; i += 128) put(sA, i / kK, i % kK, a_val(i / kK, i % kK));
  for (int i = tid; i < N * kK; i += 128) put(sB, i / kK, i % kK, b_val(i / kK, i % kK));
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar2)), "r"(n_issuers));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_taddr)), "n"(kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of A / B -> async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const unsigned tbase = s_taddr;
  const unsigned tD = tbase, tA = tbase + 256;
  if (kTS) {
    // A into tensor memory: lane = row, one 32-bit column per tf32 element / per bf16 pair
    const unsigned tq = tA + ((unsigned)(32 * warp) << 16);
    const int row = 32 * warp + lane;
    for (int c0 = 0; c0 < kKBytes / 4; c0 += 8) {
      unsigned v[8];
      for (int j = 0; j < 8; ++j) {
        if (kTF32) v[j] = __float_as_uint((float)a_val(row, c0 + j));
        else {
          const __nv_bfloat162 h = __floats2bfloat162_rn((float)a_val(row, 2 * (c0 + j)), (float)a_val(row, 2 * (c0 + j) + 1));
          v[j] = *reinterpret_cast<const unsigned*>(&h);
        }
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tq + c0), "r"(v[0]),
                   "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  const uint32_t idesc = instr_desc(kTF32 ? 2 : 1, N);
  const uint64_t dA = smem_desc(smem_u32(sA)), dB = smem_desc(smem_u32(sB));
  constexpr int kSteps = kKBytes / 32;  // MMAs per accumulation chain (32 bytes of K each)
  // ---- one chain, checked ----------------------------------------------------------------------
  if (tid == 0) {
    for (int s = 0; s < kSteps; ++s) {
      if (kTS) mma_ts<kTF32>(tD, tA + 8 * s, dB + (uint64_t)((2 * kLBO * s) >> 4), idesc, s > 0);
      else mma_ss<kTF32>(tD, dA + (uint64_t)((2 * kLBO * s) >> 4), dB + (uint64_t)((2 * kLBO * s) >> 4), idesc, s > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
  }
  mbar_wait(&s_bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  {
    const int row = 32 * warp + lane;
    int bad = 0;
    for (int c0 = 0; c0 < N; c0 += 8) {
      unsigned v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(tD + ((unsigned)(32 * warp) << 16) + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) {
        int want = 0;
        for (int k = 0; k < kK; ++k) want += a_val(row, k) * b_val(c0 + j, k);
        bad += __uint_as_float(v[j]) != (float)want;
      }
    }
    if (bad && blockIdx.x == 0) atomicAdd(errors, bad);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  // ---- throughput: n_mma instructions back to back (chains of kSteps) from n_issuers threads (lane 0 of
  // warps 0..n_issuers-1, each rotating over its own accumulator tiles), one commit per issuer ------------
  long long t0 = clock64();
  if (lane == 0 && warp < n_issuers) {
    const int accs = kAccs / n_issuers > 0 ? kAccs / n_issuers : 1;
    for (int i = 0, acc = 0; i < n_mma / n_issuers; i += kSteps, acc = (acc + 1) % accs) {
      const unsigned d = tD + (unsigned)(((warp * accs + acc) % kAccs) * N);
#pragma unroll
      for (int s = 0; s < kSteps; ++s) {
        if (kTS) mma_ts<kTF32>(d, tA + 8 * s, dB + (uint64_t)((2 * kLBO * s) >> 4), idesc, s > 0);
        else mma_ss<kTF32>(d, dA + (uint64_t)((2 * kLBO * s) >> 4), dB + (uint64_t)((2 * kLBO * s) >> 4), idesc, s > 0);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar2)) : "memory");
  }
  mbar_wait(&s_bar2, 0);
  if (tid == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(kCols));
}

template <bool kTF32, int N, bool kTS>
double run(const char* name, long long* cyc, int* err, int n_sm, int n_issuers = 1) {
  const int n_mma = 4096;
  const int smem = 128 * kKBytes + N * kKBytes;
  cudaFuncSetAttribute(probe<kTF32, N, kTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(err, 0, 4);
  for (int rep = 0; rep < 2; ++rep) {
    probe<kTF32, N, kTS><<<n_sm, 128, smem>>>(n_mma, n_issuers, cyc, err);
    if (cudaDeviceSynchronize() != cudaSuccess) break;
  }
  long long h = 0;
  int e = 0;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  cudaMemcpy(&e, err, sizeof(e), cudaMemcpyDeviceToHost);
  const double per = (double)h / n_mma;
  const double flop = 2.0 * 128 * N * (kTF32 ? 8 : 16);
  printf("%-34s %d issuer(s) %7.2f cycles per tcgen05.mma (M=128, N=%d, K=%d)  = %6.0f flop/clk/SM   result errors %d   %s\n", name, n_issuers, per, N,
         kTF32 ? 8 : 16, flop / per, e, cudaGetErrorString(cudaGetLastError()));
  return per;
}

int main() {
  long long* cyc;
  int* err;
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  cudaMalloc(&cyc, 8);
  cudaMalloc(&err, 4);
  printf("# tcgen05.mma issue rate at the GEMM shapes of a radix-16 x radix-16 DFT stage, one CTA on each of %d SMs\n", n_sm);
  const double tf_ss = run<true, 32, false>("tf32  N=32  A in shared memory", cyc, err, n_sm);
  const double tf_ts = run<true, 32, true>("tf32  N=32  A in tensor memory", cyc, err, n_sm);
  run<true, 64, false>("tf32  N=64  A in shared memory", cyc, err, n_sm);
  run<true, 64, true>("tf32  N=64  A in tensor memory", cyc, err, n_sm);
  const double bf_ss = run<false, 32, false>("bf16  N=32  A in shared memory", cyc, err, n_sm);
  const double bf_ts = run<false, 32, true>("bf16  N=32  A in tensor memory", cyc, err, n_sm);
  run<false, 64, false>("bf16  N=64  A in shared memory", cyc, err, n_sm);
  run<false, 64, true>("bf16  N=64  A in tensor memory", cyc, err, n_sm);
  // is the ~45-cycle cost per instruction the issuing thread or the tensor pipe?  several issuing threads:
  const double tf_ss2 = run<true, 32, false>("tf32  N=32  A in shared memory", cyc, err, n_sm, 2);
  const double tf_ss4 = run<true, 32, false>("tf32  N=32  A in shared memory", cyc, err, n_sm, 4);
  const double bf_ss4 = run<false, 32, false>("bf16  N=32  A in shared memory", cyc, err, n_sm, 4);
  run<false, 64, true>("bf16  N=64  A in tensor memory", cyc, err, n_sm, 4);
  const double tf_ts4 = run<true, 32, true>("tf32  N=32  A in tensor memory", cyc, err, n_sm, 4);
  const double bf_ts4 = run<false, 32, true>("bf16  N=32  A in tensor memory", cyc, err, n_sm, 4);
  run<true, 64, true>("tf32  N=64  A in tensor memory", cyc, err, n_sm, 4);
  run<true, 32, true>("tf32  N=32  A in tensor memory", cyc, err, n_sm, 2);
  // two stages x 16 GEMMs x (32 K-values per GEMM) per 128-frame tile:
  //   tf32: 4 K-steps per GEMM; products needed for 1e-3 parity (tools/sim_tc_dft.py): 4 (hi/lo x hi/lo)
  //   bf16: 2 K-steps per GEMM; products needed: 6 (three-way split)
  const double clk = 1.965e9;
  auto fps = [&](double cyc_per_mma, int ksteps, int products) {
    return 128.0 / (2 * 16 * ksteps * products * cyc_per_mma) * clk;
  };
  printf("# MMA-bound frames/s/SM of the two DFT stages at %.3f GHz (the FP32 FFT kernel runs at 11.5 M frames/s/SM):\n", clk / 1e9);
  printf("#   tf32 x 4 products: SS %.1f M, TS %.1f M      tf32 x 3 products (fails parity: 1.75e-3): SS %.1f M, TS %.1f M\n",
         fps(tf_ss, 4, 4) / 1e6, fps(tf_ts, 4, 4) / 1e6, fps(tf_ss, 4, 3) / 1e6, fps(tf_ts, 4, 3) / 1e6);
  printf("#   bf16 x 6 products: SS %.1f M, TS %.1f M\n", fps(bf_ss, 2, 6) / 1e6, fps(bf_ts, 2, 6) / 1e6);
  printf("#   with 2 / 4 issuing threads, A in shared memory: tf32 x 4 products %.1f M / %.1f M, bf16 x 6 products (4 issuers) %.1f M\n",
         fps(tf_ss2, 4, 4) / 1e6, fps(tf_ss4, 4, 4) / 1e6, fps(bf_ss4, 2, 6) / 1e6);
  printf("#   with 4 issuing threads, A in tensor memory (best case): tf32 x 4 products %.1f M, bf16 x 6 products %.1f M\n",
         fps(tf_ts4, 4, 4) / 1e6, fps(bf_ts4, 2, 6) / 1e6);
  return 0;
}
