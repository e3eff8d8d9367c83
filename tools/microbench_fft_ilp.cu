// Micro-benchmark (sm_100a): the FFT section of fbank_tile_kernel in isolation — DC mean, window, radix-16,
// twiddle, 16x16 exchange, radix-16, real-input split, power spectrum to shared memory — with the product's own
// device functions (the kernel file is included), in two configurations:
//   NS = 1   two CTAs x 8 warps per SM at <= 128 registers, one frame pair per half-warp   (the product's shape)
//   NS = 2   one CTA x 8 warps per SM at <= 255 registers, TWO frame pairs per half-warp, statement by statement
//            interleaved (twice the instruction-level parallelism per warp, half the warps; window / twiddle table
//            loads shared by the two pairs)
// Question: does the section, which runs at ~70 % of its FP32-pipe limit with four warps per scheduler, get closer to
// it with two fatter warps per scheduler?  (Decision gate for a 64-frame-tile kernel; results are synthetic numbers,
// only the time per frame counts.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mb_fft_ilp tools/microbench_fft_ilp.cu
//   build/mb_fft_ilp
#include "../joeys2t_b200/csrc/fbank_kernels.cu"

#include <cstdio>
#include <vector>

namespace js2t {

// radix-16 on NS independent register sets, butterfly by butterfly interleaved
template <bool kZeroTail, int NS>
__device__ __forceinline__ void fft16n(C2 (&v)[NS][16]) {
  constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    C2 b0, b1, b2, b3;
    JS2T_R4(v[s][0], v[s][4], v[s][8], v[s][12], b0, b1, b2, b3);
    v[s][0] = b0; v[s][4] = b1; v[s][8] = b2; v[s][12] = b3;
  }
#pragma unroll
  for (int i = 1; i < 4; ++i) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      C2 b0, b1, b2, b3;
      if (kZeroTail) {
        JS2T_R4_Z3(v[s][i], v[s][i + 4], v[s][i + 8], b0, b1, b2, b3);
      } else {
        JS2T_R4(v[s][i], v[s][i + 4], v[s][i + 8], v[s][i + 12], b0, b1, b2, b3);
      }
      v[s][i] = b0; v[s][i + 4] = b1; v[s][i + 8] = b2; v[s][i + 12] = b3;
    }
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    v[s][5] = cmul(v[s][5], c1, -s1);
    v[s][13] = cmul(v[s][13], s1, -c1);
    v[s][7] = cmul(v[s][7], s1, -c1);
    v[s][15] = cmul(v[s][15], -c1, s1);
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const u64 hh = bc(h), nh = bc(-h);
    C2 t = v[s][9];
    v[s][9] = C2{mul2(add2(t.re, t.im), hh), mul2(sub2(t.im, t.re), hh)};
    t = v[s][6];
    v[s][6] = C2{mul2(add2(t.re, t.im), hh), mul2(sub2(t.im, t.re), hh)};
    t = v[s][14];
    v[s][14] = C2{mul2(sub2(t.im, t.re), hh), mul2(add2(t.re, t.im), nh)};
    t = v[s][11];
    v[s][11] = C2{mul2(sub2(t.im, t.re), hh), mul2(add2(t.re, t.im), nh)};
    t = v[s][10];
    v[s][10] = C2{t.im, sub2(0ull, t.re)};
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      C2 o0, o1, o2, o3;
      JS2T_R4(v[s][4 * q], v[s][4 * q + 1], v[s][4 * q + 2], v[s][4 * q + 3], o0, o1, o2, o3);
      v[s][4 * q] = o0; v[s][4 * q + 1] = o1; v[s][4 * q + 2] = o2; v[s][4 * q + 3] = o3;
    }
  }
  // (natural order: element q + 4 j of the output sits in v[4 q + j]; transposed back below)
#pragma unroll
  for (int s = 0; s < NS; ++s) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = a + 1; b < 4; ++b) {
        const C2 t = v[s][4 * a + b];
        v[s][4 * a + b] = v[s][4 * b + a];
        v[s][4 * b + a] = t;
      }
  }
}

constexpr int kSigRow = 32;  // float2 (de, dO) per row and half-warp lane: 18 rows x 16 lanes per frame pair

// ABL (ablation, timing only): 1 = no 16x16 exchange, 2 = no table loads (window / twiddles as constants), 4 = no
// radix-16 butterflies, 8 = no partner shuffles in the real-input split, 16 = no W512 loads
template <int NS, int ABL = 0>
__global__ void __launch_bounds__(256, NS == 1 ? 2 : 1) fft_section_kernel(float* __restrict__ out, int passes) {
  extern __shared__ __align__(16) unsigned char sm[];
  float* sWin = reinterpret_cast<float*>(sm);                       // 416 floats
  float2* sTw256 = reinterpret_cast<float2*>(sm + 1664);            // 256
  float2* sTw512 = reinterpret_cast<float2*>(sm + 1664 + 2048);     // 136
  u64* sEx = reinterpret_cast<u64*>(sm + 4864);                     // [8 warps][NS][2][16*17]
  float* sPw = reinterpret_cast<float*>(sm + 4864 + 8 * NS * kExchPerWarp * 8);  // [NS][257*34]
  float2* sSig = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(sPw) + NS * kPFloats * 4);  // [18][16]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, half = lane >> 4, r = lane & 15;
  for (int i = tid; i < 416; i += 256) sWin[i] = i < 400 ? 0.5f - 0.5f * __cosf(6.2831853f * i / 399.f) : 0.f;
  sTw256[tid] = make_float2(__cosf(-6.2831853f * tid / 256.f), __sinf(-6.2831853f * tid / 256.f));
  if (tid < 136) sTw512[tid] = make_float2(__cosf(-6.2831853f * tid / 512.f), __sinf(-6.2831853f * tid / 512.f));
  for (int i = tid; i < 18 * 16 + 16; i += 256) sSig[i] = make_float2(__sinf(0.37f * i) * 1000.f, __cosf(0.11f * i) * 900.f);
  __syncthreads();
  const int partner = (lane & 16) | ((16 - r) & 15);
  const int fA = 4 * warp + 2 * half;
  float keep = 0.f;
#pragma unroll 1
  for (int pass = 0; pass < passes; ++pass) {
    C2 v[NS][16];
    const float2* sig = sSig + (pass & 1);  // (run-time address: the loads stay inside the loop)
    {
      float de[NS][18], dO[NS][18];
      float sa[NS], sb[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) sa[s] = sb[s] = 0.f;
#pragma unroll
      for (int n = 0; n < 18; ++n) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const float2 d = sig[((n + s) % 18) * 16 + r];
          de[s][n] = d.x;
          dO[s][n] = d.y;
          if (n < 13) sa[s] += d.x + d.y;
          if (n >= 5) sb[s] += d.x + d.y;
        }
      }
      u64 mc[NS];
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          sa[s] += __shfl_xor_sync(0xffffffffu, sa[s], off);
          sb[s] += __shfl_xor_sync(0xffffffffu, sb[s], off);
        }
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const float ma0 = sa[s] * 0.0025f, mb0 = sb[s] * 0.0025f;
        const float ma = fmaf(fmaf(-400.0f, ma0, sa[s]), 0.0025f, ma0);
        const float mb = fmaf(fmaf(-400.0f, mb0, sb[s]), 0.0025f, mb0);
        mc[s] = pk(ma * kDcScale, mb * kDcScale);
      }
      const float* wfr = sWin + 2 * r;
#pragma unroll
      for (int n1 = 0; n1 < 13; ++n1) {
        const float2 w = (ABL & 2) ? make_float2(0.3f + 0.01f * n1, 0.7f - 0.01f * n1)
                                   : *reinterpret_cast<const float2*>(wfr + 32 * n1);  // one table load for all streams
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          v[s][n1].re = mul2(sub2(pk(de[s][n1], de[s][n1 + 5]), mc[s]), bc(w.x));
          v[s][n1].im = mul2(sub2(pk(dO[s][n1], dO[s][n1 + 5]), mc[s]), bc(w.y));
        }
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (r >= 8) v[s][12] = C2{0ull, 0ull};
        v[s][13] = v[s][14] = v[s][15] = C2{0ull, 0ull};
      }
    }
    if (!(ABL & 4)) fft16n<true, NS>(v);
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) {
      const float2 w = (ABL & 2) ? make_float2(0.9f - 0.02f * k1, 0.1f + 0.03f * k1) : sTw256[k1 * 16 + r];
#pragma unroll
      for (int s = 0; s < NS; ++s) v[s][k1] = cmul(v[s][k1], w.x, w.y);
    }
    if (!(ABL & 1)) {
    __syncwarp();
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      u64* exch = sEx + ((warp * NS + s) * 2 + half) * (16 * kExchStride);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) exch[k1 * kExchStride + r] = v[s][k1].re;
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const u64* exch = sEx + ((warp * NS + s) * 2 + half) * (16 * kExchStride);
#pragma unroll
      for (int n2 = 0; n2 < 16; ++n2) v[s][n2].re = exch[r * kExchStride + n2];
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      u64* exch = sEx + ((warp * NS + s) * 2 + half) * (16 * kExchStride);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) exch[k1 * kExchStride + r] = v[s][k1].im;
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const u64* exch = sEx + ((warp * NS + s) * 2 + half) * (16 * kExchStride);
#pragma unroll
      for (int n2 = 0; n2 < 16; ++n2) v[s][n2].im = exch[r * kExchStride + n2];
    }
    }
    if (!(ABL & 4)) fft16n<false, NS>(v);
    const bool r0 = (r == 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = r + 16 * j;
      const float2 w = (ABL & (2 | 16)) ? make_float2(0.8f - 0.05f * j, 0.2f + 0.04f * j) : sTw512[k];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        float* Pf = sPw + s * kPFloats + fA;
        const C2 ma = v[s][(15 - j) & 15];
        const C2 mb = v[s][(16 - j) & 15];
        float s0, s1, s2, s3, t0, t1, t2, t3;
        upk(ma.re, s0, s1); upk(ma.im, s2, s3);
        upk(mb.re, t0, t1); upk(mb.im, t2, t3);
        float q0, q1, q2, q3;
        if (ABL & 8) {  // partner values from the lane's own registers: no shuffles, no selects
          q0 = s0; q1 = s1; q2 = s2; q3 = s3;
        } else {
          q0 = __shfl_sync(0xffffffffu, r0 ? t0 : s0, partner);
          q1 = __shfl_sync(0xffffffffu, r0 ? t1 : s1, partner);
          q2 = __shfl_sync(0xffffffffu, r0 ? t2 : s2, partner);
          q3 = __shfl_sync(0xffffffffu, r0 ? t3 : s3, partner);
        }
        const C2 z = v[s][j];
        const C2 zp = C2{pk(q0, q1), pk(q2, q3)};
        const u64 er = add2(z.re, zp.re), ei = sub2(z.im, zp.im);
        const u64 orr = add2(z.im, zp.im), oi = sub2(zp.re, z.re);
        const u64 tr = fma2(oi, bc(-w.y), mul2(orr, bc(w.x)));
        const u64 ti = fma2(orr, bc(w.y), mul2(oi, bc(w.x)));
        const u64 ar = add2(er, tr), ai = add2(ei, ti);
        const u64 br = sub2(er, tr), bi = sub2(ei, ti);
        *reinterpret_cast<u64*>(Pf + k * kPStride) = fma2(ar, ar, mul2(ai, ai));
        *reinterpret_cast<u64*>(Pf + (256 - k) * kPStride) = fma2(br, br, mul2(bi, bi));
      }
    }
    __syncwarp();
    keep += sPw[((pass * 7 + lane) % 257) * kPStride + fA];
  }
  out[blockIdx.x * 256 + tid] = keep;
}

}  // namespace js2t

template <int NS, int ABL = 0>
static void run(const char* name, int grid, int passes) {
  using namespace js2t;
  const size_t smem = 4864 + (size_t)8 * NS * kExchPerWarp * 8 + (size_t)NS * kPFloats * 4 + (18 * 16 + 16) * 8;
  cudaFuncSetAttribute(fft_section_kernel<NS, ABL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, fft_section_kernel<NS, ABL>);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fft_section_kernel<NS, ABL>, 256, smem);
  float* out;
  cudaMalloc(&out, (size_t)grid * 256 * 4);
  fft_section_kernel<NS, ABL><<<grid, 256, smem>>>(out, 8);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    fft_section_kernel<NS, ABL><<<grid, 256, smem>>>(out, passes);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  const double frames = (double)grid * 8 * 4 * NS * passes;
  printf("%-44s grid %4d  regs %3d  spill %3zu B  smem %6zu  CTAs/SM %d : %8.1f us  = %6.3f ns per frame  (%s)\n", name, grid,
         fa.numRegs, (size_t)fa.localSizeBytes, smem, occ, best * 1e3, best * 1e6 / frames,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // same number of frames per SM in every configuration: 2 CTAs x 1 pair x 2 P  ==  1 CTA x 2 pairs x 2 P
  run<1>("NS=1: 2 CTAs/SM, one pair per half-warp", 2 * sms, 400);
  run<1>("NS=1: 1 CTA/SM (8 warps only)", sms, 800);
  run<2>("NS=2: 1 CTA/SM, two pairs per half-warp", sms, 400);
  run<1, 1>("NS=1, no exchange", 2 * sms, 400);
  run<1, 2>("NS=1, no table loads", 2 * sms, 400);
  run<1, 3>("NS=1, no exchange, no table loads", 2 * sms, 400);
  run<1, 4>("NS=1, no butterflies", 2 * sms, 400);
  run<1, 5>("NS=1, no butterflies, no exchange", 2 * sms, 400);
  run<2, 2>("NS=2, no table loads", sms, 400);
  run<1, 8>("NS=1, no partner shuffles / selects", 2 * sms, 400);
  run<1, 16>("NS=1, no W512 loads", 2 * sms, 400);
  run<1, 24>("NS=1, no partner shuffles, no W512 loads", 2 * sms, 400);
  // the product's FFT section for comparison: 61.5 % of 173 us for 318 883 frames = 0.334 ns per frame
  return 0;
}
