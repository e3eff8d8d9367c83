// Hand-written sm_100a kernels for JoeyS2T's audio front-end hot path:
//
//   joeynmt/helpers_for_audio.py:30-37,41-68   extract_fbank_features -> torchaudio kaldi.fbank
//   torchaudio/compliance/kaldi.py:154-217      framing, DC removal, pre-emphasis, povey window
//   torchaudio/compliance/kaldi.py:616-633      rfft -> |.|^2 -> 80x257 mel -> log(max(., eps))
//   joeynmt/data_augmentation.py:96-109         CMVN           (finalize + apply kernels)
//   joeynmt/data_augmentation.py:38-73          SpecAugment    (host-drawn masks, apply kernel)
//   joeynmt/helpers_for_audio.py:130-170        pad_features   (padded (B,Tmax,80) layout, pad 1.0)
//
// Work decomposition: one CTA (256 threads, 8 warps) per tile of 32 consecutive frames of one
// utterance of the ragged batch.  Per tile:
//   1. stage  the tile's PCM (int16 or fp32, 128-bit coalesced loads) into shared memory as the
//             frame-independent pre-emphasised signal d[j] = x[j] - 0.97 x[j-1], plus 8-sample
//             partial sums that give every frame's DC mean without re-reading the samples;
//   2. fft    each half-warp transforms one frame: z[n] = y[2n] + i y[2n+1] (y = windowed frame,
//             zero-padded to 512) as a 256-point complex FFT = radix-16 in registers, twiddle,
//             16x16 transpose through shared memory, radix-16 in registers; the real-input split
//             (partner bin 256-k fetched with warp shuffles) yields the power spectrum directly;
//   3. mel    lane = frame, warp = run of consecutive mel filters; the sparsity structure of the
//             mel bank is compile-time (mel_structure.inc), the weights are __constant__ operands;
//   4. store  log-mel tile to HBM with coalesced stores (+ per-tile column sums / sums of squares
//             for CMVN, or normalisation + masking right here when the statistics are known).
// FP32 throughout: the path is a small FP32 contraction, not a tensor-core workload.
#include "js2t_internal.h"

#include "mel_structure.inc"

namespace js2t {

// mel weights in two-band form (tables.py: mel_two_band): wu[k] -> filter seg(k), wd[k] -> seg(k)-1
__constant__ float c_mel_wu[256];
__constant__ float c_mel_wd[256];

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr float kPreemph = 0.97f;
// (x_j - m) - 0.97f (x_{j-1} - m) == (x_j - 0.97f x_{j-1}) - (1 - 0.97f) m   (kaldi.py:183-198)
constexpr float kDcScale = (float)(1.0 - (double)0.97f);
constexpr float kLogFloor = 1.1920928955078125e-07f;  // FLT_EPSILON, kaldi.py:22,633

// ---- shared memory carve-up (bytes) -------------------------------------------------------------
constexpr int kDFloats = 5376;                   // >= kTileSamples (5360), multiple of 8
constexpr int kPsumFloats = kDFloats / 8;        // 672 partial sums of 8 samples
constexpr int kExchStride = 17;                  // float2 per row of the 16x16 transpose (padded)
constexpr int kExchPerWarp = 2 * 16 * kExchStride;  // float2: two half-warps
constexpr int kPStride = 33;                     // P[k][frame], padded: conflict-free both ways
constexpr int kPFloats = 257 * kPStride;
constexpr int kOutStride = 81;                   // outTile[frame][mel], padded
constexpr int kOutFloats = kTileFrames * kOutStride;

constexpr int kOffD = 0;
constexpr int kOffPsum = kOffD + kDFloats * 4;
constexpr int kOffMean = kOffPsum + kPsumFloats * 4;
constexpr int kOffWin = kOffMean + kTileFrames * 4;
constexpr int kOffTw256 = kOffWin + kFrameLen * 4;
constexpr int kOffTw512 = kOffTw256 + 256 * 8;
constexpr int kOffExch = kOffTw512 + 136 * 8;
constexpr int kOffP = kOffExch + kWarps * kExchPerWarp * 8;
constexpr int kSmemBytes = kOffP + kPFloats * 4;
static_assert(kOutFloats * 4 <= kWarps * kExchPerWarp * 8, "out tile aliases the exchange buffers");
static_assert(kOffExch % 16 == 0 && kOffP % 16 == 0 && kOffTw256 % 16 == 0, "alignment");

int fbank_smem_bytes() { return kSmemBytes; }

// ---- small complex helpers -----------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}
// a + (-i) b  and  a + (+i) b
__device__ __forceinline__ float2 add_mi(float2 a, float2 b) { return make_float2(a.x + b.y, a.y - b.x); }
__device__ __forceinline__ float2 add_pi(float2 a, float2 b) { return make_float2(a.x - b.y, a.y + b.x); }

#define JS2T_R4(a0, a1, a2, a3, b0, b1, b2, b3) \
  {                                             \
    const float2 t0 = cadd(a0, a2);             \
    const float2 t1 = csub(a0, a2);             \
    const float2 t2 = cadd(a1, a3);             \
    const float2 t3 = csub(a1, a3);             \
    b0 = cadd(t0, t2);                          \
    b2 = csub(t0, t2);                          \
    b1 = add_mi(t1, t3);                        \
    b3 = add_pi(t1, t3);                        \
  }

// 16-point complex FFT in registers, natural order in and out (4 x 4 Cooley-Tukey).
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 b0, b1, b2, b3;
    JS2T_R4(v[i], v[i + 4], v[i + 8], v[i + 12], b0, b1, b2, b3);
    v[i] = b0; v[i + 4] = b1; v[i + 8] = b2; v[i + 12] = b3;
  }
  // twiddles W16^(i*q) on v[i + 4q]
  v[5] = cmul(v[5], make_float2(c1, -s1));                        // W^1
  v[9] = make_float2((v[9].x + v[9].y) * h, (v[9].y - v[9].x) * h);      // W^2 = (1 - i)/sqrt2
  v[13] = cmul(v[13], make_float2(s1, -c1));                      // W^3
  v[6] = make_float2((v[6].x + v[6].y) * h, (v[6].y - v[6].x) * h);      // W^2
  v[10] = make_float2(v[10].y, -v[10].x);                         // W^4 = -i
  v[14] = make_float2((v[14].y - v[14].x) * h, -(v[14].x + v[14].y) * h);  // W^6 = (-1 - i)/sqrt2
  v[7] = cmul(v[7], make_float2(s1, -c1));                        // W^3
  v[11] = make_float2((v[11].y - v[11].x) * h, -(v[11].x + v[11].y) * h);  // W^6
  v[15] = cmul(v[15], make_float2(-c1, s1));                      // W^9
  float2 o[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    JS2T_R4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3], o[q], o[q + 4], o[q + 8], o[q + 12]);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = o[i];
}

// ---- streaming loads -------------------------------------------------------------------------------
__device__ __forceinline__ int4 ldg_stream_int4(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float pcm_sample(const uint8_t* base, long long idx, bool is_f32) {
  return is_f32 ? __ldg(reinterpret_cast<const float*>(base) + idx) * 32768.0f
                : (float)__ldg(reinterpret_cast<const short*>(base) + idx);
}

// ---- mel stage: one run of consecutive filters [M0, M1), lane = frame ---------------------------------
template <int M0, int M1>
__device__ __forceinline__ void mel_group(const float* __restrict__ Pl, float* __restrict__ orow) {
  float lo_acc = 0.f, hi_acc = 0.f;
#define JS2T_SEG(s, lo, hi)                                                          \
  if constexpr ((s) >= M0 && (s) <= M1) {                                            \
    _Pragma("unroll") for (int k = (lo); k <= (hi); ++k) {                           \
      const float p = Pl[k * kPStride];                                              \
      if constexpr ((s) < M1) hi_acc = fmaf(c_mel_wu[k], p, hi_acc);                 \
      if constexpr ((s) > M0) lo_acc = fmaf(c_mel_wd[k], p, lo_acc);                 \
    }                                                                                \
    if constexpr ((s) > M0) orow[(s)-1] = __logf(fmaxf(lo_acc, kLogFloor));          \
    lo_acc = hi_acc;                                                                 \
    hi_acc = 0.f;                                                                    \
  }
  JS2T_MEL_SEGMENTS(JS2T_SEG)
#undef JS2T_SEG
}

// =====================================================================================================
//  Kernel A: PCM tile -> log-mel tile (+ statistics | normalisation)
// =====================================================================================================
__global__ void __launch_bounds__(kThreads, 2) fbank_tile_kernel(const FbankLaunch p) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* sD = reinterpret_cast<float*>(smem + kOffD);
  float* sPsum = reinterpret_cast<float*>(smem + kOffPsum);
  float* sMean = reinterpret_cast<float*>(smem + kOffMean);
  float* sWin = reinterpret_cast<float*>(smem + kOffWin);
  float2* sTw256 = reinterpret_cast<float2*>(smem + kOffTw256);
  float2* sTw512 = reinterpret_cast<float2*>(smem + kOffTw512);
  float2* sExch = reinterpret_cast<float2*>(smem + kOffExch);
  float* sOut = reinterpret_cast<float*>(smem + kOffExch);  // aliases sExch after the FFT phase
  float* sP = reinterpret_cast<float*>(smem + kOffP);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  const TileDesc td = p.tiles[blockIdx.x];
  const UttDesc ud = p.utts[td.utt];
  const int frame0 = td.frame0;
  const int nf = max(0, min(kTileFrames, ud.n_frames - frame0));  // valid frames in this tile
  // rows this tile owns in the output (padded layout: up to Tmax, the tail is padding)
  const int rows = p.pad_tmax > 0 ? min(kTileFrames, p.pad_tmax - frame0) : nf;
  float* __restrict__ out_tile = p.out + (ud.out_row + frame0) * (long long)kMel;

  if (nf == 0) {  // pure padding tile
    if (p.tile_stats != nullptr && tid < kStatsPerTile)
      p.tile_stats[(long long)blockIdx.x * kStatsPerTile + tid] = 0.f;
    for (int e = tid; e < rows * kMel; e += kThreads) out_tile[e] = p.pad_value;
    return;
  }

  // ---- phase 0: tables into shared memory ---------------------------------------------------------
  for (int i = tid; i < kFrameLen; i += kThreads) sWin[i] = p.tab.window_half[i];
  sTw256[tid] = p.tab.tw256[tid];
  if (tid < 136) sTw512[tid] = p.tab.tw512[tid];

  // ---- phase 1: stage PCM as d[j] = x[j] - 0.97 x[j-1] and 8-sample partial sums ------------------
  {
    const bool is_f32 = (ud.flags & 1) != 0;
    const uint8_t* base = p.pcm + ud.pcm_byte_off;
    const long long s0 = (long long)frame0 * kHop;  // first sample of the tile
    const int n_chunks = ((nf - 1) * kHop + kFrameLen) >> 3;
    for (int c = tid; c < n_chunks; c += kThreads) {
      const long long j0 = s0 + 8ll * c;
      float x[8];
      if (is_f32) {
        const int4 a = ldg_stream_int4(reinterpret_cast<const float*>(base) + j0);
        const int4 b = ldg_stream_int4(reinterpret_cast<const float*>(base) + j0 + 4);
        x[0] = __int_as_float(a.x) * 32768.f; x[1] = __int_as_float(a.y) * 32768.f;
        x[2] = __int_as_float(a.z) * 32768.f; x[3] = __int_as_float(a.w) * 32768.f;
        x[4] = __int_as_float(b.x) * 32768.f; x[5] = __int_as_float(b.y) * 32768.f;
        x[6] = __int_as_float(b.z) * 32768.f; x[7] = __int_as_float(b.w) * 32768.f;
      } else {
        const int4 a = ldg_stream_int4(reinterpret_cast<const short*>(base) + j0);
        x[0] = (float)(short)(a.x & 0xffff); x[1] = (float)(a.x >> 16);
        x[2] = (float)(short)(a.y & 0xffff); x[3] = (float)(a.y >> 16);
        x[4] = (float)(short)(a.z & 0xffff); x[5] = (float)(a.z >> 16);
        x[6] = (float)(short)(a.w & 0xffff); x[7] = (float)(a.w >> 16);
      }
      // previous sample; at the very first sample of the utterance any finite value will do
      // (it only reaches frame position 0, where the povey window is exactly 0)
      const float xm1 = (j0 > 0) ? pcm_sample(base, j0 - 1, is_f32) : x[0];
      float4 d0, d1;
      d0.x = fmaf(-kPreemph, xm1, x[0]);
      d0.y = fmaf(-kPreemph, x[0], x[1]);
      d0.z = fmaf(-kPreemph, x[1], x[2]);
      d0.w = fmaf(-kPreemph, x[2], x[3]);
      d1.x = fmaf(-kPreemph, x[3], x[4]);
      d1.y = fmaf(-kPreemph, x[4], x[5]);
      d1.z = fmaf(-kPreemph, x[5], x[6]);
      d1.w = fmaf(-kPreemph, x[6], x[7]);
      reinterpret_cast<float4*>(sD)[2 * c] = d0;
      reinterpret_cast<float4*>(sD)[2 * c + 1] = d1;
      sPsum[c] = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
    }
  }
  __syncthreads();

  // per-frame DC mean (kaldi.py:183-186), pre-multiplied by (1 - 0.97): 8 threads per frame
  {
    const int f = tid >> 3, sub = tid & 7;
    float s = 0.f;
    if (f < nf) {
      const float* ps = sPsum + 20 * f;
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const int c = sub + 8 * i;
        if (c < 50) s += ps[c];
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (sub == 0) sMean[f] = (s / 400.0f) * kDcScale;
  }
  __syncthreads();

  // ---- phase 2: one frame per half-warp -> power spectrum P[k][frame] -----------------------------
  {
    const int half = lane >> 4;
    const int r = lane & 15;
    float2* exch = sExch + warp * kExchPerWarp + half * (16 * kExchStride);
    const int partner = (lane & 16) | ((16 - r) & 15);
    for (int it = warp; it < 16; it += kWarps) {
      if (it >= nf) break;  // both frames of this iteration are past the end (warp-uniform)
      const int f = it + 16 * half;
      const bool valid = f < nf;
      float2 v[16];
      {
        const float mc = sMean[valid ? f : 0];
        const float* dfr = sD + (valid ? f : 0) * kHop + 2 * r;
        const float* wfr = sWin + 2 * r;
#pragma unroll
        for (int n1 = 0; n1 < 12; ++n1) {
          const float2 x = *reinterpret_cast<const float2*>(dfr + 32 * n1);
          const float2 w = *reinterpret_cast<const float2*>(wfr + 32 * n1);
          v[n1] = make_float2((x.x - mc) * w.x, (x.y - mc) * w.y);
        }
        if (r < 8) {  // samples 384 + 2r, 385 + 2r < 400
          const float2 x = *reinterpret_cast<const float2*>(dfr + 384);
          const float2 w = *reinterpret_cast<const float2*>(wfr + 384);
          v[12] = make_float2((x.x - mc) * w.x, (x.y - mc) * w.y);
        } else {
          v[12] = make_float2(0.f, 0.f);
        }
        v[13] = v[14] = v[15] = make_float2(0.f, 0.f);
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 13; ++i) v[i] = make_float2(0.f, 0.f);
        }
      }
      // pass 1: DFT over n1 (lane = n2 = r), then twiddle by W_256^(n2*k1)
      fft16(v);
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmul(v[k1], sTw256[k1 * 16 + r]);
      // 16x16 transpose inside the half-warp
      __syncwarp();
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) exch[k1 * kExchStride + r] = v[k1];
      __syncwarp();
#pragma unroll
      for (int n2 = 0; n2 < 16; ++n2) v[n2] = exch[r * kExchStride + n2];
      // pass 2: DFT over n2 (lane = k1 = r): v[k2] = Z[r + 16 k2]
      fft16(v);

      // real-input split: bins k = r + 16 j and 256 - k from Z[k] and Z[256 - k] (partner lane)
      float* Pf = sP + f;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        // partner register index: 15 - j, except in lane r == 0 where it is (16 - j) & 15
        const float2 mine_a = v[(15 - j) & 15];
        const float2 mine_b = v[(16 - j) & 15];
        const float2 send = (r == 0) ? mine_b : mine_a;
        float2 zp;
        zp.x = __shfl_sync(0xffffffffu, send.x, partner);
        zp.y = __shfl_sync(0xffffffffu, send.y, partner);
        if (j == 8 && r != 0) continue;  // k = 128 exists only in lane r == 0
        const float2 z = v[j];
        const int k = r + 16 * j;
        const float2 w = sTw512[k];
        const float er = z.x + zp.x, ei = z.y - zp.y;
        const float orr = z.y + zp.y, oi = zp.x - z.x;
        const float tr = fmaf(w.x, orr, -w.y * oi);
        const float ti = fmaf(w.x, oi, w.y * orr);
        const float ar = er + tr, ai = ei + ti;
        const float br = er - tr, bi = ei - ti;
        if (valid) {
          Pf[k * kPStride] = fmaf(ar, ar, ai * ai);
          Pf[(256 - k) * kPStride] = fmaf(br, br, bi * bi);
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 3: mel filterbank + log, lane = frame, warp = run of filters --------------------------
  {
    const float* Pl = sP + lane;
    float* orow = sOut + lane * kOutStride;
    if (lane < nf) {
      switch (warp) {
#define JS2T_GRP(g, m0, m1) \
  case g:                   \
    mel_group<m0, m1>(Pl, orow); \
    break;
        JS2T_MEL_GROUPS(JS2T_GRP)
#undef JS2T_GRP
        default:
          break;
      }
    }
  }
  __syncthreads();

  // ---- phase 4: store -----------------------------------------------------------------------------
  if (p.epilogue == kEpiRaw) {
    for (int e = tid; e < rows * kMel; e += kThreads) {
      const int f = e / kMel, m = e - f * kMel;
      out_tile[e] = (f < nf) ? sOut[f * kOutStride + m] : p.pad_value;
    }
    if (p.tile_stats != nullptr && tid < kStatsPerTile) {
      // column sum (tid < 80) or sum of squares (tid >= 80) over the tile's valid frames
      const int m = tid < kMel ? tid : tid - kMel;
      const bool sq = tid >= kMel;
      float acc = 0.f;
      for (int f = 0; f < nf; ++f) {
        const float x = sOut[f * kOutStride + m];
        acc += sq ? x * x : x;
      }
      p.tile_stats[(long long)blockIdx.x * kStatsPerTile + tid] = acc;
    }
  } else {  // kEpiNormKnown: (x - mean) * istd and SpecAugment fill at store
    const int* mk = p.masks ? p.masks + (long long)td.utt * (p.n_fmask + p.n_tmask) * 2 : nullptr;
    const float mv = p.mask_value ? p.mask_value[td.utt] : 0.f;
    for (int e = tid; e < rows * kMel; e += kThreads) {
      const int f = e / kMel, m = e - f * kMel;
      float y = p.pad_value;
      if (f < nf) {
        y = (sOut[f * kOutStride + m] - p.g_mean[m]) * p.g_istd[m];
        if (mk != nullptr) {
          const int t = frame0 + f;
          bool masked = false;
          for (int i = 0; i < p.n_fmask; ++i) masked |= (unsigned)(m - mk[2 * i]) < (unsigned)mk[2 * i + 1];
          for (int i = p.n_fmask; i < p.n_fmask + p.n_tmask; ++i)
            masked |= (unsigned)(t - mk[2 * i]) < (unsigned)mk[2 * i + 1];
          if (masked) y = mv;
        }
      }
      out_tile[e] = y;
    }
  }
}

// =====================================================================================================
//  Kernel A': pre-extracted features (the reference's .npy / npy-in-zip branch,
//  joeynmt/helpers_for_audio.py:100-127) -> same raw layout + per-tile statistics, so that CMVN and
//  SpecAugment can run on feature matrices that did not come from the fbank kernel.
// =====================================================================================================
__global__ void __launch_bounds__(kThreads) feature_tile_kernel(const FbankLaunch p) {
  const TileDesc td = p.tiles[blockIdx.x];
  const UttDesc ud = p.utts[td.utt];
  const int frame0 = td.frame0;
  const int nf = max(0, min(kTileFrames, ud.n_frames - frame0));
  const int rows = p.pad_tmax > 0 ? min(kTileFrames, p.pad_tmax - frame0) : nf;
  const float* __restrict__ in_tile =
      reinterpret_cast<const float*>(p.pcm + ud.pcm_byte_off) + (long long)frame0 * kMel;
  float* __restrict__ out_tile = p.out + (ud.out_row + frame0) * (long long)kMel;
  const int tid = threadIdx.x;
  if (in_tile != out_tile || nf < rows) {
    for (int e = tid; e < rows * kMel; e += kThreads) out_tile[e] = (e < nf * kMel) ? in_tile[e] : p.pad_value;
  }
  if (p.tile_stats != nullptr && tid < kStatsPerTile) {
    const int m = tid < kMel ? tid : tid - kMel;
    const bool sq = tid >= kMel;
    float acc = 0.f;
    for (int f = 0; f < nf; ++f) {
      const float x = in_tile[f * kMel + m];
      acc += sq ? x * x : x;
    }
    p.tile_stats[(long long)blockIdx.x * kStatsPerTile + tid] = acc;
  }
}

// =====================================================================================================
//  Kernel F: per-utterance statistics -> mean / inverse std / SpecAugment fill value
//  (joeynmt/data_augmentation.py:96-109 CMVN; :43-46 mask value; tokenizers.py:488-493 order)
// =====================================================================================================
__global__ void __launch_bounds__(128) finalize_utt_kernel(const FinalizeLaunch p) {
  const int u = blockIdx.x;
  const int b = threadIdx.x;  // mel bin
  const UttDesc ud = p.utts[u];
  const int T = ud.n_frames;
  const int n_tiles = (T + kTileFrames - 1) / kTileFrames;
  __shared__ double s_red[kMel];
  __shared__ float s_mv;

  double S = 0.0, Q = 0.0;
  if (b < kMel) {
    const float* ts = p.tile_stats + (long long)ud.tile_start * kStatsPerTile;
    for (int t = 0; t < n_tiles; ++t) {  // fixed order: deterministic
      S += (double)ts[(long long)t * kStatsPerTile + b];
      Q += (double)ts[(long long)t * kStatsPerTile + kMel + b];
    }
    if (p.stats_out != nullptr) {
      p.stats_out[(long long)u * kStatsPerTile + b] = S;
      p.stats_out[(long long)u * kStatsPerTile + kMel + b] = Q;
    }
  }
  const int n_masks = p.n_fmask + p.n_tmask;
  const int* mk = (p.masks != nullptr) ? p.masks + (long long)u * n_masks * 2 : nullptr;

  double mean = 0.0, istd = 1.0;
  const bool shared = p.g_mean != nullptr;  // global CMVN: statistics are given, not computed
  const bool after = p.cmvn_enabled && p.cmvn_after && mk != nullptr && !shared;
  if (!after) {
    if (b < kMel && p.cmvn_enabled) {
      if (shared) {
        mean = (double)p.g_mean[b];
        istd = (double)p.g_istd[b];
      } else {
        const double mu = S / T;
        if (p.norm_means) mean = (double)(float)mu;
        if (p.norm_vars) {
          const double var = Q / T - mu * mu;
          istd = 1.0 / sqrt(fmax(var, 1e-10));
        }
      }
    }
    // fill value = mean of the spectrogram SpecAugment sees (data_augmentation.py:45-46)
    if (b < kMel) {
      double col = S / T;                                              // raw column mean
      if (p.cmvn_enabled && !p.cmvn_after) col = (col - mean) * istd;  // after CMVN(before)
      s_red[b] = col;
    }
    __syncthreads();
    if (b == 0) {
      double acc = 0.0;
      for (int i = 0; i < kMel; ++i) acc += s_red[i];
      s_mv = (p.mask_value_mode == 1) ? p.mask_value_const : (float)(acc / kMel);
    }
    __syncthreads();
  } else {
    // SpecAugment on the raw log-mel first, then CMVN over the *masked* spectrogram
    if (b < kMel) s_red[b] = S / T;
    __syncthreads();
    if (b == 0) {
      double acc = 0.0;
      for (int i = 0; i < kMel; ++i) acc += s_red[i];
      s_mv = (p.mask_value_mode == 1) ? p.mask_value_const : (float)(acc / kMel);
    }
    __syncthreads();
    if (b < kMel) {
      const double v = (double)s_mv;
      bool col_masked = false;
      for (int i = 0; i < p.n_fmask; ++i) col_masked |= (unsigned)(b - mk[2 * i]) < (unsigned)mk[2 * i + 1];
      if (col_masked) {
        S = v * T;
        Q = v * v * T;
      } else {
        // rows covered by any time mask: replace their contribution by the fill value
        int lo = T, hi = 0;
        for (int i = p.n_fmask; i < n_masks; ++i) {
          if (mk[2 * i + 1] > 0) {
            lo = min(lo, mk[2 * i]);
            hi = max(hi, mk[2 * i] + mk[2 * i + 1]);
          }
        }
        const float* x = p.raw + ud.out_row * (long long)kMel + b;
        for (int t = lo; t < min(hi, T); ++t) {
          bool m = false;
          for (int i = p.n_fmask; i < n_masks; ++i) m |= (unsigned)(t - mk[2 * i]) < (unsigned)mk[2 * i + 1];
          if (m) {
            const double xv = (double)x[(long long)t * kMel];
            S += v - xv;
            Q += v * v - xv * xv;
          }
        }
      }
      const double mu = S / T;
      if (p.norm_means) mean = (double)(float)mu;
      if (p.norm_vars) {
        const double var = Q / T - mu * mu;
        istd = 1.0 / sqrt(fmax(var, 1e-10));
      }
    }
  }
  if (b < kMel && !shared) {
    p.mean[(long long)u * kMel + b] = (float)mean;
    p.istd[(long long)u * kMel + b] = (float)istd;
  }
  if (b == 0) p.mask_value[u] = s_mv;
}

// =====================================================================================================
//  Kernel C: in-place CMVN + SpecAugment fill (+ padding rows of the padded layout)
// =====================================================================================================
__global__ void __launch_bounds__(kThreads) apply_kernel(const ApplyLaunch p) {
  const TileDesc td = p.tiles[blockIdx.x];
  const UttDesc ud = p.utts[td.utt];
  const int frame0 = td.frame0;
  const int nf = max(0, min(kTileFrames, ud.n_frames - frame0));
  const int rows = p.pad_tmax > 0 ? min(kTileFrames, p.pad_tmax - frame0) : nf;
  float4* __restrict__ o4 = reinterpret_cast<float4*>(p.out + (ud.out_row + frame0) * (long long)kMel);
  const long long so = p.shared_stats ? 0 : (long long)td.utt * kMel;
  const float4* mean4 = reinterpret_cast<const float4*>(p.mean + so);
  const float4* istd4 = reinterpret_cast<const float4*>(p.istd + so);
  const int n_masks = p.n_fmask + p.n_tmask;
  const int* mk = (p.masks != nullptr) ? p.masks + (long long)td.utt * n_masks * 2 : nullptr;
  const float mv = (p.mask_value != nullptr) ? p.mask_value[td.utt] : 0.f;

  for (int e = threadIdx.x; e < rows * (kMel / 4); e += kThreads) {
    const int f = e / (kMel / 4), c4 = e - f * (kMel / 4);
    float4 y;
    if (f < nf) {
      float4 x = o4[e];
      const float4 mu = mean4[c4], is = istd4[c4];
      bool m0 = false, m1 = false, m2 = false, m3 = false;
      if (mk != nullptr) {
        const int t = frame0 + f;
        bool trow = false;
        for (int i = p.n_fmask; i < n_masks; ++i) trow |= (unsigned)(t - mk[2 * i]) < (unsigned)mk[2 * i + 1];
        m0 = m1 = m2 = m3 = trow;
        for (int i = 0; i < p.n_fmask; ++i) {
          const int f0 = mk[2 * i];
          const unsigned w = (unsigned)mk[2 * i + 1];
          m0 |= (unsigned)(4 * c4 + 0 - f0) < w;
          m1 |= (unsigned)(4 * c4 + 1 - f0) < w;
          m2 |= (unsigned)(4 * c4 + 2 - f0) < w;
          m3 |= (unsigned)(4 * c4 + 3 - f0) < w;
        }
      }
      if (p.cmvn_after) {  // SpecAugment saw the raw log-mel; CMVN normalises the filled cells too
        if (m0) x.x = mv;
        if (m1) x.y = mv;
        if (m2) x.z = mv;
        if (m3) x.w = mv;
      }
      y = make_float4((x.x - mu.x) * is.x, (x.y - mu.y) * is.y, (x.z - mu.z) * is.z, (x.w - mu.w) * is.w);
      if (!p.cmvn_after) {
        if (m0) y.x = mv;
        if (m1) y.y = mv;
        if (m2) y.z = mv;
        if (m3) y.w = mv;
      }
    } else {
      y = make_float4(p.pad_value, p.pad_value, p.pad_value, p.pad_value);
    }
    o4[e] = y;
  }
}

// =====================================================================================================
//  Global CMVN statistics (extension named by north_star; formula of data_augmentation.py:98-105)
// =====================================================================================================
__global__ void __launch_bounds__(kStatsPerTile + 32) global_accumulate_kernel(
    const double* __restrict__ utt_stats, const UttDesc* __restrict__ utts, int n_utts,
    double* __restrict__ accum) {
  const int t = threadIdx.x;
  if (t < kStatsPerTile) {
    double a = 0.0;
    for (int u = 0; u < n_utts; ++u) a += utt_stats[(long long)u * kStatsPerTile + t];
    accum[t] += a;
  } else if (t == kStatsPerTile) {
    double n = 0.0;
    for (int u = 0; u < n_utts; ++u) n += (double)utts[u].n_frames;
    accum[kStatsPerTile] += n;
  }
}

__global__ void __launch_bounds__(128) global_finalize_kernel(const double* __restrict__ accum,
                                                              int norm_means, int norm_vars,
                                                              float* __restrict__ mean,
                                                              float* __restrict__ istd) {
  const int b = threadIdx.x;
  if (b >= kMel) return;
  const double n = accum[kStatsPerTile];
  const double mu = accum[b] / n;
  const double var = accum[kMel + b] / n - mu * mu;
  mean[b] = norm_means ? (float)mu : 0.f;
  istd[b] = norm_vars ? (float)(1.0 / sqrt(fmax(var, 1e-10))) : 1.f;
}

__global__ void fill_value_kernel(float* __restrict__ dst, int n, float value) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = value;
}

cudaError_t launch_fill_value(float* dst, int n, float value, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  fill_value_kernel<<<(n + 255) / 256, 256, 0, s>>>(dst, n, value);
  return cudaGetLastError();
}

// ---- launchers -----------------------------------------------------------------------------------
cudaError_t upload_mel_weights(const float* wu256, const float* wd256, cudaStream_t s) {
  cudaError_t e = cudaMemcpyToSymbolAsync(c_mel_wu, wu256, 256 * sizeof(float), 0, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbolAsync(c_mel_wd, wd256, 256 * sizeof(float), 0, cudaMemcpyHostToDevice, s);
}

cudaError_t launch_fbank(const FbankLaunch& p, cudaStream_t s) {
  static bool configured = false;  // per process; the attribute is per device function
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fbank_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (p.n_tiles <= 0) return cudaSuccess;
  fbank_tile_kernel<<<p.n_tiles, kThreads, kSmemBytes, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_features(const FbankLaunch& p, cudaStream_t s) {
  if (p.n_tiles <= 0) return cudaSuccess;
  feature_tile_kernel<<<p.n_tiles, kThreads, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_finalize(const FinalizeLaunch& p, cudaStream_t s) {
  if (p.n_utts <= 0) return cudaSuccess;
  finalize_utt_kernel<<<p.n_utts, 128, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_apply(const ApplyLaunch& p, cudaStream_t s) {
  if (p.n_tiles <= 0) return cudaSuccess;
  apply_kernel<<<p.n_tiles, kThreads, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_global_accumulate(const double* utt_stats, const UttDesc* utts, int n_utts,
                                     double* accum, cudaStream_t s) {
  global_accumulate_kernel<<<1, kStatsPerTile + 32, 0, s>>>(utt_stats, utts, n_utts, accum);
  return cudaGetLastError();
}

cudaError_t launch_global_finalize(const double* accum, int norm_means, int norm_vars, float* mean,
                                   float* istd, cudaStream_t s) {
  global_finalize_kernel<<<1, 128, 0, s>>>(accum, norm_means, norm_vars, mean, istd);
  return cudaGetLastError();
}

}  // namespace js2t
