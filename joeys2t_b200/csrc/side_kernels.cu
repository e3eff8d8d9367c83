// Side kernel of JoeyS2T's audio front-end: the in-place CMVN + SpecAugment pass
//   joeynmt/data_augmentation.py:96-109  CMVN.__call__      (x - mean) / std per mel bin
//   joeynmt/data_augmentation.py:38-73   SpecAugment        host-drawn masks filled with the mask value
//   joeynmt/helpers_for_audio.py:130-170 pad_features       padding rows of the padded layout
// as a small persistent kernel that runs ON THE SAME SMs, AT THE SAME TIME as the fbank kernel of the next
// batch (pipelined plans: consecutive batches alternate between streams).  The fbank kernel is bound by the
// SM's FP32 pipe and shared-memory crossbar and leaves HBM 90 % idle; this pass is a pure HBM stream, so the
// two overlap instead of queueing behind each other.  What the measurements on the B200 say it takes
// (profiles/r2_corun_ab.txt):
//   * footprint: the fbank kernel's two resident CTAs (112 registers per thread) leave 2 048 registers per
//     scheduler and 31 KB of shared memory: 128 threads = one warp per scheduler at <= 64 registers, no
//     shared memory; every kernel of the path asks for the same (maximum) shared-memory carve-out;
//   * one CTA of a launch per SM (residency gate below): the block scheduler otherwise stacks several on an SM
//     that is momentarily empty, which then cannot take its two fbank CTAs; no programmatic dependent launch
//     (early-launched dependents squat on the free resources);
//   * every warp-wide load / store covers 512 CONTIGUOUS bytes.  A first version with 20 active lanes per
//     80-float row (+ 12 lanes shadowing the last one) cost the co-resident fbank kernel 23 us per batch, this
//     one 5-8 us; a TMA-fed shared-memory ring (data in flight in shared memory instead of registers) cost
//     35-50 us — its shared-memory traffic and its code footprint both land on the fbank kernel's bottlenecks
//     (shared-memory crossbar, instruction fetch);
//   * a loop body of a few hundred bytes, kept rolled.
#include "js2t_internal.h"

namespace js2t {

#ifndef JS2T_PDL
#define JS2T_PDL 1
#endif
__device__ __forceinline__ void pdl_wait() {
#if JS2T_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_launch() {
#if JS2T_PDL
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
#ifndef JS2T_SIDE_CTAS_DEFAULT
#define JS2T_SIDE_CTAS_DEFAULT 2  // CTAs launched per SM; with a residency limit the surplus ones leave at once
#endif

// =====================================================================================================
//  Every WARP is an independent worker: it claims a tile (dynamic claims, newest tile first — see apply_kernel),
//  normalises its 32 x 80 floats in place and claims the next.  No block barrier, no shared memory; the data
//  in flight sits in registers (four 16-byte loads per lane).
// =====================================================================================================
#define JS2T_SIDE_LOAD(ptr) (*(ptr))
#define JS2T_SIDE_STORE(ptr, v) (*(ptr) = (v))
constexpr int kWarpSideThreads = 128;  // four independent warps, one per scheduler
__global__ void __launch_bounds__(kWarpSideThreads, 8) apply_warp_kernel(const ApplyLaunch p) {
  __shared__ int s_go;
  const int tid = threadIdx.x, lane = tid & 31;
  // At most side_limit CTAs of this launch per SM: the block scheduler puts several on an SM that happens to be
  // empty, and that SM could then not take its two fbank CTAs.  The surplus CTAs leave at once; tiles are
  // claimed dynamically, so the ones that stay do all the work (the first CTA to arrive on an SM always
  // stays, so every launch has workers).
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (tid == 0) {
    int go = 1;
    if (p.side_limit > 0) {
      go = atomicAdd(p.side_occ + smid, 1) < p.side_limit ? 1 : 0;
      if (!go) atomicSub(p.side_occ + smid, 1);
    }
    s_go = go;
  }
  __syncthreads();
  const bool go = s_go != 0;
  pdl_wait();
  const int n_masks = p.masks != nullptr ? p.n_fmask + p.n_tmask : 0;
  const bool after = p.cmvn_after != 0;
  const float4 pad = make_float4(p.pad_value, p.pad_value, p.pad_value, p.pad_value);
  if (go) {
    // claim pipeline (lane 0): `cur` is processed now, `nxt` was claimed one iteration ago
    int cur = 0, nxt = 0;
    if (lane == 0) {
      cur = atomicAdd(p.side_sched, 2);
      nxt = cur + 1;
    }
    cur = __shfl_sync(0xffffffffu, cur, 0);
    nxt = __shfl_sync(0xffffffffu, nxt, 0);
#pragma unroll 1
    while (cur < p.n_tiles) {
      int nn = 0;
      if (lane == 0) nn = atomicAdd(p.side_sched, 1);  // consumed at the end of this iteration
      const TileDesc td = p.tiles[p.n_tiles - 1 - cur];  // newest tiles first
      const int nf = td.nf, rows = td.rows;
      float4* t4 = reinterpret_cast<float4*>(p.out + td.out_row0 * (long long)kMel);
      const long long so = p.shared_stats ? 0 : (long long)td.utt * kMel;
      const float4* mean4 = reinterpret_cast<const float4*>(p.mean + so);
      const float4* istd4 = reinterpret_cast<const float4*>(p.istd + so);
      // SpecAugment: bit c of cmask = mel bin c is inside a frequency mask (80 bits), bit f of tmask = row f of
      // the tile is inside a time mask
      unsigned cmask0 = 0, cmask1 = 0, cmask2 = 0, tmask = 0;
      float mv = 0.f;
      if (n_masks > 0 && nf > 0) {
        const int* mk = p.masks + (long long)td.utt * n_masks * 2;
#pragma unroll 1
        for (int m = 0; m < n_masks; ++m) {
          const int m0 = __ldg(mk + 2 * m);
          const int w = __ldg(mk + 2 * m + 1);
          const int base = m < p.n_fmask ? 0 : td.frame0;
          const int lo = max(m0 - base, 0), hi = min(m0 + w - base, m < p.n_fmask ? kMel : kTileFrames);
          if (hi > lo) {
            // bits lo .. hi - 1 of a 96-bit field
            const unsigned long long lowpart = (hi - lo >= 64 ? ~0ull : ((1ull << (hi - lo)) - 1ull));
            if (m < p.n_fmask) {
              // at most 80 bits: split the run over three words
#pragma unroll
              for (int wd = 0; wd < 3; ++wd) {
                const int l2 = max(lo - 32 * wd, 0), h2 = min(hi - 32 * wd, 32);
                if (h2 > l2) {
                  const unsigned bits = (h2 - l2 >= 32 ? 0xffffffffu : ((1u << (h2 - l2)) - 1u) << l2);
                  if (wd == 0) cmask0 |= bits;
                  if (wd == 1) cmask1 |= bits;
                  if (wd == 2) cmask2 |= bits;
                }
              }
            } else {
              tmask |= (unsigned)(lowpart << lo);
            }
          }
        }
        mv = __ldg(p.mask_value + td.utt);
      }
      // The tile is rows x 20 float4, contiguous.  Element e + 160 m (m = 0..3) of lane-element e = lane + 32 k
      // (k = 0..4) lies 8 rows below element e in the SAME column: the four loads of one k share mean / 1/std
      // and the column's mask bits, and every load / store instruction of the warp covers 512 contiguous bytes.
#pragma unroll 1
      for (int k = 0; k < 5; ++k) {
        const int e = lane + 32 * k;
        const int col = e % (kMel / 4), rk = e / (kMel / 4);
        if (rk >= rows) break;  // (rows of a later k are further down)
        const float4 mu = __ldg(mean4 + col), is = __ldg(istd4 + col);
        const unsigned cword = col < 8 ? cmask0 : (col < 16 ? cmask1 : cmask2);
        const unsigned cm = (cword >> (4 * (col & 7))) & 0xfu;
        float4 x[4];
#pragma unroll
        for (int m = 0; m < 4; ++m)
          if (rk + 8 * m < nf) x[m] = JS2T_SIDE_LOAD(t4 + e + 160 * m);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int f = rk + 8 * m;
          float4 y = pad;
          if (f < nf) {
            float4 v = x[m];
            const unsigned mm = ((tmask >> f) & 1u) ? 0xfu : cm;
            if (after) {  // SpecAugment saw the raw log-mel; CMVN normalises the filled cells too
              v.x = (mm & 1u) ? mv : v.x;
              v.y = (mm & 2u) ? mv : v.y;
              v.z = (mm & 4u) ? mv : v.z;
              v.w = (mm & 8u) ? mv : v.w;
            }
            y = make_float4((v.x - mu.x) * is.x, (v.y - mu.y) * is.y, (v.z - mu.z) * is.z, (v.w - mu.w) * is.w);
            if (!after) {
              y.x = (mm & 1u) ? mv : y.x;
              y.y = (mm & 2u) ? mv : y.y;
              y.z = (mm & 4u) ? mv : y.z;
              y.w = (mm & 8u) ? mv : y.w;
            }
          }
          if (f < rows) JS2T_SIDE_STORE(t4 + e + 160 * m, y);
        }
      }
      cur = nxt;
      nxt = __shfl_sync(0xffffffffu, nn, 0);
    }
  }
  pdl_launch();
  __syncthreads();
  // the last CTA to leave re-arms the claim counter for the next launch; the SM slot is given back
  if (tid == 0) {
    if (go && p.side_limit > 0) atomicSub(p.side_occ + smid, 1);
    if (atomicAdd(p.side_sched + 1, 1) == (int)gridDim.x - 1) {
      p.side_sched[0] = 0;
      p.side_sched[1] = 0;
    }
  }
}

cudaError_t launch_apply_warp(const ApplyLaunch& p, cudaStream_t s) {
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(apply_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, JS2T_SIDE_CARVEOUT);
  }
  const int per_sm = p.side_ctas_per_sm > 0 ? p.side_ctas_per_sm : JS2T_SIDE_CTAS_DEFAULT;
  const int warps = (p.n_tiles + 3) / 4;
  const int grid = warps < per_sm * n_sm ? warps : per_sm * n_sm;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kWarpSideThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (JS2T_PDL && p.pdl) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, apply_warp_kernel, p);
}

}  // namespace js2t
