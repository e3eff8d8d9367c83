// Internal declarations shared by fbank_kernels.cu (device code + launchers) and capi.cu
// (the extern "C" surface declared in include/joeys2t_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace js2t {

// ---- frame geometry at 16 kHz (torchaudio kaldi.py:138-140; the reference never overrides it,
//      joeynmt/helpers_for_audio.py:34-36) -----------------------------------------------------
constexpr int kFrameLen = 400;   // 25 ms
constexpr int kHop = 160;        // 10 ms
constexpr int kFFT = 512;        // round_to_power_of_two
constexpr int kMel = 80;         // num_mel_bins
constexpr int kTileFrames = 32;  // frames per CTA tile
constexpr int kTileSamples = (kTileFrames - 1) * kHop + kFrameLen;  // 5360
constexpr int kStatsPerTile = 2 * kMel;                             // column sum | sum of squares

// One utterance of the ragged batch (device + host layout, 32 bytes).
struct UttDesc {
  long long pcm_byte_off;  // byte offset of sample 0 in the packed PCM buffer, 16-byte aligned
  long long out_row;       // first output row of this utterance (rows are 80 floats)
  int n_samples;           // samples available
  int n_frames;            // frames to emit: min(1+(n-400)/160, max_length truncation)
  int tile_start;          // index of the utterance's first tile
  int flags;               // bit0: PCM is fp32 in [-1,1) (scaled by 2^15 on load), else int16
};

// One tile = up to 32 consecutive frames of one utterance (32 bytes; everything the fbank kernel
// needs, so that a persistent CTA fetches its next work item with a single load).
struct TileDesc {
  long long src_byte_off;  // PCM: byte offset of the tile's first sample; features: of its first row
  long long out_row0;      // first output row of the tile
  int utt;
  int frame0;              // first frame of the tile inside the utterance
  int stats_slot;          // row of the per-tile statistics array: the tile's index in utterance order
                           // (tiles are PROCESSED full ones first, see js2t_plan_create)
  unsigned char nf;        // valid frames in the tile (0 => pure padding tile of the padded layout)
  unsigned char rows;      // output rows the tile owns (>= nf; the rest is padding)
  unsigned char flags;     // bit0: PCM is fp32
  unsigned char pad_;
};
static_assert(sizeof(TileDesc) == 32, "TileDesc is loaded with two 16-byte loads");

// device-side tables owned by a context
struct DeviceTables {
  const float* window_half;  // [400]  0.5 * povey window (the 0.5 folds the real-FFT 1/2 factors)
  const float2* tw256;       // [16*16] W_256^(n2*k1) at [k1*16+n2]
  const float2* tw512;       // [136]  W_512^k, k = 0..128
};

enum EpilogueMode : int {
  kEpiRaw = 0,        // store raw log-mel (+ optional per-tile column statistics)
  kEpiNormKnown = 1,  // statistics known up front (global CMVN): normalise + mask at store
};

struct FbankLaunch {
  const uint8_t* pcm;
  const UttDesc* utts;
  const TileDesc* tiles;
  int n_tiles;
  int* sched;         // [2] dynamic tile scheduler: next tile to hand out | CTAs that have left (self-resetting)
  DeviceTables tab;
  float* out;
  float* tile_stats;  // [n_tiles][160] or nullptr
  // padded (B, Tmax, 80) layout: rows >= n_frames of every utterance are filled with pad_value
  int pad_tmax;       // 0 => ragged layout
  float pad_value;
  int epilogue;       // EpilogueMode
  // kEpiNormKnown: per-bin mean / inverse std shared by all utterances, mask table + value
  const float* g_mean;  // [80]
  const float* g_istd;  // [80]
  const int* masks;     // [n_utts][n_masks][2] (start, width); first n_fmask are frequency masks
  int n_fmask, n_tmask;
  const float* mask_value;  // [n_utts]
  // dither compatibility mode (kaldi.py:179-181): noise[(sum T)][400] added to the frames before DC removal;
  // dither_row0[u] = first row of utterance u in that array.  nullptr = off (the reference's dither = 0).
  const float* dither;
  const long long* dither_row0;
  int grid_limit;  // tuning only: cap on the persistent grid (0 = all resident CTAs)
  int dbg_skip;  // tuning only: bit0 staging, bit1 FFT phase, bit2 mel, bit3 store, bit4 butterflies, bit5 exchange
  unsigned long long* dbg_times;  // [n_tiles][4] globaltimer stamps (debug / tuning only) or nullptr
};

struct FinalizeLaunch {
  const UttDesc* utts;
  int n_utts;
  const float* tile_stats;
  const float* raw;      // raw log-mel (needed only for CMVN-after-SpecAugment)
  int norm_means, norm_vars;
  int cmvn_enabled;      // 0: no CMVN (mean 0, istd 1), statistics still produced
  int cmvn_after;        // 1: SpecAugment is applied before CMVN (cmvn.before == False)
  const float* g_mean;   // [80] shared (global CMVN) statistics; nullptr => per-utterance statistics
  const float* g_istd;
  const int* masks;
  int n_fmask, n_tmask;
  int mask_value_mode;   // 0: mean of the spectrogram SpecAugment sees, 1: constant
  float mask_value_const;
  float* mean;           // [n_utts][80]
  float* istd;           // [n_utts][80]
  float* mask_value;     // [n_utts]
  double* stats_out;     // [n_utts][160] raw per-utterance sum | sumsq in fp64, or nullptr
};

struct ApplyLaunch {
  const UttDesc* utts;
  const TileDesc* tiles;
  int n_tiles;
  float* out;            // in place
  const float* mean;     // [n_utts][80] (or [80] when shared_stats)
  const float* istd;
  int shared_stats;
  const int* masks;
  int n_fmask, n_tmask;
  const float* mask_value;  // [n_utts]
  int cmvn_after;           // 1: fill first, then normalise (cmvn.before == False)
  int pad_tmax;
  float pad_value;
};

cudaError_t upload_mel_weights(const float* wu256, const float* wd256, cudaStream_t s);
void reference_mel_weights(float* wu256, float* wd256);  // the compiled-in two-band weights
int check_mel_weights(const float* wu256, const float* wd256);  // first bin that differs from the compiled-in weights, or -1
cudaError_t launch_fbank(const FbankLaunch& p, cudaStream_t s);
cudaError_t launch_features(const FbankLaunch& p, cudaStream_t s);  // p.pcm = feature rows
cudaError_t launch_finalize(const FinalizeLaunch& p, cudaStream_t s);
cudaError_t launch_apply(const ApplyLaunch& p, cudaStream_t s);
// accum[0..79] += sum, accum[80..159] += sumsq, accum[160] += frames  (fixed order => deterministic)
cudaError_t launch_global_accumulate(const double* utt_stats, const UttDesc* utts, int n_utts,
                                     double* accum, cudaStream_t s);
// mean/istd (float[80] each) from accum[161] with the reference CMVN formula
cudaError_t launch_global_finalize(const double* accum, int norm_means, int norm_vars,
                                   float* mean, float* istd, cudaStream_t s);
cudaError_t launch_fill_value(float* dst, int n, float value, cudaStream_t s);
// 48 kHz -> 16 kHz ingest (scripts/gradio_demo.py:35-45); ws = one int of device scratch
cudaError_t launch_reformat_48k_to_16k(const void* src, int is_f32, long long n_samples, short* dst, int* ws,
                                       cudaStream_t s);
int fbank_smem_bytes();
int fbank_persistent_grid();  // CTAs of the persistent fbank kernel on the current device

}  // namespace js2t
