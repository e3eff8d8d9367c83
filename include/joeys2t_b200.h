/*
 * joeys2t_b200 — C ABI of the B200-native audio front-end (fbank -> CMVN -> SpecAugment).
 *
 * This is the drop-in boundary for JoeyS2T's front-end hot path.  The reference is pure Python; the
 * functions it would bind through ctypes/cffi are listed below with the reference interface each
 * one replaces (paths relative to the reference repository root; "TA:" = the un-vendored PyPI
 * dependency torchaudio/compliance/kaldi.py that holds the arithmetic).  INTEGRATION.md shows the
 * ctypes stub a maintainer would add to joeynmt/helpers_for_audio.py / data_augmentation.py.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types;
 *   - every function returns an int status (JS2T_OK == 0); js2t_last_error() gives the text of the
 *     last failure on the calling thread; no exceptions cross the boundary;
 *   - pointers named *_dev are CUDA device pointers owned by the caller, everything else is host
 *     memory; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - all GPU work is enqueued asynchronously on `stream`; nothing synchronises the whole device except
 *     js2t_ctx_set_tables (once).  js2t_plan_create uploads its descriptors asynchronously on a stream of
 *     the context and orders every stream the plan is later used on behind that upload; js2t_plan_destroy
 *     waits only for the streams the plan was used on;
 *   - every call runs on the context's device and restores the caller's current device before returning;
 *   - there is no CPU fallback: without a usable CUDA device every call fails with JS2T_ERR_CUDA.
 *
 * Geometry is the reference's fixed one: 16 kHz, 25 ms / 10 ms frames (400 / 160 samples),
 * 512-point FFT, 80 mel bins, dither 0 (joeynmt/helpers_for_audio.py:34-36 passes only
 * num_mel_bins and sample_frequency; TA:514-541 defaults).
 */
#ifndef JOEYS2T_B200_H_
#define JOEYS2T_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JS2T_VERSION 100 /* 0.1.0 */

enum {
  JS2T_OK = 0,
  JS2T_ERR_INVALID = 1,     /* bad argument */
  JS2T_ERR_CUDA = 2,        /* CUDA runtime error (text in js2t_last_error) */
  JS2T_ERR_SHORT_INPUT = 3, /* an utterance is shorter than one 400-sample window (TA:142-144) */
  JS2T_ERR_TABLES = 4,      /* uploaded mel bank does not match the compiled-in structure */
  JS2T_ERR_NCCL = 5,        /* NCCL unavailable or failed */
  JS2T_ERR_STATE = 6        /* call sequence error (e.g. normalise before statistics exist) */
};

/* CMVN modes of a plan */
enum {
  JS2T_CMVN_NONE = 0,      /* raw log-mel (extract_fbank_features only) */
  JS2T_CMVN_UTTERANCE = 1, /* joeynmt/data_augmentation.py:83-115 CMVN, per utterance */
  JS2T_CMVN_GLOBAL = 2,    /* corpus-level statistics supplied with js2t_plan_set_global_stats
                              or js2t_global_stats_finalize (extension) */
  JS2T_CMVN_STATS_ONLY = 3 /* raw log-mel + per-utterance fp64 sum / sum-of-squares */
};

/* output layouts */
enum {
  JS2T_LAYOUT_RAGGED = 0, /* (sum_u T_u, 80): utterances back to back */
  JS2T_LAYOUT_PADDED = 1  /* (B, Tmax, 80) filled with pad_value, joeynmt/helpers_for_audio.py:130-170 */
};

enum { JS2T_MASK_VALUE_MEAN = 0, JS2T_MASK_VALUE_CONST = 1 };

typedef struct js2t_ctx js2t_ctx;   /* per device: tables */
typedef struct js2t_plan js2t_plan; /* per batch geometry: descriptors + workspace */

int js2t_version(void);
const char* js2t_last_error(void);

/* Frame count for n_samples at 16 kHz with snip_edges=True: 1 + (n - 400) / 160, 0 if n < 400.
 * Replaces the shape logic of TA:44-83 (_get_strided); cf. joeynmt/helpers_for_audio.py:93-96. */
int64_t js2t_num_frames(int64_t n_samples);

/* ---- context ------------------------------------------------------------------------------- */
int js2t_ctx_create(int device, js2t_ctx** out);
int js2t_ctx_destroy(js2t_ctx* ctx);

/* Upload the povey window (400 floats, TA:98-100) and the dense mel bank (80 x 256 floats,
 * row-major, TA:436-511).  The bank must have the two-adjacent-filters-per-bin structure of the
 * reference configuration, otherwise JS2T_ERR_TABLES.  Must be called once before any execute. */
int js2t_ctx_set_tables(js2t_ctx* ctx, const float* window400, const float* mel80x256);
/* The bank the kernels were compiled with (80 x 256 floats, row-major): torchaudio's
 * get_mel_banks(80, 512, 16000, 20, 0) (TA:436-511), bit for bit.  For hosts whose own float32 `log`
 * rounds differently from the build machine's: they can compare their bank against this one and decide,
 * instead of being locked out by js2t_ctx_set_tables' bit-exact check. */
int js2t_reference_mel_bank(float* mel80x256_out);

/* ---- plan: one ragged batch ------------------------------------------------------------------
 * n_utts utterances packed in one PCM buffer.  pcm_byte_off[u] (multiple of 16) is where utterance
 * u starts, n_samples[u] its length, is_f32[u] != 0 means float32 samples in [-1, 1) (the loader's
 * output, scaled by 2^15 on the fly: joeynmt/helpers_for_audio.py:54), else int16 PCM.
 * max_frames[u] > 0 truncates to the first max_frames[u] frames *before* CMVN
 * (joeynmt/tokenizers.py:477-484, evaluation branch); pass NULL for no truncation.
 * Fails with JS2T_ERR_SHORT_INPUT if any utterance has fewer than 400 samples. */
int js2t_plan_create(js2t_ctx* ctx, int n_utts, const int64_t* pcm_byte_off,
                     const int64_t* n_samples, const uint8_t* is_f32, const int32_t* max_frames,
                     int layout, int pad_tmax, float pad_value, js2t_plan** out);
/* Same, for pre-extracted features (the .npy / npy-in-zip branch of get_features,
 * joeynmt/helpers_for_audio.py:100-127): the input of js2t_features_execute is (sum_u n_frames[u], 80)
 * float32, utterances back to back.  Lets CMVN / SpecAugment / pad_features run on their own. */
int js2t_plan_create_features(js2t_ctx* ctx, int n_utts, const int32_t* n_frames,
                              const int32_t* max_frames, int layout, int pad_tmax, float pad_value,
                              js2t_plan** out);
int js2t_plan_destroy(js2t_plan* plan);
/* Same without waiting: the caller guarantees that everything it enqueued with this plan has completed
 * (for example it waited on an event recorded behind the last execute).  For hosts that retire plans
 * behind events while newer batches are already in flight on the same stream. */
int js2t_plan_destroy_completed(js2t_plan* plan);

int64_t js2t_plan_total_frames(const js2t_plan* plan); /* sum of emitted frames */
int64_t js2t_plan_out_rows(const js2t_plan* plan);     /* rows of 80 floats the output needs */
int js2t_plan_get_frames(const js2t_plan* plan, int32_t* n_frames_out);   /* [n_utts] */
int js2t_plan_get_out_rows(const js2t_plan* plan, int64_t* first_row_out); /* [n_utts] */

/* CMVN(norm_means, norm_vars, before): joeynmt/data_augmentation.py:89-109 and the call order of
 * joeynmt/tokenizers.py:488-493 (before != 0: CMVN then SpecAugment; else SpecAugment then CMVN). */
int js2t_plan_set_cmvn(js2t_plan* plan, int mode, int norm_means, int norm_vars, int before);

/* Global CMVN statistics known up front: per-bin mean and 1/std (80 doubles each, host). */
int js2t_plan_set_global_stats(js2t_plan* plan, const double* mean80, const double* istd80,
                               void* stream);

/* SpecAugment fill (joeynmt/data_augmentation.py:54-70) for masks drawn on the host:
 * table is int32 [n_utts][n_fmask + n_tmask][2] = (start, width), frequency masks first; width 0
 * is a no-op (the reference still consumes the RNG draws for it).  value_mode MEAN reproduces
 * `mask_value = spectrogram.mean()` (:45-46), CONST uses value_const.  table == NULL clears. */
int js2t_plan_set_masks(js2t_plan* plan, int n_fmask, int n_tmask, const int32_t* table,
                        int value_mode, float value_const, void* stream);

/* Dither COMPATIBILITY / TEST mode (TA:179-181: strided_input + randn(T, 400) * dither).  The reference's
 * call site never enables dither (joeynmt/helpers_for_audio.py:34-36 -> 0.0, the default of every other
 * entry point here); this reproduces torchaudio's behaviour for callers who do, with the noise drawn on the
 * host so that results are reproducible: noise_dev is float32 (js2t_plan_total_frames(), 400) on the device,
 * already multiplied by the dither constant, frames of the batch in utterance order; it is added to each
 * frame's samples (int16-range values) before DC removal.  The pointer is borrowed — it must stay valid
 * for every later execute; NULL switches the mode off.  Slower than the default path (the two frames a
 * lane processes no longer share their samples); not on the benchmarked path. */
int js2t_plan_set_dither(js2t_plan* plan, const float* noise_dev);

/* ---- the hot path ---------------------------------------------------------------------------
 * pcm_dev -> out_dev according to the plan (fbank [-> CMVN] [-> SpecAugment], padded or ragged).
 * Replaces the per-utterance chain
 *   joeynmt/helpers_for_audio.py:41-68 extract_fbank_features -> TA:514-645 fbank
 *   joeynmt/data_augmentation.py:96-109 CMVN.__call__, :38-73 SpecAugment.__call__
 *   joeynmt/helpers_for_audio.py:130-170 pad_features
 * for a whole batch.  out_dev must hold js2t_plan_out_rows() * 80 floats. */
int js2t_fbank_execute(js2t_plan* plan, const void* pcm_dev, float* out_dev, void* stream);

/* feats_dev -> out_dev: [CMVN] [SpecAugment] [padding] on pre-extracted features
 * (joeynmt/data_augmentation.py:96-109, :38-73; joeynmt/helpers_for_audio.py:130-170).
 * feats_dev == out_dev (ragged layout) is allowed. */
int js2t_features_execute(js2t_plan* plan, const float* feats_dev, float* out_dev, void* stream);

/* Instrumentation for bench.py: bracket the dominant (fbank) kernel of every execute with CUDA
 * events on the launching stream (ring of n_slots; 0 disables) and read the durations back. */
int js2t_plan_enable_profiling(js2t_plan* plan, int n_slots);
int js2t_plan_kernel_times_ms(js2t_plan* plan, float* ms_out, int n, int* n_written);

/* Tuning switches: "max_ctas" caps the persistent grid, "debug_times" records per-tile time stamps,
 * "debug_skip" skips kernel phases in -DJS2T_DBG=1 builds.  Unknown names are an error. */
int js2t_plan_set_option(js2t_plan* plan, const char* name, int value);
/* With option "debug_times" = 1: per-tile %globaltimer stamps [n_tiles][4] (tile start, stored,
 * published, normalised-older-tile) of the last execute, copied to host memory (synchronous). */
int js2t_plan_debug_times(const js2t_plan* plan, unsigned long long* host_out, int64_t n_values);

/* Device pointer to the per-utterance fp64 statistics [n_utts][160] (sum | sum of squares of the
 * raw log-mel) produced by the last execute in UTTERANCE / STATS_ONLY / masked modes. */
int js2t_plan_utt_stats(const js2t_plan* plan, const double** stats_dev);
/* Same, copied (device to device, on `stream`) into a caller-owned buffer of n_utts * 160 doubles. */
int js2t_plan_copy_utt_stats(const js2t_plan* plan, double* dst_dev, void* stream);

/* ---- global CMVN (multi-GPU) ----------------------------------------------------------------
 * accum_dev: 161 doubles on the device = per-bin sum[80] | sumsq[80] | frame count.
 * accumulate adds this plan's statistics (fixed summation order, deterministic);
 * allreduce sums accum_dev over all ranks with one ncclAllReduce (comm is a ncclComm_t);
 * finalize turns accum_dev into the plan's mean / inverse std with the reference formula
 * (var = sumsq/n - mean^2, std = sqrt(max(var, 1e-10)), joeynmt/data_augmentation.py:98-105);
 * normalize applies them (and the SpecAugment fill) in place to raw log-mel in out_dev. */
int js2t_global_stats_accumulate(js2t_plan* plan, double* accum_dev, void* stream);
int js2t_global_stats_allreduce(void* nccl_comm, double* accum_dev, void* stream);
/* The communicator for it.  The reference reduces with torch.distributed (joeynmt/helpers_for_ddp.py:
 * 157-174 ddp_reduce: dist.all_reduce(SUM)) over ranks that shard the data as indices[rank::world]
 * (helpers_for_ddp.py:319); a C host has no process group, so the ABI carries the three NCCL calls
 * it needs: rank 0 draws a 128-byte unique id and ships it to the other ranks by whatever channel the
 * host has (torch.distributed broadcast, MPI, a file), then every rank creates its communicator
 * (collective call).  libnccl is bound at run time from the host process. */
int js2t_nccl_unique_id(void* id128_out);
int js2t_nccl_comm_create(const void* id128, int world_size, int rank, int device, void** comm_out);
int js2t_nccl_comm_destroy(void* nccl_comm);
int js2t_global_stats_finalize(js2t_plan* plan, const double* accum_dev, void* stream);
int js2t_normalize_execute(js2t_plan* plan, float* out_dev, void* stream);
/* The plan's global statistics as the kernels use them: mean[80] | 1/std[80] in float32, copied device to
 * device on `stream` (for checks across ranks: they must be bit-identical everywhere). */
int js2t_plan_copy_global_stats(const js2t_plan* plan, float* dst_dev, void* stream);

/* ---- host-side packing ------------------------------------------------------------------------
 * Gathers the per-utterance waveform arrays of one batch (what the reference's per-item API hands over:
 * joeynmt/helpers_for_audio.py:41-47 `waveform`, one array per call) into the packed staging buffer the
 * plan describes: utterance u's n_bytes[u] bytes go to dst + dst_byte_off[u].  Pure host code, spread
 * over a small persistent pool of threads (n_threads <= 0: the pool's default; 1: the calling thread
 * only).  The pool holds half of the cores the process may run on, at most 8 threads with the caller (the
 * gather is bound by host memory bandwidth: 102 MB of cold PCM take 3.0 / 1.85 / 1.65 ms on 4 / 8 / 16 threads
 * of the measured host); the environment variable JS2T_COPY_THREADS, read once, overrides the size.  dst should
 * be pinned memory so that the following H2D copy is asynchronous. */
int js2t_pack_pcm(int n_utts, const void* const* src, const int64_t* n_bytes, const int64_t* dst_byte_off,
                  void* dst, int64_t dst_capacity, int n_threads);

/* ---- one call per batch, utterances in HOST memory ----------------------------------------------------
 * What a per-batch caller of the reference's data path needs (collate_fn -> Batch: joeynmt/datasets.py:221-225,
 * batch.py:114-121; SpeechProcessor per item: tokenizers.py:458-494): n_utts waveforms in host memory in,
 * features on the device out.  The call gathers the utterances into a pinned staging slot owned by the context
 * (js2t_pack_pcm's copy pool), lays the batch's descriptors, mask table and global statistics out next to the
 * PCM, uploads all of it with ONE transfer on `stream` and enqueues the kernels behind it; it returns without
 * waiting for the device.  The waveform arrays may be reused as soon as the call returns.
 *   pcm_host[u]   n_samples[u] samples: int16, or float32 in [-1, 1) where is_f32[u] != 0 (is_f32 may be NULL)
 *   opts          layout / CMVN / SpecAugment of the batch, same meaning as js2t_plan_create, js2t_plan_set_cmvn,
 *                 js2t_plan_set_global_stats (GLOBAL: both statistics arrays required) and js2t_plan_set_masks
 *   out_dev       float32 rows of 80 on the device, out_capacity_rows of them: sum of the utterances' frames
 *                 (ragged) or n_utts * Tmax (padded; Tmax = pad_tmax, or the longest utterance when pad_tmax = 0)
 *   plan_out      NULL: the context keeps the batch's plan and destroys it once the event behind its last kernel
 *                 has completed; otherwise the caller owns it (statistics read-back, js2t_plan_destroy)
 * Errors as the calls it combines (too-short utterance: JS2T_ERR_SHORT_INPUT, ...); nothing is enqueued then. */
typedef struct js2t_batch_opts {
  int layout;                  /* JS2T_LAYOUT_RAGGED | JS2T_LAYOUT_PADDED */
  int pad_tmax;                /* padded layout: rows per utterance, 0 = the longest utterance */
  float pad_value;             /* padded layout: fill value (the reference pads with float(pad_index) = 1.0) */
  int cmvn_mode;               /* JS2T_CMVN_NONE | JS2T_CMVN_UTTERANCE | JS2T_CMVN_GLOBAL */
  int norm_means, norm_vars, before;
  const double* global_mean80; /* JS2T_CMVN_GLOBAL: mean[80] and 1/std[80] */
  const double* global_istd80;
  int n_fmask, n_tmask;        /* SpecAugment: masks per utterance */
  const int32_t* mask_table;   /* [n_utts][n_fmask + n_tmask][2] = (start, width), frequency masks first; NULL = none */
  int mask_value_mode;         /* JS2T_MASK_VALUE_MEAN | JS2T_MASK_VALUE_CONST */
  float mask_value_const;
  const int32_t* max_frames;   /* [n_utts] truncation (tokenizers.py:480-481), entries <= 0 = none; NULL = none */
} js2t_batch_opts;
int js2t_batch_fbank(js2t_ctx* ctx, int n_utts, const void* const* pcm_host, const int64_t* n_samples,
                     const uint8_t* is_f32, const js2t_batch_opts* opts, float* out_dev, int64_t out_capacity_rows,
                     void* stream, js2t_plan** plan_out);

/* ---- SpecAugment draws of a whole batch (host code) --------------------------------------------------------
 * The reference draws its masks per item from the global np.random stream (joeynmt/data_augmentation.py:48-70:
 * f = randint(0, F), f0 = randint(0, num_freqs - f) per frequency mask, then t = randint(0, min(T_max,
 * floor(num_frames * p))), t0 = randint(0, num_frames - t) per time mask), and a seeded run must mask the same
 * cells.  numpy's legacy generator serves randint(0, hi) by masked rejection on the 32-bit outputs of its bit
 * generator; this function performs the draws of a whole batch, in utterance order, on the caller's generator
 * (numpy's own: BitGenerator.ctypes.next_uint32 / .state), so values AND stream position are exactly what the
 * reference's loop gives — 128 Python-level randint calls per 16-utterance batch cost more host time than the
 * batch's kernels.  time_mask_p < 0 or freq_mask_f > num_freqs etc. are left to the caller's per-item route.
 *   table_out     int32 [n_utts][freq_mask_n + time_mask_n][2] = (start, width); utterances the reference leaves
 *                 untouched (no frames) and width-0 masks are all-zero rows */
typedef uint32_t (*js2t_next_uint32_fn)(void* rng_state);
int js2t_specaug_replay(js2t_next_uint32_fn next_uint32, void* rng_state, int n_utts, const int32_t* n_frames,
                        int num_freqs, int freq_mask_n, int freq_mask_f, int time_mask_n, int time_mask_t,
                        double time_mask_p, int32_t* table_out);

/* ---- PCM ingest: 48 kHz -> 16 kHz -------------------------------------------
 * Replaces scripts/gradio_demo.py:35-45 reformat_freq for sr == 48000:
 *     y = ((y / max(np.max(y), 1)) * 32767).reshape((-1, 3)).mean(axis=1).astype("int16")
 * i.e. peak-normalise by the (signed) maximum, average blocks of three samples, truncate to int16.
 * Same arithmetic, same order and precision as numpy evaluates it: float64 for int16 input,
 * float32 for float32 input — results are bit-identical to the reference expression.
 *   src_dev       n_samples samples (int16 when is_f32 == 0, float32 otherwise); n_samples % 3 == 0
 *                 (the reference's reshape raises otherwise -> JS2T_ERR_INVALID)
 *   dst_dev       n_samples / 3 int16 samples
 *   workspace_dev 8 bytes of device scratch (the maximum), owned by the caller
 * Asynchronous on `stream`. */
int js2t_reformat_48k_to_16k(js2t_ctx* ctx, const void* src_dev, int is_f32, int64_t n_samples,
                             int16_t* dst_dev, void* workspace_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JOEYS2T_B200_H_ */
