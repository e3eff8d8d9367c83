// extern "C" surface of libjoeys2t_b200.so — see include/joeys2t_b200.h for the contract and the
// reference interfaces each entry point replaces.
#include "../../include/joeys2t_b200.h"

#include <dlfcn.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "js2t_internal.h"
#include "mel_structure.inc"

using namespace js2t;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define JS2T_CUDA(expr)                                                                    \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return fail(JS2T_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                  __FILE__, __LINE__);                                                     \
  } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Entry points run on the context's device and leave the caller's current device as they found it
// (a caller driving several GPUs from one thread must not have its device changed under it).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err;
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
    else if (err == cudaSuccess) prev = -1;  // nothing to restore
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace

struct js2t_ctx {
  int device = 0;
  float* d_tables = nullptr;  // window_half[400] | tw256[256 x float2] | tw512[136 x float2]
  bool tables_set = false;
  // plan descriptors are uploaded on this (non-blocking) stream; every plan records an event behind its
  // upload and every execute stream waits for it, so the upload is formally ordered before the kernels
  // on whatever stream the caller launches them and plan creation never synchronises the device
  cudaStream_t upload_stream = nullptr;
  // Device-buffer pool for plan workspaces: the per-item / per-batch callers of the reference API create
  // and destroy one plan per call, and cudaMalloc + cudaFree were a third of such a call.
  std::mutex pool_mu;
  std::vector<std::pair<size_t, void*>> pool;  // (capacity, pointer) of idle buffers
  size_t pool_bytes = 0;
  // js2t_batch_fbank: pinned staging slots (PCM + descriptors of one batch, uploaded with one transfer; a slot
  // is free again once the event behind its transfer has completed) and the plans of earlier calls that the
  // context destroys once the event behind their last kernel has completed
  struct StagingSlot {
    char* host = nullptr;
    size_t cap = 0;
    cudaEvent_t ev = nullptr;
    // free -> filling (a call owns it: gather + upload being issued) -> pending (event recorded behind the upload;
    // free again once it has completed).  Only the owner touches a filling slot; only recorded events are waited on.
    bool filling = false, pending = false;
  };
  std::mutex batch_mu;
  std::vector<StagingSlot> slots;
  std::vector<std::pair<cudaEvent_t, js2t_plan*>> retired;
  std::vector<cudaEvent_t> spare_events;
};

namespace {

constexpr size_t kPoolMaxEntries = 16;
constexpr size_t kPoolMaxBytes = size_t(1) << 30;

// smallest idle buffer that fits without wasting more than 4x, else a fresh allocation (64 KB granules)
cudaError_t pool_alloc(js2t_ctx* c, size_t bytes, void** out, size_t* cap) {
  {
    std::lock_guard<std::mutex> lk(c->pool_mu);
    size_t best = c->pool.size();
    for (size_t i = 0; i < c->pool.size(); ++i)
      if (c->pool[i].first >= bytes && c->pool[i].first <= 4 * bytes + (1 << 20) &&
          (best == c->pool.size() || c->pool[i].first < c->pool[best].first))
        best = i;
    if (best != c->pool.size()) {
      *cap = c->pool[best].first;
      *out = c->pool[best].second;
      c->pool_bytes -= *cap;
      c->pool.erase(c->pool.begin() + (long)best);
      return cudaSuccess;
    }
  }
  *cap = align_up(bytes, 64 * 1024);
  return cudaMalloc(out, *cap);
}

// The caller guarantees (as with cudaFree) that nothing in flight uses the buffer any more.
void pool_free(js2t_ctx* c, void* p, size_t cap) {
  if (p == nullptr) return;
  {
    std::lock_guard<std::mutex> lk(c->pool_mu);
    if (c->pool.size() < kPoolMaxEntries && c->pool_bytes + cap <= kPoolMaxBytes) {
      c->pool.emplace_back(cap, p);
      c->pool_bytes += cap;
      return;
    }
  }
  cudaFree(p);
}

}  // namespace

struct js2t_plan {
  js2t_ctx* ctx = nullptr;
  int n_utts = 0, n_tiles = 0;
  int layout = JS2T_LAYOUT_RAGGED, pad_tmax = 0;
  float pad_value = 1.0f;
  long long total_frames = 0, out_rows = 0;
  std::vector<UttDesc> h_utts;
  // device workspace (one allocation)
  void* d_ws = nullptr;
  size_t ws_cap = 0;
  UttDesc* d_utts = nullptr;
  TileDesc* d_tiles = nullptr;
  float* d_tile_stats = nullptr;
  float* d_mean = nullptr;
  float* d_istd = nullptr;
  float* d_mask_value = nullptr;
  float* d_gmean = nullptr;
  float* d_gistd = nullptr;
  double* d_utt_stats = nullptr;
  int* d_sched = nullptr;    // [2] tile scheduler counters of the persistent kernel (self-resetting)
  long long* d_frame_row0 = nullptr;  // [n_utts] frames emitted before utterance u (rows of the dither noise array)
  const float* dither = nullptr;      // compatibility mode: caller-owned noise (total_frames, 400), or NULL
  int max_utt_tiles = 0;
  int* d_masks = nullptr;
  size_t masks_cap = 0;
  bool masks_in_ws = false;  // js2t_batch_fbank: the mask table lives in the workspace (uploaded with the batch)
  const uint8_t* d_pcm = nullptr;  // js2t_batch_fbank: the batch's PCM inside the workspace
  // configuration
  int cmvn_mode = JS2T_CMVN_NONE, norm_means = 1, norm_vars = 1, before = 1;
  int n_fmask = 0, n_tmask = 0, mask_value_mode = JS2T_MASK_VALUE_MEAN;
  float mask_value_const = 0.f;
  bool has_masks = false, global_stats_set = false, stats_valid = false;
  bool feature_input = false;  // rows of 80 floats instead of PCM (js2t_plan_create_features)
  int grid_limit = 0;          // tuning only: option "max_ctas"
  int dbg_skip = 0;            // tuning only: phases of the fbank kernel to skip (results are wrong)
  unsigned long long* d_dbg = nullptr;  // [n_tiles][4] debug time stamps (option "debug_times")
  // optional instrumentation: CUDA events around the fbank kernel of each execute (ring of slots)
  std::vector<cudaEvent_t> prof_ev;  // 2 per slot
  long long prof_calls = 0;
  cudaEvent_t ready_ev = nullptr;  // recorded behind the descriptor upload (ctx->upload_stream)
  std::vector<cudaStream_t> streams;  // streams the plan has enqueued work on (normally one)
};

namespace {

// Everything the plan enqueues on `stream` comes after its descriptors have landed: the first time a
// stream is used it waits for ready_ev; afterwards stream order guarantees it.  (Nothing is inserted
// between the kernels of consecutive steps, which would undo their programmatic dependent launch.)
cudaError_t plan_begin(js2t_plan* plan, cudaStream_t stream) {
  for (cudaStream_t s : plan->streams)
    if (s == stream) return cudaSuccess;
  plan->streams.push_back(stream);
  return cudaStreamWaitEvent(stream, plan->ready_ev, 0);
}
// Wait (host side) until nothing in flight reads the plan's workspace any more: the streams the plan was
// used on, not the whole device — other streams keep running.  A stream the caller has destroyed in the
// meantime had to be drained for that, so its error is ignored.
void plan_quiesce(js2t_plan* plan) {
  if (plan->ready_ev) cudaEventSynchronize(plan->ready_ev);
  for (cudaStream_t s : plan->streams)
    if (cudaStreamSynchronize(s) != cudaSuccess) cudaGetLastError();
}

}  // namespace

extern "C" {

int js2t_version(void) { return JS2T_VERSION; }
const char* js2t_last_error(void) { return g_err; }

int64_t js2t_num_frames(int64_t n_samples) {
  return n_samples < kFrameLen ? 0 : 1 + (n_samples - kFrameLen) / kHop;
}

// ---------------------------------------------------------------------------------------------
int js2t_ctx_create(int device, js2t_ctx** out) {
  if (out == nullptr) return fail(JS2T_ERR_INVALID, "js2t_ctx_create: out is NULL");
  int n_dev = 0;
  JS2T_CUDA(cudaGetDeviceCount(&n_dev));
  if (device < 0 || device >= n_dev)
    return fail(JS2T_ERR_CUDA, "js2t_ctx_create: device %d not available (%d CUDA devices)", device, n_dev);
  cudaDeviceProp prop;
  JS2T_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(JS2T_ERR_CUDA, "js2t_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
  DeviceGuard guard(device);
  JS2T_CUDA(guard.err);
  js2t_ctx* c = new (std::nothrow) js2t_ctx();
  if (c == nullptr) return fail(JS2T_ERR_INVALID, "out of host memory");
  c->device = device;
  const size_t bytes = (400 + 2 * 256 + 2 * 136) * sizeof(float);
  cudaError_t e = cudaMalloc(&c->d_tables, bytes);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    if (c->d_tables) cudaFree(c->d_tables);
    delete c;
    return fail(JS2T_ERR_CUDA, "context allocation failed: %s", cudaGetErrorString(e));
  }
  *out = c;
  return JS2T_OK;
}

int js2t_ctx_destroy(js2t_ctx* ctx) {
  if (ctx == nullptr) return JS2T_OK;
  DeviceGuard guard(ctx->device);
  if (ctx->upload_stream) {
    cudaStreamSynchronize(ctx->upload_stream);
    cudaStreamDestroy(ctx->upload_stream);
  }
  for (auto& r : ctx->retired) {  // plans of js2t_batch_fbank calls the context still owns
    cudaEventSynchronize(r.first);
    cudaEventDestroy(r.first);
    js2t_plan_destroy_completed(r.second);
  }
  for (cudaEvent_t ev : ctx->spare_events) cudaEventDestroy(ev);
  for (auto& sl : ctx->slots) {
    if (sl.ev) {
      cudaEventSynchronize(sl.ev);
      cudaEventDestroy(sl.ev);
    }
    if (sl.host) cudaFreeHost(sl.host);
  }
  if (ctx->d_tables) cudaFree(ctx->d_tables);
  for (auto& b : ctx->pool) cudaFree(b.second);
  delete ctx;
  return JS2T_OK;
}

int js2t_reference_mel_bank(float* mel80x256_out) {
  if (mel80x256_out == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  float wu[256], wd[256];
  reference_mel_weights(wu, wd);
  memset(mel80x256_out, 0, sizeof(float) * 80 * 256);
  for (int s = 0; s < JS2T_MEL_NUM_SEGMENTS; ++s)
    for (int k = kMelSegLo[s]; k <= kMelSegHi[s]; ++k) {
      if (s < 80) mel80x256_out[s * 256 + k] = wu[k];
      if (s >= 1) mel80x256_out[(s - 1) * 256 + k] = wd[k];
    }
  return JS2T_OK;
}

int js2t_ctx_set_tables(js2t_ctx* ctx, const float* window400, const float* mel80x256) {
  if (ctx == nullptr || window400 == nullptr || mel80x256 == nullptr)
    return fail(JS2T_ERR_INVALID, "js2t_ctx_set_tables: NULL argument");
  // ---- mel bank -> two-band form, validated against the compiled-in structure ----------------
  float wu[256], wd[256];
  memset(wu, 0, sizeof(wu));
  memset(wd, 0, sizeof(wd));
  std::vector<float> dense(80 * 256, 0.f);
  for (int s = 0; s < JS2T_MEL_NUM_SEGMENTS; ++s) {
    for (int k = kMelSegLo[s]; k <= kMelSegHi[s]; ++k) {
      if (s < 80) { wu[k] = mel80x256[s * 256 + k]; dense[s * 256 + k] = wu[k]; }
      if (s >= 1) { wd[k] = mel80x256[(s - 1) * 256 + k]; dense[(s - 1) * 256 + k] = wd[k]; }
    }
  }
  for (int i = 0; i < 80 * 256; ++i) {
    if (dense[i] != mel80x256[i])
      return fail(JS2T_ERR_TABLES,
                  "mel bank entry (filter %d, bin %d) = %g is outside the compiled-in two-band structure "
                  "(16 kHz, 512-point FFT, 80 bins, 20 Hz..Nyquist)", i / 256, i % 256, (double)mel80x256[i]);
  }
  {
    const int bad = check_mel_weights(wu, wd);
    if (bad >= 0)
      return fail(JS2T_ERR_TABLES,
                  "mel bank weight of FFT bin %d differs from the compiled-in reference bank "
                  "(torchaudio get_mel_banks(80, 512, 16000, 20, 0)); rebuild with gen_mel_structure.py", bad);
  }
  // ---- window (x 0.5: folds the 1/2 of the real-FFT split; exact power-of-two scaling) --------
  std::vector<float> host(400 + 2 * 256 + 2 * 136);
  if (window400[0] != 0.f)
    return fail(JS2T_ERR_TABLES, "window[0] must be 0 (povey); the staging of x[j-1] relies on it");
  for (int i = 0; i < 400; ++i) host[i] = 0.5f * window400[i];
  const double two_pi = 6.283185307179586476925286766559;
  float* tw256 = host.data() + 400;
  for (int k1 = 0; k1 < 16; ++k1)
    for (int n2 = 0; n2 < 16; ++n2) {
      const double a = -two_pi * (double)(n2 * k1) / 256.0;
      tw256[2 * (k1 * 16 + n2)] = (float)cos(a);
      tw256[2 * (k1 * 16 + n2) + 1] = (float)sin(a);
    }
  float* tw512 = tw256 + 2 * 256;
  for (int k = 0; k < 136; ++k) {
    const double a = -two_pi * (double)k / 512.0;
    tw512[2 * k] = (float)cos(a);
    tw512[2 * k + 1] = (float)sin(a);
  }
  DeviceGuard guard(ctx->device);
  JS2T_CUDA(guard.err);
  // once per context; synchronous on purpose: the tables are complete before any stream can use them
  JS2T_CUDA(cudaMemcpy(ctx->d_tables, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
  JS2T_CUDA(upload_mel_weights(wu, wd, 0));
  JS2T_CUDA(cudaDeviceSynchronize());
  ctx->tables_set = true;
  return JS2T_OK;
}

// ---------------------------------------------------------------------------------------------
// js2t_batch_fbank: everything the device needs from the host for one batch sits at the head of the plan's
// workspace — descriptors, scheduler counters, global statistics, mask table, PCM — so that ONE transfer from a
// pinned staging slot uploads it; plan_create_common then only lays the workspace out and hands the descriptors
// back instead of uploading them itself.
struct BatchHead {
  size_t mask_bytes = 0, pcm_bytes = 0;  // in: room to reserve
  size_t head_bytes = 0;                 // out: workspace bytes [0, head_bytes) = the upload region
  size_t o_utts = 0, o_tiles = 0, o_row0 = 0, o_sched = 0, o_g = 0, o_masks = 0, o_pcm = 0;
  std::vector<TileDesc> tiles;
  std::vector<long long> row0;
};

// n_frames_in != NULL: feature-input plan (rows of 80 floats, packed back to back)
static int plan_create_common(js2t_ctx* ctx, int n_utts, const int64_t* pcm_byte_off, const int64_t* n_samples,
                              const uint8_t* is_f32, const int32_t* max_frames, const int32_t* n_frames_in,
                              int layout, int pad_tmax, float pad_value, js2t_plan** out,
                              BatchHead* head = nullptr) {
  const bool feat = n_frames_in != nullptr;
  if (ctx == nullptr || out == nullptr || (!feat && (pcm_byte_off == nullptr || n_samples == nullptr)))
    return fail(JS2T_ERR_INVALID, "js2t_plan_create: NULL argument");
  if (n_utts <= 0) return fail(JS2T_ERR_INVALID, "js2t_plan_create: n_utts = %d", n_utts);
  if (layout != JS2T_LAYOUT_RAGGED && layout != JS2T_LAYOUT_PADDED)
    return fail(JS2T_ERR_INVALID, "js2t_plan_create: unknown layout %d", layout);
  js2t_plan* p = new (std::nothrow) js2t_plan();
  if (p == nullptr) return fail(JS2T_ERR_INVALID, "out of host memory");
  p->ctx = ctx;
  p->n_utts = n_utts;
  p->layout = layout;
  p->pad_value = pad_value;
  p->h_utts.resize(n_utts);
  long long tmax = 0;
  p->feature_input = feat;
  long long in_row = 0;
  for (int u = 0; feat && u < n_utts; ++u) {
    long long T = n_frames_in[u];
    if (T <= 0) {
      delete p;
      return fail(JS2T_ERR_INVALID, "utterance %d: empty feature! (%lld frames)", u, T);
    }
    UttDesc& d = p->h_utts[u];
    d.pcm_byte_off = in_row * (long long)(kMel * sizeof(float));
    in_row += T;
    if (max_frames != nullptr && max_frames[u] > 0 && T > max_frames[u]) T = max_frames[u];
    d.n_samples = 0;
    d.n_frames = (int)T;
    d.flags = 2;
    p->total_frames += T;
    if (T > tmax) tmax = T;
  }
  for (int u = 0; !feat && u < n_utts; ++u) {
    if (pcm_byte_off[u] % 16 != 0) {
      delete p;
      return fail(JS2T_ERR_INVALID, "utterance %d: pcm_byte_off %lld is not 16-byte aligned", u,
                  (long long)pcm_byte_off[u]);
    }
    long long T = js2t_num_frames(n_samples[u]);
    if (T <= 0) {
      delete p;
      return fail(JS2T_ERR_SHORT_INPUT, "utterance %d: choose a window size 400 that is [2, %lld]", u,
                  (long long)n_samples[u]);
    }
    if (n_samples[u] > 0x7fffffffLL) {
      delete p;
      return fail(JS2T_ERR_INVALID, "utterance %d: %lld samples exceed the 2^31-1 limit", u,
                  (long long)n_samples[u]);
    }
    if (max_frames != nullptr && max_frames[u] > 0 && T > max_frames[u]) T = max_frames[u];
    UttDesc& d = p->h_utts[u];
    d.pcm_byte_off = pcm_byte_off[u];
    d.n_samples = (int)n_samples[u];
    d.n_frames = (int)T;
    d.flags = (is_f32 != nullptr && is_f32[u]) ? 1 : 0;
    p->total_frames += T;
    if (T > tmax) tmax = T;
  }
  if (layout == JS2T_LAYOUT_PADDED) {
    if (pad_tmax <= 0) pad_tmax = (int)tmax;
    if (pad_tmax < tmax) {
      delete p;
      return fail(JS2T_ERR_INVALID, "padded layout: pad_tmax %d < longest utterance %lld frames", pad_tmax, tmax);
    }
    p->pad_tmax = pad_tmax;
  }
  std::vector<TileDesc> tiles;
  {
    size_t n_est = 0;
    for (int u = 0; u < n_utts; ++u)
      n_est += (size_t)((layout == JS2T_LAYOUT_PADDED ? p->pad_tmax : p->h_utts[u].n_frames) + kTileFrames - 1) / kTileFrames;
    tiles.reserve(n_est + 1);
  }
  long long row = 0;
  for (int u = 0; u < n_utts; ++u) {
    UttDesc& d = p->h_utts[u];
    d.tile_start = (int)tiles.size();
    const int span = layout == JS2T_LAYOUT_PADDED ? p->pad_tmax : d.n_frames;
    d.out_row = row;
    row += span;
    for (int f0 = 0; f0 < span; f0 += kTileFrames) {
      TileDesc t;
      const long long per_frame = feat ? (long long)(kMel * sizeof(float))
                                       : (long long)kHop * ((d.flags & 1) ? 4 : 2);
      t.src_byte_off = d.pcm_byte_off + per_frame * f0;
      t.out_row0 = d.out_row + f0;
      t.utt = u;
      t.frame0 = f0;
      t.stats_slot = (int)tiles.size();
      const int utt_tiles = (d.n_frames + kTileFrames - 1) / kTileFrames;
      const int nfv = d.n_frames - f0 < 0 ? 0 : (d.n_frames - f0 > kTileFrames ? kTileFrames : d.n_frames - f0);
      t.nf = (unsigned char)nfv;
      t.rows = (unsigned char)(span - f0 > kTileFrames ? kTileFrames : span - f0);
      t.flags = (unsigned char)(d.flags & 1);
      t.pad_ = 0;
      if (utt_tiles > p->max_utt_tiles) p->max_utt_tiles = utt_tiles;
      tiles.push_back(t);
    }
  }
  p->out_rows = row;
  p->n_tiles = (int)tiles.size();
  // Processing order: the persistent kernel hands tiles out in array order, so the cheap ones — the
  // partial last tile of every utterance, then pure padding — go to the end, where they shorten the
  // tail in which CTAs run out of work (full tiles keep their utterance order).  Where a tile's rows
  // and statistics land does not depend on this order (out_row0, stats_slot).
  const char* probe_order = getenv("JS2T_PROBE_TILE_ORDER");
  if (probe_order != nullptr && atoi(probe_order) > 0 && !feat) {
    // TIMING PROBE ONLY (tools/cluster_probe.py, -DJS2T_PROBE_STATIC=1 builds, profiles/r2_cluster_probe.txt): the
    // tile array as a STATIC schedule of "one cluster of cs CTAs per utterance" — position i * G + b holds the i-th
    // tile of CTA b (G = persistent grid): cluster b / cs takes utterances c, c + n_clusters, ...; CTA b % cs of the
    // cluster takes tiles r, r + cs, ... of each of them; holes are empty tiles.  cs = 1: plain static round-robin
    // of the processing order below.
    const int cs = atoi(probe_order);
    const int G = fbank_persistent_grid();
    std::vector<TileDesc> sched;
    if (cs == 1) {
      std::stable_sort(tiles.begin(), tiles.end(), [](const TileDesc& a, const TileDesc& b) { return a.nf > b.nf; });
      sched = tiles;
    } else {
      const int n_clusters = G / cs;
      std::vector<std::vector<TileDesc>> seq((size_t)G);
      const bool barrier_slots = getenv("JS2T_PROBE_CLUSTER_SLOTS") != nullptr;  // every CTA of a cluster takes the same number of slots per utterance
      for (int u = 0; u < n_utts; ++u) {
        const int c = u % n_clusters;
        const int first = p->h_utts[u].tile_start;
        const int nt = (u + 1 < n_utts ? p->h_utts[u + 1].tile_start : (int)tiles.size()) - first;
        const int per = (nt + cs - 1) / cs;
        for (int r = 0; r < cs; ++r) {
          int got = 0;
          for (int j = r; j < nt; j += cs, ++got) seq[(size_t)(c * cs + r)].push_back(tiles[(size_t)(first + j)]);
          if (barrier_slots)
            for (; got < per; ++got) {
              TileDesc e = tiles[(size_t)first];
              e.nf = 0;
              e.rows = 0;
              e.stats_slot = (int)tiles.size();  // dummy statistics row (the array below is sized by the schedule)
              seq[(size_t)(c * cs + r)].push_back(e);
            }
        }
      }
      size_t L = 0;
      for (auto& s : seq) L = std::max(L, s.size());
      TileDesc empty = tiles[0];
      empty.nf = 0;
      empty.rows = 0;
      empty.stats_slot = (int)tiles.size();
      sched.assign(L * (size_t)G, empty);
      for (int b = 0; b < G; ++b)
        for (size_t i = 0; i < seq[(size_t)b].size(); ++i) sched[i * (size_t)G + (size_t)b] = seq[(size_t)b][i];
      sched.push_back(empty);  // room for the dummy statistics row
    }
    tiles.swap(sched);
    p->n_tiles = (int)tiles.size();
  } else {
    // stable, by valid frames descending: a counting sort over the 33 possible values (a batch of 256 utterances
    // has ~10 k tiles; a comparison sort was a third of the plan's host time)
    size_t start[kTileFrames + 2] = {0};
    for (const TileDesc& t : tiles) ++start[kTileFrames - t.nf + 1];
    for (int i = 1; i <= kTileFrames + 1; ++i) start[i] += start[i - 1];
    std::vector<TileDesc> sorted(tiles.size());
    for (const TileDesc& t : tiles) sorted[start[kTileFrames - t.nf]++] = t;
    tiles.swap(sorted);
  }

  // one device allocation, carved up
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t o_utts = carve(sizeof(UttDesc) * n_utts);
  const size_t o_tiles = carve(sizeof(TileDesc) * tiles.size());
  size_t o_sched = 0, o_row0 = 0, o_g = 0, o_masks = 0, o_pcm = 0;
  if (head != nullptr) {  // the upload region first (see BatchHead)
    o_row0 = carve(sizeof(long long) * n_utts);
    o_sched = carve(sizeof(int) * 2);
    o_g = carve(sizeof(float) * 2 * kMel);
    o_masks = carve(head->mask_bytes);
    o_pcm = carve(head->pcm_bytes);
    head->head_bytes = off;
  }
  const size_t o_tstats = carve(sizeof(float) * kStatsPerTile * tiles.size());
  const size_t o_mean = carve(sizeof(float) * kMel * n_utts);
  const size_t o_istd = carve(sizeof(float) * kMel * n_utts);
  const size_t o_mv = carve(sizeof(float) * n_utts);
  const size_t o_ustats = carve(sizeof(double) * kStatsPerTile * n_utts);
  if (head == nullptr) {
    o_g = carve(sizeof(float) * 2 * kMel);
    o_sched = carve(sizeof(int) * 2);
    o_row0 = carve(sizeof(long long) * n_utts);
  }
  DeviceGuard guard(ctx->device);
  cudaError_t e = guard.err;
  if (e == cudaSuccess) e = pool_alloc(ctx, off, &p->d_ws, &p->ws_cap);
  if (e != cudaSuccess) {
    delete p;
    return fail(JS2T_ERR_CUDA, "cudaMalloc(%zu bytes of plan workspace) failed: %s", off, cudaGetErrorString(e));
  }
  char* base = static_cast<char*>(p->d_ws);
  p->d_utts = reinterpret_cast<UttDesc*>(base + o_utts);
  p->d_tiles = reinterpret_cast<TileDesc*>(base + o_tiles);
  p->d_tile_stats = reinterpret_cast<float*>(base + o_tstats);
  p->d_mean = reinterpret_cast<float*>(base + o_mean);
  p->d_istd = reinterpret_cast<float*>(base + o_istd);
  p->d_mask_value = reinterpret_cast<float*>(base + o_mv);
  p->d_gmean = reinterpret_cast<float*>(base + o_g);
  p->d_gistd = p->d_gmean + kMel;
  p->d_utt_stats = reinterpret_cast<double*>(base + o_ustats);
  p->d_sched = reinterpret_cast<int*>(base + o_sched);
  p->d_frame_row0 = reinterpret_cast<long long*>(base + o_row0);
  std::vector<long long> row0(n_utts);
  {
    long long acc = 0;
    for (int u = 0; u < n_utts; ++u) {
      row0[u] = acc;
      acc += p->h_utts[u].n_frames;
    }
  }
  if (head != nullptr) {  // js2t_batch_fbank uploads the head itself, on the execute stream
    head->o_utts = o_utts;
    head->o_tiles = o_tiles;
    head->o_row0 = o_row0;
    head->o_sched = o_sched;
    head->o_g = o_g;
    head->o_masks = o_masks;
    head->o_pcm = o_pcm;
    head->tiles.swap(tiles);
    head->row0.swap(row0);
    p->d_pcm = reinterpret_cast<const uint8_t*>(base + o_pcm);
    if (head->mask_bytes) {
      p->d_masks = reinterpret_cast<int*>(base + o_masks);
      p->masks_in_ws = true;
    }
    *out = p;
    return JS2T_OK;
  }
  // Asynchronous upload on the context's own stream.  The sources are pageable: cudaMemcpyAsync returns
  // once they have been staged (so `tiles` may go out of scope), and the DMA itself is ordered on
  // upload_stream in front of ready_ev, which every execute stream waits for (plan_begin).
  e = cudaEventCreateWithFlags(&p->ready_ev, cudaEventDisableTiming);
  cudaStream_t us = ctx->upload_stream;
  if (e == cudaSuccess) e = cudaMemsetAsync(p->d_sched, 0, sizeof(int) * 2, us);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(p->d_utts, p->h_utts.data(), sizeof(UttDesc) * n_utts, cudaMemcpyHostToDevice, us);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(p->d_tiles, tiles.data(), sizeof(TileDesc) * tiles.size(), cudaMemcpyHostToDevice, us);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(p->d_frame_row0, row0.data(), sizeof(long long) * n_utts, cudaMemcpyHostToDevice, us);
  if (e == cudaSuccess) e = cudaEventRecord(p->ready_ev, us);
  if (e != cudaSuccess) {
    cudaStreamSynchronize(us);
    if (p->ready_ev) cudaEventDestroy(p->ready_ev);
    pool_free(ctx, p->d_ws, p->ws_cap);
    delete p;
    return fail(JS2T_ERR_CUDA, "plan descriptor upload failed: %s", cudaGetErrorString(e));
  }
  *out = p;
  return JS2T_OK;
}

int js2t_plan_create(js2t_ctx* ctx, int n_utts, const int64_t* pcm_byte_off, const int64_t* n_samples,
                     const uint8_t* is_f32, const int32_t* max_frames, int layout, int pad_tmax,
                     float pad_value, js2t_plan** out) {
  return plan_create_common(ctx, n_utts, pcm_byte_off, n_samples, is_f32, max_frames, nullptr, layout, pad_tmax,
                            pad_value, out);
}

int js2t_plan_create_features(js2t_ctx* ctx, int n_utts, const int32_t* n_frames, const int32_t* max_frames,
                              int layout, int pad_tmax, float pad_value, js2t_plan** out) {
  if (n_frames == nullptr) return fail(JS2T_ERR_INVALID, "js2t_plan_create_features: n_frames is NULL");
  return plan_create_common(ctx, n_utts, nullptr, nullptr, nullptr, max_frames, n_frames, layout, pad_tmax,
                            pad_value, out);
}

static int plan_destroy_impl(js2t_plan* plan, bool wait) {
  if (plan == nullptr) return JS2T_OK;
  DeviceGuard guard(plan->ctx->device);
  for (cudaEvent_t e : plan->prof_ev) cudaEventDestroy(e);
  // like the cudaFree it replaces, destroying a plan waits for whatever still uses its workspace —
  // but only for the streams the plan was used on, not for the whole device; a caller that already
  // knows the plan's work has completed (js2t_plan_destroy_completed) skips even that
  if (wait) plan_quiesce(plan);
  else if (plan->ready_ev) cudaEventSynchronize(plan->ready_ev);
  if (plan->ready_ev) cudaEventDestroy(plan->ready_ev);
  if (plan->d_masks && !plan->masks_in_ws) pool_free(plan->ctx, plan->d_masks, plan->masks_cap);
  if (plan->d_dbg) cudaFree(plan->d_dbg);
  if (plan->d_ws) pool_free(plan->ctx, plan->d_ws, plan->ws_cap);
  delete plan;
  return JS2T_OK;
}

int js2t_plan_destroy(js2t_plan* plan) { return plan_destroy_impl(plan, /*wait=*/true); }
int js2t_plan_destroy_completed(js2t_plan* plan) { return plan_destroy_impl(plan, /*wait=*/false); }

int64_t js2t_plan_total_frames(const js2t_plan* plan) { return plan ? plan->total_frames : -1; }
int64_t js2t_plan_out_rows(const js2t_plan* plan) { return plan ? plan->out_rows : -1; }

int js2t_plan_get_frames(const js2t_plan* plan, int32_t* n_frames_out) {
  if (plan == nullptr || n_frames_out == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  for (int u = 0; u < plan->n_utts; ++u) n_frames_out[u] = plan->h_utts[u].n_frames;
  return JS2T_OK;
}

int js2t_plan_get_out_rows(const js2t_plan* plan, int64_t* first_row_out) {
  if (plan == nullptr || first_row_out == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  for (int u = 0; u < plan->n_utts; ++u) first_row_out[u] = plan->h_utts[u].out_row;
  return JS2T_OK;
}

int js2t_plan_set_cmvn(js2t_plan* plan, int mode, int norm_means, int norm_vars, int before) {
  if (plan == nullptr) return fail(JS2T_ERR_INVALID, "NULL plan");
  if (mode < JS2T_CMVN_NONE || mode > JS2T_CMVN_STATS_ONLY)
    return fail(JS2T_ERR_INVALID, "unknown CMVN mode %d", mode);
  plan->cmvn_mode = mode;
  plan->norm_means = norm_means != 0;
  plan->norm_vars = norm_vars != 0;
  plan->before = before != 0;
  return JS2T_OK;
}

int js2t_plan_set_global_stats(js2t_plan* plan, const double* mean80, const double* istd80, void* stream) {
  if (plan == nullptr || mean80 == nullptr || istd80 == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  float h[2 * kMel];
  for (int b = 0; b < kMel; ++b) {
    h[b] = (float)mean80[b];
    h[kMel + b] = (float)istd80[b];
  }
  DeviceGuard guard(plan->ctx->device);
  JS2T_CUDA(guard.err);
  // pageable source: the call returns once h has been staged, the DMA is ordered on `stream`
  JS2T_CUDA(plan_begin(plan, (cudaStream_t)stream));
  JS2T_CUDA(cudaMemcpyAsync(plan->d_gmean, h, sizeof(h), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  plan->global_stats_set = true;
  return JS2T_OK;
}

int js2t_plan_set_masks(js2t_plan* plan, int n_fmask, int n_tmask, const int32_t* table, int value_mode,
                        float value_const, void* stream) {
  if (plan == nullptr) return fail(JS2T_ERR_INVALID, "NULL plan");
  if (table == nullptr || n_fmask + n_tmask == 0) {
    plan->has_masks = false;
    plan->n_fmask = plan->n_tmask = 0;
    return JS2T_OK;
  }
  if (n_fmask < 0 || n_tmask < 0) return fail(JS2T_ERR_INVALID, "negative mask count");
  if (value_mode != JS2T_MASK_VALUE_MEAN && value_mode != JS2T_MASK_VALUE_CONST)
    return fail(JS2T_ERR_INVALID, "unknown mask value mode %d", value_mode);
  const size_t bytes = sizeof(int32_t) * 2 * (size_t)(n_fmask + n_tmask) * plan->n_utts;
  DeviceGuard guard(plan->ctx->device);
  JS2T_CUDA(guard.err);
  if (bytes > plan->masks_cap) {
    if (plan->d_masks) {
      plan_quiesce(plan);  // an earlier execute may still read the old table
      pool_free(plan->ctx, plan->d_masks, plan->masks_cap);
    }
    plan->d_masks = nullptr;
    plan->masks_cap = 0;
    void* m = nullptr;
    size_t cap = 0;
    JS2T_CUDA(pool_alloc(plan->ctx, bytes, &m, &cap));
    plan->d_masks = static_cast<int*>(m);
    plan->masks_cap = cap;
  }
  // pageable source: staged when the call returns (the caller may reuse `table`), DMA ordered on `stream`
  JS2T_CUDA(plan_begin(plan, (cudaStream_t)stream));
  JS2T_CUDA(cudaMemcpyAsync(plan->d_masks, table, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  plan->n_fmask = n_fmask;
  plan->n_tmask = n_tmask;
  plan->mask_value_mode = value_mode;
  plan->mask_value_const = value_const;
  plan->has_masks = true;
  return JS2T_OK;
}

int js2t_plan_set_dither(js2t_plan* plan, const float* noise_dev) {
  if (plan == nullptr) return fail(JS2T_ERR_INVALID, "NULL plan");
  if (plan->feature_input && noise_dev != nullptr)
    return fail(JS2T_ERR_STATE, "dither applies to PCM plans (js2t_plan_create) only");
  plan->dither = noise_dev;
  return JS2T_OK;
}

// ---------------------------------------------------------------------------------------------
static DeviceTables tables_of(const js2t_ctx* c) {
  DeviceTables t;
  t.window_half = c->d_tables;
  t.tw256 = reinterpret_cast<const float2*>(c->d_tables + 400);
  t.tw512 = reinterpret_cast<const float2*>(c->d_tables + 400 + 512);
  return t;
}

static ApplyLaunch make_apply(const js2t_plan* plan, float* out_dev, bool shared) {
  ApplyLaunch a;
  a.utts = plan->d_utts;
  a.tiles = plan->d_tiles;
  a.n_tiles = plan->n_tiles;
  a.out = out_dev;
  a.mean = shared ? plan->d_gmean : plan->d_mean;
  a.istd = shared ? plan->d_gistd : plan->d_istd;
  a.shared_stats = shared ? 1 : 0;
  a.masks = plan->has_masks ? plan->d_masks : nullptr;
  a.n_fmask = plan->n_fmask;
  a.n_tmask = plan->n_tmask;
  a.mask_value = plan->d_mask_value;
  a.cmvn_after = (plan->cmvn_mode != JS2T_CMVN_NONE && !plan->before) ? 1 : 0;
  a.pad_tmax = plan->pad_tmax;
  a.pad_value = plan->pad_value;
  return a;
}

static FinalizeLaunch make_finalize(const js2t_plan* plan, const float* raw, bool shared) {
  FinalizeLaunch z;
  memset(&z, 0, sizeof(z));
  z.utts = plan->d_utts;
  z.n_utts = plan->n_utts;
  z.tile_stats = plan->d_tile_stats;
  z.raw = raw;
  z.norm_means = plan->norm_means;
  z.norm_vars = plan->norm_vars;
  z.cmvn_enabled = (plan->cmvn_mode == JS2T_CMVN_UTTERANCE || plan->cmvn_mode == JS2T_CMVN_GLOBAL) ? 1 : 0;
  z.cmvn_after = plan->before ? 0 : 1;
  z.g_mean = shared ? plan->d_gmean : nullptr;
  z.g_istd = shared ? plan->d_gistd : nullptr;
  z.masks = plan->has_masks ? plan->d_masks : nullptr;
  z.n_fmask = plan->n_fmask;
  z.n_tmask = plan->n_tmask;
  z.mask_value_mode = plan->mask_value_mode;
  z.mask_value_const = plan->mask_value_const;
  z.mean = plan->d_mean;
  z.istd = plan->d_istd;
  z.mask_value = plan->d_mask_value;
  z.stats_out = plan->d_utt_stats;
  return z;
}

// in_dev (PCM or feature rows) -> out_dev according to the plan
static int run_pipeline(js2t_plan* plan, const void* in_dev, float* out_dev, cudaStream_t stream,
                        bool from_pcm) {
  if (plan == nullptr || in_dev == nullptr || out_dev == nullptr)
    return fail(JS2T_ERR_INVALID, "execute: NULL argument");
  if (from_pcm != !plan->feature_input)
    return fail(JS2T_ERR_STATE, from_pcm ? "js2t_fbank_execute needs a plan made by js2t_plan_create"
                                         : "js2t_features_execute needs a plan made by js2t_plan_create_features");
  if (from_pcm && !plan->ctx->tables_set)
    return fail(JS2T_ERR_STATE, "js2t_ctx_set_tables has not been called");
  DeviceGuard guard(plan->ctx->device);
  JS2T_CUDA(guard.err);
  const int mode = plan->cmvn_mode;
  const bool masks = plan->has_masks;
  if (mode == JS2T_CMVN_GLOBAL && !plan->global_stats_set)
    return fail(JS2T_ERR_STATE, "global CMVN requested but no statistics were set");
  JS2T_CUDA(plan_begin(plan, stream));

  FbankLaunch f;
  memset(&f, 0, sizeof(f));
  f.pcm = static_cast<const uint8_t*>(in_dev);
  f.utts = plan->d_utts;
  f.tiles = plan->d_tiles;
  f.n_tiles = plan->n_tiles;
  f.sched = plan->d_sched;
  if (from_pcm) f.tab = tables_of(plan->ctx);
  f.out = out_dev;
  f.pad_tmax = plan->pad_tmax;
  f.pad_value = plan->pad_value;
  f.epilogue = kEpiRaw;
  f.dither = from_pcm ? plan->dither : nullptr;
  f.dither_row0 = plan->d_frame_row0;
  f.dbg_skip = plan->dbg_skip;
  f.grid_limit = plan->grid_limit;
  f.dbg_times = plan->d_dbg;

  // the dominant kernel, optionally bracketed by profiling events on the launching stream
  auto launch_main = [&](const FbankLaunch& fl) -> cudaError_t {
    const size_t slots = plan->prof_ev.size() / 2;
    const size_t slot = slots ? (size_t)(plan->prof_calls % (long long)slots) : 0;
    if (slots) cudaEventRecord(plan->prof_ev[2 * slot], stream);
    cudaError_t e = from_pcm ? launch_fbank(fl, stream) : launch_features(fl, stream);
    if (slots) {
      cudaEventRecord(plan->prof_ev[2 * slot + 1], stream);
      plan->prof_calls++;
    }
    return e;
  };
  // (1) nothing data-dependent after the log-mel: one kernel, one pass over HBM
  if (mode == JS2T_CMVN_NONE && !masks) {
    JS2T_CUDA(launch_main(f));
    plan->stats_valid = false;
    return JS2T_OK;
  }
  // (2) global CMVN with known statistics and no data-dependent fill value: normalise (+ mask) in
  //     the fbank epilogue, still one pass
  const bool mean_fill = masks && plan->mask_value_mode == JS2T_MASK_VALUE_MEAN;
  if (from_pcm && mode == JS2T_CMVN_GLOBAL && !mean_fill && plan->before && plan->dither == nullptr &&
      (!masks || plan->n_fmask + plan->n_tmask <= 16)) {
    f.epilogue = kEpiNormKnown;
    f.g_mean = plan->d_gmean;
    f.g_istd = plan->d_gistd;
    if (masks) {
      f.masks = plan->d_masks;
      f.n_fmask = plan->n_fmask;
      f.n_tmask = plan->n_tmask;
      JS2T_CUDA(launch_fill_value(plan->d_mask_value, plan->n_utts, plan->mask_value_const, stream));
      f.mask_value = plan->d_mask_value;
    }
    JS2T_CUDA(launch_main(f));
    plan->stats_valid = false;
    return JS2T_OK;
  }
  f.tile_stats = plan->d_tile_stats;
  f.dbg_times = plan->d_dbg;
  // (3) raw log-mel + per-tile statistics -> per-utterance finalize -> in-place apply
  JS2T_CUDA(launch_main(f));
  const bool shared = (mode == JS2T_CMVN_GLOBAL);
  FinalizeLaunch z = make_finalize(plan, out_dev, shared);
  JS2T_CUDA(launch_finalize(z, stream));
  plan->stats_valid = true;
  if (mode != JS2T_CMVN_STATS_ONLY) {
    ApplyLaunch a = make_apply(plan, out_dev, shared);
    JS2T_CUDA(launch_apply(a, stream));
  }
  return JS2T_OK;
}

int js2t_fbank_execute(js2t_plan* plan, const void* pcm_dev, float* out_dev, void* stream) {
  return run_pipeline(plan, pcm_dev, out_dev, (cudaStream_t)stream, /*from_pcm=*/true);
}

int js2t_features_execute(js2t_plan* plan, const float* feats_dev, float* out_dev, void* stream) {
  return run_pipeline(plan, feats_dev, out_dev, (cudaStream_t)stream, /*from_pcm=*/false);
}

int js2t_plan_enable_profiling(js2t_plan* plan, int n_slots) {
  if (plan == nullptr || n_slots < 0) return fail(JS2T_ERR_INVALID, "bad argument");
  DeviceGuard guard(plan->ctx->device);
  JS2T_CUDA(guard.err);
  for (cudaEvent_t e : plan->prof_ev) cudaEventDestroy(e);
  plan->prof_ev.clear();
  plan->prof_calls = 0;
  for (int i = 0; i < 2 * n_slots; ++i) {
    cudaEvent_t e;
    JS2T_CUDA(cudaEventCreate(&e));
    plan->prof_ev.push_back(e);
  }
  return JS2T_OK;
}

int js2t_plan_kernel_times_ms(js2t_plan* plan, float* ms_out, int n, int* n_written) {
  if (plan == nullptr || ms_out == nullptr || n_written == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  const long long slots = (long long)plan->prof_ev.size() / 2;
  const long long have = plan->prof_calls < slots ? plan->prof_calls : slots;
  int w = 0;
  for (long long i = 0; i < have && w < n; ++i) {
    float ms = 0.f;
    JS2T_CUDA(cudaEventSynchronize(plan->prof_ev[2 * i + 1]));
    JS2T_CUDA(cudaEventElapsedTime(&ms, plan->prof_ev[2 * i], plan->prof_ev[2 * i + 1]));
    ms_out[w++] = ms;
  }
  *n_written = w;
  return JS2T_OK;
}

int js2t_plan_set_option(js2t_plan* plan, const char* name, int value) {
  if (plan == nullptr || name == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  if (strcmp(name, "max_ctas") == 0) {
    plan->grid_limit = value;
    return JS2T_OK;
  }
  if (strcmp(name, "debug_skip") == 0) {
    plan->dbg_skip = value;
    return JS2T_OK;
  }
  if (strcmp(name, "debug_times") == 0) {
    DeviceGuard guard(plan->ctx->device);
  JS2T_CUDA(guard.err);
    if (value != 0 && plan->d_dbg == nullptr) {
      // [n_tiles][4] per-tile stamps + [n_tiles][8 warps][8] per-warp stamps (JS2T_DBG builds)
      JS2T_CUDA(cudaMalloc(&plan->d_dbg, sizeof(unsigned long long) * 68 * plan->n_tiles));
      JS2T_CUDA(cudaMemset(plan->d_dbg, 0, sizeof(unsigned long long) * 68 * plan->n_tiles));
    } else if (value == 0 && plan->d_dbg != nullptr) {
      cudaFree(plan->d_dbg);
      plan->d_dbg = nullptr;
    }
    return JS2T_OK;
  }
  return fail(JS2T_ERR_INVALID, "unknown option '%s'", name);
}

int js2t_plan_debug_times(const js2t_plan* plan, unsigned long long* host_out, int64_t n_values) {
  if (plan == nullptr || host_out == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  if (plan->d_dbg == nullptr) return fail(JS2T_ERR_STATE, "option debug_times is off");
  if (n_values > 68ll * plan->n_tiles) n_values = 68ll * plan->n_tiles;
  DeviceGuard guard(plan->ctx->device);
  JS2T_CUDA(guard.err);
  JS2T_CUDA(cudaMemcpy(host_out, plan->d_dbg, sizeof(unsigned long long) * n_values, cudaMemcpyDeviceToHost));
  return JS2T_OK;
}

int js2t_plan_copy_utt_stats(const js2t_plan* plan, double* dst_dev, void* stream) {
  if (plan == nullptr || dst_dev == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  if (!plan->stats_valid) return fail(JS2T_ERR_STATE, "no statistics: run js2t_fbank_execute in a statistics mode first");
  DeviceGuard guard(plan->ctx->device);
  JS2T_CUDA(guard.err);
  JS2T_CUDA(plan_begin(const_cast<js2t_plan*>(plan), (cudaStream_t)stream));
  JS2T_CUDA(cudaMemcpyAsync(dst_dev, plan->d_utt_stats, sizeof(double) * kStatsPerTile * plan->n_utts,
                            cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return JS2T_OK;
}

int js2t_plan_utt_stats(const js2t_plan* plan, const double** stats_dev) {
  if (plan == nullptr || stats_dev == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  if (!plan->stats_valid) return fail(JS2T_ERR_STATE, "no statistics: run js2t_fbank_execute in a statistics mode first");
  *stats_dev = plan->d_utt_stats;
  return JS2T_OK;
}

// ---------------------------------------------------------------------------------------------
int js2t_global_stats_accumulate(js2t_plan* plan, double* accum_dev, void* stream) {
  if (plan == nullptr || accum_dev == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  if (!plan->stats_valid) return fail(JS2T_ERR_STATE, "no statistics to accumulate");
  DeviceGuard guard(plan->ctx->device);
  JS2T_CUDA(guard.err);
  JS2T_CUDA(plan_begin(plan, (cudaStream_t)stream));
  JS2T_CUDA(launch_global_accumulate(plan->d_utt_stats, plan->d_utts, plan->n_utts, accum_dev,
                                     (cudaStream_t)stream));
  return JS2T_OK;
}

// ---- NCCL, bound at run time --------------------------------------------------------------------
// libnccl is not linked: it is looked up in the process when the first communicator call is made, so the
// library of the host application (torch bundles its own libnccl.so.2) is the one that is used, and
// single-GPU users need no NCCL at all.
namespace {

struct NcclUniqueId {
  char internal[128];  // NCCL_UNIQUE_ID_BYTES
};
struct NcclApi {
  int (*get_unique_id)(NcclUniqueId*) = nullptr;
  int (*comm_init_rank)(void**, int, NcclUniqueId, int) = nullptr;
  int (*comm_destroy)(void*) = nullptr;
  int (*all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*get_error_string)(int) = nullptr;
  bool ok = false;
};

const NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) return;
    api.get_unique_id = reinterpret_cast<decltype(api.get_unique_id)>(dlsym(h, "ncclGetUniqueId"));
    api.comm_init_rank = reinterpret_cast<decltype(api.comm_init_rank)>(dlsym(h, "ncclCommInitRank"));
    api.comm_destroy = reinterpret_cast<decltype(api.comm_destroy)>(dlsym(h, "ncclCommDestroy"));
    api.all_reduce = reinterpret_cast<decltype(api.all_reduce)>(dlsym(h, "ncclAllReduce"));
    api.get_error_string = reinterpret_cast<decltype(api.get_error_string)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_reduce;
  });
  return api.ok ? &api : nullptr;
}

int nccl_fail(const NcclApi* api, const char* what, int rc) {
  return fail(JS2T_ERR_NCCL, "%s failed: %s (code %d)", what,
              api->get_error_string ? api->get_error_string(rc) : "?", rc);
}

}  // namespace

int js2t_nccl_unique_id(void* id128_out) {
  if (id128_out == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  const NcclApi* api = nccl_api();
  if (api == nullptr) return fail(JS2T_ERR_NCCL, "cannot load libnccl: %s", dlerror());
  NcclUniqueId id;
  const int rc = api->get_unique_id(&id);
  if (rc != 0) return nccl_fail(api, "ncclGetUniqueId", rc);
  memcpy(id128_out, &id, sizeof(id));
  return JS2T_OK;
}

int js2t_nccl_comm_create(const void* id128, int world_size, int rank, int device, void** comm_out) {
  if (id128 == nullptr || comm_out == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  if (world_size < 1 || rank < 0 || rank >= world_size)
    return fail(JS2T_ERR_INVALID, "rank %d of %d", rank, world_size);
  const NcclApi* api = nccl_api();
  if (api == nullptr) return fail(JS2T_ERR_NCCL, "cannot load libnccl: %s", dlerror());
  DeviceGuard guard(device);
  JS2T_CUDA(guard.err);
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  void* comm = nullptr;
  const int rc = api->comm_init_rank(&comm, world_size, id, rank);  // collective over all ranks
  if (rc != 0) return nccl_fail(api, "ncclCommInitRank", rc);
  *comm_out = comm;
  return JS2T_OK;
}

int js2t_nccl_comm_destroy(void* nccl_comm) {
  if (nccl_comm == nullptr) return JS2T_OK;
  const NcclApi* api = nccl_api();
  if (api == nullptr) return fail(JS2T_ERR_NCCL, "cannot load libnccl");
  const int rc = api->comm_destroy(nccl_comm);
  if (rc != 0) return nccl_fail(api, "ncclCommDestroy", rc);
  return JS2T_OK;
}

int js2t_global_stats_allreduce(void* nccl_comm, double* accum_dev, void* stream) {
  if (nccl_comm == nullptr || accum_dev == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  const NcclApi* api = nccl_api();
  if (api == nullptr) return fail(JS2T_ERR_NCCL, "cannot load libnccl: %s", dlerror());
  const int kNcclFloat64 = 8, kNcclSum = 0;
  const int rc = api->all_reduce(accum_dev, accum_dev, (size_t)(kStatsPerTile + 1), kNcclFloat64, kNcclSum,
                                 nccl_comm, (cudaStream_t)stream);
  if (rc != 0) return nccl_fail(api, "ncclAllReduce", rc);
  return JS2T_OK;
}

int js2t_global_stats_finalize(js2t_plan* plan, const double* accum_dev, void* stream) {
  if (plan == nullptr || accum_dev == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  DeviceGuard guard(plan->ctx->device);
  JS2T_CUDA(guard.err);
  JS2T_CUDA(plan_begin(plan, (cudaStream_t)stream));
  JS2T_CUDA(launch_global_finalize(accum_dev, plan->norm_means, plan->norm_vars, plan->d_gmean, plan->d_gistd,
                                   (cudaStream_t)stream));
  plan->global_stats_set = true;
  return JS2T_OK;
}

int js2t_plan_copy_global_stats(const js2t_plan* plan, float* dst_dev, void* stream) {
  if (plan == nullptr || dst_dev == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  if (!plan->global_stats_set) return fail(JS2T_ERR_STATE, "no global statistics set");
  DeviceGuard guard(plan->ctx->device);
  JS2T_CUDA(guard.err);
  JS2T_CUDA(plan_begin(const_cast<js2t_plan*>(plan), (cudaStream_t)stream));
  // d_gmean | d_gistd are adjacent (one carve of 160 floats)
  JS2T_CUDA(cudaMemcpyAsync(dst_dev, plan->d_gmean, sizeof(float) * 2 * kMel, cudaMemcpyDeviceToDevice,
                            (cudaStream_t)stream));
  return JS2T_OK;
}

int js2t_normalize_execute(js2t_plan* plan, float* out_dev, void* stream_) {
  if (plan == nullptr || out_dev == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  if (!plan->global_stats_set) return fail(JS2T_ERR_STATE, "no global statistics set");
  if (!plan->stats_valid)
    return fail(JS2T_ERR_STATE, "js2t_normalize_execute follows a STATS_ONLY js2t_fbank_execute on the same plan");
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceGuard guard(plan->ctx->device);
  JS2T_CUDA(guard.err);
  JS2T_CUDA(plan_begin(plan, stream));
  const int saved = plan->cmvn_mode;
  plan->cmvn_mode = JS2T_CMVN_GLOBAL;
  // per-utterance fill value under the global normalisation (re-reads the per-tile statistics of the
  // plan's last execute; only SpecAugment needs it)
  cudaError_t e = cudaSuccess;
  if (plan->has_masks) {
    FinalizeLaunch z = make_finalize(plan, out_dev, /*shared=*/true);
    e = launch_finalize(z, stream);
  }
  ApplyLaunch a = make_apply(plan, out_dev, /*shared=*/true);
  if (e == cudaSuccess) e = launch_apply(a, stream);
  plan->cmvn_mode = saved;
  JS2T_CUDA(e);
  return JS2T_OK;
}

// ---- host-side packing of a ragged batch into one (pinned) staging buffer ------------------------
// The per-batch callers of the reference API hand over one array per utterance (pageable memory); the
// device path wants them back to back, 16-byte aligned, in pinned memory.  A single-threaded copy of a
// 20 000-frame batch (6.4 MB) costs more than its H2D transfer, so the copy is spread over a small
// persistent pool of worker threads (chunks of 256 KB, claimed with an atomic counter).
namespace {

// Copy into the staging buffer with NON-TEMPORAL stores.  The destination is pinned memory that only the
// GPU's DMA engine reads next: with ordinary stores the lines stay dirty in the caches of whichever cores
// ran the copy threads, and the H2D transfer that follows has to snoop them out one by one — measured on
// the B200 host: 6 GB/s instead of 51 GB/s for the copy after a pooled memcpy (profiles/r2h_*).
void stream_copy(char* dst, const char* src, size_t n) {
#if defined(__x86_64__)
  size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
  if (head > n) head = n;
  if (head) memcpy(dst, src, head);
  size_t i = head;
  for (; i + 64 <= n; i += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
  }
  for (; i + 16 <= n; i += 16)
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i)));
  if (i < n) memcpy(dst + i, src + i, n - i);
  _mm_sfence();
#else
  memcpy(dst, src, n);
#endif
}

class CopyPool {
 public:
  struct Chunk {
    const char* src;
    char* dst;
    size_t n;
  };
  explicit CopyPool(int n_workers) {
    for (int i = 0; i < n_workers; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }
  // copies all chunks; the calling thread works too; returns when everything is done
  void run(const std::vector<Chunk>& chunks, int n_helpers) {
    std::lock_guard<std::mutex> serial(run_mu_);  // one batch at a time per pool
    {
      std::lock_guard<std::mutex> lk(mu_);
      chunks_ = &chunks;
      next_.store(0);
      done_.store(0);
      wanted_ = std::min(n_helpers, (int)workers_.size());
      ++epoch_;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return done_.load() == chunks.size() && active_ == 0; });
    chunks_ = nullptr;
  }

 private:
  void work() {
    const std::vector<Chunk>& c = *chunks_;
    for (;;) {
      const size_t i = next_.fetch_add(1);
      if (i >= c.size()) break;
      stream_copy(c[i].dst, c[i].src, c[i].n);
      done_.fetch_add(1);
    }
  }
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || (epoch_ != seen && chunks_ != nullptr && wanted_ > 0); });
        if (stop_) return;
        seen = epoch_;
        --wanted_;
        ++active_;
      }
      work();
      {
        std::lock_guard<std::mutex> lk(mu_);
        --active_;
      }
      done_cv_.notify_all();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_cv_;
  const std::vector<Chunk>* chunks_ = nullptr;
  std::atomic<size_t> next_{0}, done_{0};
  unsigned long long epoch_ = 0;
  int wanted_ = 0, active_ = 0;
  bool stop_ = false;
};

CopyPool* copy_pool() {
  static CopyPool* pool = [] {
    // half of the cores this process may run on, at most 8 threads including the caller (host memory bandwidth is
    // what limits the gather, and N ranks share the box); JS2T_COPY_THREADS overrides it
    unsigned hw = std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) hw = (unsigned)CPU_COUNT(&set);
    int n = hw > 2 ? (int)std::min(hw / 2, 8u) - 1 : 0;  // + the calling thread
    if (const char* e = getenv("JS2T_COPY_THREADS")) n = std::max(1, std::min(atoi(e), 64)) - 1;
    return new CopyPool(n < 0 ? 0 : n);
  }();
  return pool;
}

}  // namespace

int js2t_pack_pcm(int n_utts, const void* const* src, const int64_t* n_bytes, const int64_t* dst_byte_off,
                  void* dst, int64_t dst_capacity, int n_threads) {
  if (n_utts < 0 || (n_utts > 0 && (src == nullptr || n_bytes == nullptr || dst_byte_off == nullptr || dst == nullptr)))
    return fail(JS2T_ERR_INVALID, "js2t_pack_pcm: NULL argument");
  constexpr size_t kChunk = 256 * 1024;
  std::vector<CopyPool::Chunk> chunks;
  size_t total = 0;
  for (int u = 0; u < n_utts; ++u) {
    if (n_bytes[u] < 0 || dst_byte_off[u] < 0 || dst_byte_off[u] + n_bytes[u] > dst_capacity)
      return fail(JS2T_ERR_INVALID, "js2t_pack_pcm: utterance %d (%lld bytes at %lld) does not fit in %lld bytes", u,
                  (long long)n_bytes[u], (long long)dst_byte_off[u], (long long)dst_capacity);
    if (n_bytes[u] > 0 && src[u] == nullptr) return fail(JS2T_ERR_INVALID, "js2t_pack_pcm: src[%d] is NULL", u);
    for (size_t o = 0; o < (size_t)n_bytes[u]; o += kChunk)
      chunks.push_back({static_cast<const char*>(src[u]) + o, static_cast<char*>(dst) + dst_byte_off[u] + o,
                        std::min(kChunk, (size_t)n_bytes[u] - o)});
    total += (size_t)n_bytes[u];
  }
  if (chunks.empty()) return JS2T_OK;
  CopyPool* pool = copy_pool();
  int helpers = n_threads > 0 ? n_threads - 1 : pool->size();
  if (total < 4 * kChunk) helpers = 0;  // small batches: waking the workers costs more than the copy
  if (helpers <= 0) {
    for (const auto& c : chunks) stream_copy(c.dst, c.src, c.n);
    return JS2T_OK;
  }
  pool->run(chunks, helpers);
  return JS2T_OK;
}

// One call per batch for callers that hold the utterances in host memory (the per-batch route of the
// reference's data path: collate_fn -> Batch, joeynmt/datasets.py:221-225, batch.py:114-121).
int js2t_batch_fbank(js2t_ctx* ctx, int n_utts, const void* const* pcm_host, const int64_t* n_samples,
                     const uint8_t* is_f32, const js2t_batch_opts* o, float* out_dev, int64_t out_capacity_rows,
                     void* stream_, js2t_plan** plan_out) {
  if (ctx == nullptr || pcm_host == nullptr || n_samples == nullptr || o == nullptr || out_dev == nullptr)
    return fail(JS2T_ERR_INVALID, "js2t_batch_fbank: NULL argument");
  if (n_utts <= 0) return fail(JS2T_ERR_INVALID, "js2t_batch_fbank: n_utts = %d", n_utts);
  if (!ctx->tables_set) return fail(JS2T_ERR_STATE, "js2t_ctx_set_tables has not been called");
  if (o->cmvn_mode != JS2T_CMVN_NONE && o->cmvn_mode != JS2T_CMVN_UTTERANCE && o->cmvn_mode != JS2T_CMVN_GLOBAL)
    return fail(JS2T_ERR_INVALID, "js2t_batch_fbank: CMVN mode %d", o->cmvn_mode);
  if (o->cmvn_mode == JS2T_CMVN_GLOBAL && (o->global_mean80 == nullptr || o->global_istd80 == nullptr))
    return fail(JS2T_ERR_STATE, "global CMVN requested but no statistics were given");
  const int n_masks = o->mask_table != nullptr ? o->n_fmask + o->n_tmask : 0;
  if (o->n_fmask < 0 || o->n_tmask < 0) return fail(JS2T_ERR_INVALID, "negative mask count");
  if (n_masks > 0 && o->mask_value_mode != JS2T_MASK_VALUE_MEAN && o->mask_value_mode != JS2T_MASK_VALUE_CONST)
    return fail(JS2T_ERR_INVALID, "unknown mask value mode %d", o->mask_value_mode);
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceGuard guard(ctx->device);
  JS2T_CUDA(guard.err);
  static const bool trace = getenv("JS2T_BATCH_TRACE") != nullptr;  // tuning aid: host time per phase to stderr
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto t_start = now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    const auto t1 = now();
    fprintf(stderr, "  js2t_batch_fbank %-10s %7.1f us\n", what,
            std::chrono::duration<double, std::micro>(t1 - t_start).count());
    t_start = t1;
  };

  // ---- geometry: utterances back to back, 16-byte aligned ---------------------------------------------
  std::vector<int64_t> off((size_t)n_utts), bytes((size_t)n_utts);
  int64_t pcm_bytes = 0;
  for (int u = 0; u < n_utts; ++u) {
    if (n_samples[u] < 0) return fail(JS2T_ERR_INVALID, "utterance %d: negative length", u);
    if (pcm_host[u] == nullptr && n_samples[u] > 0) return fail(JS2T_ERR_INVALID, "js2t_batch_fbank: pcm_host[%d] is NULL", u);
    off[(size_t)u] = pcm_bytes;
    bytes[(size_t)u] = n_samples[u] * ((is_f32 != nullptr && is_f32[u]) ? 4 : 2);
    pcm_bytes += (bytes[(size_t)u] + 15) / 16 * 16;
  }
  BatchHead head;
  head.pcm_bytes = (size_t)std::max<int64_t>(pcm_bytes, 16);
  head.mask_bytes = sizeof(int32_t) * 2 * (size_t)n_masks * (size_t)n_utts;
  js2t_plan* plan = nullptr;
  const int rc = plan_create_common(ctx, n_utts, off.data(), n_samples, is_f32, o->max_frames, nullptr, o->layout,
                                    o->pad_tmax, o->pad_value, &plan, &head);
  if (rc != JS2T_OK) return rc;
  if (plan->out_rows > out_capacity_rows) {
    const long long need = plan->out_rows;
    plan_destroy_impl(plan, false);
    return fail(JS2T_ERR_INVALID, "js2t_batch_fbank: out_dev holds %lld rows, the batch needs %lld",
                (long long)out_capacity_rows, need);
  }

  lap("plan");
  // ---- reap what earlier calls left behind, take a staging slot -----------------------------------------
  js2t_ctx::StagingSlot* slot = nullptr;
  {
    std::lock_guard<std::mutex> lk(ctx->batch_mu);
    while (!ctx->retired.empty() &&
           (ctx->retired.size() > 8 || cudaEventQuery(ctx->retired.front().first) == cudaSuccess)) {
      cudaEventSynchronize(ctx->retired.front().first);
      ctx->spare_events.push_back(ctx->retired.front().first);
      plan_destroy_impl(ctx->retired.front().second, false);
      ctx->retired.erase(ctx->retired.begin());
    }
    cudaGetLastError();  // cudaEventQuery's cudaErrorNotReady is not an error
    lap("reap");
  }
  constexpr size_t kMaxSlots = 8;                       // enough for a few threads with a couple of batches in flight
  constexpr size_t kMaxPinnedBytes = (size_t)512 << 20;  // ... as long as the pinned memory stays bounded
  for (;;) {
    js2t_ctx::StagingSlot* wait_for = nullptr;
    {
      std::lock_guard<std::mutex> lk(ctx->batch_mu);
      if (ctx->slots.capacity() < kMaxSlots) ctx->slots.reserve(kMaxSlots);  // pointers into the vector stay valid
      for (auto& sl : ctx->slots)
        if (sl.pending && cudaEventQuery(sl.ev) == cudaSuccess) sl.pending = false;
      cudaGetLastError();  // cudaEventQuery's cudaErrorNotReady is not an error
      for (auto& sl : ctx->slots)
        if (!sl.filling && !sl.pending && sl.cap >= head.head_bytes && (slot == nullptr || sl.cap < slot->cap)) slot = &sl;
      if (slot == nullptr) {
        for (auto& sl : ctx->slots)
          if (!sl.filling && !sl.pending) slot = &sl;  // a free slot that is too small: re-allocated below
      }
      size_t pinned = 0;
      for (auto& sl : ctx->slots) pinned += sl.cap;
      if (slot == nullptr && ctx->slots.size() < kMaxSlots &&
          (ctx->slots.size() < 2 || pinned + head.head_bytes <= kMaxPinnedBytes)) {
        ctx->slots.emplace_back();
        slot = &ctx->slots.back();
      }
      if (slot == nullptr) {  // every slot is busy: wait for an upload that has been issued (outside the lock)
        for (auto& sl : ctx->slots)
          if (sl.pending && !sl.filling) {
            wait_for = &sl;
            break;
          }
        if (wait_for != nullptr) wait_for->filling = true;  // ours from here on
      } else {
        slot->filling = true;
      }
    }
    if (slot != nullptr) break;
    if (wait_for != nullptr) {
      cudaEventSynchronize(wait_for->ev);
      wait_for->pending = false;  // (filling: no other call looks at it)
      slot = wait_for;
      lap("slot-wait");
      break;
    }
    std::this_thread::yield();  // all slots are being filled by other threads
  }
  auto release_slot = [&](bool recorded) {
    std::lock_guard<std::mutex> lk(ctx->batch_mu);
    slot->pending = recorded;
    slot->filling = false;
  };
  cudaError_t e = cudaSuccess;
  if (slot->cap < head.head_bytes) {
    if (slot->host) cudaFreeHost(slot->host);
    slot->host = nullptr;
    slot->cap = 0;
    const size_t cap = align_up(head.head_bytes + head.head_bytes / 4, 1 << 20);
    e = cudaHostAlloc(reinterpret_cast<void**>(&slot->host), cap, cudaHostAllocDefault);
    if (e == cudaSuccess) slot->cap = cap;
  }
  if (e == cudaSuccess && slot->ev == nullptr) e = cudaEventCreateWithFlags(&slot->ev, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    release_slot(false);
    plan_destroy_impl(plan, false);
    return fail(JS2T_ERR_CUDA, "js2t_batch_fbank: staging slot: %s", cudaGetErrorString(e));
  }

  lap("slot");
  // ---- fill the slot: descriptors, zeroed scheduler counters, statistics, masks, PCM ----------------------
  char* h = slot->host;
  memcpy(h + head.o_utts, plan->h_utts.data(), sizeof(UttDesc) * (size_t)n_utts);
  memcpy(h + head.o_tiles, head.tiles.data(), sizeof(TileDesc) * head.tiles.size());
  memcpy(h + head.o_row0, head.row0.data(), sizeof(long long) * (size_t)n_utts);
  memset(h + head.o_sched, 0, sizeof(int) * 2);
  if (o->cmvn_mode == JS2T_CMVN_GLOBAL) {
    float* g = reinterpret_cast<float*>(h + head.o_g);
    for (int b = 0; b < kMel; ++b) {
      g[b] = (float)o->global_mean80[b];
      g[kMel + b] = (float)o->global_istd80[b];
    }
    plan->global_stats_set = true;
  }
  if (n_masks > 0) memcpy(h + head.o_masks, o->mask_table, head.mask_bytes);
  const int prc = js2t_pack_pcm(n_utts, pcm_host, bytes.data(), off.data(), h + head.o_pcm, (int64_t)head.pcm_bytes, 0);
  if (prc != JS2T_OK) {
    release_slot(false);
    plan_destroy_impl(plan, false);
    return prc;
  }

  lap("pack");
  // ---- one transfer, then the kernels ---------------------------------------------------------------------
  // The transfer runs on the context's own copy stream and the kernels wait for its event: the upload of this
  // batch overlaps the kernels of the previous one (same caller stream), the device is then bound by the larger
  // of the two instead of their sum.  The workspace it writes is not in use: a plan's buffer returns to the pool
  // only after the event behind its last kernel has completed.
  e = cudaMemcpyAsync(plan->d_ws, h, head.head_bytes, cudaMemcpyHostToDevice, ctx->upload_stream);
  if (e == cudaSuccess) e = cudaEventRecord(slot->ev, ctx->upload_stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, slot->ev, 0);
  if (e != cudaSuccess) {
    cudaStreamSynchronize(ctx->upload_stream);  // whatever was issued no longer reads the slot
    cudaStreamSynchronize(stream);
    release_slot(false);
    plan_destroy_impl(plan, false);
    return fail(JS2T_ERR_CUDA, "js2t_batch_fbank: upload: %s", cudaGetErrorString(e));
  }
  release_slot(true);
  lap("memcpy");
  plan->streams.push_back(stream);  // stream order puts the kernels behind the upload: no ready event
  plan->cmvn_mode = o->cmvn_mode;
  plan->norm_means = o->norm_means != 0;
  plan->norm_vars = o->norm_vars != 0;
  plan->before = o->before != 0;
  if (n_masks > 0) {
    plan->n_fmask = o->n_fmask;
    plan->n_tmask = o->n_tmask;
    plan->mask_value_mode = o->mask_value_mode;
    plan->mask_value_const = o->mask_value_const;
    plan->has_masks = true;
  }
  const int xrc = run_pipeline(plan, plan->d_pcm, out_dev, stream, /*from_pcm=*/true);
  lap("launch");
  if (xrc != JS2T_OK || plan_out != nullptr) {
    if (xrc != JS2T_OK) {
      cudaStreamSynchronize(stream);
      plan_destroy_impl(plan, false);
      return xrc;
    }
    *plan_out = plan;  // the caller destroys it (js2t_plan_destroy, or _completed behind its own event)
    return JS2T_OK;
  }
  // the context keeps the plan until the event behind its last kernel has completed
  {
    std::lock_guard<std::mutex> lk(ctx->batch_mu);
    cudaEvent_t ev = nullptr;
    if (!ctx->spare_events.empty()) {
      ev = ctx->spare_events.back();
      ctx->spare_events.pop_back();
    } else if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
      ev = nullptr;
    }
    if (ev == nullptr || cudaEventRecord(ev, stream) != cudaSuccess) {
      cudaStreamSynchronize(stream);
      if (ev) cudaEventDestroy(ev);
      plan_destroy_impl(plan, false);
    } else {
      ctx->retired.emplace_back(ev, plan);
    }
  }
  return JS2T_OK;
}

// SpecAugment mask tables of a whole batch, drawn from the caller's 32-bit generator — an exact replay of the
// reference's per-item draws (joeynmt/data_augmentation.py:48-70: np.random.randint(0, hi) four times per mask
// pair), which numpy's legacy RandomState serves with masked rejection on 32-bit outputs: hi - 1 == 0 consumes
// nothing, otherwise outputs are drawn until (output & mask) <= hi - 1, mask = the smallest 2^k - 1 >= hi - 1.
// next_uint32 / rng_state are numpy's own (BitGenerator.ctypes), so the stream advances exactly as if the
// reference's loop had run.  Pure host code.
int js2t_specaug_replay(js2t_next_uint32_fn next_uint32, void* rng_state, int n_utts, const int32_t* n_frames,
                        int num_freqs, int freq_mask_n, int freq_mask_f, int time_mask_n, int time_mask_t,
                        double time_mask_p, int32_t* table_out) {
  if (next_uint32 == nullptr || n_frames == nullptr || table_out == nullptr)
    return fail(JS2T_ERR_INVALID, "js2t_specaug_replay: NULL argument");
  if (n_utts < 0 || freq_mask_n < 0 || time_mask_n < 0 || freq_mask_f < 1 || num_freqs < freq_mask_f)
    return fail(JS2T_ERR_INVALID, "js2t_specaug_replay: mask configuration outside the replayed range");
  auto bounded = [&](long long hi) -> long long {  // np.random.randint(0, hi), hi >= 1
    const unsigned long long rng = (unsigned long long)(hi - 1);
    if (rng == 0) return 0;
    unsigned long long mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    for (;;) {
      const unsigned long long v = next_uint32(rng_state) & mask;
      if (v <= rng) return (long long)v;
    }
  };
  const int n_masks = freq_mask_n + time_mask_n;
  for (int u = 0; u < n_utts; ++u) {
    int32_t* row = table_out + (size_t)u * n_masks * 2;
    for (int i = 0; i < 2 * n_masks; ++i) row[i] = 0;
    const long long T = n_frames[u];
    if (T <= 0) continue;                    // :48-52 the spectrogram is returned untouched
    for (int m = 0; m < freq_mask_n; ++m) {  // :54-58 (num_freqs - f >= 1 because f < freq_mask_f <= num_freqs)
      const long long f = bounded(freq_mask_f);
      const long long f0 = bounded(num_freqs - f);
      row[2 * m] = (int32_t)f0;
      row[2 * m + 1] = (int32_t)f;
    }
    long long max_t = (long long)floor((double)T * time_mask_p);  // :60-62
    if (time_mask_t < max_t) max_t = time_mask_t;
    if (max_t < 1) continue;                                       // :63-64 frequency masks only
    for (int m = 0; m < time_mask_n; ++m) {                        // :66-70 (T - t >= 1 because t < max_t <= T)
      const long long tt = bounded(max_t);
      const long long t0 = bounded(T - tt);
      row[2 * (freq_mask_n + m)] = (int32_t)t0;
      row[2 * (freq_mask_n + m) + 1] = (int32_t)tt;
    }
  }
  return JS2T_OK;
}

int js2t_reformat_48k_to_16k(js2t_ctx* ctx, const void* src_dev, int is_f32, int64_t n_samples,
                             int16_t* dst_dev, void* workspace_dev, void* stream_) {
  if (ctx == nullptr || workspace_dev == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  if (n_samples < 0 || n_samples % 3 != 0)
    return fail(JS2T_ERR_INVALID, "cannot reshape array of size %lld into shape (-1, 3)", (long long)n_samples);
  if (n_samples == 0) return JS2T_OK;
  if (src_dev == nullptr || dst_dev == nullptr) return fail(JS2T_ERR_INVALID, "NULL argument");
  if ((reinterpret_cast<uintptr_t>(src_dev) & 15) != 0 || (reinterpret_cast<uintptr_t>(dst_dev) & 15) != 0)
    return fail(JS2T_ERR_INVALID, "src_dev and dst_dev must be 16-byte aligned");
  DeviceGuard guard(ctx->device);
  JS2T_CUDA(guard.err);
  JS2T_CUDA(launch_reformat_48k_to_16k(src_dev, is_f32, (long long)n_samples, dst_dev,
                                       static_cast<int*>(workspace_dev), (cudaStream_t)stream_));
  return JS2T_OK;
}

}  // extern "C"
