# coding: utf-8
"""
Batched host API of the B200 front-end: ragged batch of waveforms (or pre-extracted features) in,
``(sum T, 80)`` / ``(B, Tmax, 80)`` CUDA tensor out, through the C ABI of
``include/joeys2t_b200.h``.  torch is used for device memory, pinned staging and streams only.

The per-item wrappers with the reference's signatures (``helpers_for_audio.py``,
``data_augmentation.py``, ``speech_processor.py``) are batch-of-one calls into this module.
"""
import ctypes
import os
import threading
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from joeys2t_b200 import _lib, tables

NUM_MEL = tables.NUM_MEL_BINS
Array = Union[np.ndarray, torch.Tensor]

_contexts = {}
_contexts_lock = threading.Lock()
_owner_pid: Optional[int] = None  # the process that created the device contexts


def _check_not_forked():
    """The reference's data path may run inside ``DataLoader(num_workers > 0)`` workers, which are forked
    (SURVEY 8b).  A CUDA context does not survive ``fork``: a child that inherited this module's contexts would
    fail inside the driver, or hang.  Fail here instead, with what to do about it."""
    if _owner_pid is not None and _owner_pid != os.getpid():
        raise RuntimeError(
            "joeys2t_b200 was initialised in process %d and is now called from its forked child %d: CUDA "
            "contexts do not survive fork().  Use DataLoader(num_workers=0) — the GPU front-end is faster than a "
            "pool of CPU workers — or multiprocessing_context='spawn'." % (_owner_pid, os.getpid()))


def _require_cuda():
    _check_not_forked()  # (before any CUDA query: in a forked child even those may fail)
    if not torch.cuda.is_available():
        raise RuntimeError(
            "joeys2t_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")


def _stream_ptr(stream: Optional[torch.cuda.Stream] = None, device: Optional[int] = None) -> int:
    """Raw ``cudaStream_t`` of ``stream``, or of torch's current stream ON ``device`` (a stream handle of
    another device is an invalid resource handle for the launches of a plan on ``device``)."""
    return (stream or torch.cuda.current_stream(device)).cuda_stream


class Context:
    """One per device: owns the window / twiddle / mel tables (``js2t_ctx``)."""

    def __init__(self, device: int):
        _require_cuda()
        lib = _lib.load()
        self.device = device
        self._h = ctypes.c_void_p()
        _lib.check(lib.js2t_ctx_create(device, ctypes.byref(self._h)))
        win = np.ascontiguousarray(tables.povey_window(), np.float32)
        mel = np.ascontiguousarray(tables.mel_banks(NUM_MEL), np.float32)
        # The mel weights are compile-time immediates and js2t_ctx_set_tables checks the uploaded bank bit for
        # bit.  `mel` comes from this host's float32 log(): if it rounds differently from the build machine's
        # (a last-place difference), upload the compiled-in bank instead of locking the host out — but only
        # for last-place differences; a bank that is really different (rate, bin count, low_freq) still fails.
        ref = np.zeros_like(mel)
        _lib.check(lib.js2t_reference_mel_bank(ref.ctypes.data))
        if not np.array_equal(mel, ref):
            ulp = np.abs(mel.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
            if ((mel == 0) == (ref == 0)).all() and int(ulp.max()) <= 2:
                import warnings
                warnings.warn(f"joeys2t_b200: this host's mel bank differs from the compiled-in one by {int(ulp.max())} "
                              f"ulp in {int((ulp > 0).sum())} weights (float32 log rounding); using the compiled-in bank")
                mel = ref
        _lib.check(lib.js2t_ctx_set_tables(self._h, win.ctypes.data, mel.ctypes.data))

    @property
    def handle(self):
        return self._h

    def __del__(self):
        try:
            if self._h:
                _lib.load().js2t_ctx_destroy(self._h)
        except Exception:  # pylint: disable=broad-except
            pass


def get_context(device: Optional[int] = None) -> Context:
    global _owner_pid
    _check_not_forked()
    _require_cuda()
    if device is None:
        device = torch.cuda.current_device()
    with _contexts_lock:  # per-thread staging implies threaded callers
        if device not in _contexts:
            _contexts[device] = Context(device)
            _owner_pid = os.getpid()
        return _contexts[device]


def bind_host_thread_to_gpu(device: Optional[int] = None) -> bool:
    """Pin the calling thread to the CPU cores next to ``device`` (NVML's ideal affinity), so that
    the pinned staging buffers it allocates afterwards are first-touched on the GPU's own NUMA node
    and host<->device copies do not cross the socket interconnect.  Matters for the host-resident
    (PCIe-bound) path, mostly with several GPUs per box.  Returns False when NVML is unavailable."""
    try:
        import pynvml  # nvidia-ml-py
        dev = torch.cuda.current_device() if device is None else int(device)
        pr = torch.cuda.get_device_properties(dev)
        pynvml.nvmlInit()
        bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:  # pylint: disable=broad-except
        return False


# ------------------------------------------------------------------------------------------
# packing
# ------------------------------------------------------------------------------------------
def _as_1d_pcm(w: Array) -> np.ndarray:
    """→ 1-D int16 or float32 numpy array, channel 0 of multi-channel input (quirk Q1:
    ``helpers_for_audio.py:53-54`` discards the mono mix and torchaudio takes row 0)."""
    if isinstance(w, torch.Tensor):
        w = w.detach().cpu().numpy()
    w = np.asarray(w)
    if w.ndim == 2:
        w = w[0]
    if w.ndim != 1:
        raise ValueError(f"waveform must be (C, N) or (N,), got shape {w.shape}")
    if w.dtype == np.int16:
        return np.ascontiguousarray(w)
    # the kernels compute in float32 (quirk Q3: float64 input is rounded to float32 first)
    return np.ascontiguousarray(w, dtype=np.float32)


_I16, _F32 = np.dtype(np.int16), np.dtype(np.float32)
_addressof, _c_char = ctypes.addressof, ctypes.c_char


class PackedPCM:
    """Ragged batch packed into one pinned byte buffer, every utterance 16-byte aligned."""

    def __init__(self, waveforms: Sequence[Array], host=None):
        """:param host: optional pinned uint8 tensor to pack into (reused staging buffer of a
            streaming caller; must be large enough), or a callable ``nbytes -> tensor`` that provides one"""
        arrs = [_as_1d_pcm(w) for w in waveforms]
        self.n_samples = np.array([a.shape[0] for a in arrs], np.int64)
        self.is_f32 = np.array([a.dtype != np.int16 for a in arrs], np.uint8)
        sizes = np.array([a.nbytes for a in arrs], np.int64)
        aligned = (sizes + 15) // 16 * 16
        self.byte_off = np.concatenate([[0], np.cumsum(aligned)[:-1]]).astype(np.int64)
        self.nbytes = int(aligned.sum())
        if callable(host):
            host = host(max(self.nbytes, 16))
        if host is not None:
            if host.dtype != torch.uint8 or host.numel() < self.nbytes:
                raise ValueError(f"staging buffer too small: {host.numel()} < {self.nbytes} bytes")
            self.host = host
        else:
            pin = torch.cuda.is_available()
            self.host = torch.empty(max(self.nbytes, 16), dtype=torch.uint8, pin_memory=pin)
        # gather on the C side (js2t_pack_pcm: a persistent pool of copy threads) — the numpy slice loop it
        # replaces was a third of a per-batch call
        n = len(arrs)
        ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrs])
        _lib.check(_lib.load().js2t_pack_pcm(n, ptrs, sizes.ctypes.data, self.byte_off.ctypes.data,
                                             self.host.data_ptr(), self.host.numel(), 0))
        self._keep = arrs  # the sources stay alive until the copy above has returned (it has)

    def to_device(self, device=None, non_blocking: bool = True) -> torch.Tensor:
        return self.host.to(device or "cuda", non_blocking=non_blocking)


# ------------------------------------------------------------------------------------------
# plan
# ------------------------------------------------------------------------------------------
class Plan:
    """Geometry + workspace of one ragged batch (``js2t_plan``)."""

    def __init__(self, n_samples=None, byte_off=None, is_f32=None, *, n_frames_in=None,
                 max_frames=None, layout: str = "ragged", pad_tmax: int = 0,
                 pad_value: float = 1.0, device: Optional[int] = None):
        self.ctx = get_context(device)
        lib = _lib.load()
        self._lib = lib
        self._h = ctypes.c_void_p()
        lay = {"ragged": _lib.LAYOUT_RAGGED, "padded": _lib.LAYOUT_PADDED}[layout]
        self.layout = layout
        mf = None if max_frames is None else np.ascontiguousarray(max_frames, np.int32)
        mf_p = None if mf is None else mf.ctypes.data
        if n_frames_in is not None:
            nfi = np.ascontiguousarray(n_frames_in, np.int32)
            self.n_utts = len(nfi)
            _lib.check(lib.js2t_plan_create_features(self.ctx.handle, self.n_utts, nfi.ctypes.data,
                                                     mf_p, lay, int(pad_tmax), float(pad_value),
                                                     ctypes.byref(self._h)))
            self.feature_input = True
        else:
            ns = np.ascontiguousarray(n_samples, np.int64)
            bo = np.ascontiguousarray(byte_off, np.int64)
            f32 = np.ascontiguousarray(
                np.zeros(len(ns), np.uint8) if is_f32 is None else is_f32, np.uint8)
            self.n_utts = len(ns)
            _lib.check(lib.js2t_plan_create(self.ctx.handle, self.n_utts, bo.ctypes.data,
                                            ns.ctypes.data, f32.ctypes.data, mf_p, lay,
                                            int(pad_tmax), float(pad_value),
                                            ctypes.byref(self._h)))
            self.feature_input = False
        self.total_frames = int(lib.js2t_plan_total_frames(self._h))
        self.out_rows = int(lib.js2t_plan_out_rows(self._h))
        nf = np.zeros(self.n_utts, np.int32)
        _lib.check(lib.js2t_plan_get_frames(self._h, nf.ctypes.data))
        self.n_frames = nf
        rows = np.zeros(self.n_utts, np.int64)
        _lib.check(lib.js2t_plan_get_out_rows(self._h, rows.ctypes.data))
        self.out_row = rows
        self.pad_tmax = self.out_rows // self.n_utts if layout == "padded" else 0
        self._keepalive = None

    def _stream(self) -> int:
        """torch's current stream on the PLAN's device (not on whatever device is current)."""
        return _stream_ptr(device=self.ctx.device)

    # ---- configuration ----------------------------------------------------------------------
    def set_cmvn(self, mode: str = "utterance", norm_means: bool = True, norm_vars: bool = True,
                 before: bool = True) -> "Plan":
        m = {"none": _lib.CMVN_NONE, "utterance": _lib.CMVN_UTTERANCE, "global": _lib.CMVN_GLOBAL,
             "stats": _lib.CMVN_STATS_ONLY}[mode]
        _lib.check(self._lib.js2t_plan_set_cmvn(self._h, m, int(norm_means), int(norm_vars),
                                                int(before)))
        return self

    def set_global_stats(self, mean: np.ndarray, istd: np.ndarray) -> "Plan":
        mean = np.ascontiguousarray(mean, np.float64)
        istd = np.ascontiguousarray(istd, np.float64)
        assert mean.shape == (NUM_MEL,) and istd.shape == (NUM_MEL,)
        _lib.check(self._lib.js2t_plan_set_global_stats(self._h, mean.ctypes.data, istd.ctypes.data,
                                                        self._stream()))
        return self

    def set_masks(self, table: Optional[np.ndarray], n_fmask: int = 0, n_tmask: int = 0,
                  mask_value: Optional[float] = None) -> "Plan":
        """``table``: int32 (B, n_fmask + n_tmask, 2) = (start, width), frequency masks first."""
        if table is None:
            _lib.check(self._lib.js2t_plan_set_masks(self._h, 0, 0, None, 0, 0.0, self._stream()))
            return self
        t = np.ascontiguousarray(table, np.int32)
        assert t.shape == (self.n_utts, n_fmask + n_tmask, 2), t.shape
        mode = _lib.MASK_VALUE_MEAN if mask_value is None else _lib.MASK_VALUE_CONST
        # the table is pageable host memory: the C side returns once it has been staged
        _lib.check(self._lib.js2t_plan_set_masks(self._h, n_fmask, n_tmask, t.ctypes.data, mode,
                                                 0.0 if mask_value is None else float(mask_value),
                                                 self._stream()))
        return self

    def set_dither(self, noise: Optional[torch.Tensor]) -> "Plan":
        """Compatibility / test mode (torchaudio ``fbank(dither=d)``, kaldi.py:179-181): ``noise`` is the
        host-drawn ``randn(total_frames, 400) * d`` as a float32 CUDA tensor, added to the frames before DC
        removal.  ``None`` switches it off (the reference's call site never enables dither)."""
        if noise is None:
            self._dither = None
            _lib.check(self._lib.js2t_plan_set_dither(self._h, None))
            return self
        assert noise.is_cuda and noise.dtype == torch.float32 and noise.is_contiguous()
        assert tuple(noise.shape) == (self.total_frames, tables.FRAME_LENGTH), noise.shape
        self._dither = noise  # borrowed by the C side: keep it alive
        _lib.check(self._lib.js2t_plan_set_dither(self._h, noise.data_ptr()))
        return self

    # ---- execution --------------------------------------------------------------------------
    def empty_output(self) -> torch.Tensor:
        shape = (self.n_utts, self.pad_tmax, NUM_MEL) if self.layout == "padded" else \
            (self.out_rows, NUM_MEL)
        return torch.empty(shape, dtype=torch.float32, device=f"cuda:{self.ctx.device}")

    def execute(self, src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """PCM bytes (or feature rows for a feature plan) on the device → features on the device.
        Asynchronous on the current stream."""
        assert src.is_cuda and src.is_contiguous()
        if out is None:
            out = self.empty_output()
        assert out.is_cuda and out.is_contiguous() and out.dtype == torch.float32
        assert out.numel() >= self.out_rows * NUM_MEL
        fn = self._lib.js2t_features_execute if self.feature_input else self._lib.js2t_fbank_execute
        _lib.check(fn(self._h, src.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def set_option(self, name: str, value: int) -> "Plan":
        _lib.check(self._lib.js2t_plan_set_option(self._h, name.encode(), int(value)))
        return self

    def debug_times(self) -> np.ndarray:
        """(n_tiles, 4) uint64 %globaltimer stamps of the last execute (option "debug_times")."""
        n_tiles = int(np.sum((np.maximum(self.n_frames, 1) + 31) // 32)) if self.layout == "ragged" \
            else self.n_utts * ((self.pad_tmax + 31) // 32)
        buf = np.zeros((n_tiles, 4), np.uint64)
        torch.cuda.synchronize()
        _lib.check(self._lib.js2t_plan_debug_times(self._h, buf.ctypes.data, buf.size))
        return buf

    def enable_profiling(self, n_slots: int) -> None:
        _lib.check(self._lib.js2t_plan_enable_profiling(self._h, int(n_slots)))

    def kernel_times_ms(self, n: int) -> np.ndarray:
        """Durations of the fbank kernel of the profiled executes (synchronises on their events)."""
        buf = np.zeros(n, np.float32)
        got = ctypes.c_int(0)
        _lib.check(self._lib.js2t_plan_kernel_times_ms(self._h, buf.ctypes.data, n, ctypes.byref(got)))
        return buf[:got.value]

    def utt_stats(self) -> torch.Tensor:
        """(B, 160) float64 per-utterance sum | sum of squares of the raw log-mel (a device copy)."""
        out = torch.empty((self.n_utts, 2 * NUM_MEL), dtype=torch.float64,
                          device=f"cuda:{self.ctx.device}")
        _lib.check(self._lib.js2t_plan_copy_utt_stats(self._h, out.data_ptr(), self._stream()))
        return out

    def accumulate_global(self, accum: torch.Tensor) -> None:
        """accum (161,) float64 cuda: sum[80] | sumsq[80] | frames  += this batch."""
        assert accum.is_cuda and accum.dtype == torch.float64 and accum.numel() == 2 * NUM_MEL + 1
        _lib.check(self._lib.js2t_global_stats_accumulate(self._h, accum.data_ptr(), self._stream()))

    def finalize_global(self, accum: torch.Tensor) -> None:
        _lib.check(self._lib.js2t_global_stats_finalize(self._h, accum.data_ptr(), self._stream()))

    def global_mean_istd(self) -> torch.Tensor:
        """(160,) float32 on the device: the global mean[80] | 1/std[80] the kernels normalise with."""
        out = torch.empty(2 * NUM_MEL, dtype=torch.float32, device=f"cuda:{self.ctx.device}")
        _lib.check(self._lib.js2t_plan_copy_global_stats(self._h, out.data_ptr(), self._stream()))
        return out

    def normalize(self, out: torch.Tensor) -> torch.Tensor:
        _lib.check(self._lib.js2t_normalize_execute(self._h, out.data_ptr(), self._stream()))
        return out

    def split(self, out: torch.Tensor) -> List[torch.Tensor]:
        """Views of the per-utterance (T_u, 80) blocks of an output tensor."""
        flat = out.reshape(-1, NUM_MEL)
        return [flat[r:r + t] for r, t in zip(self.out_row.tolist(), self.n_frames.tolist())]

    def close(self, completed: bool = False):
        """Destroy the plan.  ``completed=True``: the caller has already waited for the plan's last work
        (an event), so nothing is synchronised — newer batches on the same stream keep running."""
        if self._h:
            (self._lib.js2t_plan_destroy_completed if completed else self._lib.js2t_plan_destroy)(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass


# ------------------------------------------------------------------------------------------
# one-call batched entry point
# ------------------------------------------------------------------------------------------
_tls = threading.local()
_RING = 4          # pinned staging buffers per calling thread
_MAX_PENDING = 8   # plans whose work may still be in flight before the oldest is waited for


class _Staging:
    """Per-thread ring of grow-only pinned staging buffers.  The one-call entry points below return as
    soon as their work is ENQUEUED (the output tensor is ordered on the caller's stream like any torch
    result): a slot is reused only after the H2D copy that read it has completed (an event per slot), and
    a plan is destroyed only after the kernels that use its workspace have (``pending``).  Allocating
    pinned memory, creating / destroying a plan and synchronising per call used to cost 60x the kernels
    of a 20 000-frame batch."""

    def __init__(self):
        self.bufs = [None] * _RING
        self.events = [None] * _RING
        self.next = 0
        self.pending = []  # (event, plan, device tensors to keep alive)

    def acquire(self, nbytes: int) -> Tuple[int, torch.Tensor]:
        j = self.next
        self.next = (j + 1) % _RING
        if self.events[j] is not None:
            self.events[j].synchronize()  # normally long complete
        buf = self.bufs[j]
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, pin_memory=True)
            self.bufs[j] = buf
        return j, buf

    def release_after(self, j: int, stream: torch.cuda.Stream) -> None:
        ev = self.events[j] or torch.cuda.Event()
        ev.record(stream)
        self.events[j] = ev

    def retire(self, plan: "Plan", keep, stream: torch.cuda.Stream) -> None:
        ev = torch.cuda.Event()
        ev.record(stream)
        self.pending.append((ev, plan, keep))
        self.reap(limit=_MAX_PENDING)

    def reap(self, limit: int = 0) -> None:
        while self.pending and (len(self.pending) > limit or self.pending[0][0].query()):
            ev, plan, _ = self.pending.pop(0)
            ev.synchronize()
            plan.close(completed=True)


def _staging_state() -> _Staging:
    st = getattr(_tls, "state", None)
    if st is None:
        st = _tls.state = _Staging()
    return st


def fbank_cmvn_specaug_ragged(
    waveforms: Sequence[Array],
    *,
    cmvn: Optional[dict] = None,
    masks: Optional[np.ndarray] = None,
    n_fmask: int = 0,
    n_tmask: int = 0,
    mask_value: Optional[float] = None,
    max_frames: Optional[Sequence[int]] = None,
    layout: str = "ragged",
    pad_value: float = 1.0,
    global_stats: Optional[Tuple[np.ndarray, np.ndarray]] = None,
    dither_noise: Optional[Array] = None,
) -> Tuple[torch.Tensor, np.ndarray]:
    """Whole batch in one call: pack → H2D → fbank [→ CMVN] [→ SpecAugment] on the GPU.

    :param waveforms: list of (C, N) / (N,) arrays; int16 PCM, or float in [-1, 1) as produced by
        ``torchaudio.load`` (scaled by 2**15 on the device, ``helpers_for_audio.py:54``)
    :param cmvn: ``{"norm_means", "norm_vars", "before"}`` like the reference's ``CMVN(**cfg)``;
        ``None`` = raw log-mel
    :param masks: int32 (B, n_fmask + n_tmask, 2) host-drawn SpecAugment table (see
        :func:`joeys2t_b200.data_augmentation.draw_masks`)
    :param dither_noise: compatibility / test mode — host-drawn ``randn(sum T, 400) * dither`` (float32),
        see :meth:`Plan.set_dither`; ``None`` = the reference's ``dither = 0``
    :returns: (features on the GPU — ragged ``(sum T, 80)`` or padded ``(B, Tmax, 80)`` —, n_frames)
    """
    _require_cuda()
    if dither_noise is None:
        return _batch_fbank(waveforms, cmvn, masks, n_fmask, n_tmask, mask_value, max_frames, layout,
                            pad_value, global_stats)
    # dither compatibility / test mode: the step-by-step route (plan, separate uploads)
    st = _staging_state()
    st.reap()
    slot = []

    def take(nbytes):
        j, buf = st.acquire(nbytes)
        slot.append(j)
        return buf

    packed = PackedPCM(waveforms, host=take)
    plan = Plan(packed.n_samples, packed.byte_off, packed.is_f32, max_frames=max_frames,
                layout=layout, pad_value=pad_value)
    if global_stats is not None:
        kw = dict(cmvn or {})
        plan.set_cmvn("global", kw.get("norm_means", True), kw.get("norm_vars", True),
                      kw.get("before", True))
        plan.set_global_stats(*global_stats)
    elif cmvn is not None:
        plan.set_cmvn("utterance", cmvn.get("norm_means", True), cmvn.get("norm_vars", True),
                      cmvn.get("before", True))
    if masks is not None:
        plan.set_masks(masks, n_fmask, n_tmask, mask_value)
    noise_dev = torch.as_tensor(dither_noise, dtype=torch.float32).contiguous().to(
        f"cuda:{plan.ctx.device}")
    plan.set_dither(noise_dev)
    stream = torch.cuda.current_stream(plan.ctx.device)
    dev_pcm = packed.host[:max(packed.nbytes, 16)].to(f"cuda:{plan.ctx.device}", non_blocking=True)
    st.release_after(slot[0], stream)   # the staging slot is free again once the H2D copy has read it
    out = plan.execute(dev_pcm)
    n_frames = plan.n_frames.copy()
    # asynchronous return: the plan's workspace and the device PCM outlive the enqueued kernels (retired
    # behind an event on this stream); `out` is ordered on the current stream like any torch result
    st.retire(plan, (dev_pcm, noise_dev), stream)
    return out, n_frames


def _batch_fbank(waveforms, cmvn, masks, n_fmask, n_tmask, mask_value, max_frames, layout, pad_value,
                 global_stats, device: Optional[int] = None) -> Tuple[torch.Tensor, np.ndarray]:
    """The whole batch through ONE C call (``js2t_batch_fbank``): gather into a pinned staging slot of the
    context, one H2D transfer for PCM + descriptors + masks + statistics, three kernel launches; returns as
    soon as the work is enqueued on torch's current stream (``out`` is ordered on it like any torch result)."""
    ctx = get_context(device)
    lib = _lib.load()
    i16 = _I16
    # (fast path for what the callers hand over: contiguous 1-D int16 / float32 numpy arrays)
    arrs = [w if (type(w) is np.ndarray and w.ndim == 1 and (w.dtype == i16 or w.dtype == _F32)
                  and w.flags.c_contiguous) else _as_1d_pcm(w) for w in waveforms]
    n = len(arrs)
    n_samples = np.array([a.shape[0] for a in arrs], np.int64)
    is_f32 = np.array([a.dtype != i16 for a in arrs], np.uint8)
    # (a too-short utterance gives 0 frames here; the C call rejects it with JS2T_ERR_SHORT_INPUT before
    # anything is enqueued)
    n_frames = np.where(n_samples >= 400, 1 + (n_samples - 400) // 160, 0).astype(np.int32)
    mf = None
    if max_frames is not None:
        mf = np.ascontiguousarray(max_frames, np.int32)
        n_frames = np.where(mf > 0, np.minimum(n_frames, mf), n_frames).astype(np.int32)
    padded = layout == "padded"
    tmax = int(n_frames.max()) if n else 0
    rows = n * tmax if padded else int(n_frames.sum())
    dev = f"cuda:{ctx.device}"
    out = torch.empty((n, tmax, NUM_MEL) if padded else (rows, NUM_MEL), dtype=torch.float32, device=dev)
    o = _lib.BatchOpts()
    o.layout = {"ragged": _lib.LAYOUT_RAGGED, "padded": _lib.LAYOUT_PADDED}[layout]
    o.pad_tmax = 0
    o.pad_value = float(pad_value)
    keep = [arrs, n_samples, is_f32, mf]
    if global_stats is not None:
        kw = dict(cmvn or {})
        gm = np.ascontiguousarray(global_stats[0], np.float64)
        gi = np.ascontiguousarray(global_stats[1], np.float64)
        assert gm.shape == (NUM_MEL,) and gi.shape == (NUM_MEL,)
        keep += [gm, gi]
        o.cmvn_mode = _lib.CMVN_GLOBAL
        o.global_mean80, o.global_istd80 = gm.ctypes.data, gi.ctypes.data
    else:
        kw = cmvn or {}
        o.cmvn_mode = _lib.CMVN_NONE if cmvn is None else _lib.CMVN_UTTERANCE
    o.norm_means, o.norm_vars = int(kw.get("norm_means", True)), int(kw.get("norm_vars", True))
    o.before = int(kw.get("before", True))
    if masks is not None and n_fmask + n_tmask > 0:
        tb = np.ascontiguousarray(masks, np.int32)
        assert tb.shape == (n, n_fmask + n_tmask, 2), tb.shape
        keep.append(tb)
        o.n_fmask, o.n_tmask, o.mask_table = int(n_fmask), int(n_tmask), tb.ctypes.data
        o.mask_value_mode = _lib.MASK_VALUE_MEAN if mask_value is None else _lib.MASK_VALUE_CONST
        o.mask_value_const = 0.0 if mask_value is None else float(mask_value)
    if mf is not None:
        o.max_frames = mf.ctypes.data
    try:  # (the cheapest way to an ndarray's address; read-only arrays take the slower attribute)
        ptrs = np.array([_addressof(_c_char.from_buffer(a)) for a in arrs], np.uint64)
    except (TypeError, ValueError):
        ptrs = np.array([a.ctypes.data for a in arrs], np.uint64)
    _lib.check(lib.js2t_batch_fbank(ctx.handle, n, ptrs.ctypes.data, n_samples.ctypes.data, is_f32.ctypes.data,
                                    ctypes.byref(o), out.data_ptr(), rows, _stream_ptr(device=ctx.device),
                                    None))
    del keep  # everything the call read from the host has been copied when it returns
    return out, n_frames


def features_cmvn_specaug_ragged(
    feats: Sequence[np.ndarray],
    *,
    cmvn: Optional[dict] = None,
    masks: Optional[np.ndarray] = None,
    n_fmask: int = 0,
    n_tmask: int = 0,
    mask_value: Optional[float] = None,
    max_frames: Optional[Sequence[int]] = None,
    layout: str = "ragged",
    pad_value: float = 1.0,
) -> Tuple[torch.Tensor, np.ndarray]:
    """Same for pre-extracted (T, 80) feature matrices (the .npy / zip branch of ``get_features``)."""
    arrs = [np.ascontiguousarray(f, np.float32) for f in feats]
    for a in arrs:
        if a.ndim != 2 or a.shape[1] != NUM_MEL:
            raise ValueError(f"features must be (T, {NUM_MEL}); got {a.shape}")
    n_in = np.array([a.shape[0] for a in arrs], np.int32)
    host = torch.empty((int(n_in.sum()), NUM_MEL), dtype=torch.float32, pin_memory=True)
    np.concatenate(arrs, 0, out=host.numpy())
    plan = Plan(n_frames_in=n_in, max_frames=max_frames, layout=layout, pad_value=pad_value)
    if cmvn is not None:
        plan.set_cmvn("utterance", cmvn.get("norm_means", True), cmvn.get("norm_vars", True),
                      cmvn.get("before", True))
    if masks is not None:
        plan.set_masks(masks, n_fmask, n_tmask, mask_value)
    out = plan.execute(host.to(f"cuda:{plan.ctx.device}", non_blocking=True))
    n_frames = plan.n_frames.copy()
    torch.cuda.current_stream().synchronize()
    plan.close()
    return out, n_frames


# ------------------------------------------------------------------------------------------
# PCM ingest: 48 kHz -> 16 kHz
# ------------------------------------------------------------------------------------------
def reformat_48k_to_16k(y: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Device-side ``reformat_freq`` for ``sr == 48000`` (``scripts/gradio_demo.py:35-45``):
    ``((y / max(max(y), 1)) * 32767).reshape(-1, 3).mean(1).astype(int16)`` — bit-identical to
    numpy (float64 arithmetic for int16 input, float32 for float32 input).

    :param y: 1-D CUDA tensor, int16 or float32, length divisible by 3
    :returns: int16 CUDA tensor of length ``len(y) // 3``
    """
    _require_cuda()
    if not y.is_cuda or y.dim() != 1 or y.dtype not in (torch.int16, torch.float32):
        raise ValueError("reformat_48k_to_16k needs a 1-D CUDA tensor of int16 or float32 samples")
    n = y.numel()
    if n % 3 != 0:
        raise ValueError(f"cannot reshape array of size {n} into shape (-1, 3)")
    y = y.contiguous()
    ctx = get_context(y.device.index)
    if out is None:
        out = torch.empty(n // 3, dtype=torch.int16, device=y.device)
    ws = torch.empty(2, dtype=torch.int32, device=y.device)
    lib = _lib.load()
    with torch.cuda.device(y.device):
        _lib.check(lib.js2t_reformat_48k_to_16k(
            ctx.handle, y.data_ptr(), int(y.dtype == torch.float32), n, out.data_ptr(),
            ws.data_ptr(), _stream_ptr(device=y.device.index)))
    return out


# ------------------------------------------------------------------------------------------
# host-resident streaming: H2D, kernels and D2H of consecutive batches overlap
# ------------------------------------------------------------------------------------------
class HostPipeline:
    """Host PCM in, host features out, for a stream of batches.

    Each slot owns a device staging buffer, a device output and a pinned host output.  Three CUDA
    streams (copy-in, compute, copy-out) are chained with events so that the H2D copy of batch
    ``i+1``, the kernels of batch ``i`` and the D2H copy of batch ``i-1`` run concurrently — from
    host memory the path is PCIe-bound, and PCIe is full duplex.

    ``submit(packed, plan)`` returns the slot index; ``result(slot)`` blocks until that slot's
    features are in pinned host memory and returns them as a (rows, 80) float32 tensor.
    """

    def __init__(self, n_slots: int, max_pcm_bytes: int, max_out_rows: int, device: Optional[int] = None):
        _require_cuda()
        self.device = torch.cuda.current_device() if device is None else device
        dev = torch.device("cuda", self.device)
        self.n_slots = n_slots
        self.s_in, self.s_compute, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.stage = [torch.empty(max_pcm_bytes, dtype=torch.uint8, device=dev) for _ in range(n_slots)]
        self.out = [torch.empty((max_out_rows, NUM_MEL), dtype=torch.float32, device=dev)
                    for _ in range(n_slots)]
        self.host_out = [torch.empty((max_out_rows, NUM_MEL), dtype=torch.float32, pin_memory=True)
                         for _ in range(n_slots)]
        self.ev_in = [torch.cuda.Event() for _ in range(n_slots)]
        self.ev_compute = [torch.cuda.Event() for _ in range(n_slots)]
        self.ev_out = [torch.cuda.Event() for _ in range(n_slots)]
        self.rows = [0] * n_slots
        self._next = 0
        self._used = [False] * n_slots

    def submit(self, packed: "PackedPCM", plan: "Plan") -> int:
        j = self._next
        self._next = (j + 1) % self.n_slots
        assert packed.nbytes <= self.stage[j].numel() and plan.out_rows <= self.out[j].shape[0]
        with torch.cuda.stream(self.s_in):
            if self._used[j]:
                self.s_in.wait_event(self.ev_compute[j])  # the slot's previous kernels have read the stage
            self.stage[j][:packed.nbytes].copy_(packed.host[:packed.nbytes], non_blocking=True)
            self.ev_in[j].record(self.s_in)
        with torch.cuda.stream(self.s_compute):
            self.s_compute.wait_event(self.ev_in[j])
            if self._used[j]:
                self.s_compute.wait_event(self.ev_out[j])  # the slot's previous output has left the device
            plan.execute(self.stage[j], self.out[j])
            self.ev_compute[j].record(self.s_compute)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_compute[j])
            self.host_out[j][:plan.out_rows].copy_(self.out[j][:plan.out_rows], non_blocking=True)
            self.ev_out[j].record(self.s_out)
        self.rows[j] = plan.out_rows
        self._used[j] = True
        return j

    def result(self, slot: int) -> torch.Tensor:
        self.ev_out[slot].synchronize()
        return self.host_out[slot][:self.rows[slot]]

    def synchronize(self) -> None:
        for s in (self.s_in, self.s_compute, self.s_out):
            s.synchronize()
