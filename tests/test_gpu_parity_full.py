# coding: utf-8
"""
Full-size parity of the CUDA path on BASELINE.json's own configs — EVERY utterance of configs 2, 3 and 4
against the reference's CPU path (``oracle/ref_cpu_path``: the unmodified reference when ``oracle/_ref`` is
installed, else torchaudio's ``kaldi.fbank`` + restated glue), computed on all host cores — plus the
fused global-CMVN + SpecAugment epilogue (config 3's global variant) and the remaining drop-in entry
points.  ``-m gpu`` only.

Tolerances (BASELINE.json north_star / SURVEY.md §8c):
  * raw log-mel:            max |Δ| <= 1e-3
  * CMVN-normalised output: |Δ| <= 5e-4 + 1e-4 |ref|
  * SpecAugment:            masked-cell positions bit-exact, fill value within 1e-6
"""
import argparse
import multiprocessing as mp
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import fbank_numpy as O  # noqa: E402

LOGMEL_ATOL = 1e-3
MUSTC_SA = dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=1.0)


@pytest.fixture(scope="module")
def fe():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    from joeys2t_b200 import _lib, frontend
    if _lib.is_stale():
        _lib.build()
    return frontend


def reference_logmel(waves):
    """Raw log-mel of every utterance by the reference CPU path, one process per host core (spawned: the
    test process holds a CUDA context and torch thread pools that must not be forked)."""
    from oracle import ref_cpu_path
    cores = min(os.cpu_count() or 1, 32)
    with mp.get_context("spawn").Pool(cores, initializer=ref_cpu_path.single_thread) as pool:
        return pool.map(ref_cpu_path.fbank, waves, chunksize=max(1, len(waves) // (4 * cores)))


def tables_to_masks(table, nf, u):
    fm = [(int(a), int(b)) for a, b in table[u, :nf]]
    tm = [(int(a), int(b)) for a, b in table[u, nf:]]
    return fm, tm


def assert_cmvn_close(got, ref, what):
    err = np.abs(got - ref)
    tol = 5e-4 + 1e-4 * np.abs(ref)
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} cells out of tolerance, max err {err.max():.3e}"


def _bench_batch(workload, seed):
    import bench
    return bench.make_batch(argparse.Namespace(workload=workload, utts=0, sweep_hours=0.0), seed)


# ------------------------------------------------------------------------------------------------
# configs 2, 3, 4 at BASELINE.json's full sizes: every utterance, every cell
# ------------------------------------------------------------------------------------------------
def test_config2_every_utterance_vs_reference_cpu_path(fe):
    """256 x U(10,15) s int16 (the headline workload of bench.py, same generator and seed): raw log-mel and
    utterance CMVN of all 256 utterances (~319 k frames)."""
    waves = _bench_batch("cfg2", 1234)
    assert len(waves) == 256
    refs = reference_logmel(waves)
    raw, nf = fe.fbank_cmvn_specaug_ragged(waves)
    out, nf2 = fe.fbank_cmvn_specaug_ragged(waves, cmvn={})
    raw, out = raw.cpu().numpy(), out.cpu().numpy()
    assert nf.tolist() == [r.shape[0] for r in refs] == nf2.tolist()
    off, worst = 0, 0.0
    for u, ref in enumerate(refs):
        t = ref.shape[0]
        d = float(np.abs(raw[off:off + t] - ref).max())
        assert d <= LOGMEL_ATOL, f"utterance {u}: log-mel max err {d:.3e}"
        worst = max(worst, d)
        assert_cmvn_close(out[off:off + t], O.cmvn(ref), f"utterance {u}")
        off += t
    assert off == raw.shape[0]
    print(f"config 2: 256 utterances, {off} frames, worst log-mel |err| {worst:.2e}")


@pytest.mark.parametrize("variant", ["utterance", "global"])
def test_config3_every_utterance_with_specaugment(fe, variant):
    """512 ragged utterances (1-30 s) + SpecAugment (configs/mustc_st.yaml:23-28): utterance CMVN with the
    mean fill (the reference's mustc_st.yaml:29-32), and the global-CMVN extension with a constant fill —
    the fused fbank epilogue bench.py's cfg3g times."""
    from joeys2t_b200.data_augmentation import SpecAugment, mask_tables_for_batch
    waves = _bench_batch("cfg3", 2345)
    assert len(waves) == 512
    refs = reference_logmel(waves)
    n_frames = [r.shape[0] for r in refs]
    np.random.seed(2345)
    table, nfm, ntm = mask_tables_for_batch(SpecAugment(**MUSTC_SA), n_frames)
    if variant == "utterance":
        out, nf = fe.fbank_cmvn_specaug_ragged(waves, cmvn={}, masks=table, n_fmask=nfm, n_tmask=ntm)
        want = [O.apply_specaugment_masks(O.cmvn(r), tables_to_masks(table, nfm, u)) for u, r in enumerate(refs)]
        # fill value = spectrogram.mean() of the CMVN output (data_augmentation.py:45-46), a number ~1e-8.  The
        # reference's own float32 mean of its float32-normalised rows wanders by ~1e-6 at 3 000 frames, so the
        # arbiter for the 1e-6 gate is the float64 restatement of the same formula; the reference's float32 value
        # is checked at 5e-6
        fill64 = [float(O.cmvn_fp64(r).astype(np.float64).mean()) for r in refs]
    else:
        s, q, n = O.global_cmvn_stats(refs)
        mean = s / n
        istd = 1.0 / np.sqrt(np.maximum(q / n - mean**2, 1e-10))
        out, nf = fe.fbank_cmvn_specaug_ragged(waves, cmvn={}, global_stats=(mean, istd), masks=table,
                                               n_fmask=nfm, n_tmask=ntm, mask_value=0.0)
        want = [O.apply_specaugment_masks(g, tables_to_masks(table, nfm, u), mask_value=0.0)
                for u, g in enumerate(O.global_cmvn(refs))]
        fill64 = [0.0] * len(refs)
    assert nf.tolist() == n_frames
    got = out.cpu().numpy()
    off = 0
    for u, w in enumerate(want):
        t = n_frames[u]
        blk = got[off:off + t]
        off += t
        m = np.zeros((t, 80), bool)
        for f0, f in table[u, :nfm]:
            m[:, f0:f0 + f] = True
        for t0, tt in table[u, nfm:]:
            m[t0:t0 + tt, :] = True
        # masked cells: exactly the drawn rectangles, one fill value per utterance (bit-identical cells)
        if m.any():
            vals = blk[m]
            assert np.ptp(vals) == 0.0, u
            assert abs(float(vals[0]) - fill64[u]) <= 1e-6, u
            assert abs(float(vals[0]) - float(w[m][0])) <= 5e-6, u
        assert_cmvn_close(blk[~m], w[~m], f"{variant} utterance {u}")


def test_config4_every_utterance_longform_mixed_dtype(fe):
    """64 x U(30,60) s, even utterances int16 / odd float32 in [-1, 1): the two-slot staging path."""
    waves = _bench_batch("cfg4", 3456)
    assert len(waves) == 64 and waves[0].dtype == np.int16 and waves[1].dtype == np.float32
    refs = reference_logmel(waves)
    raw, nf = fe.fbank_cmvn_specaug_ragged(waves)
    out, _ = fe.fbank_cmvn_specaug_ragged(waves, cmvn={})
    raw, out = raw.cpu().numpy(), out.cpu().numpy()
    assert nf.tolist() == [r.shape[0] for r in refs]
    off = 0
    for u, ref in enumerate(refs):
        t = ref.shape[0]
        assert np.abs(raw[off:off + t] - ref).max() <= LOGMEL_ATOL, u
        # the reference CMVN's own float32 accumulation error grows to ~1.3e-4 at these lengths
        # (SURVEY.md §8c); the float64 restatement of the same formula is the arbiter
        assert_cmvn_close(out[off:off + t], O.cmvn_fp64(ref), f"utterance {u}")
        off += t


# ------------------------------------------------------------------------------------------------
# fused global CMVN + SpecAugment epilogue (kEpiNormKnown) and Plan.normalize with masks
# ------------------------------------------------------------------------------------------------
def _global_setup(fixtures_pcm, ref_fbank, mixed):
    pcm, _ = fixtures_pcm
    waves = [x.astype(np.float32) / np.float32(32768.0) if (mixed and i % 2) else x for i, x in enumerate(pcm)]
    s, q, n = O.global_cmvn_stats(ref_fbank)
    mean = s / n
    istd = 1.0 / np.sqrt(np.maximum(q / n - mean**2, 1e-10))
    return waves, mean, istd, O.global_cmvn(ref_fbank)


@pytest.mark.parametrize("layout", ["ragged", "padded"])
@pytest.mark.parametrize("n_f,n_t", [(2, 2), (4, 12), (5, 14)])
def test_global_cmvn_with_constant_fill_masks(fe, fixtures_pcm, ref_fbank, layout, n_f, n_t):
    """global statistics known + masks + constant fill: <= 16 masks run in the fbank epilogue (single
    pass), more take the three-kernel path (capi.cu run_pipeline) — both against
    O.global_cmvn + O.apply_specaugment_masks (settings of configs/mustc_st.yaml:23-32)."""
    from joeys2t_b200.data_augmentation import SpecAugment, mask_tables_for_batch
    waves, mean, istd, want = _global_setup(fixtures_pcm, ref_fbank, mixed=True)
    n_frames = [f.shape[0] for f in ref_fbank]
    np.random.seed(7 + n_f)
    sa = SpecAugment(freq_mask_n=n_f, freq_mask_f=27, time_mask_n=n_t, time_mask_t=100, time_mask_p=1.0)
    table, nfm, ntm = mask_tables_for_batch(sa, n_frames)
    fill = -0.25
    out, nf = fe.fbank_cmvn_specaug_ragged(waves, cmvn={}, global_stats=(mean, istd), masks=table, n_fmask=nfm,
                                           n_tmask=ntm, mask_value=fill, layout=layout)
    got = out.cpu().numpy()
    off = 0
    for u, g in enumerate(want):
        t = n_frames[u]
        blk = got[u, :t] if layout == "padded" else got[off:off + t]
        off += t
        ref = O.apply_specaugment_masks(g, tables_to_masks(table, nfm, u), mask_value=fill)
        m = ref != g  # cells the oracle filled (a cell already equal to the fill value stays equal)
        assert (blk[m] == np.float32(fill)).all(), u
        assert_cmvn_close(blk[~m], ref[~m], f"utterance {u}")
        if layout == "padded":
            assert (got[u, t:] == 1.0).all()


@pytest.mark.parametrize("fill", [None, 0.5])
def test_two_pass_normalize_with_masks(fe, fixtures_pcm, ref_fbank, fill):
    """STATS_ONLY pass -> global finalize -> Plan.normalize (js2t_normalize_execute) with masks: mean fill
    (= mean of the globally normalised utterance, data_augmentation.py:45-46) and constant fill."""
    from joeys2t_b200 import distributed as D
    from joeys2t_b200.data_augmentation import SpecAugment, mask_tables_for_batch
    waves, _, _, want = _global_setup(fixtures_pcm, ref_fbank, mixed=False)
    n_frames = [f.shape[0] for f in ref_fbank]
    np.random.seed(11)
    table, nfm, ntm = mask_tables_for_batch(SpecAugment(**MUSTC_SA), n_frames)
    packed = fe.PackedPCM(waves)
    plan = fe.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
    plan.set_cmvn("stats")
    raw = plan.execute(packed.to_device())
    accum = D.new_accumulator("cuda")
    plan.accumulate_global(accum)
    plan.set_cmvn("global")
    plan.finalize_global(accum)
    plan.set_masks(table, nfm, ntm, mask_value=fill)
    got = plan.normalize(raw).cpu().numpy()
    torch.cuda.synchronize()
    plan.close()
    off = 0
    for u, g in enumerate(want):
        t = n_frames[u]
        blk = got[off:off + t]
        off += t
        ref = O.apply_specaugment_masks(g, tables_to_masks(table, nfm, u), mask_value=fill)
        m = np.zeros((t, 80), bool)
        for f0, f in table[u, :nfm]:
            m[:, f0:f0 + f] = True
        for t0, tt in table[u, nfm:]:
            m[t0:t0 + tt, :] = True
        if m.any():
            assert np.ptp(blk[m]) == 0.0
            want_fill = float(g.mean()) if fill is None else fill
            assert abs(float(blk[m][0]) - want_fill) <= (2e-5 if fill is None else 0.0), u
        assert_cmvn_close(blk[~m], ref[~m], f"utterance {u}")


# ------------------------------------------------------------------------------------------------
# remaining drop-in entry points
# ------------------------------------------------------------------------------------------------
def test_get_torchaudio_fbank_takes_int16_range_floats(fe, fixtures_pcm, ref_fbank):
    """helpers_for_audio.py:30-37: the reference calls it with the waveform already scaled by 2**15
    (helpers_for_audio.py:54); the drop-in undoes that scaling exactly before the device path."""
    from joeys2t_b200.helpers_for_audio import _get_torchaudio_fbank, extract_fbank_features
    pcm, _ = fixtures_pcm
    for i in (0, 6):
        scaled = torch.from_numpy(pcm[i].astype(np.float32))[None]  # == (int16 / 32768) * 2**15
        got = _get_torchaudio_fbank(scaled, 16000, n_bins=80)
        assert got.dtype == np.float32 and got.shape == ref_fbank[i].shape
        assert np.abs(got - ref_fbank[i]).max() <= LOGMEL_ATOL
        # same bits as the public entry point on the unscaled waveform
        direct = extract_fbank_features(scaled * 2.0**-15, 16000)
        assert np.array_equal(got, direct)
    with pytest.raises(ValueError):
        _get_torchaudio_fbank(torch.zeros(1, 1000), 8000)


# ------------------------------------------------------------------------------------------------
# dither compatibility mode (kaldi.py:179-181; the reference's call site leaves dither at 0)
# ------------------------------------------------------------------------------------------------
def test_dither_compat_mode_vs_torchaudio(fe, fixtures_pcm):
    """north_star: "Dither ... come[s] from the host so results are reproducible".  torchaudio draws
    randn(T, 400) per utterance and adds randn * dither to the frames before DC removal; the same noise, drawn
    on the host under the same seed, goes through js2t_plan_set_dither: int16 and float32 PCM, an odd frame
    count (a partner frame past the end), one utterance that spans several tiles."""
    ta_kaldi = pytest.importorskip("torchaudio.compliance.kaldi")
    pcm, _ = fixtures_pcm
    waves = [pcm[0], pcm[3].astype(np.float32) / np.float32(32768.0), pcm[6], pcm[1][:400 + 160 * 2], pcm[2]]
    dither = 1.0
    refs, noises = [], []
    for i, w in enumerate(waves):
        x = torch.from_numpy(w.astype(np.float32) if w.dtype == np.int16 else w * np.float32(32768.0))[None]
        torch.manual_seed(100 + i)
        refs.append(ta_kaldi.fbank(x, num_mel_bins=80, sample_frequency=16000, dither=dither).numpy())
        torch.manual_seed(100 + i)
        noises.append((torch.randn(refs[-1].shape[0], 400) * dither).numpy())
    noise = np.concatenate(noises, 0)
    out, nf = fe.fbank_cmvn_specaug_ragged(waves, dither_noise=noise)
    got = out.cpu().numpy()
    assert nf.tolist() == [r.shape[0] for r in refs]
    plain, _ = fe.fbank_cmvn_specaug_ragged(waves)
    plain = plain.cpu().numpy()
    off = 0
    for u, ref in enumerate(refs):
        t = ref.shape[0]
        assert np.abs(got[off:off + t] - ref).max() <= LOGMEL_ATOL, u
        assert np.abs(plain[off:off + t] - ref).max() > 1e-2, u  # the noise really went in
        off += t
    # with utterance CMVN on top (three-kernel path) and in the padded layout
    outc, _ = fe.fbank_cmvn_specaug_ragged(waves, cmvn={}, dither_noise=noise, layout="padded")
    outc = outc.cpu().numpy()
    for u, ref in enumerate(refs):
        if ref.shape[0] >= 100:  # CMVN over the 3-frame utterance divides by near-zero stds: ill-conditioned
            assert_cmvn_close(outc[u, :ref.shape[0]], O.cmvn(ref), f"utterance {u}")
        assert (outc[u, ref.shape[0]:] == 1.0).all()
    # zero noise == dither off, bit for bit
    zero, _ = fe.fbank_cmvn_specaug_ragged(waves, dither_noise=np.zeros_like(noise))
    assert np.array_equal(zero.cpu().numpy(), plain)


# ---- js2t_batch_fbank: the one-call per-batch entry point ------------------------------------------------
def _mixed_batch(rng, lens):
    waves = []
    for i, n in enumerate(lens):
        w = (np.cumsum(rng.randint(-400, 401, n)) % 20001 - 10000).astype(np.int16)
        waves.append(w.astype(np.float32) / np.float32(32768.0) if i % 3 == 1 else w)
    return waves


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["ragged", "padded"])
@pytest.mark.parametrize("mode", ["raw", "utterance", "utterance_after_masks", "global_masks"])
def test_batch_call_matches_the_step_by_step_route(layout, mode):
    """``js2t_batch_fbank`` (host waveforms in, one transfer, kernels enqueued: what
    ``frontend.fbank_cmvn_specaug_ragged`` and through it ``SpeechProcessor.process_batch`` /
    ``SpeechBatchCollator`` call) must give bit for bit what the explicit route gives — ``js2t_pack_pcm`` ->
    ``js2t_plan_create`` -> ``js2t_plan_set_*`` -> ``js2t_fbank_execute`` — which the tests above pin to the
    oracle (tokenizers.py:458-494 order: truncation -> CMVN -> SpecAugment, pad_features :130-170)."""
    import torch
    from joeys2t_b200 import frontend

    rng = np.random.RandomState(11)
    lens = [400, 559, 400 + 160 * 31, 400 + 160 * 32, 5521, 16000 * 2 + 17, 16000 * 5, 16000 * 9 + 3]
    waves = _mixed_batch(rng, lens)
    max_frames = [0, 0, 20, 0, 0, 150, 0, 0]
    T = np.array([1 + (n - 400) // 160 for n in lens])
    T = np.where(np.array(max_frames) > 0, np.minimum(T, max_frames), T)
    n_f, n_t = 2, 2
    table = np.zeros((len(waves), n_f + n_t, 2), np.int32)
    for u in range(len(waves)):
        table[u, :n_f, 0] = rng.randint(0, 70, n_f)
        table[u, :n_f, 1] = rng.randint(0, 12, n_f)
        table[u, n_f:, 0] = rng.randint(0, max(int(T[u]), 1), n_t)
        table[u, n_f:, 1] = rng.randint(0, 30, n_t)
    gstats = (rng.uniform(-8, 2, 80), rng.uniform(0.2, 1.5, 80))
    kw = dict(max_frames=max_frames, layout=layout, pad_value=1.0)
    if mode == "utterance":
        kw.update(cmvn=dict(norm_means=True, norm_vars=True, before=True))
    elif mode == "utterance_after_masks":
        kw.update(cmvn=dict(norm_means=True, norm_vars=False, before=False), masks=table, n_fmask=n_f, n_tmask=n_t)
    elif mode == "global_masks":
        kw.update(cmvn=dict(norm_means=True, norm_vars=True, before=True), global_stats=gstats, masks=table,
                  n_fmask=n_f, n_tmask=n_t, mask_value=0.0)
    got, n_frames = frontend.fbank_cmvn_specaug_ragged(waves, **kw)
    assert n_frames.tolist() == T.tolist()

    packed = frontend.PackedPCM(waves)
    plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32, max_frames=max_frames, layout=layout,
                         pad_value=1.0)
    if mode == "global_masks":
        plan.set_cmvn("global", True, True, True)
        plan.set_global_stats(*gstats)
    elif mode != "raw":
        plan.set_cmvn("utterance", **kw["cmvn"])
    if "masks" in kw:
        plan.set_masks(table, n_f, n_t, kw.get("mask_value"))
    ref = plan.execute(packed.to_device())
    torch.cuda.synchronize()
    a, b = ref.cpu().numpy().reshape(-1, 80), got.cpu().numpy().reshape(-1, 80)
    assert a.shape == b.shape and np.isfinite(a).all()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    plan.close()
    # many calls back to back: staging slots and retired plans are recycled behind their events
    outs = [frontend.fbank_cmvn_specaug_ragged(waves, **kw)[0] for _ in range(12)]
    torch.cuda.synchronize()
    for o in outs:
        assert np.array_equal(o.cpu().numpy().reshape(-1, 80).view(np.uint32), a.view(np.uint32))


@pytest.mark.gpu
def test_batch_call_errors_enqueue_nothing():
    """A too-short utterance (helpers_for_audio.py:41-68 -> kaldi.py's window-size assertion) or an output
    buffer that is too small is reported by ``js2t_batch_fbank`` before anything is enqueued."""
    import ctypes

    import torch
    from joeys2t_b200 import _lib, frontend

    waves = [np.zeros(16000, np.int16), np.zeros(399, np.int16)]
    with pytest.raises(_lib.Js2tError) as ei:
        frontend.fbank_cmvn_specaug_ragged(waves)
    assert ei.value.code == _lib.ERR_SHORT_INPUT and "choose a window size 400" in str(ei.value)
    ctx = frontend.get_context()
    lib = _lib.load()
    w = np.zeros(16000, np.int16)
    out = torch.empty((10, 80), dtype=torch.float32, device="cuda")  # 98 rows needed
    o = _lib.BatchOpts()
    o.layout, o.pad_value, o.cmvn_mode = _lib.LAYOUT_RAGGED, 1.0, _lib.CMVN_NONE
    ptrs = np.array([w.ctypes.data], np.uint64)
    ns = np.array([16000], np.int64)
    rc = lib.js2t_batch_fbank(ctx.handle, 1, ptrs.ctypes.data, ns.ctypes.data, None, ctypes.byref(o),
                              out.data_ptr(), 10, 0, None)
    assert rc == _lib.ERR_INVALID and b"98" in lib.js2t_last_error()
    # a caller-owned plan: statistics can be read back, the caller destroys it
    out = torch.empty((98, 80), dtype=torch.float32, device="cuda")
    o.cmvn_mode, o.norm_means, o.norm_vars, o.before = _lib.CMVN_UTTERANCE, 1, 1, 1
    h = ctypes.c_void_p()
    rc = lib.js2t_batch_fbank(ctx.handle, 1, ptrs.ctypes.data, ns.ctypes.data, None, ctypes.byref(o),
                              out.data_ptr(), 98, torch.cuda.current_stream().cuda_stream, ctypes.byref(h))
    assert rc == _lib.OK and h.value
    torch.cuda.synchronize()
    assert int(lib.js2t_plan_total_frames(h)) == 98
    assert lib.js2t_plan_destroy(h) == _lib.OK


@pytest.mark.gpu
def test_forked_worker_gets_a_clear_error(fe, fixtures_pcm):
    """``DataLoader(num_workers > 0)`` forks its workers (the reference's loaders do: datasets.py make_iter); a CUDA
    context does not survive ``fork``.  A forked child that calls the front-end must fail with an explanation —
    not inside the driver — and the parent must be unaffected."""
    import os

    fe.fbank_cmvn_specaug_ragged([fixtures_pcm[0][0]])  # the parent owns a context now
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:  # child: no CUDA call may be made here
        os.close(r)
        try:
            fe.fbank_cmvn_specaug_ragged([fixtures_pcm[0][0]])
            msg = b"no error"
        except RuntimeError as e:
            msg = str(e).encode()
        except BaseException as e:  # pylint: disable=broad-except
            msg = b"other: " + repr(e).encode()
        os.write(w, msg[:4000])
        os._exit(0)
    os.close(w)
    got = os.read(r, 4096).decode()
    os.waitpid(pid, 0)
    os.close(r)
    assert "do not survive fork" in got and "num_workers=0" in got, got
    out, _ = fe.fbank_cmvn_specaug_ragged([fixtures_pcm[0][0]])
    assert np.isfinite(out.cpu().numpy()).all()


@pytest.mark.gpu
def test_threads_with_their_own_streams_get_bit_identical_results(fe):
    """the reference's functions are called from several loader threads / workers at once, so the
    CUDA-backed drop-in must be re-entrant per stream.  Four threads, each on its own CUDA stream, push different
    batches through the one-call entry point (shared context: staging slots, plan pool, upload stream, retired
    plans) at the same time; every result must equal, bit for bit, what the same call gives single-threaded."""
    import threading

    import torch

    rng = np.random.RandomState(5)
    n_threads, n_rounds = 4, 12
    jobs = []
    for t in range(n_threads):
        lens = [int(x) for x in rng.randint(400, 16000 * 6, size=5 + 3 * t)]
        waves = _mixed_batch(rng, lens)
        T = np.array([1 + (n - 400) // 160 for n in lens])
        table = np.zeros((len(waves), 4, 2), np.int32)
        for u in range(len(waves)):
            table[u, :2, 0] = rng.randint(0, 60, 2)
            table[u, :2, 1] = rng.randint(0, 20, 2)
            table[u, 2:, 0] = rng.randint(0, max(int(T[u]), 1), 2)
            table[u, 2:, 1] = rng.randint(0, 25, 2)
        kw = dict(cmvn=dict(norm_means=True, norm_vars=bool(t % 2), before=bool(t < 2)), masks=table, n_fmask=2,
                  n_tmask=2, layout="padded" if t % 2 else "ragged", pad_value=1.0)
        jobs.append((waves, kw))
    expected = []
    for waves, kw in jobs:
        out, _ = fe.fbank_cmvn_specaug_ragged(waves, **kw)
        expected.append(out.cpu().numpy().copy())
    torch.cuda.synchronize()

    errors, barrier = [], threading.Barrier(n_threads)

    def work(t):
        try:
            waves, kw = jobs[t]
            stream = torch.cuda.Stream()
            barrier.wait()
            with torch.cuda.stream(stream):
                outs = [fe.fbank_cmvn_specaug_ragged(waves, **kw)[0] for _ in range(n_rounds)]
                stream.synchronize()
                for o in outs:
                    got = o.cpu().numpy()
                    if not np.array_equal(got.view(np.uint32), expected[t].view(np.uint32)):
                        errors.append(f"thread {t}: result differs from the single-threaded one")
                        break
        except Exception as e:  # pylint: disable=broad-except
            errors.append(f"thread {t}: {e!r}")

    threads = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
