# coding: utf-8
"""
Parity of the CUDA path (through the C ABI) against the reference's golden vectors and the CPU
oracle.  ``-m gpu`` only.

Tolerances (BASELINE.json north_star / SURVEY.md §8c):
  * raw log-mel:            max |Δ| <= 1e-3
  * CMVN-normalised output: |Δ| <= 5e-4 + 1e-4 |ref|
  * CMVN arithmetic alone (our CMVN on the reference's log-mel): rtol 1e-4, atol 1e-5
  * SpecAugment:            masked-cell positions bit-exact, fill value within 1e-6 of the mean
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import fbank_numpy as O  # noqa: E402

LOGMEL_ATOL = 1e-3
KNOWN_ANSWER = np.array([
    -1.0788909, -1.0076448, -1.0421542, -1.0393586, -1.0239305,
    -0.9921213, -0.95107234, -0.9340749, -0.9119267, -0.8962079,
], np.float32)  # /root/reference/test/unit/test_tokenizer.py:322-329


@pytest.fixture(scope="module")
def fe():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    from joeys2t_b200 import _lib, frontend
    if _lib.is_stale():
        _lib.build()
    return frontend


def cmvn_close(got, ref):
    err = np.abs(got - ref)
    tol = 5e-4 + 1e-4 * np.abs(ref)
    assert (err <= tol).all(), f"max err {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"


# ------------------------------------------------------------------------------------------------
# config 1: the reference's own fixture wavs
# ------------------------------------------------------------------------------------------------
def test_fixture_fbank_vs_reference_golden(fe, fixtures_pcm, ref_fbank):
    pcm, n_frames = fixtures_pcm
    out, lens = fe.fbank_cmvn_specaug_ragged(pcm)
    assert lens.tolist() == n_frames.tolist()
    got = out.cpu().numpy()
    off = 0
    worst = 0.0
    for i, ref in enumerate(ref_fbank):
        g = got[off:off + ref.shape[0]]
        off += ref.shape[0]
        d = np.abs(g - ref).max()
        worst = max(worst, d)
        assert d <= LOGMEL_ATOL, f"clip {i}: {d}"
    print("fixture log-mel max abs err vs reference", worst)


def test_fixture_float32_pcm_path(fe, fixtures_pcm, ref_fbank):
    pcm, _ = fixtures_pcm
    waves = [torch.from_numpy(x.astype(np.float32) / 32768.0).unsqueeze(0) for x in pcm[:4]]
    out, _ = fe.fbank_cmvn_specaug_ragged(waves)
    ints, _ = fe.fbank_cmvn_specaug_ragged(pcm[:4])
    # quirk Q4: float PCM * 2**15 is exactly the int16 sample
    assert torch.equal(out, ints)
    assert np.abs(out.cpu().numpy()[:215] - ref_fbank[0]).max() <= LOGMEL_ATOL


def test_known_answer_vector(fe, fixtures_pcm):
    pcm, _ = fixtures_pcm
    out, _ = fe.fbank_cmvn_specaug_ragged([pcm[1]], cmvn=dict(norm_means=True, norm_vars=True))
    np.testing.assert_allclose(out.cpu().numpy()[0, :10], KNOWN_ANSWER, atol=5e-4, rtol=1e-4)


def test_fixture_utterance_cmvn(fe, fixtures_pcm, ref_fbank, ref_cmvn):
    pcm, _ = fixtures_pcm
    out, lens = fe.fbank_cmvn_specaug_ragged(pcm, cmvn=dict(norm_means=True, norm_vars=True))
    got = out.cpu().numpy()
    off = 0
    for i, f in enumerate(ref_fbank):
        g = got[off:off + f.shape[0]]
        off += f.shape[0]
        cmvn_close(g, O.cmvn(f))
        assert np.abs(g[:4] - ref_cmvn[f"m1v1_head{i}"]).max() <= 1e-3
    cmvn_close(got[215:215 + 172], ref_cmvn["m1v1_full1"])


@pytest.mark.parametrize("nm", [True, False])
@pytest.mark.parametrize("nv", [True, False])
def test_cmvn_arithmetic_in_isolation(fe, ref_fbank, ref_cmvn, nm, nv):
    """Our CMVN kernels on the reference's own log-mel (feature-input path)."""
    tag = f"m{int(nm)}v{int(nv)}"
    out, lens = fe.features_cmvn_specaug_ragged(ref_fbank, cmvn=dict(norm_means=nm, norm_vars=nv))
    got = out.cpu().numpy()
    off = 0
    for i, f in enumerate(ref_fbank):
        g = got[off:off + f.shape[0]]
        off += f.shape[0]
        # the reference's float32 column sums drift with T (1.3e-4 at T~5000); the fp64
        # restatement of the same formula is the arbiter for the tight tolerance
        np.testing.assert_allclose(g, O.cmvn_fp64(f, nm, nv), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(g, O.cmvn(f, nm, nv), rtol=1e-4, atol=2e-4)
        np.testing.assert_allclose(g[:4], ref_cmvn[f"{tag}_head{i}"], rtol=1e-4, atol=2e-4)


def test_cmvn_class_dropin(fe, ref_fbank, ref_cmvn):
    from joeys2t_b200.data_augmentation import CMVN
    c = CMVN()
    assert repr(c) == "CMVN(norm_means=True, norm_vars=True, before=True)" and c.before is True
    x = ref_fbank[1].copy()
    y = c(x)
    assert np.array_equal(x, ref_fbank[1]), "input must not be mutated"
    assert y.dtype == np.float32 and y.flags["C_CONTIGUOUS"]
    np.testing.assert_allclose(y, ref_cmvn["m1v1_full1"], rtol=1e-4, atol=1e-4)


def test_silent_utterance(fe, ref_cmvn):
    out, _ = fe.fbank_cmvn_specaug_ragged([np.zeros(4000, np.int16)])
    assert np.array_equal(out.cpu().numpy(), ref_cmvn["silent_fbank"])
    # CMVN of a constant column is ill-conditioned: the reference's own output there is float32
    # rounding noise of the column mean divided by sqrt(noise) (-1.4e-4 everywhere); the exact answer
    # is 0.  Both must be "zero at CMVN scale".
    out, _ = fe.fbank_cmvn_specaug_ragged([np.zeros(4000, np.int16)], cmvn={})
    assert np.abs(ref_cmvn["silent_cmvn"]).max() < 1e-3
    assert np.abs(out.cpu().numpy()).max() < 1e-3


# ------------------------------------------------------------------------------------------------
# SpecAugment: positions bit-exact, fill value = mean of the spectrogram
# ------------------------------------------------------------------------------------------------
SA_CFGS = {
    "mustc": dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=1.0),
    "test": dict(freq_mask_n=1, freq_mask_f=5, time_mask_n=1, time_mask_t=10, time_mask_p=1.0),
    "default": dict(),
    "smallp": dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=0.05),
    "zerop": dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=0.001),
    "widef": dict(freq_mask_n=2, freq_mask_f=81, time_mask_n=2, time_mask_t=100, time_mask_p=1.0),
    "const": dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=1.0,
                  mask_value=0.0),
}


@pytest.mark.parametrize("cname", list(SA_CFGS))
def test_specaugment_class_vs_reference_golden(fe, ref_fbank, ref_specaugment, cname):
    from joeys2t_b200.data_augmentation import SpecAugment
    g = ref_specaugment
    cfg = SA_CFGS[cname]
    sa = SpecAugment(**cfg)
    for seed in (0, 2345):
        for i in (1, 0, 9):
            x = O.cmvn(ref_fbank[i])
            key = f"{cname}_s{seed}_c{i}"
            np.random.seed(seed)
            y = sa(x)
            assert y.shape == x.shape and y.dtype == np.float32
            # positions: bit-exact
            assert np.array_equal(np.packbits(y != x), g[key + "_changed"]), key
            changed = y != x
            # untouched cells are bit-identical to the input; filled cells hold the fill value
            assert np.array_equal(y[~changed], x[~changed])
            if changed.any():
                assert np.abs(y[changed] - g[key + "_maskvalue"]).max() <= 1e-6
            if i == 1:
                np.testing.assert_allclose(y, g[key + "_full"], rtol=0, atol=1e-6)
            # the RNG must have been consumed exactly like the reference: next draw agrees
            nxt = np.random.randint(0, 1 << 30)
            np.random.seed(seed)
            O.specaugment(x, **cfg)
            assert nxt == np.random.randint(0, 1 << 30)


def test_fused_fbank_cmvn_specaugment_batch(fe, fixtures_pcm, ref_fbank):
    """The fused batched path == reference per-item loop (CMVN -> SpecAugment), padded layout."""
    from joeys2t_b200.data_augmentation import SpecAugment, mask_tables_for_batch
    pcm, n_frames = fixtures_pcm
    cfg = SA_CFGS["mustc"]
    np.random.seed(77)
    table, nf, nt = mask_tables_for_batch(SpecAugment(**cfg), n_frames)
    out, lens = fe.fbank_cmvn_specaug_ragged(pcm, cmvn=dict(norm_means=True, norm_vars=True),
                                             masks=table, n_fmask=nf, n_tmask=nt, layout="padded")
    got = out.cpu().numpy()
    assert got.shape == (10, 1470, 80)
    np.random.seed(77)
    for u, f in enumerate(ref_fbank):
        x = O.cmvn(f)
        ref = O.specaugment(x, **cfg)
        t = f.shape[0]
        masked = ref != x
        cmvn_close(got[u, :t][~masked], ref[~masked])
        assert np.abs(got[u, :t][masked] - ref[masked]).max(initial=0) <= 1e-6
        assert (got[u, t:] == 1.0).all()


def test_speech_processor_vs_reference_golden(fe, fixtures_pcm, ref_processor, tmp_path):
    """SpeechProcessor drop-in through real wav files: filters, truncation, CMVN/SpecAugment order."""
    import wave
    from joeys2t_b200.speech_processor import SpeechProcessor
    pcm, _ = fixtures_pcm
    (tmp_path / "wav").mkdir()
    for i, x in enumerate(pcm):
        with wave.open(str(tmp_path / "wav" / f"c{i}.wav"), "wb") as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(16000)
            w.writeframes(x.tobytes())
    g = ref_processor
    sa = SA_CFGS["mustc"]
    for vname, before in (("before", True), ("after", False)):
        proc = SpeechProcessor(level="frame", num_freq=80, max_length=500, min_length=200,
                               specaugment=sa, cmvn=dict(norm_means=True, norm_vars=True, before=before))
        proc.root_path = tmp_path
        for i in range(10):
            for is_train in (True, False):
                key = f"{vname}_{'train' if is_train else 'eval'}_c{i}"
                np.random.seed(1000 + i)
                y = proc(f"wav/c{i}.wav", is_train=is_train)
                if int(g[key + "_none"]):
                    assert y is None, key
                    continue
                assert list(y.shape) == g[key + "_shape"].tolist(), key
                if key + "_full" in g.files:
                    ref = g[key + "_full"]
                    if before:
                        cmvn_close(y, ref)
                    else:
                        # CMVN after SpecAugment: a masked column is constant, the reference divides
                        # its own rounding noise by std = 1e-5 there (ill-conditioned)
                        col_const = np.ptp(ref, axis=0) < 1e-2
                        cmvn_close(y[:, ~col_const], ref[:, ~col_const])
                        assert np.abs(y[:, col_const]).max(initial=0) < 0.2


def test_truncation_before_cmvn(fe, fixtures_pcm, ref_fbank):
    pcm, _ = fixtures_pcm
    out, lens = fe.fbank_cmvn_specaug_ragged([pcm[2], pcm[4]], cmvn={}, max_frames=[500, 500])
    assert lens.tolist() == [500, 500]
    got = out.cpu().numpy()
    cmvn_close(got[:500], O.cmvn(ref_fbank[2][:500]))
    cmvn_close(got[500:], O.cmvn(ref_fbank[4][:500]))


# ------------------------------------------------------------------------------------------------
# edge cases
# ------------------------------------------------------------------------------------------------
def test_edge_lengths(fe):
    rng = np.random.default_rng(0)
    lens = [400, 401, 559, 560, 719, 720, 5359, 5360, 5361, 5519, 5520, 5521, 16000]
    waves = [rng.integers(-20000, 20000, n).astype(np.int16) for n in lens]
    out, nf = fe.fbank_cmvn_specaug_ragged(waves)
    assert nf.tolist() == [O.num_frames(n) for n in lens]
    got = out.cpu().numpy()
    off = 0
    for w, t in zip(waves, nf):
        ref = O.extract_fbank_features(w)
        assert np.abs(got[off:off + t] - ref).max() <= LOGMEL_ATOL
        off += t


def test_short_input_raises(fe):
    from joeys2t_b200._lib import Js2tError
    from joeys2t_b200.helpers_for_audio import extract_fbank_features
    with pytest.raises(Js2tError):
        fe.fbank_cmvn_specaug_ragged([np.zeros(399, np.int16)])
    with pytest.raises(ValueError):  # quirk Q2: ValueError with or without output_path
        extract_fbank_features(torch.zeros(1, 399), 16000)
    with pytest.raises(ValueError):
        extract_fbank_features(torch.zeros(1, 4000), 8000)


def test_multichannel_takes_channel0(fe, fixtures_pcm):
    pcm, _ = fixtures_pcm
    x = pcm[0][:8000].astype(np.float32) / 32768.0
    stereo = torch.from_numpy(np.stack([x, x[::-1].copy()]))
    a, _ = fe.fbank_cmvn_specaug_ragged([stereo])
    b, _ = fe.fbank_cmvn_specaug_ragged([x])
    assert torch.equal(a, b)


def test_extract_fbank_features_dropin(fe, fixtures_pcm, ref_fbank, tmp_path):
    from joeys2t_b200.helpers_for_audio import extract_fbank_features, get_features
    pcm, _ = fixtures_pcm
    w = torch.from_numpy(pcm[5].astype(np.float32) / 32768.0).unsqueeze(0)
    p = tmp_path / "f.npy"
    feats = extract_fbank_features(w, 16000, output_path=p)
    assert feats.dtype == np.float32 and feats.shape == ref_fbank[5].shape
    assert np.abs(feats - ref_fbank[5]).max() <= LOGMEL_ATOL
    assert p.is_file() and np.array_equal(np.load(p), feats)
    # cache hit returns the stored array without recomputation
    np.save(p, feats + 1)
    assert np.array_equal(extract_fbank_features(w, 16000, output_path=p), feats + 1)
    assert np.array_equal(get_features(tmp_path, "f.npy"), feats + 1)


def test_gain_invariance_of_utterance_cmvn(fe, fixtures_pcm):
    """Utterance CMVN output is independent of the input scale except through the floor
    (scripts/gradio_demo.py:52-54 relies on it)."""
    pcm, _ = fixtures_pcm
    x = pcm[3]
    half = (x // 2).astype(np.int16)
    b, _ = fe.fbank_cmvn_specaug_ragged([(half * 2).astype(np.int16)], cmvn={})
    c, _ = fe.fbank_cmvn_specaug_ragged([half], cmvn={})
    assert (b - c).abs().max().item() < 1e-4  # exact factor 2: only rounding + floor can differ


# ------------------------------------------------------------------------------------------------
# synthetic configs 2-4 at reduced size vs the oracle; full size through properties
# ------------------------------------------------------------------------------------------------
def test_librispeech_shaped_small_vs_oracle(fe):
    from joeys2t_b200 import synthetic
    waves = synthetic.librispeech_batch(6, seed=1234, lo=2.0, hi=4.0)
    waves[2][5000:5000 + 4800] = 0  # exact digital silence: exercises the floor
    out, nf = fe.fbank_cmvn_specaug_ragged(waves, cmvn={})
    got = out.cpu().numpy()
    off = 0
    for w, t in zip(waves, nf):
        raw = O.extract_fbank_features(w)
        cmvn_close(got[off:off + t], O.cmvn_fp64(raw))
        off += t


def test_longform_mixed_dtype_vs_oracle(fe):
    from joeys2t_b200 import synthetic
    waves = synthetic.longform_batch(2, seed=3456)
    waves = [w[:16000 * 31] for w in waves]
    assert waves[0].dtype == np.int16 and waves[1].dtype == np.float32
    out, nf = fe.fbank_cmvn_specaug_ragged(waves)
    got = out.cpu().numpy()
    off = 0
    for w, t in zip(waves, nf):
        ref = O.extract_fbank_features(w)
        assert np.abs(got[off:off + t] - ref).max() <= LOGMEL_ATOL
        off += t


def test_full_size_properties(fe):
    """Config-2 size (256 x 10-15 s): size-independent properties instead of the slow oracle."""
    from joeys2t_b200 import synthetic
    waves = synthetic.pooled_batch(256, seed=1234, lo=10.0, hi=15.0)
    out, nf = fe.fbank_cmvn_specaug_ragged(waves, cmvn={})
    assert int(nf.sum()) == sum(O.num_frames(len(w)) for w in waves)
    assert torch.isfinite(out).all()
    # per-utterance column means ~ 0 and stds ~ 1 after CMVN
    off = 0
    for t in nf[:32]:
        blk = out[off:off + int(t)].double()
        off += int(t)
        assert blk.mean(0).abs().max().item() < 1e-4
        assert (blk.std(0, unbiased=False) - 1).abs().max().item() < 1e-3
    # batching must not change results: utterance 17 alone == utterance 17 in the batch
    single, _ = fe.fbank_cmvn_specaug_ragged([waves[17]], cmvn={})
    start = int(nf[:17].sum())
    assert torch.equal(single, out[start:start + int(nf[17])])
    # determinism
    again, _ = fe.fbank_cmvn_specaug_ragged(waves, cmvn={})
    assert torch.equal(again, out)
    # spot-check a few utterances against the oracle
    for u in (0, 100, 255):
        start = int(nf[:u].sum())
        cmvn_close(out[start:start + int(nf[u])].cpu().numpy(),
                   O.cmvn_fp64(O.extract_fbank_features(waves[u])))


# ------------------------------------------------------------------------------------------------
# global CMVN (extension; oracle = reference CMVN formula on the frame-concatenation)
# ------------------------------------------------------------------------------------------------
def test_global_cmvn_two_pass(fe, fixtures_pcm, ref_fbank):
    from joeys2t_b200 import distributed as D
    pcm, _ = fixtures_pcm
    packed = fe.PackedPCM(pcm)
    plan = fe.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
    plan.set_cmvn("stats")
    dev = packed.to_device()
    raw = plan.execute(dev)
    accum = D.new_accumulator("cuda")
    plan.accumulate_global(accum)
    s, q, n = O.global_cmvn_stats(ref_fbank)
    a = accum.cpu().numpy()
    assert a[160] == n
    np.testing.assert_allclose(a[:80], s, rtol=2e-6)
    np.testing.assert_allclose(a[80:160], q, rtol=2e-6)
    # second pass: normalise in place with the finalised statistics
    plan.set_cmvn("global")
    plan.finalize_global(accum)
    out = plan.normalize(raw).cpu().numpy()
    want = np.concatenate(O.global_cmvn(ref_fbank), 0)
    cmvn_close(out, want)
    # single fused pass with the statistics known up front
    mean, istd = D.stats_to_mean_istd(accum)
    fused, _ = fe.fbank_cmvn_specaug_ragged(pcm, cmvn={}, global_stats=(mean, istd))
    cmvn_close(fused.cpu().numpy(), want)
    torch.cuda.synchronize()
    plan.close()


# ------------------------------------------------------------------------------------------------
# determinism: two plans over the same batch, and re-execution of one plan, give bit-identical results
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("with_masks", [False, True])
def test_utterance_cmvn_is_deterministic(fe, fixtures_pcm, with_masks):
    from joeys2t_b200 import synthetic
    from joeys2t_b200.data_augmentation import SpecAugment, mask_tables_for_batch
    pcm, _ = fixtures_pcm
    waves = list(pcm) + synthetic.librispeech_batch(12, seed=5, lo=0.5, hi=6.0)
    packed = fe.PackedPCM(waves)
    dev = packed.to_device()
    outs = []
    for _ in range(2):
        plan = fe.Plan(packed.n_samples, packed.byte_off, packed.is_f32, layout="padded")
        plan.set_cmvn("utterance", True, True, True)
        if with_masks:
            np.random.seed(99)
            table, nf, nt = mask_tables_for_batch(SpecAugment(**SA_CFGS["mustc"]), plan.n_frames)
            plan.set_masks(table, nf, nt)
        out = plan.execute(dev)
        again = plan.execute(dev, torch.empty_like(out))  # scheduler counters must be reusable
        torch.cuda.synchronize()
        assert torch.equal(out, again)
        stats = plan.utt_stats().cpu()
        outs.append((out.cpu(), stats))
        plan.close()
    assert torch.equal(outs[0][1], outs[1][1])  # identical fp64 statistics (fixed summation order)
    assert torch.equal(outs[0][0], outs[1][0])  # identical normalised features and fill values


# ------------------------------------------------------------------------------------------------
# more edge cases and properties (round 1d)
# ------------------------------------------------------------------------------------------------
def test_modified_mel_bank_is_rejected(fe):
    """The mel weights are compile-time immediates; a context must refuse any other bank."""
    import ctypes
    from joeys2t_b200 import _lib, tables
    lib = _lib.load()
    h = ctypes.c_void_p()
    _lib.check(lib.js2t_ctx_create(torch.cuda.current_device(), ctypes.byref(h)))
    try:
        win = np.ascontiguousarray(tables.povey_window(), np.float32)
        mel = np.ascontiguousarray(tables.mel_banks(80), np.float32)
        assert lib.js2t_ctx_set_tables(h, win.ctypes.data, mel.ctypes.data) == _lib.OK
        bad = mel.copy()
        k = int(np.flatnonzero(bad[40])[0])
        bad[40, k] = np.nextafter(bad[40, k], np.float32(2.0))   # one ulp off
        assert lib.js2t_ctx_set_tables(h, win.ctypes.data, bad.ctypes.data) == _lib.ERR_TABLES
        assert b"mel bank" in lib.js2t_last_error()
        bad = mel.copy()
        bad[3, 200] = 0.5                                           # outside the two-band structure
        assert lib.js2t_ctx_set_tables(h, win.ctypes.data, bad.ctypes.data) == _lib.ERR_TABLES
        w2 = win.copy()
        w2[0] = 1e-3                                                # window[0] must be exactly 0
        assert lib.js2t_ctx_set_tables(h, w2.ctypes.data, mel.ctypes.data) == _lib.ERR_TABLES
    finally:
        lib.js2t_ctx_destroy(h)


def test_extreme_amplitudes_and_dc(fe):
    """Full-scale square wave, full-scale DC (mean removal must cancel it), a lone impulse and the
    int16 extremes: log-mel within tolerance of the oracle, floored cells bit-identical."""
    n = 16000
    t = np.arange(n)
    waves = [
        np.where((t // 40) % 2 == 0, 32767, -32768).astype(np.int16),
        np.full(n, 32767, np.int16),
        np.full(n, -32768, np.int16),
        np.zeros(n, np.int16),
    ]
    waves[3][7777] = 32767
    out, nf = fe.fbank_cmvn_specaug_ragged(waves)
    got = out.cpu().numpy()
    off = 0
    floor = np.float32(np.log(np.float32(1.1920928955078125e-07)))
    for w, tt in zip(waves, nf):
        ref = O.extract_fbank_features(w)
        blk = got[off:off + tt]
        off += tt
        floored = ref == floor
        assert (blk[floored] == floor).all(), "digital silence must hit the exact float32 floor"
        assert np.abs(blk[~floored] - ref[~floored]).max(initial=0.0) <= LOGMEL_ATOL
    # constant input: DC removal leaves exact zeros -> every cell is the floor
    assert (got[nf[0]:nf[0] + nf[1]] == floor).all() and (got[nf[0] + nf[1]:nf[:3].sum()] == floor).all()


def test_float64_and_tensor_inputs_match_int16(fe, fixtures_pcm):
    """Q3/Q4: float64 / float32 tensors in [-1, 1) and int16 PCM give the same features."""
    pcm, _ = fixtures_pcm
    w = pcm[1]
    a, _ = fe.fbank_cmvn_specaug_ragged([w])
    b, _ = fe.fbank_cmvn_specaug_ragged([w.astype(np.float64) / 32768.0])
    c, _ = fe.fbank_cmvn_specaug_ragged([torch.from_numpy(w.astype(np.float32) / np.float32(32768.0))[None]])
    assert torch.equal(a, b) and torch.equal(a, c)


def test_tile_boundary_utterances_in_padded_layout(fe):
    """Utterances of exactly 32/33/64/65 frames (tile boundaries), padded layout: rows past each
    utterance hold the pad value, rows inside match the ragged layout bit for bit."""
    rng = np.random.default_rng(3)
    frames = [1, 31, 32, 33, 63, 64, 65, 96]
    waves = [rng.integers(-9000, 9000, 400 + 160 * (f - 1)).astype(np.int16) for f in frames]
    rag, nf = fe.fbank_cmvn_specaug_ragged(waves, cmvn={})
    pad, nf2 = fe.fbank_cmvn_specaug_ragged(waves, cmvn={}, layout="padded", pad_value=-7.0)
    assert nf.tolist() == frames == nf2.tolist() and pad.shape == (len(frames), 96, 80)
    off = 0
    for u, f in enumerate(frames):
        assert torch.equal(pad[u, :f], rag[off:off + f])
        assert (pad[u, f:] == -7.0).all()
        off += f


def test_host_pipeline_matches_direct_execution(fe):
    """frontend.HostPipeline (three streams, pinned host buffers) returns what Plan.execute does."""
    from joeys2t_b200 import synthetic
    batches = [synthetic.pooled_batch(12, seed=50 + i, lo=1.0, hi=3.0) for i in range(5)]
    packs = [fe.PackedPCM(b) for b in batches]
    plans = [fe.Plan(p.n_samples, p.byte_off, p.is_f32).set_cmvn("utterance") for p in packs]
    want = [pl.execute(pk.to_device()).cpu() for pl, pk in zip(plans, packs)]
    pipe = fe.HostPipeline(n_slots=2, max_pcm_bytes=max(p.nbytes for p in packs),
                           max_out_rows=max(p.out_rows for p in plans))
    for i in range(len(packs)):
        slot = pipe.submit(packs[i], plans[i])
        got = pipe.result(slot).clone()
        assert torch.equal(got, want[i]), i
    # back-to-back submissions without reading in between (slots are recycled safely)
    slots = [pipe.submit(packs[i], plans[i]) for i in range(2)]
    for i, s in enumerate(slots):
        assert torch.equal(pipe.result(s), want[i])
    for pl in plans:
        pl.close()


def test_mustc_and_longform_full_size_properties(fe):
    """Configs 3 and 4 at full size: frame counts, finiteness, CMVN moments, masks where drawn."""
    import argparse
    import bench
    from joeys2t_b200.data_augmentation import SpecAugment, mask_tables_for_batch
    for wl in ("cfg3", "cfg4"):
        args = argparse.Namespace(workload=wl, utts=0, sweep_hours=0.0)
        waves = bench.make_batch(args, 2345)
        n_frames = [O.num_frames(len(w)) for w in waves]
        table = nf_ = nt_ = None
        if wl == "cfg3":
            np.random.seed(2345)
            table, nf_, nt_ = mask_tables_for_batch(SpecAugment(2, 27, 2, 100, 1.0), n_frames)
        out, nf = fe.fbank_cmvn_specaug_ragged(waves, cmvn={}, masks=table, n_fmask=nf_ or 0, n_tmask=nt_ or 0)
        assert nf.tolist() == n_frames and torch.isfinite(out).all()
        off = 0
        for u, t in enumerate(n_frames[:24]):
            blk = out[off:off + t].double()
            off += t
            if table is None:
                assert blk.mean(0).abs().max().item() < 2e-4
                assert (blk.std(0, unbiased=False) - 1).abs().max().item() < 2e-3
            else:  # masked cells hold one value per utterance, in exactly the drawn rectangles
                m = np.zeros((t, 80), bool)
                for f0, w in table[u, :nf_]:
                    m[:, f0:f0 + w] = True
                for t0, w in table[u, nf_:]:
                    m[t0:t0 + w, :] = True
                vals = blk.cpu().numpy()[m]
                assert vals.size == 0 or np.ptp(vals) == 0.0
        # spot-check against the oracle (unmasked utterance CMVN of the same waveform)
        u = 5
        start = int(np.sum(n_frames[:u]))
        ref = O.cmvn_fp64(O.extract_fbank_features(waves[u]))
        got = out[start:start + n_frames[u]].cpu().numpy()
        if table is not None:
            keep = np.ones_like(ref, bool)
            for f0, w in table[u, :nf_]:
                keep[:, f0:f0 + w] = False
            for t0, w in table[u, nf_:]:
                keep[t0:t0 + w, :] = False
            err = np.abs(got - ref)[keep]
            assert (err <= 5e-4 + 1e-4 * np.abs(ref[keep])).all()
        else:
            cmvn_close(got, ref)


def test_float_pcm_two_slot_path_matches_int16_bit_for_bit(fe):
    """float32 tiles are staged as two half-tiles (frames 0-15 | 16-31): frame counts around the half
    and full tile boundaries, int16 and float32 utterances interleaved in one batch — the float path
    must reproduce the int16 path bit for bit (x / 32768 * 2**15 is exact, quirk Q4)."""
    rng = np.random.default_rng(11)
    frames = [1, 2, 15, 16, 17, 18, 31, 32, 33, 47, 48, 49, 64, 65, 100]
    ints = [rng.integers(-30000, 30000, 400 + 160 * (f - 1) + int(rng.integers(0, 160))).astype(np.int16)
            for f in frames]
    floats = [x.astype(np.float32) / np.float32(32768.0) for x in ints]
    a, nf = fe.fbank_cmvn_specaug_ragged(ints)
    assert nf.tolist() == frames
    mixed = [floats[i] if i % 2 == 0 else ints[i] for i in range(len(ints))]
    mixed2 = [ints[i] if i % 2 == 0 else floats[i] for i in range(len(ints))]
    for batch in (floats, mixed, mixed2):
        b, nf2 = fe.fbank_cmvn_specaug_ragged(batch)
        assert nf2.tolist() == frames and torch.equal(a, b)
    ref = O.extract_fbank_features(ints[9])
    off = sum(frames[:9])
    assert np.abs(a[off:off + frames[9]].cpu().numpy() - ref).max() <= LOGMEL_ATOL


def test_non_finite_pcm_does_not_leak_into_other_utterances(fe):
    """A float utterance full of NaN / Inf poisons only its own rows: what is left in the on-chip PCM
    slots must never reach the frames of the utterances processed after it (their last frame reads
    up to 16 samples past the staged data, where the window is zero)."""
    rng = np.random.default_rng(12)
    good = [rng.integers(-20000, 20000, n).astype(np.int16) for n in (400, 1999, 5360, 7000, 16000)]
    good_f = [g.astype(np.float32) / np.float32(32768.0) for g in good]
    bad = np.full(16000 * 3, np.nan, np.float32)
    bad[::7] = np.inf
    clean, nf = fe.fbank_cmvn_specaug_ragged(good + good_f)
    batch = []
    for g in good + good_f:
        batch += [bad, g]
    out, nf2 = fe.fbank_cmvn_specaug_ragged(batch + [bad] * 40 + good_f)
    rows = np.concatenate([[0], np.cumsum(nf2)])
    got = torch.cat([out[rows[2 * i + 1]:rows[2 * i + 2]] for i in range(len(good) * 2)])
    assert torch.equal(got, clean)
    tail = out[rows[len(batch) + 40]:]
    assert torch.equal(tail, clean[sum(nf[:len(good)]):])
    assert torch.isfinite(clean).all()
