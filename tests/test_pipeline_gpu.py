"""End-to-end parity: drop-in entry points (wav file in, numpy / torch out) against the golden
fixtures produced by executing the reference's own functions, and the batched C2 pipeline
against the oracle clip by clip."""
import json
import os
import random

import numpy as np
import pytest
import torch

from cases import RECORDINGS, SR, check_digest, sha, sha_list
from signals import golden_signal

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
META = json.load(open(os.path.join(HERE, "golden", "ref_util.json")))["cases"]


@pytest.fixture(scope="module")
def wav_dir(tmp_path_factory):
    from heart_murmur_detection_b200 import audio_io

    d = tmp_path_factory.mktemp("wavs")
    for name, n, seed, lead, tail in RECORDINGS:
        audio_io.write_wav_f32(str(d / f"{name}.wav"), golden_signal(n, seed, SR, lead, tail), SR)
    return str(d)


def _check(key, out, spec_tol):
    g = META[key]
    if g.get("none"):
        assert out is None, key
        return
    if isinstance(out, np.ndarray):
        assert list(out.shape) == g["shape"] and str(out.dtype) == g["dtype"], (key, out.shape, out.dtype, g["shape"], g["dtype"])
        if out.ndim == 1:
            if out.dtype == np.float32:
                assert sha(out) == g["sha"], key
        else:
            check_digest(out, g["digest"], rtol=spec_tol, atol=spec_tol)
        return
    outs = [o.numpy() if hasattr(o, "numpy") else np.asarray(o) for o in out]
    assert len(outs) == g["n"] and [list(o.shape) for o in outs] == g["shapes"], key
    if outs and outs[0].ndim == 1 and outs[0].dtype == np.float32:
        assert sha_list(outs) == g["sha"], key
    for o, d in zip([o for o in outs if o.ndim == 2], g["digests"]):
        check_digest(o, d, rtol=spec_tol, atol=spec_tol)


def test_dropin_entry_points_match_reference(wav_dir):
    """Audio outputs (index work) bit-exact; spectrogram digests within 2e-4 (normalised log-mel)
    / 2.3e-3 (fbank, natural log)."""
    from heart_murmur_detection_b200 import extract_feature as EF
    from heart_murmur_detection_b200 import util as U

    for name, *_ in RECORDINGS:
        for kw in (
            dict(input_sec=8, spectrogram=True, pad=True, types="zero", max_sec=32),
            dict(input_sec=8, spectrogram=True, pad=True),
            dict(input_sec=8, spectrogram=True),
            dict(input_sec=2, spectrogram=False, pad=True),
        ):
            tag = ",".join(f"{k}={v}" for k, v in kw.items())
            _check(f"entire/{name}/{tag}", U.get_entire_signal_librosa(wav_dir, name, **kw), 2e-4)
        for kw in (
            dict(input_sec=8.18, spectrogram=True),
            dict(input_sec=4.09, spectrogram=True, trim_tail=True),
            dict(input_sec=2, spectrogram=False),
        ):
            tag = ",".join(f"{k}={v}" for k, v in kw.items())
            _check(f"split/{name}/{tag}", U.get_split_signal_librosa(wav_dir, name, **kw), 2e-4)
        _check(f"fbank_pad/{name}/10", U.get_split_signal_fbank_pad(wav_dir, name, input_sec=10, spectrogram=True), 2.3e-3)
        _check(f"fbank_pad/{name}/2", U.get_split_signal_fbank_pad(wav_dir, name, input_sec=2, spectrogram=True), 2.3e-3)
        _check(f"fbank/{name}/10", EF.get_split_signal_fbank(wav_dir, name, input_sec=10), 2.3e-3)
        _check(f"segments/{name}/8", U.get_individual_segments_librosa(wav_dir, name, input_sec=8, spectrogram=True), 2e-4)
        _check(f"segments_audio/{name}/4", U.get_individual_segments_librosa(wav_dir, name, input_sec=4), 2e-4)


def test_bandpassed_entry_point(wav_dir):
    """butterworth_filter=5: float64 output dtype like the reference; values within 2e-4."""
    from heart_murmur_detection_b200 import util as U

    for name, *_ in RECORDINGS:
        kw = dict(input_sec=8, spectrogram=True, pad=True, types="zero", max_sec=32, butterworth_filter=5)
        tag = ",".join(f"{k}={v}" for k, v in kw.items())
        _check(f"entire/{name}/{tag}", U.get_entire_signal_librosa(wav_dir, name, **kw), 2e-4)


def test_global_rng_side_effect(wav_dir):
    """_duplicate_padding reseeds Python's RNG (src/util.py:564-565); the drop-in reproduces it."""
    from heart_murmur_detection_b200 import util as U

    random.seed(4242)
    U.get_split_signal_librosa(wav_dir, "r_mid", input_sec=2)
    a = random.random()
    random.seed(7456)
    random.random()
    assert a == random.random()
    random.seed(4242)
    U.get_entire_signal_librosa(wav_dir, "r_long", input_sec=8, spectrogram=True, pad=True, types="zero")  # no repeat pad
    b = random.random()
    random.seed(4242)
    assert b == random.random()


def test_pre_process_and_filter_dropins():
    from heart_murmur_detection_b200 import util as U
    from oracle import frontend as F

    x = golden_signal(90000, 4)
    got = U.pre_process_audio_mel_t(x, f_max=8000)
    ref = F.log_mel(x, f_max=8000)
    assert got.dtype == ref.dtype and got.shape == ref.shape and np.abs(got - ref).max() <= 2e-4
    y = U._butter_bandpass_filter(x, 200, 1800, SR, order=5)
    yr = F.butter_bandpass_filter(x, 200, 1800, SR, order=5)
    assert y.dtype == np.float64 and np.abs(y - yr).max() <= 1e-6


def test_c2_pipeline_matches_oracle_per_clip():
    """OPERA-CT linear-probe front-end (model_util.py:161-163 + band-pass): ragged batch through
    the batched pipeline equals the per-clip oracle: frame counts exact, normalised log-mel <= 2e-4."""
    from heart_murmur_detection_b200 import pipeline as pl
    from heart_murmur_detection_b200 import synth
    from oracle import frontend as F

    lens = synth.clip_lengths("c2", 24, seed=99)
    lens[:4] = [3 * SR, 7 * SR + 123, 8 * SR + 1, 40 * SR]
    wav, off = synth.make_batch(lens, base_seed=5, device="cuda")
    res = pl.entire_signal_batch(wav, off, input_sec=8, butterworth_filter=5, spectrogram=True, pad=True, types="zero",
                                 max_sec=32)
    host = wav.cpu().numpy()
    assert res.chunks.valid.all() and len(res.chunks.starts) == len(lens)
    for i in range(len(lens)):
        ref = F.entire_signal(host[off[i] : off[i + 1]], input_sec=8, butterworth_filter=5, spectrogram=True, pad=True,
                              types="zero", max_sec=32)
        got = res.chunk(i).cpu().numpy()
        assert got.shape == ref.shape, (i, got.shape, ref.shape)
        assert np.abs(got - ref).max() <= 2e-4, i


@pytest.mark.parametrize("filt", [5, None])
def test_host_buffer_pipeline_equals_device_pipeline(filt):
    """entire_signal_from_host (pinned host in / out, sub-batches over three streams, reused
    device buffers) returns exactly what entire_signal_batch computes on a resident batch."""
    from heart_murmur_detection_b200 import pipeline, synth

    lens = synth.clip_lengths("c2", 60, seed=21)
    lens[3], lens[17], lens[40] = 20000, 90000, 640000  # padded, padded, cut at max_sec
    wav, off = synth.make_batch(lens, base_seed=500, device="cuda")
    kw = dict(input_sec=8, butterworth_filter=filt, pad=True, types="zero", max_sec=32)
    ref = pipeline.entire_signal_batch(wav, off, spectrogram=True, **kw)
    rows = int(ref.row_offsets[-1])
    h_wav = torch.empty(wav.numel(), dtype=torch.float32, pin_memory=True)
    h_wav.copy_(wav)
    for chunk_bytes in (8 << 20, 3 << 20, 1 << 30):  # many / more (buffers grow and are reused) / one sub-batch
        h_out, ro, clip_ids, valid = pipeline.entire_signal_from_host(h_wav, off, chunk_bytes=chunk_bytes, **kw)
        assert np.array_equal(ro, ref.row_offsets) and np.array_equal(clip_ids, ref.chunks.clip_ids)
        assert np.array_equal(valid, ref.chunks.valid)
        # sub-batching changes the IIR chunk length (warm-up boundaries): equal up to float32 rounding
        # of the filtered samples, i.e. far below the 2e-4 budget of the normalised log-mel
        tol = 0.0 if filt is None else 5e-5
        assert (h_out[:rows] - ref.features[:rows].cpu()).abs().max().item() <= tol
    # a second call with a caller-provided output buffer reuses every device buffer
    h_out2 = torch.empty((rows, 64), dtype=torch.float32, pin_memory=True)
    pipeline.entire_signal_from_host(h_wav, off, h_out2, chunk_bytes=8 << 20, **kw)
    assert (h_out2 - ref.features[:rows].cpu()).abs().max().item() <= (0.0 if filt is None else 5e-5)


def test_logmel_from_host_equals_device():
    from heart_murmur_detection_b200 import frontend as fe
    from heart_murmur_detection_b200 import synth

    lens = synth.clip_lengths("c2", 40, seed=3)
    wav, off = synth.make_batch(lens, base_seed=900, device="cuda")
    plan = fe.logmel_plan(16000, 64, 50, 8000, 1024, 512)
    ref, fo = plan(wav, off)
    h_wav = torch.empty(wav.numel(), dtype=torch.float32, pin_memory=True)
    h_wav.copy_(wav)
    h_out, fo2 = fe.logmel_from_host(plan, h_wav, off, chunk_bytes=4 << 20)
    assert np.array_equal(fo, fo2)
    assert torch.equal(h_out, ref.cpu())


def test_individual_cycles_match_reference(wav_dir):
    """get_individual_cycles_librosa (src/util.py:374-422): labels, slice bounds and dtypes exact;
    unfiltered audio bit-exact, band-passed audio (float64) within 1e-6 of scipy's lfilter."""
    import pandas as pd

    from cases import CYCLES
    from heart_murmur_detection_b200 import util as U

    ann = pd.DataFrame(CYCLES, columns=["Start", "End", "Crackles", "Wheezes", "Disease"])
    for split, n_cls in (("cycle", 4), ("cycle", 2), ("diagnosis", 3), ("diagnosis", 2)):
        for bw in (None, 5):
            g = META[f"cycles/{split}/{n_cls}/{bw}"]
            out = U.get_individual_cycles_librosa(split, ann, wav_dir, "r_long", SR, n_cls, butterworth_filter=bw)
            assert [lab for _, lab in out] == g["labels"]
            assert [len(a) for a, _ in out] == g["lens"]
            assert sorted({str(a.dtype) for a, _ in out}) == g["dtypes"]
            if bw is None:
                assert sha_list([a for a, _ in out]) == g["sha"]
            for (a, _), d in zip(out, g["digests"]):
                check_digest(a, d, rtol=1e-6, atol=1e-6)


def test_fbank_from_host_equals_device():
    from heart_murmur_detection_b200 import frontend as fe
    from heart_murmur_detection_b200 import synth

    lens = np.concatenate([synth.clip_lengths("c3", 6, seed=1), [300, 163840 + 77, 20000, 401]])
    wav, off = synth.make_batch(lens, base_seed=40, device="cuda")
    plan = fe.fbank_plan(sample_rate=16000)
    h_wav = torch.empty(wav.numel(), dtype=torch.float32, pin_memory=True)
    h_wav.copy_(wav)
    for rows_per_clip in (0, 1024):
        ref, ro = plan(wav, off, rows_per_clip=rows_per_clip)
        h_out, ro2 = fe.fbank_from_host(plan, h_wav, off, rows_per_clip=rows_per_clip, chunk_bytes=1 << 20)
        assert np.array_equal(ro, ro2)
        assert torch.equal(h_out, ref.cpu())


def test_pcm16_host_input_equals_decoded_float_input():
    """16-bit PCM payload in host memory (decoded on the device) == the same samples decoded on the
    host the way soundfile / librosa.load do (int16 / 32768), bit for bit."""
    from heart_murmur_detection_b200 import frontend as fe
    from heart_murmur_detection_b200 import pipeline, synth

    lens = synth.clip_lengths("c2", 24, seed=8)
    wav, off = synth.make_batch(lens, base_seed=70, device="cpu")
    pcm = torch.clamp(torch.round(wav * 32768.0), -32768, 32767).to(torch.int16)
    decoded = (pcm.to(torch.float32) / 32768.0).contiguous()
    for n in (0, 1, 7, 8, 9, 4097):  # vector body + scalar tail
        d = fe.pcm16_to_f32(pcm[:n].cuda().contiguous()) if n else torch.empty(0)
        assert torch.equal(d.cpu(), decoded[:n])
    odd = fe.pcm16_to_f32(pcm.cuda()[1:].contiguous() if False else pcm[1:4098].cuda())  # fresh allocation: aligned
    assert torch.equal(odd.cpu(), decoded[1:4098])
    kw = dict(input_sec=8, butterworth_filter=5, pad=True, types="zero", max_sec=32)
    h_pcm = pcm.pin_memory()
    h_f32 = decoded.pin_memory()
    a, ro_a, ids_a, valid_a = pipeline.entire_signal_from_host(h_pcm, off, chunk_bytes=16 << 20, **kw)
    b, ro_b, ids_b, valid_b = pipeline.entire_signal_from_host(h_f32, off, chunk_bytes=16 << 20, **kw)
    assert np.array_equal(ro_a, ro_b) and np.array_equal(ids_a, ids_b) and np.array_equal(valid_a, valid_b)
    rows = int(ro_a[-1])
    assert torch.equal(a[:rows], b[:rows])
    # the log-mel and fbank host-buffer entries take the same payload
    lm = fe.logmel_plan(16000, 64, 50, 8000, 1024, 512)
    x1, _ = fe.logmel_from_host(lm, h_pcm, off, chunk_bytes=8 << 20)
    x2, _ = fe.logmel_from_host(lm, h_f32, off, chunk_bytes=8 << 20)
    assert torch.equal(x1, x2)
    fb = fe.fbank_plan(sample_rate=16000)
    y1, _ = fe.fbank_from_host(fb, h_pcm, off, chunk_bytes=8 << 20)
    y2, _ = fe.fbank_from_host(fb, h_f32, off, chunk_bytes=8 << 20)
    assert torch.equal(y1, y2)


def test_c4_cola_pairs_on_the_fly_match_oracle():
    """BASELINE config 4: recordings -> whole-recording log-mel (get_entire_signal_librosa,
    heart_pressl.py:76-81) -> COLA item = random_mask, two random_crops, two random_multiplies
    (cola_training.py:56-80), all on the GPU with host-drawn random numbers.  Crop starts and
    masked rows are exact (same Python RNG stream), values within the log-mel tolerance."""
    from heart_murmur_detection_b200 import datasets, pipeline, synth
    from oracle import frontend as F

    lens = synth.clip_lengths("c4", 6, seed=11)
    lens[2] = 7 * SR  # shorter than input_sec: the producer drops it ("audio too short")
    wav, off = synth.make_batch(lens, base_seed=4400, device="cuda")
    fb = pipeline.entire_signal_batch(wav, off, input_sec=8, spectrogram=True)
    assert list(fb.chunks.valid) == [True, True, False, True, True, True]
    store = datasets.SpecStore.from_features(fb.features, fb.row_offsets)
    host = wav.cpu().numpy()
    specs = [F.entire_signal(host[off[i] : off[i + 1]], input_sec=8, spectrogram=True) for i in range(len(lens))]
    specs = [s for s in specs if s is not None]
    assert len(specs) == 5 and [s.shape[0] for s in specs] == [store.rows(i) for i in range(5)]
    items = [3, 0, 4, 1, 2, 0]
    random.seed(2024)
    x1, x2 = datasets.cola_batch(store, items, max_len=251, augment=True)
    state_after = random.getstate()
    random.seed(2024)
    for k, idx in enumerate(items):
        x = F.random_mask(specs[idx])
        r1 = F.random_crop(x, crop_size=251)
        r2 = F.random_crop(x, crop_size=251)
        r1, r2 = F.random_multiply(r1), F.random_multiply(r2)
        for got, ref in ((x1[k], r1), (x2[k], r2)):
            got = got.cpu().numpy()
            assert got.shape == (251, 64)
            assert np.abs(got[: ref.shape[0]] - ref).max() <= 3e-4
    assert random.getstate() == state_after  # the same number of draws was consumed


def test_cache_writers_match_per_file_entry_points(wav_dir, tmp_path):
    """The batched cache writers (entire_spec_npy, spectrogram_pad8.npy, fbank_audiomae.npy) hold what
    the reference's loops would save: the per-file drop-in results, which are pinned to the reference
    by test_dropin_entry_points_match_reference."""
    from heart_murmur_detection_b200 import caches
    from heart_murmur_detection_b200 import util as U

    names = [name for name, *_ in RECORDINGS]
    files = [os.path.join(wav_dir, n + ".wav") for n in names]
    feature_dir = str(tmp_path)
    written, invalid = caches.write_entire_spec_cache(files, feature_dir, input_sec=8, batch_bytes=1 << 20)
    per_file = {n: U.get_entire_signal_librosa(wav_dir, n, spectrogram=True, input_sec=8) for n in names}
    assert invalid == sum(v is None for v in per_file.values()) and invalid >= 1
    assert [os.path.basename(w) for w in written] == [n for n in names if per_file[n] is not None]
    assert list(np.load(os.path.join(feature_dir, "entire_spec_filenames.npy"))) == written
    for w in written:
        got = np.load(w + ".npy")
        ref = per_file[os.path.basename(w)]
        assert got.dtype == np.float32 and got.shape == ref.shape and np.array_equal(got, ref)

    pad8 = caches.build_spectrogram_pad_cache(files, input_sec=8.18, batch_bytes=1 << 20, save_to=os.path.join(feature_dir, "spectrogram_pad8.npy"))
    assert pad8.shape == (len(names), 256, 64) and pad8.dtype == np.float32
    state = random.getstate()
    for i, n in enumerate(names):
        ref = U.get_split_signal_librosa(wav_dir, n, spectrogram=True, input_sec=8.18, trim_tail=False)[0]
        assert np.array_equal(pad8[i], ref), n
    random.setstate(state)
    assert np.array_equal(np.load(os.path.join(feature_dir, "spectrogram_pad8.npy")), pad8)

    fb = caches.build_fbank_cache(files, input_sec=10, batch_bytes=1 << 20)
    assert fb.shape == (len(names), 998, 128)
    for i, n in enumerate(names):
        ref = U.get_split_signal_fbank_pad(wav_dir, n, spectrogram=True, input_sec=10, trim_tail=False)[0]
        assert np.array_equal(fb[i], ref.numpy()), n


def test_cache_writers_resample_native_rate_files(tmp_path):
    """CirCor recordings are 4 kHz PCM16: every file takes the resampler.  The cache writers decode
    in worker threads but resample whole batches on the calling thread (one plan per thread, include/hmfe.h);
    the result must equal the per-file drop-in, which resamples one file at a time."""
    from heart_murmur_detection_b200 import audio_io, caches
    from heart_murmur_detection_b200 import util as U

    d = tmp_path / "wav4k"
    d.mkdir()
    names = []
    for i, (sec, native) in enumerate([(9.0, 4000), (12.5, 4000), (3.0, 4000), (10.0, 8000), (9.5, 16000), (20.0, 4000),
                                       (8.7, 2000), (11.0, 4000)]):
        n = int(sec * native)
        x = golden_signal(n * (SR // native), 900 + i, SR, 0.2, 0.3)[:: SR // native]  # decimated test tone
        audio_io.write_wav_pcm16(str(d / f"r{i}.wav"), x, native)
        names.append(f"r{i}")
    files = [str(d / (n + ".wav")) for n in names]
    for workers in (1, 8):
        feature_dir = str(tmp_path / f"feat{workers}")
        os.makedirs(feature_dir)
        written, invalid = caches.write_entire_spec_cache(files, feature_dir, input_sec=8, batch_bytes=1 << 19, workers=workers)
        per_file = {n: U.get_entire_signal_librosa(str(d), n, spectrogram=True, input_sec=8) for n in names}
        assert invalid == sum(v is None for v in per_file.values()) == 1
        assert [os.path.basename(w) for w in written] == [n for n in names if per_file[n] is not None]
        for w in written:
            got, ref = np.load(w + ".npy"), per_file[os.path.basename(w)]
            assert got.shape == ref.shape and np.array_equal(got, ref), w


def test_host_entry_from_native_rate_pcm16_matches_device_path_and_oracle():
    """pipeline.entire_signal_from_host(sr_in=4000) - the 16-bit payload at the recordings' native rate crosses PCIe,
    decode + rate conversion run on the GPU (librosa.load(path, sr=16000) of src/util.py:222) - equals the device
    path (resample, then entire_signal_batch) bit for bit over several sub-batches, and the CPU oracle
    (torchaudio.functional.resample + the restated reference path) within the log-mel tolerance."""
    from heart_murmur_detection_b200 import frontend as fe
    from heart_murmur_detection_b200 import pipeline as pl
    from oracle import frontend as F

    kw = dict(input_sec=8, butterworth_filter=5, pad=True, types="zero", max_sec=32)
    secs = [9.0, 3.1, 12.5, 40.0, 8.0, 1.0, 20.3, 5.0, 33.0]
    pcm = [np.clip(np.round(golden_signal(int(s_ * 16000), 70 + i, SR, 0.1, 0.2)[::4] * 32768.0), -32768, 32767).astype(np.int16)
           for i, s_ in enumerate(secs)]
    o4 = np.zeros(len(pcm) + 1, dtype=np.int64)
    np.cumsum([len(c) for c in pcm], out=o4[1:])
    h_pcm = torch.from_numpy(np.concatenate(pcm)).pin_memory()
    h_out, ro, clip_ids, valid = pl.entire_signal_from_host(h_pcm, o4, sr_in=4000, chunk_bytes=4 << 20, **kw)
    assert valid.all() and list(clip_ids) == list(range(len(pcm)))
    rplan = fe.resample_plan(4000, 16000, **fe.RESAMPLE_PRESETS["torchaudio"])
    wav16, o16 = rplan(h_pcm.cuda(), o4)
    ref = pl.entire_signal_batch(wav16, o16, spectrogram=True, **kw)
    rows = int(ref.row_offsets[-1])
    assert np.array_equal(ro, ref.row_offsets)
    assert torch.equal(h_out[:rows], ref.features[:rows].cpu())
    for i in (1, 3, 5):
        x16 = F.resample_torchaudio(pcm[i].astype(np.float32) / np.float32(32768.0), 4000, 16000)
        want = F.entire_signal(x16, spectrogram=True, **kw)
        got = h_out[int(ro[i]) : int(ro[i + 1])].numpy()
        assert got.shape == want.shape and np.abs(got - want).max() <= 3e-4


def test_c2_full_size_properties():
    """BASELINE config 2 at full size (5 272 ragged clips, 1.78 G samples) through size-independent
    properties: the one-pass overlap band-pass equals the exact chunked scan on every sample; trim
    indices are identical for both; every clip's normalised log-mel spans exactly [0, 1] with the
    reference's frame count; 20 clips spread over the batch equal the per-clip CPU oracle."""
    from heart_murmur_detection_b200 import frontend as fe
    from heart_murmur_detection_b200 import pipeline as pl
    from heart_murmur_detection_b200 import synth
    from oracle import frontend as F

    lens = synth.clip_lengths("c2", 5272, seed=1234)
    wav, off = synth.make_batch(lens, base_seed=0, device="cuda")
    sos = fe.butter_bandpass_sos(200, 1800, SR, 5)
    ctx_o, ctx_s = fe.Context(), fe.Context()
    ctx_s.set_iir_algo("scan")
    y_o, se_o = fe.iir_sos_trim(wav, off, sos, ctx=ctx_o)
    assert ctx_o.last_iir_plan()["algo"] == "overlap" and ctx_o.last_iir_plan()["rows"] == "vector"
    y_s, se_s = fe.iir_sos_trim(wav, off, sos, ctx=ctx_s)
    assert ctx_s.last_iir_plan()["algo"] == "scan"
    assert (y_o - y_s).abs().max().item() <= 2e-7  # both float32 outputs of float64 recurrences: last-bit differences only
    assert torch.equal(se_o, se_s)
    del y_o, y_s
    kw = dict(input_sec=8, butterworth_filter=5, pad=True, types="zero", max_sec=32)
    res = pl.entire_signal_batch(wav, off, spectrogram=True, **kw)
    assert res.chunks.valid.all() and len(res.chunks.starts) == 5272
    T = np.diff(res.row_offsets)
    n_trim = (se_o[:, 1] - se_o[:, 0]).cpu().numpy()
    expect = 1 + np.minimum(np.maximum(n_trim, 8 * SR), 32 * SR) // 512  # pad to 8 s, cut at 32 s, 1 + n // hop frames
    assert np.array_equal(T, expect)
    feats = res.features[: int(res.row_offsets[-1])]
    assert torch.isfinite(feats).all()
    seg = torch.from_numpy(np.repeat(np.arange(5272), T)).cuda()
    mx = torch.full((5272,), -1.0, device="cuda").scatter_reduce(0, seg, feats.amax(dim=1), "amax")
    mn = torch.full((5272,), 2.0, device="cuda").scatter_reduce(0, seg, feats.amin(dim=1), "amin")
    assert torch.all(mx == 1.0) and torch.all(mn == 0.0)
    host_idx = np.linspace(0, 5271, 20).astype(int)
    for i in host_idx:
        x = wav[off[i] : off[i + 1]].cpu().numpy()
        ref = F.entire_signal(x, spectrogram=True, **kw)
        got = res.chunk(int(i)).cpu().numpy()
        assert got.shape == ref.shape and np.abs(got - ref).max() <= 2e-4, int(i)


def test_c3_full_size_properties():
    """BASELINE config 3 at full size (1000 x 10.24 s -> [1024, 128]): pad rows are zero, frames are
    finite, a batch equals its clips run alone, and shifting a clip by one frame shift (160 samples)
    shifts its fbank rows by one, within the fbank tolerance (snip_edges framing has no other coupling)."""
    from heart_murmur_detection_b200 import frontend as fe
    from heart_murmur_detection_b200 import synth

    lens = synth.clip_lengths("c3", 1000)
    wav, off = synth.make_batch(lens, base_seed=21, device="cuda")
    plan = fe.fbank_plan(sample_rate=16000)
    out, ro = plan(wav, off, rows_per_clip=1024)
    out = out.view(1000, 1024, 128)
    assert torch.isfinite(out).all() and torch.all(out[:, 1022:] == 0.0) and torch.all(out[:, :1022].amax(dim=2) > -15.9)
    for i in (0, 500, 999):
        single, _ = plan(wav[off[i] : off[i + 1]].clone(), np.array([0, lens[i]]), rows_per_clip=1024)
        assert torch.equal(single.view(1024, 128), out[i])
    # the frames land in other lanes / packed pairs of the warp: equal within the fbank tolerance, not bit for bit
    shifted, _ = plan(wav[off[7] + 160 : off[8]].clone(), np.array([0, lens[7] - 160]))
    assert (shifted - out[7, 1:1022]).abs().max().item() <= 2.3e-3


@pytest.mark.parametrize("bp", [None, 5])
@pytest.mark.parametrize("pad,types,max_sec", [(True, "zero", 32), (True, "repeat", None), (False, "repeat", 32), (True, "zero", None)])
def test_device_planner_equals_host_planner(bp, pad, types, max_sec):
    """hmfe_entire_plan_batch (duration test, pad / cut decisions, padding descriptors, row and work-item offsets on the
    device) against the numpy planner that reads the trim indices back: identical chunk tables, identical features."""
    from heart_murmur_detection_b200 import pipeline as pl
    from heart_murmur_detection_b200 import synth

    lens = synth.clip_lengths("c2", 96, seed=31)
    lens[:8] = [1, 2000, 16000, 63999, 64000, 64001, 127999, 128000]  # knife edges of the pad rules (L = 128000)
    lens[8:12] = [128001, 511999, 512000, 512001]                      # and of the 32 s cut
    wav, off = synth.make_batch(lens, base_seed=9100, device="cuda")
    wav[off[20] : off[21]] = 0.0                                       # an all-silent recording: trims to nothing
    kw = dict(input_sec=8, butterworth_filter=bp, pad=pad, types=types, max_sec=max_sec, spectrogram=True)
    host = pl.entire_signal_batch(wav, off, planner="host", **kw)
    dev = pl.entire_signal_batch(wav, off, planner="device", **kw)
    np.testing.assert_array_equal(dev.row_offsets, host.row_offsets)
    np.testing.assert_array_equal(dev.chunks.clip_ids, host.chunks.clip_ids)
    np.testing.assert_array_equal(dev.chunks.valid, host.chunks.valid)
    np.testing.assert_array_equal(dev.chunks.trim, host.chunks.trim)
    np.testing.assert_array_equal(dev.chunks.lengths, host.chunks.lengths)
    np.testing.assert_array_equal(dev.chunks.is_view, host.chunks.is_view)
    assert dev.chunks.used_duplicate_padding == host.chunks.used_duplicate_padding
    rows = int(host.row_offsets[-1])
    assert torch.equal(dev.features[:rows], host.features[:rows])
    for k in range(len(host.chunks.starts)):  # the padded copies themselves
        if not host.chunks.is_view[k]:
            assert torch.equal(dev.chunks.samples(k), host.chunks.samples(k))
