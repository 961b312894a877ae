"""Parity of the remaining CUDA stages (through the C ABI) with the oracle / live third-party
libraries / golden fixtures: IIR band-pass, silence trim, pad-split gather, Kaldi fbank,
resampler, spectrogram-domain ops."""
import json
import os
import random

import numpy as np
import pytest
import torch

from cases import PAD_SPLIT_LENGTHS, PAD_SPLIT_SECS, RECORDINGS, SR, hash_spec, sha, sha_list
from signals import golden_signal

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
META = json.load(open(os.path.join(HERE, "golden", "ref_util.json")))["cases"]


def _batch(clips):
    off = np.zeros(len(clips) + 1, dtype=np.int64)
    np.cumsum([len(c) for c in clips], out=off[1:])
    return torch.from_numpy(np.concatenate(clips).astype(np.float32)).cuda(), off


def rng_fingerprint():
    return sha(np.array(random.getstate()[1], dtype=np.uint64))


# ------------------------------------------------------------------------------------------- IIR


@pytest.mark.parametrize("order", [5, 3, 1])
def test_iir_matches_scipy_lfilter(order):
    """<= 1e-4 * max|y|; in practice ~1e-8 (float64 cascade vs float64 TF form)."""
    from scipy.signal import butter, lfilter

    from heart_murmur_detection_b200 import frontend as fe

    lens = [1, 2, 127, 128, 129, 255, 256, 257, 4095, 4096, 4097, 128000, 90001, 12, 300000]
    clips = [golden_signal(n, seed=3 + i) for i, n in enumerate(lens)]
    clips[3] = np.zeros(128, np.float32)
    clips[3][0] = 1.0  # impulse
    clips[4] = np.ones(129, np.float32)  # step
    wav, off = _batch(clips)
    sos = fe.butter_bandpass_sos(200, 1800, SR, order)
    b, a = butter(order, [200 / 8000, 1800 / 8000], btype="band")
    y64 = fe.iir_sos(wav, off, sos, out_dtype=torch.float64).cpu().numpy()
    y32 = fe.iir_sos(wav, off, sos, out_dtype=torch.float32).cpu().numpy()
    # scipy's own SOS pairing has numerators like (1, 2, 1) / (1, -2, 1): the general (non band-pass-form) path
    y64g = fe.iir_sos(wav, off, butter(order, [200 / 8000, 1800 / 8000], btype="band", output="sos"),
                      out_dtype=torch.float64).cpu().numpy()
    for i, x in enumerate(clips):
        ref = lfilter(b, a, x)
        assert ref.dtype == np.float64
        tol = 1e-4 * max(np.abs(ref).max(), 1e-12)
        assert np.abs(y64[off[i] : off[i + 1]] - ref).max() <= min(tol, 1e-6)
        assert np.abs(y64g[off[i] : off[i + 1]] - ref).max() <= min(tol, 1e-6)
        assert np.abs(y32[off[i] : off[i + 1]] - ref).max() <= tol


def test_iir_large_batch_uses_long_chunks_and_is_linear():
    """> 32 Mi samples switches to 512-sample chunks; check against scipy on a few clips and
    linearity (filter(a*x + b*y) == a*filter(x) + b*filter(y)) on the whole batch."""
    from scipy.signal import butter, lfilter

    from heart_murmur_detection_b200 import frontend as fe
    from heart_murmur_detection_b200 import synth

    lens = synth.clip_lengths("c2", 120, seed=5)
    lens[0] = 40 * SR * 16  # one very long clip to cross 32 Mi samples in total
    assert lens.sum() > (32 << 20)
    wav, off = synth.make_batch(lens, base_seed=77, device="cuda")
    sos = fe.butter_bandpass_sos(200, 1800, SR, 5)
    y = fe.iir_sos(wav, off, sos, out_dtype=torch.float64)
    b, a = butter(5, [200 / 8000, 1800 / 8000], btype="band")
    for i in (0, 1, 57, 119):
        ref = lfilter(b, a, wav[off[i] : off[i + 1]].cpu().numpy())
        assert np.abs(y[off[i] : off[i + 1]].cpu().numpy() - ref).max() <= 1e-7 * max(1.0, np.abs(ref).max())
    wav2 = torch.roll(wav, 12345)
    y2 = fe.iir_sos(wav2, off, sos, out_dtype=torch.float64)
    y3 = fe.iir_sos((0.5 * wav + 0.25 * wav2), off, sos, out_dtype=torch.float64)
    assert (y3 - (0.5 * y + 0.25 * y2)).abs().max().item() <= 1e-6


@pytest.mark.parametrize("order", [5, 2])
def test_iir_overlap_matches_scan_and_scipy(order):
    """The one-pass overlap kernel (chunks warmed up over W samples) against scipy.lfilter and
    against the exact scan, on lengths around the chunk / warm-up / hop boundaries."""
    from scipy.signal import butter, lfilter

    from heart_murmur_detection_b200 import frontend as fe

    sos = fe.butter_bandpass_sos(200, 1800, SR, order)
    b, a = butter(order, [200 / 8000, 1800 / 8000], btype="band")
    ctx_o, ctx_s = fe.Context(), fe.Context()
    ctx_o.set_iir_algo("overlap")
    ctx_s.set_iir_algo("scan")
    probe, poff = _batch([golden_signal(64000, 1)])
    fe.iir_sos(probe, poff, sos, ctx=ctx_o)
    plan = ctx_o.last_iir_plan()
    assert plan["algo"] == "overlap" and plan["warmup"] % 32 == 0 and plan["chunk"] % 32 == 0
    for _ in range(3):  # the chunk length depends on the batch: iterate until the lengths straddle it
        Cc, W = plan["chunk"], plan["warmup"]
        lens = [1, 2, 31, 32, 33, 799, 800, 801, 1599, 1601, W - 1, W, W + 1, Cc - 1, Cc, Cc + 1, 2 * Cc + 5, 3 * Cc,
                128000, 90001, 300000, 5 * Cc + W + 17]
        clips = [golden_signal(n, seed=3 + i) for i, n in enumerate(lens)]
        clips[4] = np.zeros(33, np.float32)
        clips[4][0] = 1.0  # impulse
        clips[5] = np.ones(799, np.float32)  # step
        wav, off = _batch(clips)
        y64 = fe.iir_sos(wav, off, sos, out_dtype=torch.float64, ctx=ctx_o).cpu().numpy()
        assert ctx_o.last_launches == 1
        plan = ctx_o.last_iir_plan()
        assert plan["algo"] == "overlap"
        if plan["chunk"] == Cc:
            break
    y32 = fe.iir_sos(wav, off, sos, out_dtype=torch.float32, ctx=ctx_o).cpu().numpy()
    assert ctx_o.last_iir_plan()["rows"] == "vector", ctx_o.last_iir_plan()
    # a source view that starts 1 element into the allocation: x and y on different 16-byte phases -> scalar rows
    shifted = torch.empty(wav.numel() + 1, dtype=torch.float32, device="cuda")
    shifted[1:].copy_(wav)
    y32s = fe.iir_sos(shifted[1:], off, sos, out_dtype=torch.float32, ctx=ctx_o).cpu().numpy()
    assert ctx_o.last_iir_plan()["rows"] == "scalar"
    # ... and on the same (odd) phase: vector rows again, clips starting at every alignment slot
    outs = torch.zeros(wav.numel() + 1, dtype=torch.float32, device="cuda")
    fe.iir_sos(shifted[1:], off, sos, out=outs[1:], ctx=ctx_o)
    assert ctx_o.last_iir_plan()["rows"] == "vector"
    assert float(outs[0]) == 0.0  # nothing written in front of the first clip
    assert np.abs(outs[1:].cpu().numpy() - y32).max() <= 1e-8  # other chunk boundaries: equal up to the last rounding
    ys = fe.iir_sos(wav, off, sos, out_dtype=torch.float64, ctx=ctx_s).cpu().numpy()
    assert ctx_s.last_iir_plan()["algo"] == "scan"
    # scipy's own pairing -> the general (non band-pass-form) cascade code
    y64g = fe.iir_sos(wav, off, butter(order, [200 / 8000, 1800 / 8000], btype="band", output="sos"),
                      out_dtype=torch.float64, ctx=ctx_o).cpu().numpy()
    for i, x in enumerate(clips):
        ref = lfilter(b, a, x)
        scale = max(np.abs(ref).max(), 1e-12)
        seg = slice(off[i], off[i + 1])
        assert np.abs(y64[seg] - ref).max() <= min(1e-4 * scale, 1e-6), (i, lens[i])
        assert np.abs(y64g[seg] - ref).max() <= min(1e-4 * scale, 1e-6), (i, lens[i])
        assert np.abs(y32[seg] - ref).max() <= 1e-4 * scale, (i, lens[i])
        assert np.abs(y64[seg] - ys[seg]).max() <= 1e-10 * max(scale, 1e-3), (i, lens[i])
        # the float32 output takes the 128-bit-row kernel: same values up to the final rounding
        assert np.abs(y32[seg] - ys[seg]).max() <= 2e-7 * max(scale, 1e-3), (i, lens[i])
        assert np.abs(y32s[seg] - ys[seg]).max() <= 2e-7 * max(scale, 1e-3), (i, lens[i])


def test_iir_overlap_refused_for_slow_filters():
    """A pole at radius 0.99995 needs far more than 8192 warm-up samples: forced overlap is an
    error, auto falls back to the exact scan and still matches scipy."""
    from scipy.signal import lfilter

    from heart_murmur_detection_b200 import _lib
    from heart_murmur_detection_b200 import frontend as fe

    r, th = 0.99995, 0.3
    sos = np.array([[1.0, 0.0, 0.0, 1.0, -2 * r * np.cos(th), r * r]])
    x = golden_signal(50000, 5)
    wav, off = _batch([x])
    ctx = fe.Context()
    y = fe.iir_sos(wav, off, sos, out_dtype=torch.float64, ctx=ctx).cpu().numpy()
    assert ctx.last_iir_plan()["algo"] == "scan"
    ref = lfilter(sos[0, :3], sos[0, 3:], x)
    assert np.abs(y - ref).max() <= 1e-9 * np.abs(ref).max()
    ctx.set_iir_algo("overlap")
    with pytest.raises(_lib.HmfeError):
        fe.iir_sos(wav, off, sos, ctx=ctx)


@pytest.mark.parametrize("algo", ["overlap", "overlap-scalar", "scan", "auto"])
def test_iir_trim_fused_indices_exact(algo):
    """hmfe_iir_sos_trim_batch: indices equal librosa.effects.trim of scipy's lfilter output."""
    from heart_murmur_detection_b200 import frontend as fe
    from oracle import frontend as F

    clips = [golden_signal(n, seed, SR, lead, tail) for _, n, seed, lead, tail in RECORDINGS]
    for lead, tail in [(0, 0), (799, 801), (800, 1600), (2400, 0), (0, 4000), (5, 17)]:
        clips.append(golden_signal(48000, 11, SR, lead, tail))
    clips.append(np.zeros(5000, np.float32))
    clips.append(golden_signal(30000, 12, SR, 0, 0))
    clips.append(golden_signal(700, 13, SR, 0, 0))
    clips.append(golden_signal(800, 14, SR, 0, 0))
    clips.append(golden_signal(25600 + 799, 15, SR, 3000, 0))
    clips.append(golden_signal(300000, 16, SR, 20000, 33000))
    # in-band tone bursts between very quiet edges: after the band-pass the edges fall > 60 dB below the
    # burst, so the trim indices are non-trivial (the heart tones of golden_signal are out of band)
    for k, (lead, body, tail) in enumerate([(4000, 30000, 9000), (801, 16000, 1599), (0, 52000, 20000), (12345, 3200, 0)]):
        t = np.arange(lead + body + tail)
        x = 1e-6 * ((t * 7919 + k) % 13 - 6.0)
        x[lead : lead + body] += 0.4 * np.sin(2 * np.pi * (450 + 100 * k) * t[lead : lead + body] / SR)
        clips.append(x.astype(np.float32))
    wav, off = _batch(clips)
    ctx = fe.Context()
    ctx.set_iir_algo(algo.split("-")[0])
    if algo.endswith("scalar"):
        ctx.set_iir_rows("scalar")
    sos = fe.butter_bandpass_sos(200, 1800, SR, 5)
    y, se = fe.iir_sos_trim(wav, off, sos, ctx=ctx)
    if algo.startswith("overlap"):
        assert ctx.last_launches == 2  # one filter pass + the index kernel: the signal is not re-read
        assert ctx.last_iir_plan()["rows"] == ("scalar" if algo.endswith("scalar") else "vector")
    y, se = y.cpu().numpy(), se.cpu().numpy()
    for i, x in enumerate(clips):
        ref = F.butter_bandpass_filter(x, 200, 1800, SR, 5)
        _, idx = F.trim_silence(ref, SR)
        assert [int(se[i, 0]), int(se[i, 1])] == [int(idx[0]), int(idx[1])], (i, len(x))
        assert np.abs(y[off[i] : off[i + 1]] - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-12)


def test_iir_auto_picks_overlap_on_large_batches():
    from scipy.signal import butter, lfilter

    from heart_murmur_detection_b200 import frontend as fe
    from heart_murmur_detection_b200 import synth

    lens = synth.clip_lengths("c2", 600, seed=9)
    wav, off = synth.make_batch(lens, base_seed=31, device="cuda")
    sos = fe.butter_bandpass_sos(200, 1800, SR, 5)
    ctx = fe.Context()
    y = fe.iir_sos(wav, off, sos, out_dtype=torch.float64, ctx=ctx)
    assert ctx.last_iir_plan()["algo"] == "overlap", ctx.last_iir_plan()
    b, a = butter(5, [200 / 8000, 1800 / 8000], btype="band")
    for i in (0, 1, 299, 599):
        ref = lfilter(b, a, wav[off[i] : off[i + 1]].cpu().numpy())
        assert np.abs(y[off[i] : off[i + 1]].cpu().numpy() - ref).max() <= 1e-7 * max(1.0, np.abs(ref).max())
    ctx.set_iir_algo("scan")
    ys = fe.iir_sos(wav, off, sos, out_dtype=torch.float64, ctx=ctx)
    assert (y - ys).abs().max().item() <= 1e-10


def test_butter_design_matches_scipy():
    from scipy.signal import butter

    from heart_murmur_detection_b200.util import _butter_bandpass

    for order, lo, hi in [(5, 200, 1800), (3, 25, 400), (2, 100, 1000)]:
        b, a = _butter_bandpass(lo, hi, SR, order)
        bs, as_ = butter(order, [lo / 8000, hi / 8000], btype="band")
        np.testing.assert_allclose(b, bs, rtol=1e-9, atol=1e-15)
        np.testing.assert_allclose(a, as_, rtol=1e-9, atol=1e-15)


# ------------------------------------------------------------------------------------------- trim


def test_trim_indices_exact():
    from heart_murmur_detection_b200 import frontend as fe
    from oracle import frontend as F

    clips = [golden_signal(n, seed, SR, lead, tail) for _, n, seed, lead, tail in RECORDINGS]
    for lead, tail in [(0, 0), (799, 801), (800, 1600), (2400, 0), (0, 4000), (5, 17)]:
        clips.append(golden_signal(48000, 11, SR, lead, tail))
    clips.append(np.zeros(5000, np.float32))                     # all silent -> (0, 0)
    clips.append(golden_signal(30000, 12, SR, 0, 0))             # all loud
    clips.append(golden_signal(700, 13, SR, 0, 0))               # shorter than one hop
    clips.append((golden_signal(20000, 14, SR, 0, 0) * 1e-4).astype(np.float32))  # quiet everywhere (relative threshold)
    wav, off = _batch(clips)
    se = fe.trim_indices(wav, off).cpu().numpy()
    for i, x in enumerate(clips):
        _, idx = F.trim_silence(x, SR)
        assert [int(se[i, 0]), int(se[i, 1])] == [int(idx[0]), int(idx[1])], i
    for k, (name, *_r) in enumerate(RECORDINGS):
        assert [int(se[k, 0]), int(se[k, 1])] == META[f"trim/{name}"]


# ------------------------------------------------------------------------------------------- pad / split


@pytest.mark.parametrize("sec", PAD_SPLIT_SECS)
def test_split_pad_sample_bit_exact_vs_reference(sec):
    """Golden hashes were produced by the reference's own split_pad_sample / split_sample."""
    from heart_murmur_detection_b200 import util as U
    from heart_murmur_detection_b200.extract_feature import split_sample

    for n in PAD_SPLIT_LENGTHS:
        x = golden_signal(n, seed=n % 97, sr=SR, lead=0, tail=0)
        for types_ in ("repeat", "zero"):
            random.seed(99)
            out = U.split_pad_sample([x, 0, 0], sec, SR, types=types_)
            g = META[f"split_pad/{sec}/{n}/{types_}"]
            assert len(out) == g["n_chunks"]
            assert sha_list([o[0] for o in out]) == g["sha"], (sec, n, types_)
            assert sorted({str(o[0].dtype) for o in out}) == g["dtypes"]
            assert rng_fingerprint() == g["rng_after"]
        out = split_sample(x, sec, SR)
        g = META[f"split_sample/{sec}/{n}"]
        assert [len(o) for o in out] == g["lens"] and sha_list(out) == g["sha"]
        assert bool(U.decide_droplast(x, SR, sec)) == META[f"droplast/{sec}/{n}"]


# ------------------------------------------------------------------------------------------- fbank


def _kaldi_ref(x):
    import torchaudio

    w = torch.tensor(x - x.mean()).reshape(1, -1)
    return torchaudio.compliance.kaldi.fbank(
        w, channel=0, frame_length=25, htk_compat=True, sample_frequency=SR, use_energy=False, window_type="hanning",
        num_mel_bins=128, dither=0.0, frame_shift=10,
    ).numpy()


def test_fbank_matches_torchaudio_kaldi():
    """<= 2.3e-3 nat (= 1e-2 dB); floor entries (empty mel row 3) exact; frame counts exact."""
    from heart_murmur_detection_b200.frontend import FbankPlan

    plan = FbankPlan()
    lens = [401, 560, 32000, 128000, 160000, 163840, 400, 719, 720, 721]
    clips = [golden_signal(n, seed=21 + i, lead=0, tail=0) for i, n in enumerate(lens)]
    clips.append((golden_signal(50000, 40, lead=0, tail=0) + 0.3).astype(np.float32))  # DC offset
    wav, off = _batch(clips)
    out, ro = plan(wav, off)
    out = out.cpu().numpy()
    floor = np.float32(np.log(np.float32(1.1920929e-7)))
    for i, x in enumerate(clips):
        ref = _kaldi_ref(x)
        got = out[ro[i] : ro[i + 1]]
        assert got.shape == ref.shape == (1 + (len(x) - 400) // 160, 128)
        assert np.abs(got - ref).max() <= 2.3e-3
        np.testing.assert_array_equal(got[:, 3], ref[:, 3])
        assert (got[:, 3] == floor).all()
    # fewer than one frame -> no rows; padded layout [n, 1024, 128] with zero rows
    w2, o2 = _batch([clips[5], golden_signal(399, 1), clips[2]])
    padded, ro2 = plan(w2, o2, rows_per_clip=1024)
    padded = padded.view(3, 1024, 128).cpu().numpy()
    assert np.abs(padded[0, :1022] - _kaldi_ref(clips[5])).max() <= 2.3e-3
    assert not padded[0, 1022:].any() and not padded[1].any() and not padded[2, 198:].any()
    # silence: every bin at the floor
    z, oz = _batch([np.zeros(4000, np.float32)])
    fz, _ = plan(z, oz)
    assert (fz.cpu().numpy() == floor).all()


def test_fbank_mel_basis_matches_torchaudio():
    import torchaudio

    from heart_murmur_detection_b200.frontend import FbankPlan

    ours = FbankPlan().mel_basis()
    ref, _ = torchaudio.compliance.kaldi.get_mel_banks(128, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    ref = torch.nn.functional.pad(ref, (0, 1)).numpy()
    # torchaudio evaluates the mel scale in float32 (torch.log on float32 tensors); our bank is the
    # float64 formula rounded once, so the two differ by float32 round-off of the mel values
    assert np.abs(ours - ref).max() <= 5e-5
    assert int((ours != 0).sum()) == int((ref != 0).sum()) == 504 and not ours[3].any()


# ------------------------------------------------------------------------------------------- resample


@pytest.mark.parametrize("sr_in", [4000, 2000, 8000, 44100, 22050, 48000])
def test_resample_matches_torchaudio(sr_in):
    """Same-algorithm oracle (torchaudio.transforms.Resample, src/model/models_eval.py:964-968):
    <= 1e-6 * max|x|; output length ceil(n * 16000 / sr) exact.  soxr parity is unpinned."""
    import torchaudio

    from heart_murmur_detection_b200.frontend import ResamplePlan

    plan = ResamplePlan(sr_in, 16000)
    lens = [1, 17, 1000, 12345, 3 * sr_in + 7]
    clips = [golden_signal(n, seed=5 + i, sr=sr_in, lead=0, tail=0) for i, n in enumerate(lens)]
    wav, off = _batch(clips)
    out, no = plan(wav, off)
    out = out.cpu().numpy()
    tr = torchaudio.transforms.Resample(sr_in, 16000)
    for i, x in enumerate(clips):
        ref = tr(torch.from_numpy(x)).numpy()
        got = out[no[i] : no[i + 1]]
        assert len(got) == len(ref) == int(np.ceil(len(x) * 16000 / sr_in))
        assert np.abs(got - ref).max() <= 1e-6 * max(1.0, np.abs(x).max())
        ref_f = torchaudio.functional.resample(torch.from_numpy(x), sr_in, 16000).numpy()  # float32-built taps
        assert np.abs(got - ref_f).max() <= 2e-5


@pytest.mark.parametrize("preset", ["torchaudio", "soxr_hq_like"])
def test_resample_presets_and_pcm16_input(preset):
    """The named filter designs against torchaudio's own kernel builder with the same parameters, and the fused
    PCM16 path (int16 in, / 32768 and rate conversion in one pass) bit-identical to decode-then-resample."""
    import torchaudio

    from heart_murmur_detection_b200 import frontend as fe

    kw = fe.RESAMPLE_PRESETS[preset]
    for sr_in in (4000, 2000, 8000):
        plan = fe.ResamplePlan(sr_in, 16000, **kw)
        lens = [1, 17, 1000, 12345, 3 * sr_in + 7]
        clips = [golden_signal(n, seed=15 + i, sr=sr_in, lead=0, tail=0) for i, n in enumerate(lens)]
        pcm = [np.clip(np.round(c * 32768.0), -32768, 32767).astype(np.int16) for c in clips]
        deq = [(q.astype(np.float32) / np.float32(32768.0)) for q in pcm]
        wav, off = _batch(deq)
        out, no = plan(wav, off)
        out16, no16 = plan(torch.from_numpy(np.concatenate(pcm)).cuda(), off)
        assert np.array_equal(no, no16) and torch.equal(out, out16)
        out = out.cpu().numpy()
        for i, x in enumerate(deq):
            ref = torchaudio.functional.resample(torch.from_numpy(x).double(), sr_in, 16000, lowpass_filter_width=kw["lowpass_filter_width"],
                                                 rolloff=kw["rolloff"], resampling_method=kw["method"], beta=kw.get("beta")).numpy()
            got = out[no[i] : no[i + 1]]
            assert len(got) == len(ref) == int(np.ceil(len(x) * 16000 / sr_in))
            assert np.abs(got - ref).max() <= 2e-6 * max(1.0, np.abs(x).max()), (preset, sr_in, i)


# ------------------------------------------------------------------------------------------- spectrogram ops


@pytest.mark.parametrize("T,Fq,crop", [(251, 64, 251), (400, 64, 251), (1022, 128, 512), (63, 64, 32), (3750, 64, 251)])
def test_cola_batcher_matches_reference_draws(T, Fq, crop):
    """Crop starts, gains and masked rows are bit-exact with the reference's RNG stream
    (golden hashes from the reference's random_mask/random_crop/random_multiply); masked rows
    hold the float32 mean within 1 ulp-level tolerance (numpy pairwise vs float64 accumulate)."""
    from heart_murmur_detection_b200 import datasets as D
    from oracle import frontend as F

    g = META[f"specops/{T}x{Fq}/{crop}"]
    spec = hash_spec(T, Fq, seed=T)
    store = D.SpecStore([spec])
    random.seed(1000 + T)
    x1, x2 = D.cola_batch(store, [0], max_len=crop, augment=True)
    assert rng_fingerprint() == g["rng_after"]
    random.seed(1000 + T)
    m = F.random_mask(spec)
    c1 = F.random_crop(m, crop_size=crop)
    c2 = F.random_crop(m, crop_size=crop)
    r1, r2 = F.random_multiply(c1), F.random_multiply(c2)
    assert [sha(r1), sha(r2)] == g["mul_sha"]  # the oracle replay reproduces the reference outputs
    for got, ref in ((x1, r1), (x2, r2)):
        got = got[0].cpu().numpy()
        unmasked = ~np.all(ref == ref[:, :1], axis=1)
        np.testing.assert_array_equal(got[unmasked], ref[unmasked])  # bit exact
        assert np.abs(got - ref).max() <= 2e-7
    first = D.pad_or_crop_batch(store, [0], max_len=crop)[0].cpu().numpy()
    assert sha(first[: min(T, crop)]) == g["first_sha"]


def test_pad_to_model_size():
    from heart_murmur_detection_b200 import datasets as D
    from oracle import frontend as F

    specs = [hash_spec(998, 128, 1), hash_spec(1022, 128, 2), hash_spec(1300, 128, 3), hash_spec(5, 128, 4)]
    store = D.SpecStore(specs)
    out = D.pad_or_crop_batch(store, [0, 1, 2, 3], max_len=1024).cpu().numpy()
    for i, s in enumerate(specs):
        np.testing.assert_array_equal(out[i], F.pad_to_model(s, 1024, 128))


def test_htsat_input_stage_matches_reference():
    """hmfe_htsat_input_batch (bn0 + bicubic time resize + fold, htsat.py:889-891, 829-858) against the
    torch-CPU oracle (<= 2e-5) and the fixtures produced by the reference's reshape_wav2img."""
    from cases import HTSAT_T, check_digest, htsat_bn_params
    from heart_murmur_detection_b200 import frontend as fe
    from oracle import frontend as F

    w, b, m, v = htsat_bn_params(64)
    specs = [hash_spec(T, 64, seed=700 + T) for T in HTSAT_T]
    ro = np.zeros(len(specs) + 1, dtype=np.int64)
    np.cumsum([s.shape[0] for s in specs], out=ro[1:])
    dev = torch.from_numpy(np.concatenate(specs)).cuda()
    out = fe.htsat_input(dev, ro[:-1], np.diff(ro), w, b, m, v).cpu().numpy()
    assert out.shape == (len(specs), 1, 256, 256)
    for i, (T, s) in enumerate(zip(HTSAT_T, specs)):
        ref = F.htsat_input(s, w, b, m, v)
        assert np.abs(out[i, 0] - ref).max() <= 2e-5, T
        check_digest(out[i, 0], META[f"htsat_input/{T}"], rtol=2e-5, atol=2e-5)
    # a crop of a longer spectrogram (row range inside the batch) and the full-length identity path
    crop = fe.htsat_input(dev, [int(ro[3]) + 17], [900], w, b, m, v).cpu().numpy()[0, 0]
    assert np.abs(crop - F.htsat_input(specs[3][17:917], w, b, m, v)).max() <= 2e-5
    with pytest.raises(Exception):
        fe.htsat_input(dev, [0], [1025], w, b, m, v)


@pytest.mark.parametrize("order", [5, 2])
def test_sosfiltfilt_matches_scipy(order):
    """Zero-phase mode (named in BASELINE.json's north_star; not what the reference calls):
    <= 1e-4 * max|y| against scipy.signal.sosfiltfilt, in practice ~1e-7 (float32 intermediate)."""
    from scipy.signal import butter, sosfiltfilt

    from heart_murmur_detection_b200 import _lib
    from heart_murmur_detection_b200 import frontend as fe

    sos_scipy = butter(order, [200 / 8000, 1800 / 8000], btype="band", output="sos")
    edge = 3 * (2 * len(sos_scipy) + 1)
    lens = [edge + 1, edge + 2, 100, 1000, 4097, 128000, 90001, 300000]
    clips = [golden_signal(n, seed=9 + i) for i, n in enumerate(lens)]
    clips[3] = np.ones(1000, np.float32) * 0.25  # constant: the steady-state initial conditions matter
    wav, off = _batch(clips)
    for sos in (sos_scipy, fe.butter_bandpass_sos(200, 1800, SR, order)):
        y = fe.sosfiltfilt(wav, off, sos).cpu().numpy()
        assert y.dtype == np.float64
        y32 = fe.sosfiltfilt(wav, off, sos, out_dtype=torch.float32).cpu().numpy()
        for i, x in enumerate(clips):
            ref = sosfiltfilt(sos_scipy, x)
            scale = max(np.abs(ref).max(), np.abs(x).max() * 1e-3, 1e-12)
            assert np.abs(y[off[i] : off[i + 1]] - ref).max() <= 2e-6 * scale, (i, lens[i])
            assert np.abs(y32[off[i] : off[i + 1]] - ref).max() <= 1e-4 * scale, (i, lens[i])
    short, soff = _batch([golden_signal(edge, 1)])
    with pytest.raises(_lib.HmfeError):
        fe.sosfiltfilt(short, soff, sos_scipy)  # scipy: "The length of the input vector x must be greater than padlen"


def test_vggish_examples_match_reference():
    """vggish_input.waveform_to_examples (sibling front-end, vggish_input.py:52-125) on the fbank kernel with
    VGGish constants: shapes / dtype exact, values within 1e-4 of the float64 reference (fixtures produced by
    executing the reference's vggish_input + mel_features).  The log offset 0.01 turns the float32 FFT's absolute
    error (~5e-7 for unit-scale frames) into 5e-5 in the log domain near the floor log(0.01)."""
    from cases import RECORDINGS, check_digest
    from heart_murmur_detection_b200 import vggish_input as V

    arr = {k.replace("|", "/"): v for k, v in np.load(os.path.join(HERE, "golden", "ref_util.npz")).items()}
    rec = {name: golden_signal(n, seed, SR, lead, tail) for name, n, seed, lead, tail in RECORDINGS}
    for name in ("r_short", "r_mid", "r_long"):
        g = META[f"vggish/{name}"]
        ex = V.waveform_to_examples(rec[name], 16000)
        assert list(ex.shape) == g["shape"] and str(ex.dtype) == g["dtype"]
        check_digest(ex, g["digest"], rtol=1e-4, atol=1e-4)
    assert np.abs(V.waveform_to_examples(rec["r_short"], 16000) - arr["vggish/r_short"]).max() <= 1e-4
    assert V.waveform_to_examples(rec["r_short"][:15000], 16000).shape == (0, 96, 64)  # shorter than one example


# ---------------------------------------------------------------------------------------------------------------
# Dataset __getitem__ over batches (a18): fixtures from the executed reference classes (make_golden_datasets.py)
# ---------------------------------------------------------------------------------------------------------------


def _dataset_fixture():
    import json as _json

    from cases import DATASET_SPECS

    g = _json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_datasets.json")))["cases"]
    specs64 = [hash_spec(T, 64, seed=300 + T) for T in DATASET_SPECS["rows64"]]
    specs128 = [hash_spec(T, 128, seed=500 + T) for T in DATASET_SPECS["rows128"]]
    return g, specs64, specs128, DATASET_SPECS["order"]


def _close_masked(got, ref):
    """unmasked rows bit exact; rows replaced by the mean within one float32 rounding of the mean"""
    same = (got == ref).all(axis=1)
    assert np.abs(got - ref).max() <= 2e-7
    assert same.sum() >= 0.7 * len(ref)


@pytest.mark.gpu
def test_cola_batch_with_windowing_matches_executed_reference():
    """datasets.cola_batch (native draw planner + crop kernel) against mae_training.py's AudioDataset, method 'cola',
    with and without the 3 * max_len window and augmentation: same Python RNG stream, same crops, same gains."""
    from heart_murmur_detection_b200 import datasets as D
    from oracle import frontend as F

    g, specs64, _, order = _dataset_fixture()
    store = D.SpecStore(specs64)
    for windowing in (False, True):
        for augment in (False, True):
            fx = g[f"mae_training/cola/w{int(windowing)}a{int(augment)}"]
            random.seed(4242)
            x1, x2 = D.cola_batch(store, order, max_len=251, augment=augment, windowing=windowing)
            assert rng_fingerprint() == fx["rng_after"]
            random.seed(4242)
            refs = [F.dataset_cola_item(specs64[i], 251, augment, windowing) for i in order]
            assert [sha(a) for a, _ in refs] == fx["x1"]
            for k, (r1, r2) in enumerate(refs):
                for got, ref in ((x1[k], r1), (x2[k], r2)):
                    got = got.cpu().numpy()
                    if augment:
                        _close_masked(got, ref)
                    else:
                        np.testing.assert_array_equal(got, ref)


@pytest.mark.gpu
def test_mae_batch_matches_executed_reference():
    """datasets.mae_batch against AudioDataset methods 'mae' (max_len 256, 64 mel) and 'audiomae' (1024, 128 mel):
    random_crop draws only for the items that are too long, zero padding for the short ones; bit exact."""
    from heart_murmur_detection_b200 import datasets as D

    g, specs64, specs128, order = _dataset_fixture()
    random.seed(77)
    out = D.mae_batch(D.SpecStore(specs64), order, max_len=256).cpu().numpy()
    assert [sha(np.ascontiguousarray(x)) for x in out] == g["mae_training/mae/256"]["x"]
    assert rng_fingerprint() == g["mae_training/mae/256"]["rng_after"]
    random.seed(78)
    out = D.mae_batch(D.SpecStore(specs128), list(range(len(specs128))), max_len=1024).cpu().numpy()
    assert [sha(np.ascontiguousarray(x)) for x in out] == g["mae_training/audiomae/1024"]["x"]
    assert rng_fingerprint() == g["mae_training/audiomae/1024"]["rng_after"]


@pytest.mark.gpu
def test_finetune_batch_matches_executed_reference():
    """datasets.finetune_batch against finetuning.py's AudioDataset: crop -> random_mask (mean of the CROPPED item) ->
    random_multiply -> SpecAugmentation stripes; Python and torch generators end in the reference's states."""
    from heart_murmur_detection_b200 import datasets as D
    from oracle import frontend as F

    g, specs64, _, order = _dataset_fixture()
    store = D.SpecStore(specs64)
    for name in ("first", "random_aug", "specaug", "specaug_only"):
        fx = g[f"finetuning/{name}"]
        random.seed(990)
        torch.manual_seed(991)
        out = D.finetune_batch(store, order, **fx["kw"]).cpu().numpy()
        assert rng_fingerprint() == fx["rng_after"], name
        assert sha(torch.get_rng_state().numpy()) == fx["torch_rng_after"], name
        random.seed(990)
        torch.manual_seed(991)
        for k, i in enumerate(order):
            ref = F.dataset_finetune_item(specs64[i], **fx["kw"])
            assert list(out[k].shape) == fx["shape"][k]
            assert [int(r) for r in np.flatnonzero((out[k] == 0).all(axis=1))] == fx["zero_rows"][k]
            assert [int(c) for c in np.flatnonzero((out[k] == 0).all(axis=0))] == fx["zero_cols"][k]
            assert np.abs(out[k] - ref).max() <= 2e-7, name
            assert abs(float(np.asarray(out[k], np.float64).sum()) - fx["sum"][k]) <= 1e-6 * max(1.0, abs(fx["sum"][k]))


@pytest.mark.gpu
def test_trim_near_the_threshold_stress():
    """Knife edge of librosa.effects.trim (src/util.py:170-172): 100 000 frames whose level lies within +-1e-3 dB of
    the -60 dB threshold (20 000 clips x 5 frames).  The compare `db > -60` is a float32 compare on a float32 sum
    whose order differs between numpy (pairwise) and the GPU (hop energies), so a frame that sits within rounding of
    the threshold may flip.  Counted here, and bounded: at most 0.2 % of the clips, every flip by exactly one frame."""
    from heart_murmur_detection_b200 import frontend as fe
    from oracle import librosa_restated as lr

    rng = np.random.default_rng(7)
    n_clips, hop = 20000, 800
    blocks = 2 + 6  # two loud hop blocks (the reference level), six quiet ones -> five quiet-quiet frames
    amp = np.ones((n_clips, blocks), dtype=np.float64)
    amp[:, 2:] = 1e-3 * 10.0 ** (rng.uniform(-1e-3, 1e-3, size=(n_clips, blocks - 2)) / 20.0)
    sign = np.where(np.arange(hop) % 2 == 0, 1.0, -1.0)
    x = (amp[:, :, None] * sign[None, None, :]).reshape(n_clips, blocks * hop).astype(np.float32)
    off = np.arange(n_clips + 1, dtype=np.int64) * (blocks * hop)
    se = fe.trim_indices(torch.from_numpy(x.reshape(-1)).cuda(), off, frame_length=1600, hop_length=hop).cpu().numpy()
    ref = np.stack([lr.trim(x[i], top_db=60, frame_length=1600, hop_length=hop)[1] for i in range(n_clips)])
    # the frames really are at the knife edge: the oracle's own decisions split roughly evenly
    ends = ref[:, 1] // hop
    assert len(np.unique(ends)) >= 5
    assert np.array_equal(se[:, 0], ref[:, 0])
    diff = np.flatnonzero(se[:, 1] != ref[:, 1])
    print(f"trim knife edge: {diff.size} of {n_clips} clips differ from the numpy restatement")
    assert diff.size <= n_clips // 500
    assert np.all(np.abs(se[diff, 1] - ref[diff, 1]) <= hop * blocks)


@pytest.mark.gpu
def test_caller_provided_workspace_contract():
    """hmfe_ctx_set_workspace / hmfe_ctx_reserve (never allocate, caller workspace): with a workspace of the
    queried size the band-pass + trim and the stand-alone trim give the results of the allocating context; a
    workspace that is too small is refused with the size that is needed, and so is a batch beyond the reservation."""
    from heart_murmur_detection_b200 import frontend as fe
    from heart_murmur_detection_b200._lib import HmfeError

    clips = [golden_signal(n, seed=60 + i) for i, n in enumerate((90000, 20000, 128000, 300000, 513, 47999))]
    wav, off = _batch(clips)
    sos = fe.butter_bandpass_sos(200, 1800, 16000, order=5)
    y0, se0 = fe.iir_sos_trim(wav, off, sos)
    t0 = fe.trim_indices(wav, off)
    ctx = fe.Context()
    need = max(fe.iir_workspace_bytes(off, 5, 800), fe.trim_workspace_bytes(off, 1600, 800))
    ctx.use_workspace(torch.empty(need, dtype=torch.uint8, device="cuda"), max_clips=len(clips))
    for algo in ("overlap", "scan"):
        ctx.set_iir_algo(algo)
        y1, se1 = fe.iir_sos_trim(wav, off, sos, ctx=ctx)
        assert torch.equal(se1, se0)
        assert (y1 - y0).abs().max().item() <= 1e-6
    assert torch.equal(fe.trim_indices(wav, off, ctx=ctx), t0)
    small = fe.Context()
    small.use_workspace(torch.empty(256, dtype=torch.uint8, device="cuda"))
    with pytest.raises(HmfeError, match="too small"):
        fe.iir_sos_trim(wav, off, sos, ctx=small)
    # the reservation for 2 clips carries a few KB of slack for the fixed-size records: a 700-clip batch is beyond it
    many = [golden_signal(1000, seed=70 + i % 5) for i in range(700)]
    wav_m, off_m = _batch(many)
    tight = fe.Context()
    tight.use_workspace(torch.empty(fe.trim_workspace_bytes(off_m, 1600, 800), dtype=torch.uint8, device="cuda"), max_clips=2)
    with pytest.raises(HmfeError, match="reserved"):
        fe.trim_indices(wav_m, off_m, ctx=tight)
