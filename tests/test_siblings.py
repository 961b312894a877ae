"""Sibling front-ends of the thesis baselines: CLAP (torchlibrosa log-mel at 44.1 kHz) and
HeAR (mel-PCEN).  CPU tests pin the oracles; ``-m gpu`` tests compare the CUDA path with them."""
import os
import random

import numpy as np
import pytest
import torch

from signals import golden_signal

from oracle import frontend as F

HERE = os.path.dirname(os.path.abspath(__file__))

CLAP_SR = 44100


def _clap_clip(n, seed):
    return golden_signal(n, seed=seed, sr=CLAP_SR)


# --------------------------------------------------------------------------------------------- CLAP, CPU


def test_clap_oracle_vs_torchaudio():
    """torchlibrosa is absent (parity unpinned against it); torchaudio's MelSpectrogram with the same constants is an
    independent float32 implementation of the same definition."""
    import torchaudio

    tf = torchaudio.transforms.MelSpectrogram(CLAP_SR, n_fft=1024, hop_length=320, f_min=50.0, f_max=14000.0, n_mels=64,
                                              norm="slaney", mel_scale="slaney", center=True, pad_mode="reflect", power=2.0)
    for n, seed in ((5 * CLAP_SR, 11), (30000, 12)):
        x = _clap_clip(n, seed)
        db, S = F.clap_logmel(x, return_power=True)
        St = tf(torch.from_numpy(x)).numpy().T
        assert S.shape == St.shape == (1 + n // 320, 64)
        assert np.abs(S - St).max() <= 1e-4 * np.abs(S).max()
        dbt = 10 * np.log10(np.maximum(1e-10, St))
        loud = S >= 1e-6 * S.max()  # below that the float32 STFT of torchaudio is rounding noise
        assert np.abs(db - dbt)[loud].max() <= 1e-2


def test_clap_fixed_duration_plan_matches_oracle():
    from heart_murmur_detection_b200 import clap_input as ci

    def apply(ch, x):  # what the gather kernel does with a plan record (include/hmfe.h: hmfe_gather_desc)
        if ch[0] == "view":
            return x[ch[1] : ch[1] + ch[2]]
        _, length, src_start, period, a_end, a_phase, b_end, b_start = ch
        i = np.arange(length)
        out = np.where(i < a_end, x[src_start + (a_phase + i) % period], 0.0)
        return out.astype(np.float32)

    for n in (1, 7, 1000, 110249, 110250, 220499, 220500, 220501, 300000, 1000000):
        x = (np.arange(n) % 977).astype(np.float32)
        random.seed(n)
        want = F.clap_fixed_duration(x)
        state_after = random.getstate()
        random.seed(n)
        got = apply(ci.plan_fixed_duration(n, 5, CLAP_SR), x)
        assert random.getstate() == state_after  # same number of draws
        np.testing.assert_array_equal(got, want)
    with pytest.raises(ValueError):
        ci.plan_fixed_duration(0, 5, CLAP_SR)


# --------------------------------------------------------------------------------------------- CLAP, GPU


@pytest.mark.gpu
def test_clap_logmel_matches_oracle():
    """dB <= 1e-2 wherever the band is within 60 dB of the clip maximum (BASELINE tolerance), linear mel power
    <= 1e-4 * max; quieter bands are float32-FFT rounding noise in any float32 implementation (torchlibrosa's
    conv-STFT included) and are bounded loosely."""
    from heart_murmur_detection_b200 import clap_input as ci
    from heart_murmur_detection_b200 import frontend as fe

    clips = np.stack([_clap_clip(5 * CLAP_SR, 21 + i) for i in range(3)])
    clips[2, :2000] = 0.0  # digital silence: bands at the 1e-10 floor
    got = ci.logmel_batch(torch.from_numpy(clips).cuda())
    assert got.shape == (3, 1, 690, 64)
    got = got.cpu().numpy()
    plan = fe.logmel_plan(CLAP_SR, 64, 50, 14000, 1024, 320, pad_mode="reflect")
    off = np.arange(4, dtype=np.int64) * clips.shape[1]
    power, _ = plan(torch.from_numpy(clips.reshape(-1)).cuda(), off, mode="power")
    power = power.cpu().numpy().reshape(3, 690, 64)
    for i in range(3):
        db, S = F.clap_logmel(clips[i], return_power=True)
        assert np.abs(power[i] - S).max() <= 1e-4 * S.max()
        loud = S >= 1e-6 * S.max()
        assert np.abs(got[i, 0] - db)[loud].max() <= 1e-2
        assert np.abs(got[i, 0] - db).max() <= 1.0
        assert got[i, 0].min() >= -100.0
    # ragged batch with clips barely longer than the reflect minimum, and the error below it
    lens = [513, 514, 1023, 1024, 1025, 5000, 44100]
    xs = [_clap_clip(n, 40 + n % 7) for n in lens]
    off = np.zeros(len(xs) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    out, fo = plan(torch.from_numpy(np.concatenate(xs)).cuda(), off, mode="db_abs")
    out = out.cpu().numpy()
    for i, x in enumerate(xs):
        db, S = F.clap_logmel(x, return_power=True)
        assert fo[i + 1] - fo[i] == 1 + len(x) // 320 == db.shape[0]
        loud = S >= 1e-6 * S.max()
        assert np.abs(out[fo[i] : fo[i + 1]] - db)[loud].max() <= 1e-2
    with pytest.raises(Exception, match="reflect"):
        plan(torch.zeros(512, device="cuda"), np.array([0, 512]), mode="db_abs")


@pytest.mark.gpu
def test_clap_load_audio_batch_matches_oracle():
    """Resample (torchaudio algorithm) + repeat / crop with the reference's RNG draws."""
    import torchaudio

    from heart_murmur_detection_b200 import clap_input as ci

    sr = 16000
    lens = [20000, 80000, 81000, 200000]  # 80 000 -> exactly 220 500 after resampling; longer ones are cropped
    xs = [golden_signal(n, seed=60 + i, sr=sr) for i, n in enumerate(lens)]
    off = np.zeros(len(xs) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    random.seed(99)
    want = []
    for x in xs:
        y = torchaudio.transforms.Resample(sr, CLAP_SR)(torch.from_numpy(x)).numpy()  # CLAPWrapper.py:270-271
        want.append(F.clap_fixed_duration(y))
    state_after = random.getstate()
    random.seed(99)
    got = ci.load_audio_batch(torch.from_numpy(np.concatenate(xs)).cuda(), off, sr).cpu().numpy()
    assert random.getstate() == state_after
    assert got.shape == (4, 5 * CLAP_SR)
    for g, w in zip(got, want):
        assert np.abs(g - w).max() <= 1e-5  # resampler tolerance of test_resample_matches_torchaudio
    # preprocess_audio: list of decoded clips -> [n, 1, L]
    random.seed(99)
    pa = ci.preprocess_audio(xs, [sr] * len(xs))
    assert pa.shape == (4, 1, 5 * CLAP_SR)
    np.testing.assert_array_equal(pa[:, 0].cpu().numpy(), got)


# --------------------------------------------------------------------------------------------- HeAR, CPU

HEAR = {k.replace("|", "/"): v for k, v in np.load(os.path.join(HERE, "golden", "ref_hear.npz")).items()}
HEAR_PCEN_TOL = 1e-4  # absolute, on PCEN values of 0..6 (quiet bands amplify the float32 FFT noise: x / ema^0.8)


def _hear_batch(name):
    from cases import HEAR_CASES

    n, seeds = HEAR_CASES[name]
    return np.stack([golden_signal(n, seed=s) for s in seeds])


def test_hear_oracle_matches_the_reference_fixtures():
    """oracle.hear_* restates audio_utils.py with the torch ops the reference calls; the fixtures were produced by
    executing the reference's own preprocess_audio / _linear_to_mel_weight_matrix / _pcen_function.  On the machine
    that generated them the agreement is bit-exact; the bounds leave room for another CPU's FFT / pow kernels."""
    from cases import HEAR_CASES

    np.testing.assert_allclose(F.hear_mel_matrix(), HEAR["mel_matrix"], rtol=0, atol=1e-6)
    for name in HEAR_CASES:
        np.testing.assert_allclose(F.hear_preprocess_audio(_hear_batch(name)), HEAR[f"out/{name}"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(F.hear_mel_power(_hear_batch("b3")), HEAR["mel/b3"], rtol=0, atol=1e-5 * HEAR["mel/b3"].max())
    np.testing.assert_allclose(F.hear_pcen(HEAR["mel/b3"]), HEAR["pcen/b3"], rtol=0, atol=1e-5)
    with pytest.raises(ValueError):
        F.hear_preprocess_audio(np.zeros((1, 32001), np.float32))


def test_hear_host_mirror_tables_and_errors():
    from heart_murmur_detection_b200 import hear_input as hi

    np.testing.assert_allclose(hi.linear_to_mel_weight_matrix().numpy(), HEAR["mel_matrix"], rtol=0, atol=1e-6)
    with pytest.raises(ValueError, match="Nyquist"):
        hi.linear_to_mel_weight_matrix(upper_edge_hertz=9000.0)
    with pytest.raises(ValueError, match="rank 2"):
        hi.preprocess_audio(torch.zeros(32000))
    with pytest.raises(ValueError, match="32000 samples"):
        hi.preprocess_audio(torch.zeros(1, 32001))


def test_hear_kernel_emulation_matches_reference(tmp_path):
    """csrc/host_check.cu runs the kernel's lane-level code (25 x 16 FFT-400, separation, banded mel, PCEN + resize) on
    the CPU; compared with the outputs of the reference itself."""
    import subprocess

    from cases import HEAR_CASES
    from heart_murmur_detection_b200 import build

    hc = build.build_host_check()
    torch.hann_window(400).numpy().tofile(tmp_path / "w.f32")
    HEAR["mel_matrix"].astype(np.float32).tofile(tmp_path / "m.f32")
    for name, (n, seeds) in HEAR_CASES.items():
        x = _hear_batch(name)
        x.tofile(tmp_path / "a.f32")
        subprocess.check_call([hc, "hear", str(n), "32000", "192"] + [str(tmp_path / f) for f in
                                                                      ("a.f32", "w.f32", "m.f32", "om.f32", "op.f32")])
        mel = np.fromfile(tmp_path / "om.f32", dtype=np.float32).reshape(len(seeds), 200, 128)
        img = np.fromfile(tmp_path / "op.f32", dtype=np.float32).reshape(len(seeds), 1, 192, 128)
        assert np.abs(img - HEAR[f"out/{name}"]).max() <= HEAR_PCEN_TOL
        ref_mel = F.hear_mel_power(np.pad(x, ((0, 0), (0, 32000 - n))))
        assert np.abs(mel - ref_mel).max() <= 1e-4 * ref_mel.max()  # BASELINE: <= 1e-4 relative on linear spectra


# --------------------------------------------------------------------------------------------- HeAR, GPU


@pytest.mark.gpu
def test_hear_preprocess_audio_matches_reference():
    from cases import HEAR_CASES
    from heart_murmur_detection_b200 import hear_input as hi

    plan = hi.hear_plan()
    for name, (n, seeds) in HEAR_CASES.items():
        x = _hear_batch(name)
        got = hi.preprocess_audio(torch.from_numpy(x))  # host tensor in -> host tensor out, like the reference
        assert got.shape == (len(seeds), 1, 192, 128) and got.dtype == torch.float32 and not got.is_cuda
        assert np.abs(got.numpy() - HEAR[f"out/{name}"]).max() <= HEAR_PCEN_TOL
        assert plan.last_launches == 4
        mel = plan.mel_power(torch.from_numpy(x).cuda()).cpu().numpy()
        ref_mel = F.hear_mel_power(np.pad(x, ((0, 0), (0, 32000 - n))))
        assert mel.shape == ref_mel.shape == (len(seeds), 200, 128)
        assert np.abs(mel - ref_mel).max() <= 1e-4 * ref_mel.max()
    # CUDA tensor in -> CUDA tensor out
    assert hi.preprocess_audio(torch.from_numpy(_hear_batch("single")).cuda()).is_cuda
    # a larger batch against the oracle (different clips per warp / CTA, grid-stride path), odd clip length
    x = np.stack([golden_signal(31999, seed=300 + i) for i in range(70)])
    got = hi.preprocess_audio(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.abs(got - F.hear_preprocess_audio(x)).max() <= HEAR_PCEN_TOL
    # empty batch
    assert hi.hear_plan()(torch.zeros((0, 32000), device="cuda")).shape == (0, 192, 128)
