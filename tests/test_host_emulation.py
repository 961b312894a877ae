"""CPU emulation of the warp algorithm (csrc/host_check.cu: same __host__ __device__ templates
as the kernels, lanes looped on the host) against the oracle.  Validates the FFT index math,
twiddle planes, partner-lane separation, banded mel and float32 accuracy without a GPU."""
import subprocess

import numpy as np
import pytest

from signals import golden_signal


@pytest.fixture(scope="module")
def host_check():
    from heart_murmur_detection_b200 import build

    return build.build_host_check()


@pytest.mark.parametrize("variant", ["scalar", "packed", "pair"])
def test_logmel_power_emulation(host_check, variant, tmp_path):
    from oracle import frontend as F

    for n, seed, fmax in [(128000, 5, 8000), (20000, 3, 8000), (1, 1, 8000), (513, 2, 2000), (40001, 4, 2000)]:
        x = golden_signal(n, seed)
        fin, fout = tmp_path / "in.f32", tmp_path / "out.f32"
        x.tofile(fin)
        subprocess.check_call([host_check, "logmel", variant, "512", "64", "50", str(fmax), str(fin), str(fout)])
        P = np.fromfile(fout, dtype=np.float32).reshape(-1, 64)
        _, db, S = F.log_mel(x, f_max=fmax, return_parts=True)
        assert P.shape == S.shape
        assert np.abs(P - S).max() <= 1e-4 * np.abs(S).max()
        dbp = 10 * np.log10(np.maximum(1e-10, P)) - 10 * np.log10(max(1e-10, P.max()))
        dbp = np.maximum(dbp, dbp.max() - 80)
        assert np.abs(dbp - db).max() <= 1e-2


def test_banded_mel_lane_assignment(host_check):
    """tables.h build_banded for the Kaldi bank (kaldi_fbank.cu): aligning the windows modulo 8 cuts the loop trips from 72
    to 40, and the minimum-cost lane re-assignment (prefer = 16) makes those 40 trips bank-conflict free again; every
    placement must reproduce the dense mel matrix exactly (verify_banded), which is what plan creation enforces."""
    out = subprocess.check_output([host_check, "banded"], text=True).strip().splitlines()
    rows = {}
    for line in out[:3]:
        f = line.split()
        rows[(int(f[1]), int(f[3]))] = dict(trip=int(f[5]), wf=int(f[7]), ideal=int(f[9]), verify=int(f[11]))
    assert all(r["verify"] == 1 for r in rows.values())
    assert rows[(16, 0)]["trip"] == 72 and rows[(16, 0)]["wf"] == rows[(16, 0)]["ideal"]
    assert rows[(8, 0)]["trip"] == 40 and rows[(8, 0)]["wf"] > rows[(8, 0)]["ideal"]
    assert rows[(8, 16)]["trip"] == 40 and rows[(8, 16)]["wf"] == rows[(8, 16)]["ideal"] == 80
    assert out[3].split() == ["assign", "total", "0", "distinct", "32"]


def _bf16_rn(x):
    """float32 -> bfloat16 (round to nearest even) -> float32, like cvt.rn.bf16x2.f32."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def test_tensor_core_split_precision():
    """The arithmetic of HMFE_VARIANT_TC without a GPU (csrc/logmel_tc.cu): weights and powers as bf16 (hi, lo) pairs,
    mel = W_hi P_hi + W_hi P_lo + W_lo P_hi + W_lo P_lo accumulated in float32, against the float64 product of the float32
    operands.  This is the bound the GPU test relies on (3e-5 relative per mel bin, 1e-4 of the clip maximum overall) on
    powers that span the 80 dB the dB stage keeps."""
    from oracle import frontend as F
    from oracle import librosa_restated as lr

    W = lr.mel_filterbank(16000, 1024, n_mels=64, fmin=50, fmax=8000).astype(np.float32)[:, :512]
    w_hi = _bf16_rn(W)
    w_lo = _bf16_rn(W - w_hi)
    rng = np.random.default_rng(3)
    for x in (golden_signal(128000, 5), golden_signal(40001, 4)):
        _, _, S = F.log_mel(x, f_max=8000, return_parts=True)  # only for the scale of real mel powers
        P = (rng.random((64, 512)) ** 6 * float(S.max()) * 4).astype(np.float32)  # 64 frames of bin powers, ~100 dB of range
        p_hi = _bf16_rn(P)
        p_lo = _bf16_rn(P - p_hi)
        acc = np.zeros((64, 64), dtype=np.float32)
        for k0 in range(0, 512, 16):  # K = 16 per MMA, float32 accumulation across the 32 instructions
            sl = slice(k0, k0 + 16)
            part = (p_hi[:, sl].astype(np.float64) @ w_hi[:, sl].T.astype(np.float64) + p_lo[:, sl].astype(np.float64) @ w_hi[:, sl].T.astype(np.float64)
                    + p_hi[:, sl].astype(np.float64) @ w_lo[:, sl].T.astype(np.float64) + p_lo[:, sl].astype(np.float64) @ w_lo[:, sl].T.astype(np.float64))
            acc = (acc.astype(np.float64) + part).astype(np.float32)
        exact = P.astype(np.float64) @ W.T.astype(np.float64)
        rel = np.abs(acc - exact) / np.maximum(exact, 1e-300)
        assert rel[exact > 0].max() <= 3e-5, rel[exact > 0].max()
        assert np.abs(acc - exact).max() <= 1e-5 * exact.max()
