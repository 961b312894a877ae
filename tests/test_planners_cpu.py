"""Host index planners (frontend.plan_*, pipeline chunkers) against the oracle restatement of the
reference's pad / split functions, over randomly drawn lengths (hypothesis).  No GPU: the gather
records are applied with a numpy emulation of the semantics documented in include/hmfe.h
(hmfe_gather_desc), which is exactly what gather_kernel implements."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from heart_murmur_detection_b200 import frontend as fe
from heart_murmur_detection_b200 import pipeline as pl
from oracle import frontend as F

SR = 16000


def apply_chunk(clip, ch):
    """numpy emulation of one chunk plan: view or gather record (include/hmfe.h)."""
    if ch[0] == "view":
        return clip[ch[1] : ch[1] + ch[2]]
    _, length, src_start, period, a_end, a_phase, b_end, b_start = ch
    src = clip[src_start:]
    i = np.arange(length)
    out = np.zeros(length, dtype=np.float32)
    a = i < a_end
    out[a] = src[(a_phase + i[a]) % period]
    bsel = (~a) & (i < b_end)
    out[bsel] = src[b_start + (i[bsel] - a_end)]
    return out


def signal(n, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(n).astype(np.float32)


lengths = st.one_of(st.integers(1, 400_000), st.sampled_from([15999, 16000, 16001, 31999, 32000, 32001, 65439, 65440, 65441,
                                                              127999, 128000, 128001, 130880, 196320, 261760]))
secs = st.sampled_from([1, 2, 4.09, 8, 8.18])


@settings(max_examples=120, deadline=None, derandomize=True)
@given(n=lengths, sec=secs, types=st.sampled_from(["repeat", "zero"]))
def test_plan_split_pad_equals_reference_layout(n, sec, types):
    """split_pad_sample / _duplicate_padding / _equally_slice_pad_sample / _zero_padding (src/util.py:504-620)."""
    x = signal(n, n)
    ref = F.split_pad_sample(x, sec, SR, types)
    plan = fe.plan_split_pad(n, sec, SR, types)
    assert len(plan) == len(ref)
    for ch, r in zip(plan, ref):
        got = apply_chunk(x, ch)
        assert got.shape == r.shape and np.array_equal(got, r)


@settings(max_examples=60, deadline=None, derandomize=True)
@given(n=lengths, sec=st.sampled_from([2, 10]))
def test_plan_split_sample_equals_reference(n, sec):
    """split_sample (extract_feature.py:250-259)."""
    x = signal(n, n + 1)
    ref = F.split_sample(x, sec, SR)
    plan = fe.plan_split_sample(n, sec, SR)
    assert len(plan) == len(ref)
    for ch, r in zip(plan, ref):
        assert np.array_equal(apply_chunk(x, ch), r)


@settings(max_examples=120, deadline=None, derandomize=True)
@given(n=st.integers(0, 700_000), pad=st.booleans(), types=st.sampled_from(["repeat", "zero"]),
       max_sec=st.sampled_from([None, 32, 10]))
def test_entire_signal_chunker_equals_reference_control_flow(n, pad, types, max_sec):
    """Control flow of get_entire_signal_librosa after the trim (src/util.py:248-259): too short ->
    None / padded; longer than max_sec -> cut."""
    x = signal(n, 7)
    chunker = pl.entire_signal_chunker(8, SR, pad, types, max_sec)
    chunker.dup_called = False
    chunks = chunker(n)
    # the oracle's entire_signal trims first: feed it through the same steps without the trim
    duration = n / SR
    if duration < 8:
        ref = None if (not pad or n == 0) else F.split_pad_sample(x, 8, SR, types)[0]
    else:
        ref = x
    if ref is not None and max_sec and duration > max_sec:
        ref = ref[: int(max_sec * SR)]
    if ref is None:
        assert chunks is None
    else:
        assert len(chunks) == 1
        assert np.array_equal(apply_chunk(x, chunks[0]), ref)


@settings(max_examples=80, deadline=None, derandomize=True)
@given(n=st.integers(1, 600_000), sec=st.sampled_from([2, 4.09, 8.18, 10]), trim_tail=st.booleans())
def test_split_signal_chunker_drop_last(n, sec, trim_tail):
    """get_split_signal_librosa's chunk list incl. decide_droplast (src/util.py:348-354, 369-371)."""
    x = signal(n, 3)
    chunker = pl.split_signal_chunker(sec, SR, trim_tail)
    chunker.dup_called = False
    chunks = chunker(n)
    ref = F.split_pad_sample(x, sec, SR)
    if trim_tail and F.decide_droplast(x, SR, sec):
        ref.pop()
    assert len(chunks) == len(ref)
    for ch, r in zip(chunks, ref):
        assert np.array_equal(apply_chunk(x, ch), r)
