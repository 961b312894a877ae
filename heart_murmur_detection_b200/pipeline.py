"""Batched front-end pipelines: the reference's per-file functions, run on ragged batches.

Each function takes the decoded 16 kHz batch (``wav``: float32 CUDA tensor with all clips back
to back, ``offsets``: int64 host array) and runs
    [band-pass] -> silence trim -> pad / split / cut (host index planning) -> log-mel | fbank
on the GPU.  The only host round trip is the (start, end) trim indices (16 bytes per clip),
which the reference's own control flow needs (durations decide padding and skipping).

Reference functions (``/root/reference``):
    get_entire_signal_librosa      src/util.py:205-267
    get_split_signal_librosa       src/util.py:309-364
    get_split_signal_fbank_pad     src/util.py:794-860
    get_split_signal_fbank         src/benchmark/baseline/extract_feature.py:213-247
    get_individual_segments_librosa src/util.py:141-202
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import frontend as fe


@dataclass
class ChunkBatch:
    """Chunks (views into ``work``) produced from a batch of recordings."""

    work: torch.Tensor
    starts: np.ndarray      # [n_chunks] element offsets into work
    lengths: np.ndarray     # [n_chunks]
    clip_ids: np.ndarray    # [n_chunks] source recording of each chunk
    n_clips: int
    valid: np.ndarray       # [n_clips] False where the reference returns None / [] ("too short")
    trim: np.ndarray        # [n_clips, 2] trim indices (start, end), clip relative
    used_duplicate_padding: bool = False
    launches: int = 0
    is_view: np.ndarray | None = None   # [n_chunks] True for straight views (no padding applied)
    filtered: bool = False              # band-pass applied (the reference's data is float64 then)
    alt: torch.Tensor | None = None     # padded copies kept apart from a read-only signal: start s < 0 -> alt[-s-1:]
    stored: np.ndarray | None = None    # [n_chunks] samples of the chunk that exist in memory; the rest of ``lengths`` is zeros
                                        # that were never written (zero padding as a view, entire_signal_device)

    def samples(self, k: int) -> torch.Tensor:
        """Device view of chunk k (a zero-padded copy when its trailing zeros were never materialised)."""
        s, n = int(self.starts[k]), int(self.lengths[k])
        m = n if self.stored is None else int(self.stored[k])
        x = self.work[s : s + m] if s >= 0 else self.alt[-s - 1 : -s - 1 + m]
        return x if m == n else torch.cat([x, x.new_zeros(n - m)])

    def ref_dtype(self, k: int):
        """dtype the reference produces for chunk k: float64 for untouched views of band-passed
        audio (lfilter output), float32 for padded copies (np.zeros(..., float32)) and unfiltered audio."""
        return np.float64 if self.filtered and (self.is_view is None or bool(self.is_view[k])) else np.float32

    def chunks_of(self, clip: int) -> np.ndarray:
        return np.flatnonzero(self.clip_ids == clip)


@dataclass
class FeatureBatch:
    features: torch.Tensor      # [sum rows, n_mels]
    row_offsets: np.ndarray     # [n_chunks + 1]
    chunks: ChunkBatch
    launches: int = 0

    def chunk(self, k: int) -> torch.Tensor:
        return self.features[int(self.row_offsets[k]) : int(self.row_offsets[k + 1])]

    def of_clip(self, clip: int):
        return [self.chunk(int(k)) for k in self.chunks.chunks_of(clip)]


def _sos_for(butterworth_filter, lowcut, highcut, sample_rate):
    if not butterworth_filter:
        return None
    return fe.butter_bandpass_sos(lowcut, highcut, sample_rate, order=int(butterworth_filter))


def prepare_chunks(wav: torch.Tensor, offsets, chunker, *, sample_rate=16000, butterworth_filter=None, lowcut=200,
                   highcut=1800, pad_hint: int = 0, ctx: fe.Context | None = None) -> ChunkBatch:
    """[band-pass] -> trim -> ``chunker(n_trimmed)`` per clip -> materialised chunk views.

    ``chunker`` returns a list of chunk plans (frontend.plan_*) or None to drop the clip.
    ``pad_hint`` = expected padded-chunk length, used to size the spare capacity of the work
    buffer so that padded copies land behind the signal without re-allocation.
    """
    ctx = ctx or fe.default_ctx()
    o = fe._as_offsets(offsets)
    n = o.size - 1
    total = int(o[-1])
    launches = 0
    sos = _sos_for(butterworth_filter, lowcut, highcut, sample_rate)
    lens = np.diff(o)
    spare = int(pad_hint) * (int((lens < pad_hint).sum()) + 16) if pad_hint else 0
    frame_len = int(sample_rate / 10)
    if sos is not None:  # band-pass and the trim of the filtered signal in one call
        work = torch.empty(total + spare, dtype=torch.float32, device=wav.device)
        _, se = fe.iir_sos_trim(wav, o, sos, out=work, frame_length=frame_len, hop_length=int(frame_len / 2), ctx=ctx)
    else:
        work = wav
        se = fe.trim_indices(work, o, frame_length=frame_len, hop_length=int(frame_len / 2), ctx=ctx)
    launches += ctx.last_launches
    se = se.cpu().numpy()  # the one host round trip: durations drive the reference's control flow
    chunk_lists, valid = [], np.ones(n, dtype=bool)
    chunker.dup_called = False  # set by chunkers whenever the reference would call _duplicate_padding
    for i in range(n):
        chunks = chunker(int(se[i, 1] - se[i, 0]))
        if chunks is None:
            valid[i] = False
            chunks = []
        chunk_lists.append(chunks)
    dup = bool(getattr(chunker, "dup_called", False))
    clip_starts = o[:-1] + se[:, 0]
    work, starts, lengths, clip_ids, is_view = fe.materialise_chunks(work, total, clip_starts, chunk_lists, ctx=ctx,
                                                                     owned=sos is not None)
    if not is_view.all():
        launches += ctx.last_launches
    return ChunkBatch(work, starts, lengths, clip_ids, n, valid, se, dup, launches, is_view, sos is not None)


def _truncate(chunk, new_len):
    if chunk[0] == "view":
        return fe._view(chunk[1], min(chunk[2], new_len))
    _, length, src_start, period, a_end, a_phase, b_end, b_start = chunk
    new_len = min(length, new_len)
    return fe._gather(new_len, src_start, period, min(a_end, new_len), a_phase, min(b_end, new_len), b_start)


def entire_signal_chunker(input_sec=8, sample_rate=16000, pad=False, types="repeat", max_sec=None):
    """Control flow of get_entire_signal_librosa after the trim (src/util.py:248-259)."""

    def chunker(n):
        duration = n / sample_rate
        chunk = fe._view(0, n)
        if duration < input_sec:
            if not pad:
                return None
            if n == 0:
                return None  # the reference would loop forever / divide by zero on an empty clip
            chunk = fe.plan_split_pad(n, input_sec, sample_rate, types)[0]
            chunker.dup_called = chunker.dup_called or types != "zero"
        if max_sec and duration > max_sec:
            chunk = _truncate(chunk, int(max_sec * sample_rate))
        return [chunk]

    return chunker


def split_signal_chunker(input_sec=8, sample_rate=16000, trim_tail=False, types="repeat", first_only=False):
    """get_split_signal_librosa / get_split_signal_fbank_pad after the trim (src/util.py:348-354).
    ``first_only`` keeps chunk 0 only - what the cache writers take with ``[...][0]``
    (finetuning.py:973-975, 1126-1133; heart_pressl.py:35-37)."""

    def chunker(n):
        if n == 0:
            return None
        chunks = fe.plan_split_pad(n, input_sec, sample_rate, types)
        chunker.dup_called = chunker.dup_called or types != "zero"
        duration = n / sample_rate
        if trim_tail and duration > input_sec and (duration % input_sec) * 2 < input_sec:  # decide_droplast
            chunks.pop()
        return chunks[:1] if first_only else chunks

    return chunker


def log_mel_features(cb: ChunkBatch, f_max=8000, n_mels=64, f_min=50, nfft=1024, hop=512, sample_rate=16000,
                     mode="normalised", out: torch.Tensor | None = None) -> FeatureBatch:
    """``sample_rate`` is the rate of the MEL BASIS.  The reference's callers never forward their own ``sample_rate``
    to pre_process_audio_mel_t (src/util.py:198,261-263,358-360): the basis is always built for 16 kHz, whatever rate
    the audio was loaded at; the batch entry points below do the same."""
    plan = fe.logmel_plan(sample_rate, n_mels, f_min, f_max, nfft, hop)
    if len(cb.starts) == 0:
        return FeatureBatch(torch.empty((0, n_mels), device=cb.work.device), np.zeros(1, np.int64), cb, cb.launches)
    if out is not None and out.numel() < int((1 + cb.lengths // hop).sum()) * n_mels:
        out = None  # caller's buffer is too small for this batch: allocate
    out, fo = fe.logmel_views(plan, cb.work, cb.starts, cb.lengths, mode=mode, out=out, alt=cb.alt)
    return FeatureBatch(out, fo, cb, cb.launches + plan.last_launches)


def fbank_features(cb: ChunkBatch, sample_rate=16000, rows_per_chunk=0, min_samples_exclusive=400) -> FeatureBatch:
    """kaldi fbank per chunk; chunks of <= 400 samples are skipped as in the reference
    (``if waveform.shape[1] > 400``, src/util.py:844; extract_feature.py:231)."""
    plan = fe.fbank_plan(sample_rate=sample_rate)
    keep = cb.lengths > min_samples_exclusive
    kept = ChunkBatch(cb.work, cb.starts[keep], cb.lengths[keep], cb.clip_ids[keep], cb.n_clips, cb.valid, cb.trim,
                      cb.used_duplicate_padding, cb.launches, None if cb.is_view is None else cb.is_view[keep],
                      cb.filtered)
    if len(kept.starts) == 0:
        return FeatureBatch(torch.empty((0, plan.n_mels), device=cb.work.device), np.zeros(1, np.int64), kept, cb.launches)
    out, ro = plan.views(cb.work, kept.starts, kept.lengths, rows_per_clip=rows_per_chunk)
    return FeatureBatch(out, ro, kept, cb.launches + plan.last_launches)


# ----------------------------------------------------------------------------------------------
# batch versions of the reference entry points
# ----------------------------------------------------------------------------------------------


def _entire_signal_fast(wav, offsets, input_sec, sample_rate, butterworth_filter, pad, types, lowcut, highcut, max_sec,
                        ctx=None, work=None) -> ChunkBatch:
    """Vectorised planning for get_entire_signal_librosa (one output chunk per recording): the
    same integer formulas as frontend.plan_* evaluated with numpy over the whole batch."""
    ctx = ctx or fe.default_ctx()
    o = fe._as_offsets(offsets)
    n_clips, total = o.size - 1, int(o[-1])
    L = int(input_sec * sample_rate)
    launches = 0
    sos = _sos_for(butterworth_filter, lowcut, highcut, sample_rate)
    lens = np.diff(o)
    spare = L * (int((lens < L).sum()) + 16) if pad else 0
    frame_len = int(sample_rate / 10)
    alt, pad_buf = None, None
    if sos is not None:  # band-pass and the trim of the filtered signal in one call
        if work is None or work.numel() < total:  # a caller-provided work buffer is reused across calls
            work = torch.empty(total + spare, dtype=torch.float32, device=wav.device)
        _, se = fe.iir_sos_trim(wav, o, sos, out=work, frame_length=frame_len, hop_length=int(frame_len / 2), ctx=ctx)
    else:
        pad_buf = work if (work is not None and work.data_ptr() != wav.data_ptr()) else None  # caller's scratch for the padded copies
        work = wav
        se = fe.trim_indices(work, o, frame_length=frame_len, hop_length=int(frame_len / 2), ctx=ctx)
    launches += ctx.last_launches
    se = se.cpu().numpy()
    n = se[:, 1] - se[:, 0]
    dur = n / sample_rate
    short = dur < input_sec
    valid = ~short | (bool(pad) & (n > 0))
    starts = o[:-1] + se[:, 0]
    lengths = n.copy()
    if max_sec:
        cut = dur > max_sec
        lengths[cut] = np.minimum(lengths[cut], int(max_sec * sample_rate))
    is_view = np.ones(n_clips, dtype=bool)
    padded = np.flatnonzero(short & valid)
    dup = False
    if padded.size:
        npad = n[padded]
        d = np.zeros(padded.size, dtype=fe.GATHER_DTYPE)
        d["src_off"] = starts[padded]
        d["dst_off"] = total + L * np.arange(padded.size, dtype=np.int64)
        d["len"] = L
        d["period"] = npad
        if types == "zero":  # _equally_slice_pad_sample -> one slice -> _zero_padding
            tile = npad / L < 0.5
            copies = (L - 1) // npad
            d["a_end"] = np.where(tile, copies * npad, 0)
            d["b_end"] = np.where(tile, copies * npad, npad)
        else:  # _duplicate_padding: source at the end, tail of the doubled clip in front
            left = L - npad
            k = np.ceil(np.log2(np.maximum(1.0, left / npad))).astype(np.int64)
            len_aug = npad << k
            len_aug = np.where(len_aug < left, len_aug * 2, len_aug)  # guard log2 round-off
            half = len_aug // 2
            len_aug = np.where((len_aug > npad) & (half >= left), half, len_aug)
            d["a_end"] = left
            d["a_phase"] = (len_aug - left) % npad
            d["b_end"] = L
            dup = True
        need = total + L * padded.size
        if sos is not None and need <= work.numel():  # spare capacity behind the band-passed copy (a buffer of ours or
            # the caller's explicit ``work=``); the caller's SIGNAL tensor is never written, however roomy it is
            fe.gather(work, work, d, ctx=ctx)
            starts[padded] = d["dst_off"]
        else:  # read-only signal without room: padded copies go to their own buffer, addressed by negative starts
            if pad_buf is None or pad_buf.numel() < L * padded.size:
                pad_buf = torch.empty(L * padded.size, dtype=torch.float32, device=work.device)
            d["dst_off"] -= total
            fe.gather(work, pad_buf, d, ctx=ctx)
            alt = pad_buf
            starts[padded] = -(d["dst_off"] + 1)
        launches += ctx.last_launches
        lengths[padded] = L
        is_view[padded] = False
    keep = np.flatnonzero(valid)
    return ChunkBatch(work, starts[keep], lengths[keep], keep.astype(np.int64), n_clips, valid, se, dup, launches,
                      is_view[keep], sos is not None, alt)


class DeviceFeatures:
    """Result of ``entire_signal_device``: everything has been ENQUEUED, nothing has been read back.  ``features`` holds
    the rows of the kept recordings back to back (its row count is only an upper bound until ``resolve``);
    ``resolve()`` waits for the small descriptor copy and returns the usual FeatureBatch with host-side offsets."""

    def __init__(self, features, work, alt, h_meta, event, n_clips, total, L, filtered, types, launches, keep_alive, max_len=0):
        self.features, self.work, self.alt = features, work, alt
        self._h_meta, self._event, self._n, self._total, self._L, self._max_len = h_meta, event, n_clips, total, L, max_len
        self._filtered, self._types, self.launches, self._keep = filtered, types, launches, keep_alive
        self._fb = None

    def resolve(self) -> FeatureBatch:
        if self._fb is None:
            self._event.synchronize()
            n = self._n
            m = self._h_meta.numpy()
            se = m[: 2 * n].reshape(n, 2).copy()
            starts, lengths = m[2 * n : 3 * n], m[3 * n : 4 * n]
            frame_off = m[4 * n : 5 * n + 1]
            rows = np.diff(frame_off)
            keep = np.flatnonzero(rows > 0)
            nn = se[:, 1] - se[:, 0]
            valid = np.zeros(n, dtype=bool)
            valid[keep] = True
            nk = np.minimum(nn[keep], self._max_len) if self._max_len else nn[keep]
            padded = nk < self._L                        # short recordings that were kept: chunk length L
            stored = lengths[keep].copy()                # < L where only trailing zeros were added and nothing was copied
            chunk_len = np.where(padded, self._L, stored)
            is_view = ~padded
            dup = bool(padded.any()) and self._types != "zero"
            cb = ChunkBatch(self.work, starts[keep].copy(), chunk_len, keep.astype(np.int64), n, valid, se, dup,
                            self.launches, is_view, self._filtered, self.alt, stored if bool((stored < chunk_len).any()) else None)
            ro = np.zeros(keep.size + 1, dtype=np.int64)
            np.cumsum(rows[keep], out=ro[1:])
            self._fb = FeatureBatch(self.features, ro, cb, self.launches)
        return self._fb


class _Resolved:
    """A FeatureBatch that was planned on the host, behind the DeviceFeatures interface."""

    def __init__(self, fb):
        self.features, self._fb = fb.features, fb

    def resolve(self):
        return self._fb


_dev_scratch: dict = {}


def entire_signal_device(wav, offsets, input_sec=8, sample_rate=16000, butterworth_filter=None, pad=False, types="repeat",
                         lowcut=200, highcut=1800, max_sec=None, f_max=8000, work=None, out=None, ctx=None) -> DeviceFeatures:
    """get_entire_signal_librosa(spectrogram=True) over a batch with NO host round trip: band-pass + trim, the
    planner kernel (duration test, pad / cut decisions, row and work-item offsets: ``hmfe_entire_plan_batch``), the
    padded copies, the log-mel kernels and an asynchronous copy of the descriptors are enqueued back to back on the
    current stream.  The reference's control flow (src/util.py:248-259) runs on the device; the host learns the row
    offsets when it asks (``DeviceFeatures.resolve``).  ``out`` must hold sum(1 + max(len_i, L) // 512) rows."""
    import ctypes as C

    from . import _lib

    if max_sec and max_sec < input_sec:
        raise ValueError("max_sec < input_sec (pad, then cut) is planned on the host: use entire_signal_batch")
    ctx = ctx or fe.default_ctx()
    o = fe._as_offsets(offsets)
    n, total = o.size - 1, int(o[-1])
    dev = wav.device
    L = int(input_sec * sample_rate)
    hop = 512
    plan = fe.logmel_plan(16000, 64, 50, f_max, 1024, hop)
    lens = np.diff(o)
    rows_ub = int((1 + np.maximum(lens, L if pad else 0) // hop).sum())
    launches = 0
    sos = _sos_for(butterworth_filter, lowcut, highcut, sample_rate)
    frame_len = int(sample_rate / 10)
    se = torch.empty((n, 2), dtype=torch.int64, device=dev)
    pad_elems = L * n if pad else 0  # upper bound: every recording may turn out short after the trim
    alt = None
    with torch.cuda.device(dev):
        if sos is not None:
            if work is None or work.numel() < total + pad_elems:
                work = torch.empty(total + pad_elems, dtype=torch.float32, device=dev)
            _lib.check(_lib.hmfe_iir_sos_trim_batch(ctx._h, C.c_void_p(wav.data_ptr()), o.ctypes.data_as(C.c_void_p), n,
                                                    np.ascontiguousarray(sos, dtype=np.float64).ctypes.data_as(C.c_void_p),
                                                    sos.shape[0], C.c_void_p(work.data_ptr()), C.c_void_p(), frame_len,
                                                    int(frame_len / 2), 60.0, C.c_void_p(se.data_ptr()), fe._stream_ptr()),
                       "hmfe_iir_sos_trim_batch")
            signal, dst, dst_base, use_alt = work, work, total, 0
        else:
            if pad and (work is None or work.data_ptr() == wav.data_ptr() or work.numel() < pad_elems):
                work = torch.empty(max(pad_elems, 1), dtype=torch.float32, device=dev)
            _lib.check(_lib.hmfe_trim_batch(ctx._h, C.c_void_p(wav.data_ptr()), o.ctypes.data_as(C.c_void_p), n, frame_len,
                                            int(frame_len / 2), 60.0, C.c_void_p(se.data_ptr()), fe._stream_ptr()), "hmfe_trim_batch")
            signal, dst, dst_base, use_alt = wav, work, 0, 1
            alt = work if pad else None
        launches += ctx.last_launches
        key = (dev, n, torch.cuda.current_stream().cuda_stream)  # scratch is reused by stream-ordered calls only
        sc = _dev_scratch.get(key)
        if sc is None:
            ws = int(_lib.hmfe_logmel_device_workspace_bytes(n))
            sc = _dev_scratch[key] = {"desc": torch.empty(4 * n + 3, dtype=torch.int64, device=dev),
                                      "gather": torch.empty(n * 40, dtype=torch.uint8, device=dev),
                                      "ws": torch.empty(ws, dtype=torch.uint8, device=dev)}
            if len(_dev_scratch) > 64:
                _dev_scratch.pop(next(iter(_dev_scratch)))
        _lib.check(_lib.hmfe_entire_plan_batch(ctx._h, o.ctypes.data_as(C.c_void_p), n, C.c_void_p(se.data_ptr()), int(sample_rate),
                                               float(input_sec), int(bool(pad)), 2 if types == "zero" else 0, float(max_sec or 0.0), hop, 4,
                                               dst_base, use_alt, C.c_void_p(sc["desc"].data_ptr()),
                                               C.c_void_p(sc["gather"].data_ptr()), fe._stream_ptr()), "hmfe_entire_plan_batch")
        launches += 1
        if pad:
            _lib.check(_lib.hmfe_gather_device(ctx._h, C.c_void_p(signal.data_ptr()), C.c_void_p(dst.data_ptr()),
                                               C.c_void_p(sc["gather"].data_ptr()),
                                               C.c_void_p(sc["desc"].data_ptr() + 8 * (4 * n + 2)), n, L, fe._stream_ptr()),
                       "hmfe_gather_device")
            launches += 1
        if out is None or out.numel() < rows_ub * 64:
            out = torch.empty((rows_ub, 64), dtype=torch.float32, device=dev)
        _lib.check(_lib.hmfe_logmel_batch_device(plan._h, C.c_void_p(signal.data_ptr()),
                                                 C.c_void_p(alt.data_ptr()) if alt is not None else C.c_void_p(),
                                                 C.c_void_p(sc["desc"].data_ptr()), n, C.c_void_p(out.data_ptr()), 0,
                                                 C.c_void_p(sc["ws"].data_ptr()), fe._stream_ptr()), "hmfe_logmel_batch_device")
        launches += plan.last_launches
        h_meta = torch.empty(5 * n + 1, dtype=torch.int64, pin_memory=True)
        h_meta[: 2 * n].copy_(se.view(-1), non_blocking=True)
        h_meta[2 * n :].copy_(sc["desc"][: 3 * n + 1], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
    return DeviceFeatures(out, signal, alt, h_meta, ev, n, total, L, sos is not None, types, launches, (se, sc),
                          int(max_sec * sample_rate) if max_sec else 0)


def entire_signal_batch(wav, offsets, input_sec=8, sample_rate=16000, butterworth_filter=None, spectrogram=False,
                        pad=False, types="repeat", lowcut=200, highcut=1800, max_sec=None, f_max=8000, work=None,
                        out=None, planner="device"):
    """get_entire_signal_librosa over a batch.  Returns FeatureBatch (spectrogram=True) or ChunkBatch.

    ``planner="device"`` (default for spectrograms): the pad / cut / offset planning runs in a kernel and the host reads
    the descriptors back once, after everything has been enqueued (``entire_signal_device``); ``"host"``: the trim
    indices are read back in the middle of the step and numpy plans (the first implementation, kept as the
    cross-check of the planner kernel).

    ``work`` / ``out`` are optional pre-allocated device buffers (filtered signal + padded copies,
    feature rows) for callers that run many batches back to back."""
    if spectrogram and planner == "device" and (not max_sec or max_sec >= input_sec) and len(np.atleast_1d(offsets)) > 1:
        return entire_signal_device(wav, offsets, input_sec, sample_rate, butterworth_filter, pad, types, lowcut, highcut,
                                    max_sec, f_max, work=work, out=out).resolve()
    if not max_sec or max_sec >= input_sec:
        cb = _entire_signal_fast(wav, offsets, input_sec, sample_rate, butterworth_filter, pad, types, lowcut, highcut,
                                 max_sec, work=work)
    else:  # padded-then-cut corner: generic per-clip planner
        L = int(input_sec * sample_rate)
        cb = prepare_chunks(wav, offsets, entire_signal_chunker(input_sec, sample_rate, pad, types, max_sec),
                            sample_rate=sample_rate, butterworth_filter=butterworth_filter, lowcut=lowcut,
                            highcut=highcut, pad_hint=L if pad else 0)
    return log_mel_features(cb, f_max=f_max, out=out) if spectrogram else cb


def split_signal_batch(wav, offsets, input_sec=8, sample_rate=16000, butterworth_filter=None, spectrogram=False,
                       trim_tail=False, lowcut=200, highcut=1800, f_max=8000, first_only=False):
    """get_split_signal_librosa over a batch."""
    cb = prepare_chunks(wav, offsets, split_signal_chunker(input_sec, sample_rate, trim_tail, first_only=first_only),
                        sample_rate=sample_rate,
                        butterworth_filter=butterworth_filter, lowcut=lowcut, highcut=highcut,
                        pad_hint=int(input_sec * sample_rate))
    return log_mel_features(cb, f_max=f_max) if spectrogram else cb


def split_signal_fbank_pad_batch(wav, offsets, input_sec=8, sample_rate=16000, butterworth_filter=None,
                                 spectrogram=False, trim_tail=False, rows_per_chunk=0, first_only=False):
    """get_split_signal_fbank_pad over a batch."""
    cb = prepare_chunks(wav, offsets, split_signal_chunker(input_sec, sample_rate, trim_tail, first_only=first_only),
                        sample_rate=sample_rate,
                        butterworth_filter=butterworth_filter, lowcut=200, highcut=1800,
                        pad_hint=int(input_sec * sample_rate))
    return fbank_features(cb, sample_rate, rows_per_chunk) if spectrogram else cb


def split_signal_fbank_batch(wav, offsets, input_sec=10, sample_rate=16000, rows_per_chunk=0):
    """get_split_signal_fbank (extract_feature.py:213-247) over a batch."""

    def chunker(n):
        return fe.plan_split_sample(n, input_sec, sample_rate)

    cb = prepare_chunks(wav, offsets, chunker, sample_rate=sample_rate)
    return fbank_features(cb, sample_rate, rows_per_chunk)


def individual_segments_batch(wav, offsets, input_sec=8, sample_rate=16000, hop_sec=2, butterworth_filter=None,
                              spectrogram=False):
    """get_individual_segments_librosa (src/util.py:141-202) over a batch (default f_max=2000 log-mel)."""

    def chunker(n):
        duration = n / sample_rate
        if duration < 2:
            return None

        def cut(t0, t1):
            a, b = min(int(t0 * sample_rate), n), min(int(t1 * sample_rate), n)
            return a, b - a

        chunks, start, end = [], 0, input_sec
        while end <= duration:
            chunks.append(fe._view(*cut(start, end)))
            start += hop_sec
            end += hop_sec
        if start + 2 < duration:
            a, ln = cut(start, end)
            pad = fe.plan_split_pad(ln, 8, sample_rate)[0]
            chunker.dup_called = True
            if pad[0] == "view":
                pad = fe._view(a + pad[1], pad[2])
            else:  # re-base the gather on the tail segment
                _, length, src_start, period, a_end, a_phase, b_end, b_start = pad
                pad = fe._gather(length, a + src_start, period, a_end, a_phase, b_end, b_start)
            chunks.append(pad)
        return chunks

    cb = prepare_chunks(wav, offsets, chunker, sample_rate=sample_rate, butterworth_filter=butterworth_filter,
                        lowcut=200, highcut=1800, pad_hint=8 * sample_rate)
    return log_mel_features(cb, f_max=2000) if spectrogram else cb


# ----------------------------------------------------------------------------------------------
# host-buffer entry (what an extractor loop holding decoded audio in host memory calls)
# ----------------------------------------------------------------------------------------------


_host_pipes: dict = {}


def entire_signal_from_host(h_wav: torch.Tensor, offsets, h_out: torch.Tensor | None = None, *, sr_in: int | None = None,
                            resample: str = "torchaudio", chunk_bytes=512 << 20, device=None, **kw):
    """get_entire_signal_librosa(spectrogram=True) over a batch held in (pinned) HOST memory.

    ``h_wav`` holds the decoded files back to back: float32 samples or the int16 PCM payload, at ``sr_in`` Hz
    (default: already at ``sample_rate``).  With a native rate (CirCor: 4 kHz, PhysioNet 2016: 2 kHz) the batch
    crosses PCIe as it is on disk and ``librosa.load(path, sr=16000)``'s rate conversion (src/util.py:222) runs on the
    GPU (``frontend.RESAMPLE_PRESETS[resample]``), fused with the PCM16 decode.

    Sub-batches of about ``chunk_bytes`` of 16 kHz float32 samples flow through three streams - copy-in (+ decode /
    resample), compute, copy-out - so the PCIe transfers of neighbouring sub-batches overlap the kernels.  All device
    buffers (two input buffers, two sample buffers, two work buffers, two feature buffers) are allocated once per
    device and reused across sub-batches and calls.
    Returns (h_out [sum T, 64] host tensor, row_offsets [n_chunks+1], clip_ids, valid).
    """
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    o_in = fe._as_offsets(offsets)
    n = o_in.size - 1
    sr = kw.get("sample_rate", 16000)
    native = sr_in is not None and int(sr_in) != int(sr)
    if native:
        with torch.cuda.device(dev):
            rplan = fe.resample_plan(int(sr_in), int(sr), **fe.RESAMPLE_PRESETS[resample])
        o = np.zeros(o_in.size, dtype=np.int64)
        np.cumsum(rplan.out_lengths(np.diff(o_in)), out=o[1:])
    else:
        o = o_in
    bounds = [0]
    while bounds[-1] < n:
        c0 = bounds[-1]
        c1 = int(np.searchsorted(o, o[c0] + chunk_bytes // 4, side="right")) - 1
        bounds.append(min(n, max(c0 + 1, c1)))
    subs = list(zip(bounds[:-1], bounds[1:]))
    hop = 512
    input_sec = kw.get("input_sec", 8)
    L = int(input_sec * sr)
    pad = bool(kw.get("pad", False))
    lens = np.diff(o)
    rows_ub = 1 + np.maximum(lens, L if pad else 0) // hop
    max_samples = max(int(o[b] - o[a]) for a, b in subs)
    max_native = max(int(o_in[b] - o_in[a]) for a, b in subs)
    max_work = max(int(o[b] - o[a]) + (L * (b - a + 16) if pad else 0) for a, b in subs)
    max_rows = max(int(rows_ub[a:b].sum()) for a, b in subs)
    if h_out is None:
        h_out = torch.empty((int(rows_ub.sum()), 64), dtype=torch.float32, pin_memory=True)

    pcm = h_wav.dtype == torch.int16  # 16-bit WAV payload: half the PCIe bytes, decoded on the device
    if not pcm and h_wav.dtype != torch.float32:
        raise TypeError("h_wav must be float32 samples or int16 PCM")
    staged = pcm or native  # the host payload lands in a staging buffer first
    key = (dev, h_wav.dtype, staged)
    pipe = _host_pipes.get(key)
    need = (max_samples, max_work, max_rows, max_native if staged else 0)
    if pipe is None or any(c < x for c, x in zip(pipe["cap"], need)):
        cap = need if pipe is None else tuple(max(x, y) for x, y in zip(pipe["cap"], need))
        pipe = _host_pipes[key] = {
            "cap": cap,
            "d_raw": [torch.empty(cap[3], dtype=h_wav.dtype, device=dev) for _ in range(2)],
            # sample buffers carry the padding spare too: without a band-pass they double as the work buffer
            "d_in": [torch.empty(cap[1], dtype=torch.float32, device=dev) for _ in range(2)],
            "work": [torch.empty(cap[1], dtype=torch.float32, device=dev) for _ in range(2)],
            "feat": [torch.empty((cap[2], 64), dtype=torch.float32, device=dev) for _ in range(2)],
            "streams": [torch.cuda.Stream(device=dev) for _ in range(3)],
        }
    s_in, s_cmp, s_out = pipe["streams"]
    d_in, work, feat, d_raw = pipe["d_in"], pipe["work"], pipe["feat"], pipe["d_raw"]
    cur = torch.cuda.current_stream(dev)
    for s in pipe["streams"]:
        s.wait_stream(cur)
    ev_in = [torch.cuda.Event() for _ in range(2)]     # sample buffer filled
    ev_free = [torch.cuda.Event() for _ in range(2)]   # sample buffer consumed by the compute stream
    ev_out = [torch.cuda.Event() for _ in range(2)]    # feature buffer copied out

    def copy_in(i):
        a, b = subs[i]
        with torch.cuda.stream(s_in):
            if i >= 2:
                s_in.wait_event(ev_free[i % 2])
            n_in = int(o_in[b] - o_in[a])
            src = h_wav[int(o_in[a]) : int(o_in[b])]
            if native:  # (PCM16 decode +) rate conversion in one pass, straight into the sample buffer
                d_raw[i % 2][:n_in].copy_(src, non_blocking=True)
                rplan(d_raw[i % 2][:n_in], o_in[a : b + 1] - o_in[a], stream=s_in, out=d_in[i % 2])
            elif pcm:
                d_raw[i % 2][:n_in].copy_(src, non_blocking=True)
                fe.pcm16_to_f32(d_raw[i % 2][:n_in], out=d_in[i % 2], stream=s_in)
            else:
                d_in[i % 2][:n_in].copy_(src, non_blocking=True)
            ev_in[i % 2].record(s_in)

    # Every sub-batch is enqueued without waiting for its own trim indices (device-side planner).  Its copy-out needs
    # its row count and its position in h_out, so it is enqueued one iteration later, after its descriptors have
    # arrived - by then the next sub-batch's kernels are already queued and the GPU never idles.
    pending = []
    row_offsets, clip_ids, valid = [np.zeros(1, np.int64)], [], np.zeros(n, dtype=bool)
    base = 0

    def copy_out(j):  # descriptors of sub-batch j -> host offsets, rows -> h_out
        nonlocal base
        a_, b_, res_, done_ = pending[j]
        fb = res_.resolve()
        rows = int(fb.row_offsets[-1])
        with torch.cuda.stream(s_out):
            s_out.wait_event(done_)
            h_out[base : base + rows].copy_(fb.features[:rows], non_blocking=True)
            ev_out[j % 2].record(s_out)
        row_offsets.append(base + fb.row_offsets[1:])
        clip_ids.append(a_ + fb.chunks.clip_ids)
        valid[a_:b_] = fb.chunks.valid
        base += rows

    copy_in(0)
    for i, (a, b) in enumerate(subs):
        if i + 1 < len(subs):
            copy_in(i + 1)
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in[i % 2])
            if i >= 2:
                s_cmp.wait_event(ev_out[i % 2])  # feat[i % 2] still being copied out by sub-batch i - 2
            if kw.get("max_sec") and kw["max_sec"] < input_sec:  # pad-then-cut corner: host planner
                res = _Resolved(entire_signal_batch(d_in[i % 2], o[a : b + 1] - o[a], spectrogram=True, work=work[i % 2],
                                                    out=feat[i % 2], **kw))
            else:
                res = entire_signal_device(d_in[i % 2], o[a : b + 1] - o[a], work=work[i % 2], out=feat[i % 2], **kw)
            ev_free[i % 2].record(s_cmp)
            done = torch.cuda.Event()
            done.record(s_cmp)
        pending.append((a, b, res, done))
        if i >= 1:
            copy_out(i - 1)
    copy_out(len(subs) - 1)
    s_out.synchronize()
    s_cmp.synchronize()
    return h_out, np.concatenate(row_offsets), (np.concatenate(clip_ids) if clip_ids else np.zeros(0, np.int64)), valid
