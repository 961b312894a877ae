/* hmfe - heart-murmur front-end: C ABI of the B200 (sm_100a) audio preprocessing path.
 *
 * Drop-in boundary for the hot path of carla-biermann/heart-murmur-detection
 * (reference = /root/reference).  The reference is pure Python and has no FFI; each entry
 * point below cites the reference call it replaces, and INTEGRATION.md shows the ctypes
 * binding a maintainer would add to src/util.py / extract_feature.py.
 *
 * Conventions
 *   - plain C types only; `d_` pointers are device memory owned by the caller, `h_` pointers
 *     are host memory; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - a ragged batch of clips is one concatenated float32 buffer plus int64 offsets[n_clips+1]
 *     on the HOST (clip lengths are always known to the host: they come from file decode);
 *   - every function returns HMFE_OK (0) or a negative HMFE_ERR_*; the message is available
 *     from hmfe_last_error() (thread local).  Nothing throws, nothing falls back to the CPU;
 *   - calls are asynchronous with respect to `stream` unless stated otherwise.
 */
#ifndef HMFE_H_
#define HMFE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMFE_OK 0
#define HMFE_ERR_INVALID (-1)
#define HMFE_ERR_CUDA (-2)
#define HMFE_ERR_UNSUPPORTED (-3)

int hmfe_version(void);
const char* hmfe_last_error(void);
int hmfe_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * Log-mel spectrogram: replaces pre_process_audio_mel_t (src/util.py:481-501), i.e.
 * librosa.feature.melspectrogram(n_fft=1024, hop, Hann periodic, centre zero pad, power 2,
 * Slaney mel/norm) -> librosa.power_to_db(ref=np.max, amin 1e-10, top_db 80) -> per-clip
 * min-max normalisation -> transpose.  Output is [sum_i T_i, n_mels] float32 row-major with
 * T_i = 1 + n_i / hop frames per clip, clips packed back to back in batch order.
 * ------------------------------------------------------------------------------------------ */
typedef struct hmfe_logmel_plan hmfe_logmel_plan;

#define HMFE_LOGMEL_OUT_NORMALISED 0 /* reference output: dB min-max normalised to [0,1]   */
#define HMFE_LOGMEL_OUT_DB 1         /* power_to_db(ref=max, top_db=80) before normalising  */
#define HMFE_LOGMEL_OUT_POWER 2      /* mel power (linear)                                  */
#define HMFE_LOGMEL_OUT_DB_ABS 3     /* 10 log10(max(1e-10, S)): power_to_db(ref=1.0, top_db=None), the torchlibrosa
                                        LogmelFilterBank of the CLAP baseline (msclap/models/audio.py:146-175)    */

#define HMFE_VARIANT_AUTO 0
#define HMFE_VARIANT_SCALAR 1 /* one complex FFT (2 frames) per warp iteration            */
#define HMFE_VARIANT_PACKED 2 /* two complex FFTs (4 frames), packed FP32 (FFMA2)          */
#define HMFE_VARIANT_PAIR 3   /* one complex FFT, FFMA2 across element pairs, 20 warps / SM */
#define HMFE_VARIANT_TC 4     /* FFT as PACKED; mel projection on the tensor cores (tcgen05.mma, accumulators in tensor
                                 memory, bf16 hi/lo split of weights and powers), frames staged by bulk asynchronous
                                 copies; warp-specialised (FFT + MMA issue / epilogue warps).  n_mels <= 64, hop <= 512.
                                 The bulk copies move whole 16-byte granules: up to 12 bytes in front of the first and
                                 behind the last sample of a clip are READ (never used) when the clip is not 16-byte
                                 aligned - always inside a granule that holds samples of the buffer */

/* n_fft = 1024 (the only value the reference uses) runs the register-resident kernels selected by `variant`; other powers of
 * two from 64 to 4096 run a plain shared-memory FFT kernel (pre_process_audio_mel_t takes nfft as an argument,
 * src/util.py:482); anything else is HMFE_ERR_UNSUPPORTED.  n_mels a multiple of 4, <= 256; the register-resident kernels
 * take multiples of 32 (the reference uses 64), other counts run the plain kernel too. */
int hmfe_logmel_plan_create(hmfe_logmel_plan** plan, int sample_rate, int n_fft, int hop, int n_mels, double f_min,
                            double f_max, int variant);
void hmfe_logmel_plan_destroy(hmfe_logmel_plan* plan);
/* Centre padding of the STFT.  HMFE_PAD_CONSTANT (default): zeros, what librosa 0.10's melspectrogram does for
 * src/util.py:484.  HMFE_PAD_REFLECT: numpy 'reflect' (torchlibrosa Spectrogram(pad_mode="reflect") of the CLAP
 * baseline, msclap/models/audio.py:147,155-163); every clip must then be longer than n_fft / 2 samples. */
#define HMFE_PAD_CONSTANT 0
#define HMFE_PAD_REFLECT 1
int hmfe_logmel_plan_set_pad_mode(hmfe_logmel_plan* plan, int pad_mode);
/* frames librosa produces for a clip of n samples (centre=True): 1 + n / hop */
int64_t hmfe_logmel_num_frames(int64_t n_samples, int hop);
/* copies the float32 mel basis [n_mels][n_fft/2+1] the plan uses to host memory (tests) */
int hmfe_logmel_mel_basis(const hmfe_logmel_plan* plan, float* h_out);
int hmfe_logmel_batch(hmfe_logmel_plan* plan, const float* d_wav, const int64_t* h_offsets, int64_t n_clips,
                      float* d_out, int out_mode, void* stream);
/* same, but clip i is the view d_wav[h_starts[i] : h_starts[i] + h_lengths[i]] (clips may overlap
 * or leave gaps: trimmed recordings, 50 %-overlap chunks and padded copies are all views) */
int hmfe_logmel_batch_views(hmfe_logmel_plan* plan, const float* d_wav, const int64_t* h_starts,
                            const int64_t* h_lengths, int64_t n_clips, float* d_out, int out_mode, void* stream);
/* same with a second sample buffer: a clip with a negative start s is the view
 * d_wav_alt[-s - 1 : -s - 1 + length] (padded copies kept apart from a read-only signal buffer) */
int hmfe_logmel_batch_views2(hmfe_logmel_plan* plan, const float* d_wav, const float* d_wav_alt, const int64_t* h_starts,
                             const int64_t* h_lengths, int64_t n_clips, float* d_out, int out_mode, void* stream);
/* same with the per-clip descriptors already in device memory (written by hmfe_entire_plan_batch): d_desc[4 n + 2] =
 * clip_start | clip_len | frame_off | item_prefix.  Nothing is read back: the call is asynchronous and the step it
 * belongs to can be captured in a CUDA graph.  d_out must hold the caller's upper bound of rows. */
int hmfe_logmel_batch_device(hmfe_logmel_plan* plan, const float* d_wav, const float* d_wav_alt, const int64_t* d_desc,
                             int64_t n_clips, float* d_out, int out_mode, void* d_workspace, void* stream);
/* bytes of d_workspace for n_clips (per-clip statistics and the work queue) */
int64_t hmfe_logmel_device_workspace_bytes(int64_t n_clips);
/* HMFE_VARIANT_TC only: protocol-error word of the kernel's bounded mbarrier waits (0 = no error); synchronises. */
int hmfe_logmel_tc_status(hmfe_logmel_plan* plan, uint32_t* h_status);
/* number of kernel launches the last hmfe_logmel_batch call on this plan issued */
int hmfe_logmel_last_launches(const hmfe_logmel_plan* plan);
/* Measurement hook: when enabled, every hmfe_logmel_batch call records CUDA events on its
 * launch stream around the STFT+mel kernel and the dB/min-max kernel.  hmfe_logmel_profile_ms
 * synchronises on them, returns the summed durations since the last query and resets. */
int hmfe_logmel_set_profile(hmfe_logmel_plan* plan, int enable);
int hmfe_logmel_profile_ms(hmfe_logmel_plan* plan, double* power_ms, double* finalize_ms, int* n_calls);

/* ------------------------------------------------------------------------------------------
 * Context for the plan-less stages: owns descriptor staging and device scratch.  One context
 * per host thread / stream of work; calls on one context are ordered by the caller.
 * ------------------------------------------------------------------------------------------ */
typedef struct hmfe_ctx hmfe_ctx;
int hmfe_ctx_create(hmfe_ctx** ctx);
void hmfe_ctx_destroy(hmfe_ctx* ctx);
/* Caller-provided memory (the "never allocate" contract): by default a context grows its device scratch and its
 * pinned / device descriptor staging on demand (convenient, but a call may then allocate and, when it has to grow,
 * synchronise).  With
 *   hmfe_ctx_set_workspace(ctx, d_ws, bytes)   the stages use the caller's device workspace as scratch (256-byte
 *                                               aligned; NULL returns to the internal scratch) and fail with
 *                                               HMFE_ERR_INVALID, naming the size they need, when it is too small;
 *   hmfe_ctx_reserve(ctx, max_clips)           descriptor staging for batches of up to max_clips clips is allocated
 *                                               now, once; later calls allocate nothing and larger batches fail.
 * Sizes: hmfe_trim_workspace_bytes (hmfe_trim_batch), hmfe_iir_workspace_bytes (hmfe_iir_sos_batch and
 * hmfe_iir_sos_trim_batch: an upper bound over the algorithms they may pick; hop_length = 0 without the trim),
 * hmfe_sosfiltfilt_workspace_bytes, hmfe_hear_workspace_bytes, hmfe_logmel_device_workspace_bytes.  The gather,
 * planner and spectrogram stages need no scratch. */
int hmfe_ctx_set_workspace(hmfe_ctx* ctx, void* d_workspace, size_t bytes);
int hmfe_ctx_reserve(hmfe_ctx* ctx, int64_t max_clips);
int64_t hmfe_trim_workspace_bytes(const int64_t* h_offsets, int64_t n_clips, int frame_length, int hop_length);
int64_t hmfe_iir_workspace_bytes(const int64_t* h_offsets, int64_t n_clips, int n_sections, int hop_length);
/* number of kernel launches issued by the last call made on this context */
int hmfe_ctx_last_launches(const hmfe_ctx* ctx);
/* Measurement hook: with profiling enabled every kernel a ctx stage launches is bracketed by
 * CUDA events on its launch stream.  hmfe_ctx_profile_ms synchronises on them, writes the
 * summed milliseconds and launch counts per kernel id (arrays of HMFE_KERNEL_COUNT) and resets.
 * ids: 0 iir zero-state pass, 1 iir carry scan, 2 iir final pass, 3 trim frame power,
 *      4 trim first/last index, 5 pad-split gather, 6 spectrogram mean, 7 spectrogram crop,
 *      8 iir one-pass overlap kernel */
#define HMFE_KERNEL_COUNT 9
int hmfe_ctx_set_profile(hmfe_ctx* ctx, int enable);
int hmfe_ctx_profile_ms(hmfe_ctx* ctx, double* ms_by_kernel, int* launches_by_kernel);

/* ------------------------------------------------------------------------------------------
 * PCM16 decode: the sample conversion of librosa.load / soundfile for 16-bit WAV payloads
 * (src/util.py:153,222,323,391,805; extract_feature.py:214): d_out[i] = d_pcm[i] / 32768, exact.
 * Lets a loader ship the 2-byte file payload over PCIe instead of 4-byte floats.
 * ------------------------------------------------------------------------------------------ */
int hmfe_pcm16_decode(const int16_t* d_pcm, int64_t n, float* d_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Silence trim: replaces librosa.effects.trim(y, top_db=60, frame_length=sr/10, hop_length=sr/20)
 * (src/util.py:170-172, 237-244, 338-340, 820-822; extract_feature.py:219-221).
 * Writes int64 (start, end) per clip, clip-relative, to d_start_end[n_clips][2]; (0,0) when the
 * whole clip is silent.  Indices are exact (integer work); the RMS reduction is float32.
 * ------------------------------------------------------------------------------------------ */
int64_t hmfe_trim_num_frames(int64_t n_samples, int frame_length, int hop_length);
int hmfe_trim_batch(hmfe_ctx* ctx, const float* d_wav, const int64_t* h_offsets, int64_t n_clips, int frame_length,
                    int hop_length, float top_db, int64_t* d_start_end, void* stream);

/* ------------------------------------------------------------------------------------------
 * Pad / split / tile / repeat gather: applies host-computed index tables for
 * split_pad_sample, _duplicate_padding, _equally_slice_pad_sample/_zero_padding, the max_sec
 * cut and split_sample (src/util.py:504-620, 257-259; extract_feature.py:250-259).
 * Output chunk element i (0 <= i < len):
 *   i < a_end : src[src_off + (a_phase + i) mod period]
 *   i < b_end : src[src_off + b_start + (i - a_end)]
 *   else      : 0
 * ------------------------------------------------------------------------------------------ */
typedef struct hmfe_gather_desc {
    int64_t src_off; /* first sample of the source clip in d_src            */
    int64_t dst_off; /* first sample of this chunk in d_dst                 */
    int32_t len;     /* chunk length                                        */
    int32_t period;  /* source length used for the modular (repeat) part    */
    int32_t a_end;
    int32_t a_phase;
    int32_t b_end;
    int32_t b_start;
} hmfe_gather_desc;
int hmfe_gather_batch(hmfe_ctx* ctx, const float* d_src, float* d_dst, const hmfe_gather_desc* h_descs,
                      int64_t n_chunks, void* stream);

/* Device-side planner: the control flow of get_entire_signal_librosa between the silence trim and the log-mel
 * (src/util.py:248-259: duration test, "too short" -> dropped or padded to input_sec by _zero_padding /
 * _duplicate_padding (src/util.py:504-575), cut at max_sec) evaluated ON THE DEVICE from the trim indices
 * d_start_end[n_clips][2], so that the host does not wait for them.  Writes, for hmfe_logmel_batch_device,
 *   d_desc[4 n + 3] = clip_start[n] | clip_len[n] | frame_off[n + 1] | item_prefix[n + 1] | n_padded   (int64)
 * (dropped clips get zero rows) and the compact list d_gather[0 .. n_padded) of the padded copies for hmfe_gather_device
 * (room for n records; padded copy k of the batch goes to element dst_base + k * int(input_sec * sample_rate) of the destination;
 * with `alt` the clip start is encoded as -(offset + 1), the second-buffer convention of hmfe_logmel_batch_views2).
 * item_frames = frames per work item of the log-mel variant (4).  All outputs are caller-provided device memory.
 * pad_zero: 0 = _duplicate_padding ("repeat"), 1 = _zero_padding, 2 = _zero_padding without copying the clips that only get
 * trailing zeros (at least half of input_sec long): clip_len keeps their own length while frame_off counts the padded one -
 * hmfe_logmel_batch_device reads zeros behind the end of a clip, so the features are the same and the copy is saved. */
int hmfe_entire_plan_batch(hmfe_ctx* ctx, const int64_t* h_offsets, int64_t n_clips, const int64_t* d_start_end,
                           int sample_rate, double input_sec, int pad, int pad_zero, double max_sec, int hop, int item_frames,
                           int64_t dst_base, int alt, int64_t* d_desc, hmfe_gather_desc* d_gather, void* stream);
int hmfe_gather_device(hmfe_ctx* ctx, const float* d_src, float* d_dst, const hmfe_gather_desc* d_descs,
                       const int64_t* d_n_descs, int64_t max_chunks, int max_len, void* stream);

/* ------------------------------------------------------------------------------------------
 * Band-pass IIR: replaces scipy.signal.lfilter(b, a, x) in _butter_bandpass_filter
 * (src/util.py:113-126) with the equivalent second-order-section cascade, float64 arithmetic,
 * causal, zero initial state.  h_sos is [n_sections][6] = (b0 b1 b2 a0 a1 a2) as produced by
 * scipy.signal.butter(..., output="sos") / frontend.butter_bandpass_sos.  Either output may
 * be NULL: d_y32 feeds the device pipeline, d_y64 reproduces lfilter's float64 result.
 * ------------------------------------------------------------------------------------------ */
int hmfe_iir_sos_batch(hmfe_ctx* ctx, const float* d_x, const int64_t* h_offsets, int64_t n_clips,
                       const double* h_sos, int n_sections, float* d_y32, double* d_y64, void* stream);
/* Band-pass followed by the silence trim of the FILTERED signal, i.e. lines 226-244 of
 * get_entire_signal_librosa (src/util.py; same pair at :157-172, :327-340, :809-822) in one call:
 * y = lfilter(b, a, x) as above, and d_start_end[n_clips][2] = librosa.effects.trim indices of y
 * as hmfe_trim_batch would return them.  When the one-pass overlap kernel runs and
 * frame_length == 2 * hop_length, the frame energies are accumulated while y is produced and y
 * is not read back.  d_y32 is required. */
int hmfe_iir_sos_trim_batch(hmfe_ctx* ctx, const float* d_x, const int64_t* h_offsets, int64_t n_clips,
                            const double* h_sos, int n_sections, float* d_y32, double* d_y64, int frame_length,
                            int hop_length, float top_db, int64_t* d_start_end, void* stream);
/* Zero-phase variant: scipy.signal.sosfiltfilt(sos, x) with the default odd padding (padlen < 0 ->
 * scipy's default 3 * ntaps; hmfe_sosfiltfilt_padlen returns it).  BASELINE.json's north_star names
 * sosfiltfilt; the reference itself calls the causal lfilter (src/util.py:113-126), which is what
 * hmfe_iir_sos_batch reproduces - this is the additional mode, not the drop-in default.  Clips must be
 * longer than padlen (scipy raises otherwise).  d_workspace: hmfe_sosfiltfilt_workspace_bytes() bytes of
 * device memory.  Either output may be NULL. */
int hmfe_sosfiltfilt_padlen(const double* h_sos, int n_sections);
int64_t hmfe_sosfiltfilt_workspace_bytes(const int64_t* h_offsets, int64_t n_clips, int padlen);
int hmfe_sosfiltfilt_batch(hmfe_ctx* ctx, const float* d_x, const int64_t* h_offsets, int64_t n_clips,
                           const double* h_sos, int n_sections, int padlen, void* d_workspace, int64_t workspace_bytes,
                           float* d_y32, double* d_y64, void* stream);
/* Two realisations of the same filter.  SCAN: exact chunked scan (zero-state pass, carry scan with
 * the chunk transition matrix, final pass), any stable or unstable cascade.  OVERLAP: one pass in
 * which every chunk starts W samples early from a zero state, W chosen so that the cascade's
 * zero-input response has decayed below 1e-13 (infinity norm); refused for filters that need
 * W > 8192.  AUTO (default) picks the cheaper one for the batch at hand. */
#define HMFE_IIR_ALGO_AUTO 0
#define HMFE_IIR_ALGO_SCAN 1
#define HMFE_IIR_ALGO_OVERLAP 2
int hmfe_ctx_set_iir_algo(hmfe_ctx* ctx, int algo);
/* The overlap kernel moves rows with 16-byte accesses (VECTOR) when d_x and d_y32 sit on the same
 * 16-byte phase and no float64 output is requested, with 4-byte accesses (SCALAR) otherwise;
 * hmfe_ctx_set_iir_rows(ctx, HMFE_IIR_ROWS_SCALAR) forces the 4-byte variant (tests). */
#define HMFE_IIR_ROWS_AUTO 0
#define HMFE_IIR_ROWS_SCALAR 1
#define HMFE_IIR_ROWS_VECTOR 2
int hmfe_ctx_set_iir_rows(hmfe_ctx* ctx, int rows);
/* algorithm, chunk length, warm-up length and row mode the last IIR call on this context used */
int hmfe_ctx_last_iir_plan(const hmfe_ctx* ctx, int* algo, int* chunk, int* warmup, int* rows);

/* ------------------------------------------------------------------------------------------
 * Kaldi fbank: replaces torchaudio.compliance.kaldi.fbank(w, htk_compat=True,
 * sample_frequency=16000, use_energy=False, window_type="hanning", num_mel_bins=128, dither=0,
 * frame_shift=10, frame_length=25) (src/util.py:845-856; extract_feature.py:232-243).
 * Frames per clip: m = 1 + (n - win)/shift for n >= win, else 0 (snip_edges).
 * rows_per_clip == 0: output [sum m_i, n_mels], clips packed back to back.
 * rows_per_clip  > 0: output [n_clips, rows_per_clip, n_mels]; missing rows are zero, extra
 *                     frames dropped (model-side pad to 1024, audioMAE/models_mae.py:1178-1181).
 * ------------------------------------------------------------------------------------------ */
typedef struct hmfe_fbank_plan hmfe_fbank_plan;
int hmfe_fbank_plan_create(hmfe_fbank_plan** plan, int sample_rate, double frame_length_ms, double frame_shift_ms,
                           int n_mels, double low_freq, double high_freq, double preemph);
/* Same frame -> 512-point FFT -> banded mel -> log kernel with the caller's window [win] and dense mel
 * matrix [n_mels][257] (every row one contiguous band), for the sibling front-ends that differ from
 * Kaldi's only in these constants - e.g. VGGish (src/benchmark/baseline/vggish/mel_features.py:125-170,
 * 196-260, 342-400): periodic Hann, no DC removal / pre-emphasis, |X| instead of |X|^2, log(x + 0.01). */
#define HMFE_FB_REMOVE_DC 1   /* subtract the frame mean (Kaldi remove_dc_offset)       */
#define HMFE_FB_MAGNITUDE 2   /* mel of |X| instead of |X|^2                            */
#define HMFE_FB_LOG_OFFSET 4  /* log(x + log_offset) instead of log(max(x, FLT_EPSILON)) */
int hmfe_fbank_plan_create_custom(hmfe_fbank_plan** plan, int sample_rate, int win, int shift, int n_mels,
                                  const float* h_window, const float* h_mel, int flags, double preemph, double log_offset);
void hmfe_fbank_plan_destroy(hmfe_fbank_plan* plan);
int64_t hmfe_fbank_num_frames(const hmfe_fbank_plan* plan, int64_t n_samples);
int hmfe_fbank_mel_basis(const hmfe_fbank_plan* plan, float* h_out);
int hmfe_fbank_batch(hmfe_fbank_plan* plan, const float* d_wav, const int64_t* h_offsets, int64_t n_clips,
                     float* d_out, int rows_per_clip, void* stream);
int hmfe_fbank_batch_views(hmfe_fbank_plan* plan, const float* d_wav, const int64_t* h_starts, const int64_t* h_lengths,
                           int64_t n_clips, float* d_out, int rows_per_clip, void* stream);
int hmfe_fbank_last_launches(const hmfe_fbank_plan* plan);
/* measurement hook, as hmfe_logmel_set_profile / hmfe_logmel_profile_ms */
int hmfe_fbank_set_profile(hmfe_fbank_plan* plan, int enable);
int hmfe_fbank_profile_ms(hmfe_fbank_plan* plan, double* kernel_ms, int* n_calls);

/* ------------------------------------------------------------------------------------------
 * Polyphase resampler (the rate conversion inside librosa.load(..., sr=16000), src/util.py:153,
 * 222,323,391,805).  Implements the windowed-sinc algorithm of torchaudio.transforms.Resample
 * (used by the reference at src/model/models_eval.py:964-968): method 0 = sinc_interp_hann,
 * 1 = sinc_interp_kaiser.  librosa's own default (libsoxr HQ) is a closed third-party filter
 * design that is not installable here; see DESIGN.md ("parity unpinned").
 * Output clip i has ceil(n_i * new / orig) samples, clips packed back to back.
 * ------------------------------------------------------------------------------------------ */
typedef struct hmfe_resample_plan hmfe_resample_plan;
int hmfe_resample_plan_create(hmfe_resample_plan** plan, int orig_freq, int new_freq, int lowpass_filter_width,
                              double rolloff, int method, double beta);
void hmfe_resample_plan_destroy(hmfe_resample_plan* plan);
int64_t hmfe_resample_out_len(const hmfe_resample_plan* plan, int64_t n_in);
int hmfe_resample_taps(const hmfe_resample_plan* plan, int* n_phases, int* n_taps, float* h_out);
int hmfe_resample_batch(hmfe_resample_plan* plan, const float* d_in, const int64_t* h_in_offsets, int64_t n_clips,
                        float* d_out, void* stream);
/* same with the 16-bit PCM payload of the WAV file as input (sample = pcm / 32768, exact): decode and rate conversion
 * in one pass, for loaders that ship native-rate PCM16 over PCIe (CirCor: 4 kHz, 1/8 of the bytes of 16 kHz float32) */
int hmfe_resample_batch_pcm16(hmfe_resample_plan* plan, const int16_t* d_pcm, const int64_t* h_in_offsets, int64_t n_clips,
                              float* d_out, void* stream);
int hmfe_resample_last_launches(const hmfe_resample_plan* plan);

/* ------------------------------------------------------------------------------------------
 * Spectrogram-domain dataset ops: crop_first / random_crop / random_mask / random_multiply
 * (src/util.py:26-51) and pad-or-crop to a fixed frame count (cola_training.py:56-80,
 * mae_training.py:88-109, audioMAE/models_mae.py:1178-1181).  Random draws happen on the host
 * in the reference's order; the kernels apply them.
 * d_spec is a ragged batch of spectrograms [sum T_i, n_cols] with int64 row offsets.
 * Output item k = rows [src_row, src_row + n_rows) of d_spec (global row index); row r of the item
 * is replaced by the mean of spectrogram spec_id when d_row_mask[mask_off + r] is non-zero (every
 * item owns its slice of the mask: the reference draws a fresh mask per __getitem__), everything
 * times gain, zero padded to out_rows:  d_out[n_items][out_rows][n_cols].
 * ------------------------------------------------------------------------------------------ */
typedef struct hmfe_crop_desc {
    int64_t src_row;
    int32_t n_rows;
    int32_t spec_id;
    float gain;
    int32_t mask_off; /* first byte of this item's rows in d_row_mask (ignored without a mask) */
} hmfe_crop_desc;
int hmfe_spec_mean_batch(hmfe_ctx* ctx, const float* d_spec, const int64_t* h_row_offsets, int64_t n_specs, int n_cols,
                         float* d_mean, void* stream);
/* means over arbitrary row ranges [h_row_lo[i], h_row_hi[i]) of d_spec: the fill value of random_mask when it runs on a
 * window or a crop of a recording (mae_training.py:64-69, finetuning.py:96-102) */
int hmfe_spec_mean_ranges(hmfe_ctx* ctx, const float* d_spec, const int64_t* h_row_lo, const int64_t* h_row_hi, int64_t n,
                          int n_cols, float* d_mean, void* stream);
int hmfe_spec_crop_batch(hmfe_ctx* ctx, const float* d_spec, int n_cols, const hmfe_crop_desc* h_descs, int64_t n_items,
                         const uint8_t* d_row_mask, const float* d_mean, float* d_out, int out_rows, void* stream);

/* SpecAugmentation (torchlibrosa DropStripes as called from finetuning.py:64-69,104-116): zero rectangles
 * [row0, row0 + n_rows) x [col0, col0 + n_cols) of output item `item` of d_out[n_items][out_rows][n_cols].
 * The stripe positions are drawn on the host (torch.randint, the generator the reference consumes). */
typedef struct hmfe_rect_desc {
    int64_t item;
    int32_t row0, n_rows, col0, n_cols;
} hmfe_rect_desc;
int hmfe_spec_zero_rects(hmfe_ctx* ctx, float* d_out, int out_rows, int n_cols, int64_t n_items, const hmfe_rect_desc* h_rects,
                         int64_t n_rects, void* stream);

/* Host planner (no device work): the random draws of AudioDataset.__getitem__, method "cola"
 * (cola_training.py:56-80; mae_training.py:64-79 adds `windowing`) for n_items items of h_rows[i] frames each, from
 * the stream of uniforms h_u[n_u] in the reference's order: [window crop start] -> random_mask rows (second draw only
 * after a masked frame) -> two crop starts -> two gains.  Outputs: h_mask (concatenated per-item row masks, item i at
 * h_mask_off[i], rows counted inside the window), window / crop starts (Python int() truncation) and float32 gains.
 * Returns the number of uniforms consumed, -100 when h_u is too short (nothing is consumed then: call again with
 * more), HMFE_ERR_INVALID on bad arguments. */
int64_t hmfe_cola_draws(const double* h_u, int64_t n_u, const int64_t* h_rows, int64_t n_items, int max_len, int windowing,
                        int augment, double rate_start, double rate_seq, uint8_t* h_mask, int64_t* h_mask_off,
                        int64_t* h_win_start, int64_t* h_start1, int64_t* h_start2, float* h_gain1, float* h_gain2);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU: push `n` floats from local memory into the same offset of a buffer on EVERY GPU of the
 * node with NVLink-switch multicast stores (multimem.st).  `mc_dst` is the multicast mapping of a
 * symmetric allocation (e.g. torch.distributed._symmetric_memory: handle.multicast_ptr + byte
 * offset).  Used by dist.PeerAllGather for the all-gather of the feature blocks, the single
 * collective of the path (the reference itself is single GPU, cola_training.py:275-278).
 * ------------------------------------------------------------------------------------------ */
int hmfe_multicast_push(const float* d_src, float* mc_dst, int64_t n, int n_ctas, void* stream);

/* ------------------------------------------------------------------------------------------
 * HTS-AT input stage (the first thing the OPERA-CT encoder does to the log-mel; src/model/htsat/
 * htsat.py:889-891 bn0 in inference form, :829-858 reshape_wav2img): per mel bin y = x*scale+shift,
 * bicubic (align_corners) resize of the time axis to spec_size*ratio frames (ratio = spec_size /
 * n_cols; no resize for full-length items), fold to d_out[n_items][spec_size][spec_size] with
 * out[n*n_cols + f][t'] = y[n*(spec_size) + t'][f].  Item i = rows [h_src_row[i], +h_n_rows[i]) of
 * d_spec [rows][n_cols].  n_cols must be 64.
 * ------------------------------------------------------------------------------------------ */
int hmfe_htsat_input_batch(hmfe_ctx* ctx, const float* d_spec, int n_cols, const int64_t* h_src_row,
                           const int32_t* h_n_rows, int64_t n_items, const float* h_scale, const float* h_shift,
                           int spec_size, float* d_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * HeAR mel-PCEN front-end (sibling front-end of the thesis baselines): replaces preprocess_audio
 * (src/benchmark/baseline/hear/python/data_processing/audio_utils.py:448-476): batch-wide min / max
 * scaling to [-1, 1] (:361-365), 400-sample frames every 160 samples with a 400-POINT FFT and end
 * padding (:367-378, _compute_stft :22-115), |X|^2 @ mel[201][n_mels] (:379-382), PCEN with its EMA
 * smoother (:121-246), bilinear resize of the frame axis to out_rows (:386-445, :475).
 * d_audio is [n_clips][n_samples] float32; clips shorter than n_padded (32000 in the reference,
 * :466-468) are treated as zero padded BEFORE the scaling, as the reference does.
 * h_window [400] and h_mel [201][n_mels] (the reference's torch.hann_window(400) and
 * _linear_to_mel_weight_matrix(), :264-358) come from the caller; n_mels a multiple of 32, <= 128.
 * The workspace holds the mel power [n_clips][frames][n_mels] (+16 bytes), frames = ceil(n_padded/160).
 * ------------------------------------------------------------------------------------------ */
typedef struct hmfe_hear_plan hmfe_hear_plan;
int hmfe_hear_plan_create(hmfe_hear_plan** plan, const float* h_window, const float* h_mel, int n_mels, double alpha,
                          double smooth_coef, double delta, double root, double floor);
void hmfe_hear_plan_destroy(hmfe_hear_plan* plan);
int hmfe_hear_num_frames(int n_padded);
size_t hmfe_hear_workspace_bytes(const hmfe_hear_plan* plan, int64_t n_clips, int n_padded);
/* d_out [n_clips][out_rows][n_mels] (the reference's [B, 1, 192, 128]) */
int hmfe_hear_mel_pcen_batch(hmfe_hear_plan* plan, const float* d_audio, int64_t n_clips, int n_samples, int n_padded,
                             int out_rows, float* d_out, void* d_workspace, size_t workspace_bytes, void* stream);
/* mel power only: d_mel [n_clips][frames][n_mels]; the workspace needs 16 bytes */
int hmfe_hear_mel_batch(hmfe_hear_plan* plan, const float* d_audio, int64_t n_clips, int n_samples, int n_padded,
                        float* d_mel, void* d_workspace, size_t workspace_bytes, void* stream);
int hmfe_hear_last_launches(const hmfe_hear_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* HMFE_H_ */
