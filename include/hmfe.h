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

#define HMFE_VARIANT_AUTO 0
#define HMFE_VARIANT_SCALAR 1 /* one complex FFT (2 frames) per warp iteration            */
#define HMFE_VARIANT_PACKED 2 /* two complex FFTs (4 frames), packed FP32 (FFMA2)          */

/* n_fft must be 1024 (the only value the reference uses); n_mels a multiple of 32, <= 256. */
int hmfe_logmel_plan_create(hmfe_logmel_plan** plan, int sample_rate, int n_fft, int hop, int n_mels, double f_min,
                            double f_max, int variant);
void hmfe_logmel_plan_destroy(hmfe_logmel_plan* plan);
/* frames librosa produces for a clip of n samples (centre=True): 1 + n / hop */
int64_t hmfe_logmel_num_frames(int64_t n_samples, int hop);
/* copies the float32 mel basis [n_mels][n_fft/2+1] the plan uses to host memory (tests) */
int hmfe_logmel_mel_basis(const hmfe_logmel_plan* plan, float* h_out);
int hmfe_logmel_batch(hmfe_logmel_plan* plan, const float* d_wav, const int64_t* h_offsets, int64_t n_clips,
                      float* d_out, int out_mode, void* stream);
/* number of kernel launches the last hmfe_logmel_batch call on this plan issued */
int hmfe_logmel_last_launches(const hmfe_logmel_plan* plan);
/* Measurement hook: when enabled, every hmfe_logmel_batch call records CUDA events on its
 * launch stream around the STFT+mel kernel and the dB/min-max kernel.  hmfe_logmel_profile_ms
 * synchronises on them, returns the summed durations since the last query and resets. */
int hmfe_logmel_set_profile(hmfe_logmel_plan* plan, int enable);
int hmfe_logmel_profile_ms(hmfe_logmel_plan* plan, double* power_ms, double* finalize_ms, int* n_calls);

#ifdef __cplusplus
}
#endif
#endif /* HMFE_H_ */
