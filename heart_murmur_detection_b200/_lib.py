"""ctypes binding of libhmfe.so (the C ABI declared in include/hmfe.h).

There is no CPU fallback: importing this module without the built library raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhmfe.so")


class HmfeError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing. Build it with `python -m heart_murmur_detection_b200.build` "
        "(or __graft_entry__.build()); this package has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

c_i64p = C.POINTER(C.c_int64)
c_f32p = C.POINTER(C.c_float)
c_voidp = C.c_void_p


def _sig(name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


hmfe_version = _sig("hmfe_version", C.c_int)
hmfe_last_error = _sig("hmfe_last_error", C.c_char_p)
hmfe_device_info = _sig("hmfe_device_info", C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int))

hmfe_logmel_plan_create = _sig(
    "hmfe_logmel_plan_create", C.c_int, C.POINTER(c_voidp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
    C.c_int,
)
hmfe_logmel_plan_destroy = _sig("hmfe_logmel_plan_destroy", None, c_voidp)
hmfe_logmel_plan_set_pad_mode = _sig("hmfe_logmel_plan_set_pad_mode", C.c_int, c_voidp, C.c_int)
hmfe_logmel_num_frames = _sig("hmfe_logmel_num_frames", C.c_int64, C.c_int64, C.c_int)
hmfe_logmel_mel_basis = _sig("hmfe_logmel_mel_basis", C.c_int, c_voidp, c_voidp)
hmfe_logmel_batch = _sig(
    "hmfe_logmel_batch", C.c_int, c_voidp, c_voidp, c_voidp, C.c_int64, c_voidp, C.c_int, c_voidp
)
hmfe_logmel_batch_views = _sig(
    "hmfe_logmel_batch_views", C.c_int, c_voidp, c_voidp, c_voidp, c_voidp, C.c_int64, c_voidp, C.c_int, c_voidp
)
hmfe_logmel_batch_views2 = _sig(
    "hmfe_logmel_batch_views2", C.c_int, c_voidp, c_voidp, c_voidp, c_voidp, c_voidp, C.c_int64, c_voidp, C.c_int, c_voidp
)
hmfe_logmel_batch_device = _sig(
    "hmfe_logmel_batch_device", C.c_int, c_voidp, c_voidp, c_voidp, c_voidp, C.c_int64, c_voidp, C.c_int, c_voidp, c_voidp
)
hmfe_logmel_device_workspace_bytes = _sig("hmfe_logmel_device_workspace_bytes", C.c_int64, C.c_int64)
hmfe_logmel_last_launches = _sig("hmfe_logmel_last_launches", C.c_int, c_voidp)
hmfe_logmel_tc_status = _sig("hmfe_logmel_tc_status", C.c_int, c_voidp, C.POINTER(C.c_uint32))
hmfe_logmel_set_profile = _sig("hmfe_logmel_set_profile", C.c_int, c_voidp, C.c_int)
hmfe_logmel_profile_ms = _sig(
    "hmfe_logmel_profile_ms", C.c_int, c_voidp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)
)


hmfe_ctx_create = _sig("hmfe_ctx_create", C.c_int, C.POINTER(c_voidp))
hmfe_ctx_destroy = _sig("hmfe_ctx_destroy", None, c_voidp)
hmfe_ctx_set_workspace = _sig("hmfe_ctx_set_workspace", C.c_int, c_voidp, c_voidp, C.c_size_t)
hmfe_ctx_reserve = _sig("hmfe_ctx_reserve", C.c_int, c_voidp, C.c_int64)
hmfe_trim_workspace_bytes = _sig("hmfe_trim_workspace_bytes", C.c_int64, c_voidp, C.c_int64, C.c_int, C.c_int)
hmfe_iir_workspace_bytes = _sig("hmfe_iir_workspace_bytes", C.c_int64, c_voidp, C.c_int64, C.c_int, C.c_int)
hmfe_ctx_last_launches = _sig("hmfe_ctx_last_launches", C.c_int, c_voidp)
hmfe_ctx_set_profile = _sig("hmfe_ctx_set_profile", C.c_int, c_voidp, C.c_int)
hmfe_ctx_profile_ms = _sig("hmfe_ctx_profile_ms", C.c_int, c_voidp, C.POINTER(C.c_double), C.POINTER(C.c_int))
KERNEL_NAMES = ["iir_zero_state", "iir_carry", "iir_final", "trim_power", "trim_index", "gather", "spec_mean", "spec_crop",
                "iir_overlap"]

hmfe_multicast_push = _sig("hmfe_multicast_push", C.c_int, c_voidp, c_voidp, C.c_int64, C.c_int, c_voidp)
hmfe_pcm16_decode = _sig("hmfe_pcm16_decode", C.c_int, c_voidp, C.c_int64, c_voidp, c_voidp)

hmfe_trim_num_frames = _sig("hmfe_trim_num_frames", C.c_int64, C.c_int64, C.c_int, C.c_int)
hmfe_trim_batch = _sig(
    "hmfe_trim_batch", C.c_int, c_voidp, c_voidp, c_voidp, C.c_int64, C.c_int, C.c_int, C.c_float, c_voidp, c_voidp
)


class GatherDesc(C.Structure):
    """struct hmfe_gather_desc (include/hmfe.h)."""

    _fields_ = [
        ("src_off", C.c_int64), ("dst_off", C.c_int64), ("len", C.c_int32), ("period", C.c_int32),
        ("a_end", C.c_int32), ("a_phase", C.c_int32), ("b_end", C.c_int32), ("b_start", C.c_int32),
    ]


class CropDesc(C.Structure):
    """struct hmfe_crop_desc (include/hmfe.h)."""

    _fields_ = [("src_row", C.c_int64), ("n_rows", C.c_int32), ("spec_id", C.c_int32), ("gain", C.c_float),
                ("mask_off", C.c_int32)]


hmfe_entire_plan_batch = _sig(
    "hmfe_entire_plan_batch", C.c_int, c_voidp, c_voidp, C.c_int64, c_voidp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double,
    C.c_int, C.c_int, C.c_int64, C.c_int, c_voidp, c_voidp, c_voidp,
)
hmfe_gather_device = _sig("hmfe_gather_device", C.c_int, c_voidp, c_voidp, c_voidp, c_voidp, c_voidp, C.c_int64, C.c_int, c_voidp)
hmfe_gather_batch = _sig("hmfe_gather_batch", C.c_int, c_voidp, c_voidp, c_voidp, c_voidp, C.c_int64, c_voidp)
hmfe_iir_sos_batch = _sig(
    "hmfe_iir_sos_batch", C.c_int, c_voidp, c_voidp, c_voidp, C.c_int64, c_voidp, C.c_int, c_voidp, c_voidp, c_voidp
)

hmfe_iir_sos_trim_batch = _sig(
    "hmfe_iir_sos_trim_batch", C.c_int, c_voidp, c_voidp, c_voidp, C.c_int64, c_voidp, C.c_int, c_voidp, c_voidp, C.c_int,
    C.c_int, C.c_float, c_voidp, c_voidp,
)
hmfe_sosfiltfilt_padlen = _sig("hmfe_sosfiltfilt_padlen", C.c_int, c_voidp, C.c_int)
hmfe_sosfiltfilt_workspace_bytes = _sig("hmfe_sosfiltfilt_workspace_bytes", C.c_int64, c_voidp, C.c_int64, C.c_int)
hmfe_sosfiltfilt_batch = _sig(
    "hmfe_sosfiltfilt_batch", C.c_int, c_voidp, c_voidp, c_voidp, C.c_int64, c_voidp, C.c_int, C.c_int, c_voidp, C.c_int64,
    c_voidp, c_voidp, c_voidp,
)
IIR_ALGOS = {"auto": 0, "scan": 1, "overlap": 2}
hmfe_ctx_set_iir_algo = _sig("hmfe_ctx_set_iir_algo", C.c_int, c_voidp, C.c_int)
IIR_ROWS = {"auto": 0, "scalar": 1, "vector": 2}
hmfe_ctx_set_iir_rows = _sig("hmfe_ctx_set_iir_rows", C.c_int, c_voidp, C.c_int)
hmfe_ctx_last_iir_plan = _sig(
    "hmfe_ctx_last_iir_plan", C.c_int, c_voidp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
    C.POINTER(C.c_int),
)

hmfe_fbank_plan_create = _sig(
    "hmfe_fbank_plan_create", C.c_int, C.POINTER(c_voidp), C.c_int, C.c_double, C.c_double, C.c_int, C.c_double,
    C.c_double, C.c_double,
)
FB_REMOVE_DC, FB_MAGNITUDE, FB_LOG_OFFSET = 1, 2, 4
hmfe_fbank_plan_create_custom = _sig(
    "hmfe_fbank_plan_create_custom", C.c_int, C.POINTER(c_voidp), C.c_int, C.c_int, C.c_int, C.c_int, c_voidp, c_voidp, C.c_int,
    C.c_double, C.c_double,
)
hmfe_fbank_plan_destroy = _sig("hmfe_fbank_plan_destroy", None, c_voidp)
hmfe_fbank_num_frames = _sig("hmfe_fbank_num_frames", C.c_int64, c_voidp, C.c_int64)
hmfe_fbank_mel_basis = _sig("hmfe_fbank_mel_basis", C.c_int, c_voidp, c_voidp)
hmfe_fbank_batch = _sig("hmfe_fbank_batch", C.c_int, c_voidp, c_voidp, c_voidp, C.c_int64, c_voidp, C.c_int, c_voidp)
hmfe_fbank_batch_views = _sig(
    "hmfe_fbank_batch_views", C.c_int, c_voidp, c_voidp, c_voidp, c_voidp, C.c_int64, c_voidp, C.c_int, c_voidp
)
hmfe_fbank_last_launches = _sig("hmfe_fbank_last_launches", C.c_int, c_voidp)
hmfe_fbank_set_profile = _sig("hmfe_fbank_set_profile", C.c_int, c_voidp, C.c_int)
hmfe_fbank_profile_ms = _sig("hmfe_fbank_profile_ms", C.c_int, c_voidp, C.POINTER(C.c_double), C.POINTER(C.c_int))

hmfe_resample_plan_create = _sig(
    "hmfe_resample_plan_create", C.c_int, C.POINTER(c_voidp), C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double
)
hmfe_resample_plan_destroy = _sig("hmfe_resample_plan_destroy", None, c_voidp)
hmfe_resample_out_len = _sig("hmfe_resample_out_len", C.c_int64, c_voidp, C.c_int64)
hmfe_resample_taps = _sig("hmfe_resample_taps", C.c_int, c_voidp, C.POINTER(C.c_int), C.POINTER(C.c_int), c_voidp)
hmfe_resample_batch = _sig("hmfe_resample_batch", C.c_int, c_voidp, c_voidp, c_voidp, C.c_int64, c_voidp, c_voidp)
hmfe_resample_batch_pcm16 = _sig("hmfe_resample_batch_pcm16", C.c_int, c_voidp, c_voidp, c_voidp, C.c_int64, c_voidp, c_voidp)
hmfe_resample_last_launches = _sig("hmfe_resample_last_launches", C.c_int, c_voidp)

hmfe_spec_mean_batch = _sig(
    "hmfe_spec_mean_batch", C.c_int, c_voidp, c_voidp, c_voidp, C.c_int64, C.c_int, c_voidp, c_voidp
)
hmfe_spec_mean_ranges = _sig(
    "hmfe_spec_mean_ranges", C.c_int, c_voidp, c_voidp, c_voidp, c_voidp, C.c_int64, C.c_int, c_voidp, c_voidp
)
hmfe_spec_crop_batch = _sig(
    "hmfe_spec_crop_batch", C.c_int, c_voidp, c_voidp, C.c_int, c_voidp, C.c_int64, c_voidp, c_voidp, c_voidp, C.c_int,
    c_voidp,
)


class RectDesc(C.Structure):
    """struct hmfe_rect_desc (include/hmfe.h)."""

    _fields_ = [("item", C.c_int64), ("row0", C.c_int32), ("n_rows", C.c_int32), ("col0", C.c_int32), ("n_cols", C.c_int32)]


hmfe_spec_zero_rects = _sig(
    "hmfe_spec_zero_rects", C.c_int, c_voidp, c_voidp, C.c_int, C.c_int, C.c_int64, c_voidp, C.c_int64, c_voidp
)
hmfe_cola_draws = _sig(
    "hmfe_cola_draws", C.c_int64, c_voidp, C.c_int64, c_voidp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
    c_voidp, c_voidp, c_voidp, c_voidp, c_voidp, c_voidp, c_voidp,
)

hmfe_htsat_input_batch = _sig(
    "hmfe_htsat_input_batch", C.c_int, c_voidp, c_voidp, C.c_int, c_voidp, c_voidp, C.c_int64, c_voidp, c_voidp, C.c_int,
    c_voidp, c_voidp,
)

hmfe_hear_plan_create = _sig(
    "hmfe_hear_plan_create", C.c_int, C.POINTER(c_voidp), c_voidp, c_voidp, C.c_int, C.c_double, C.c_double, C.c_double,
    C.c_double, C.c_double,
)
hmfe_hear_plan_destroy = _sig("hmfe_hear_plan_destroy", None, c_voidp)
hmfe_hear_num_frames = _sig("hmfe_hear_num_frames", C.c_int, C.c_int)
hmfe_hear_workspace_bytes = _sig("hmfe_hear_workspace_bytes", C.c_size_t, c_voidp, C.c_int64, C.c_int)
hmfe_hear_mel_pcen_batch = _sig(
    "hmfe_hear_mel_pcen_batch", C.c_int, c_voidp, c_voidp, C.c_int64, C.c_int, C.c_int, C.c_int, c_voidp, c_voidp, C.c_size_t,
    c_voidp,
)
hmfe_hear_mel_batch = _sig(
    "hmfe_hear_mel_batch", C.c_int, c_voidp, c_voidp, C.c_int64, C.c_int, C.c_int, c_voidp, c_voidp, C.c_size_t, c_voidp
)
hmfe_hear_last_launches = _sig("hmfe_hear_last_launches", C.c_int, c_voidp)


def check(rc: int, what: str = "hmfe call"):
    if rc != 0:
        raise HmfeError(f"{what} failed (rc={rc}): {hmfe_last_error().decode(errors='replace')}")
