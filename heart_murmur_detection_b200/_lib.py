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
hmfe_logmel_num_frames = _sig("hmfe_logmel_num_frames", C.c_int64, C.c_int64, C.c_int)
hmfe_logmel_mel_basis = _sig("hmfe_logmel_mel_basis", C.c_int, c_voidp, c_voidp)
hmfe_logmel_batch = _sig(
    "hmfe_logmel_batch", C.c_int, c_voidp, c_voidp, c_voidp, C.c_int64, c_voidp, C.c_int, c_voidp
)
hmfe_logmel_last_launches = _sig("hmfe_logmel_last_launches", C.c_int, c_voidp)
hmfe_logmel_set_profile = _sig("hmfe_logmel_set_profile", C.c_int, c_voidp, C.c_int)
hmfe_logmel_profile_ms = _sig(
    "hmfe_logmel_profile_ms", C.c_int, c_voidp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)
)


def check(rc: int, what: str = "hmfe call"):
    if rc != 0:
        raise HmfeError(f"{what} failed (rc={rc}): {hmfe_last_error().decode(errors='replace')}")
