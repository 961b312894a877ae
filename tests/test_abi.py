"""The C-ABI library loads on a CPU-only box and exports every symbol include/hmfe.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hmfe.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hmfe_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from heart_murmur_detection_b200 import build

    path = build.build()
    lib = ctypes.CDLL(path)
    names = _declared_symbols()
    assert len(names) >= 8
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.hmfe_version.restype = ctypes.c_int
    assert lib.hmfe_version() >= 100


def test_python_binding_covers_header():
    from heart_murmur_detection_b200 import _lib

    for n in _declared_symbols():
        assert hasattr(_lib, n), f"_lib.py does not bind {n}"


def test_argument_validation_without_gpu():
    """Pure host-side checks run before any CUDA call."""
    from heart_murmur_detection_b200 import _lib

    h = ctypes.c_void_p()
    rc = _lib.hmfe_logmel_plan_create(ctypes.byref(h), 16000, 1000, 512, 64, 50.0, 8000.0, 0)  # not a power of two
    assert rc == -3 and b"n_fft" in _lib.hmfe_last_error()
    rc = _lib.hmfe_logmel_plan_create(ctypes.byref(h), 16000, 1024, 512, 50, 50.0, 8000.0, 0)
    assert rc == -1
    assert _lib.hmfe_logmel_num_frames(128000, 512) == 251
    assert _lib.hmfe_logmel_num_frames(0, 512) == 1
