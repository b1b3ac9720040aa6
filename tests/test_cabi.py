"""The C-ABI library loads and exports every symbol include/b200dsp.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

from cyberether_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200dsp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header():
    assert sorted(_native.SIGNATURES) == declared_symbols()


def test_pure_host_entry_points_work_without_a_gpu():
    lib = _native.load()
    assert lib.b200_version().startswith(b"b200dsp")
    out = ctypes.c_float()
    assert lib.b200_amplitude_scaling_coeff(4096, ctypes.byref(out)) == 0
    assert abs(out.value - (-72.2472)) < 1e-4
    s, o = ctypes.c_float(), ctypes.c_float()
    assert lib.b200_range_coefficients(-120.0, 0.0, ctypes.byref(s), ctypes.byref(o)) == 0
    assert abs(s.value - 1 / 120) < 1e-9 and o.value == 1.0
    assert lib.b200_range_coefficients(5.0, 5.0, ctypes.byref(s), ctypes.byref(o)) == 0 and s.value == 0.0 and o.value == 0.5
    # error convention: Result code + thread-local message, never an exception
    assert lib.b200_amplitude_scaling_coeff(0, ctypes.byref(out)) == 1
    assert b"bad argument" in lib.b200_last_error()


def test_filter_taps_host_matches_port():
    import numpy as np
    from oracle import port
    lib = _native.load()
    centers = (ctypes.c_double * 2)(0.0, 1.5e6)
    host = np.zeros((2, 33), np.complex64)
    assert lib.b200_filter_taps_host(8e6, 1e6, centers, 2, 33, host.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.array_equal(host, port.filter_taps(8e6, 1e6, [0.0, 1.5e6], 33))
    assert lib.b200_filter_taps_host(8e6, 1e6, centers, 2, 32, host.ctypes.data_as(ctypes.c_void_p)) == 1
    assert b"must be odd" in lib.b200_last_error()
