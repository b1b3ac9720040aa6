"""TEST INFRASTRUCTURE — ctypes binding over oracle/_ref/libjst_ref.so, i.e. the UNMODIFIED
reference (CyberEther/Jetstream v1.9.1) CPU compute path driven through its own
Flowgraph / scheduler_synchronous / NativeCpuRuntime (see oracle/ref_driver.cc).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module. The product path (cyberether_b200) never does.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libjst_ref.so")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(
                f"{LIB_PATH} missing: run oracle/build_ref.sh where /root/reference exists")
        L = ctypes.CDLL(LIB_PATH)
        L.jst_ref_last_error.restype = ctypes.c_char_p
        L.jst_ref_version.restype = ctypes.c_char_p
        L.jst_ref_create.restype = ctypes.c_void_p
        L.jst_ref_create.argtypes = [ctypes.c_int]
        L.jst_ref_destroy.argtypes = [ctypes.c_void_p]
        L.jst_ref_add_source.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                         ctypes.POINTER(ctypes.c_uint64), ctypes.c_int64,
                                         ctypes.c_int64, ctypes.c_int64]
        L.jst_ref_write_source.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p,
                                           ctypes.c_uint64]
        L.jst_ref_add_block.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p,
                                        ctypes.c_char_p, ctypes.c_char_p]
        L.jst_ref_reconfigure.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
        L.jst_ref_compute.argtypes = [ctypes.c_void_p]
        L.jst_ref_last_compute_seconds.argtypes = [ctypes.c_void_p]
        L.jst_ref_last_compute_seconds.restype = ctypes.c_double
        L.jst_ref_output_info.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p,
                                          ctypes.POINTER(ctypes.c_int64)]
        L.jst_ref_output_read.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p,
                                          ctypes.c_void_p, ctypes.c_uint64]
        L.jst_ref_metrics.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p,
                                      ctypes.c_uint64]
        _lib = L
    return _lib


class RefError(RuntimeError):
    pass


def _kv(d: Optional[Dict[str, object]]) -> bytes:
    if not d:
        return b""
    out = []
    for k, v in d.items():
        if isinstance(v, bool):
            v = "true" if v else "false"
        elif isinstance(v, (list, tuple)):
            v = "[" + ", ".join(repr(float(x)) if isinstance(x, float) else str(x) for x in v) + "]"
        out.append(f"{k}={v}")
    return "\n".join(out).encode()


class Session:
    """One reference Flowgraph. Sources are caller-filled; blocks are reference blocks."""

    def __init__(self, log_level: int = 1):
        self._L = lib()
        self._h = self._L.jst_ref_create(log_level)
        if not self._h:
            raise RefError(self._L.jst_ref_last_error().decode())
        self._sources: Dict[str, Tuple[np.dtype, Tuple[int, ...]]] = {}

    def close(self):
        if self._h:
            self._L.jst_ref_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise RefError(self._L.jst_ref_last_error().decode())

    _INT_CODES = {"I8": (2, np.int8), "U8": (3, np.uint8), "I16": (4, np.int16), "U16": (5, np.uint16),
                  "I32": (6, np.int32), "U32": (7, np.uint32), "CI8": (8, np.int8), "CU8": (9, np.uint8),
                  "CI16": (10, np.int16), "CU16": (11, np.uint16), "CI32": (12, np.int32), "CU32": (13, np.uint32)}

    def add_source(self, name: str, array: np.ndarray, sample_axis: int = -1, batch_axis: int = -1,
                   channel_axis: int = -1, dtype: Optional[str] = None):
        """`dtype` names an integer tensor type ("I8" ... "CU32"); complex integers are passed as an integer
        array whose last axis holds (re, im)."""
        array = np.ascontiguousarray(array)
        shape_logical = array.shape
        if dtype is not None:
            dt, np_type = self._INT_CODES[dtype]
            if array.dtype != np_type:
                raise TypeError(f"{dtype} source needs a {np_type} array")
            if dtype.startswith("C"):
                assert array.shape[-1] == 2
                shape_logical = array.shape[:-1]
        elif array.dtype == np.float32:
            dt = 0
        elif array.dtype == np.complex64:
            dt = 1
        else:
            raise TypeError("source must be float32 or complex64 (or pass dtype=...)")
        shape = (ctypes.c_uint64 * len(shape_logical))(*shape_logical)
        self._check(self._L.jst_ref_add_source(self._h, name.encode(), dt, len(shape_logical), shape,
                                               sample_axis, batch_axis, channel_axis))
        self._sources[name] = (array.dtype, array.shape)
        self.write_source(name, array)

    def write_source(self, name: str, array: np.ndarray):
        dtype, shape = self._sources[name]
        array = np.ascontiguousarray(array, dtype=dtype)
        assert array.shape == shape, (array.shape, shape)
        self._check(self._L.jst_ref_write_source(self._h, name.encode(),
                                                 array.ctypes.data_as(ctypes.c_void_p), array.nbytes))

    def add_block(self, name: str, type_: str, config: Optional[Dict[str, object]] = None,
                  inputs: Optional[Dict[str, str]] = None):
        self._check(self._L.jst_ref_add_block(self._h, name.encode(), type_.encode(), _kv(config),
                                              _kv(inputs)))

    def reconfigure(self, name: str, config: Dict[str, object]):
        self._check(self._L.jst_ref_reconfigure(self._h, name.encode(), _kv(config)))

    def compute(self) -> float:
        self._check(self._L.jst_ref_compute(self._h))
        return self._L.jst_ref_last_compute_seconds(self._h)

    def output_info(self, block: str, port: str) -> dict:
        info = (ctypes.c_int64 * 16)()
        self._check(self._L.jst_ref_output_info(self._h, block.encode(), port.encode(), info))
        rank = info[1]
        return dict(dtype={0: np.float32, 1: np.complex64}.get(info[0]), rank=rank,
                    shape=tuple(info[2 + i] for i in range(rank)), sample_axis=info[10],
                    batch_axis=info[11], channel_axis=info[12], contiguous=bool(info[13]))

    def output(self, block: str, port: str) -> np.ndarray:
        info = self.output_info(block, port)
        if info["dtype"] is None:
            raise RefError("unsupported output dtype")
        out = np.empty(info["shape"], dtype=info["dtype"])
        self._check(self._L.jst_ref_output_read(self._h, block.encode(), port.encode(),
                                                out.ctypes.data_as(ctypes.c_void_p), out.nbytes))
        return out

    def metrics(self, block: str) -> Dict[str, Tuple[int, float]]:
        buf = ctypes.create_string_buffer(1 << 14)
        n = self._L.jst_ref_metrics(self._h, block.encode(), buf, len(buf))
        if n < 0:
            raise RefError(self._L.jst_ref_last_error().decode())
        out = {}
        for line in buf.value.decode().splitlines():
            k, c, t = line.rsplit(" ", 2)
            out[k] = (int(c), float(t))
        return out


# ---------------------------------------------------------------------------------------------
# Convenience one-shot wrappers (each builds a Flowgraph, computes once, returns numpy).
# ---------------------------------------------------------------------------------------------

def _axes_for(x: np.ndarray, sample_axis: Optional[int]):
    if sample_axis is None:
        sample_axis = x.ndim - 1
    batch_axis = -1
    if x.ndim >= 2:
        batch_axis = 0 if sample_axis != 0 else 1
    return sample_axis, batch_axis


def run_block(type_: str, inputs: Dict[str, np.ndarray], config: Optional[dict] = None,
              out_port: str = "signal", sample_axis: Optional[int] = None, cycles: int = 1,
              axes: Optional[Dict[str, Tuple[int, int, int]]] = None,
              dtypes: Optional[Dict[str, str]] = None) -> np.ndarray:
    with Session() as s:
        wiring = {}
        for port, arr in inputs.items():
            dtype = (dtypes or {}).get(port)
            if axes and port in axes:
                sa, ba, ca = axes[port]
            elif dtype is not None and dtype.startswith("C"):
                sa, ba = _axes_for(arr[..., 0], sample_axis)
                ca = -1
            else:
                sa, ba = _axes_for(arr, sample_axis)
                ca = -1
            s.add_source("src_" + port, arr, sa, ba, ca, dtype=dtype)
            wiring[port] = f"src_{port}.signal"
        s.add_block("dut", type_, config, wiring)
        for _ in range(cycles):
            s.compute()
        return s.output("dut", out_port)


def window(n: int) -> np.ndarray:
    with Session() as s:
        s.add_block("w", "window", {"size": n})
        s.compute()
        return s.output("w", "window")


def fft(x: np.ndarray, forward: bool = True, sample_axis: Optional[int] = None) -> np.ndarray:
    return run_block("fft", {"signal": x}, {"forward": forward}, "signal", sample_axis)


def multiply(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return run_block("multiply", {"a": a, "b": b}, None, "product")


def amplitude(x: np.ndarray, sample_axis: Optional[int] = None) -> np.ndarray:
    return run_block("amplitude", {"signal": x}, None, "signal", sample_axis)


def range_(x: np.ndarray, lo: float, hi: float) -> np.ndarray:
    return run_block("range", {"signal": x}, {"min": float(lo), "max": float(hi)}, "signal")


def spectrum_engine(x: np.ndarray, enable_scale: bool = True, range_min: float = -120.0,
                    range_max: float = 0.0, sample_axis: Optional[int] = None) -> np.ndarray:
    return run_block("spectrum_engine", {"buffer": x},
                     {"enableScale": enable_scale, "rangeMin": range_min, "rangeMax": range_max},
                     "buffer", sample_axis)


# ---------------------------------------------------------------------------------------------
# lineplot / waterfall: the reference modules' computeSubmit() run directly
# through Registry::BuildModule + Runtime (oracle/ref_driver.cc: jst_ref_viz_*).
# ---------------------------------------------------------------------------------------------

class VizSession:
    """One reference `lineplot` or `waterfall` module on a caller-filled F32 input; state (EMA / ring) persists
    across compute() calls like in the running reference."""

    def __init__(self, type_: str, shape: Sequence[int], config: Optional[dict] = None, sample_axis: int = -1,
                 batch_axis: int = -1, channel_axis: int = -1):
        L = lib()
        L.jst_ref_viz_create.restype = ctypes.c_void_p
        L.jst_ref_viz_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int,
                                         ctypes.POINTER(ctypes.c_uint64), ctypes.c_int64, ctypes.c_int64,
                                         ctypes.c_int64]
        L.jst_ref_viz_compute.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
        L.jst_ref_viz_reconfigure.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.jst_ref_viz_read.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
        L.jst_ref_viz_read.restype = ctypes.c_int64
        L.jst_ref_viz_write_index.argtypes = [ctypes.c_void_p]
        L.jst_ref_viz_write_index.restype = ctypes.c_int64
        L.jst_ref_viz_destroy.argtypes = [ctypes.c_void_p]
        self._L, self.type = L, type_
        arr = (ctypes.c_uint64 * len(shape))(*[int(v) for v in shape])
        self._h = L.jst_ref_viz_create(type_.encode(), _kv(config), len(shape), arr, sample_axis, batch_axis,
                                       channel_axis)
        if not self._h:
            raise RefError(L.jst_ref_last_error().decode())
        self.shape = tuple(int(v) for v in shape)

    def compute(self, x: np.ndarray):
        a = np.ascontiguousarray(x, dtype=np.float32)
        assert a.shape == self.shape
        if self._L.jst_ref_viz_compute(self._h, a.ctypes.data_as(ctypes.c_void_p), a.size) != 0:
            raise RefError(self._L.jst_ref_last_error().decode())

    def reconfigure(self, config: dict):
        if self._L.jst_ref_viz_reconfigure(self._h, _kv(config)) != 0:
            raise RefError(self._L.jst_ref_last_error().decode())

    def read(self) -> np.ndarray:
        """lineplot: signalPoints [n, 2]; waterfall: the ring [height, n]."""
        n = self._L.jst_ref_viz_read(self._h, None, 0)
        out = np.empty(n, np.float32)
        self._L.jst_ref_viz_read(self._h, out.ctypes.data_as(ctypes.c_void_p), n)
        return out

    def write_index(self) -> int:
        return int(self._L.jst_ref_viz_write_index(self._h))

    def close(self):
        if self._h:
            self._L.jst_ref_viz_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
