"""ctypes binding over shim/_build/libjst_b200.so: the reference's own Flowgraph / scheduler_synchronous / runtimes
with the reference CPU provider ("generic") AND the b200 binding linked in (shim/shim_capi.cc). tests/ and bench.py
use it to run one reference flowgraph on either target; the product (cyberether_b200) never imports it.

    s = Session()
    s.add_source("src", (8, 4096), "CF32", target=B200, sampleAxis=1, batchAxis=0)
    s.add_block("spec", "spectrum_engine", {"enableScale": True}, {"buffer": "src.signal"}, target=B200)
    s.write_source("src", x); s.compute(); y = s.read("spec", "buffer")
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libjst_b200.so")
DEMO_PATH = os.path.join(_HERE, "_build", "shim_demo")

CPU = (0, "generic")       # (device code, provider): the reference's own CPU modules
B200 = (1, "b200")         # DeviceType::CUDA, provider "b200" -> libb200dsp.so

DTYPE_CODES = {"F32": 0, "CF32": 1, "I8": 2, "U8": 3, "I16": 4, "U16": 5, "I32": 6, "U32": 7,
               "CI8": 8, "CU8": 9, "CI16": 10, "CU16": 11, "CI32": 12, "CU32": 13}

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run shim/build_shim.sh where /root/reference exists")
        L = ctypes.CDLL(LIB_PATH)
        vp, cp, u64, i64 = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int64
        L.jst_shim_last_error.restype = cp
        L.jst_shim_set_cuda_device.argtypes = [ctypes.c_int]
        L.jst_shim_create.restype = vp
        L.jst_shim_create.argtypes = [ctypes.c_int]
        L.jst_shim_destroy.argtypes = [vp]
        L.jst_shim_add_source.argtypes = [vp, cp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(u64), i64, i64, i64,
                                          ctypes.c_int, cp, ctypes.c_int]
        L.jst_shim_write_source.argtypes = [vp, cp, vp, u64]
        L.jst_shim_add_block.argtypes = [vp, cp, cp, cp, cp, ctypes.c_int, cp]
        L.jst_shim_reconfigure.argtypes = [vp, cp, cp]
        L.jst_shim_compute.argtypes = [vp]
        L.jst_shim_last_compute_seconds.argtypes = [vp]
        L.jst_shim_last_compute_seconds.restype = ctypes.c_double
        L.jst_shim_output_info.argtypes = [vp, cp, cp, ctypes.POINTER(i64)]
        L.jst_shim_output_attribute_f32.argtypes = [vp, cp, cp, cp, ctypes.POINTER(ctypes.c_float)]
        L.jst_shim_output_pointer.argtypes = [vp, cp, cp]
        L.jst_shim_output_pointer.restype = vp
        L.jst_shim_output_read.argtypes = [vp, cp, cp, vp, u64]
        L.jst_shim_metrics.argtypes = [vp, cp, cp, u64]
        L.jst_shim_viz_list.argtypes = [cp, u64]
        L.jst_shim_viz_read.argtypes = [cp, vp, u64]
        L.jst_shim_viz_read.restype = i64
        L.jst_shim_viz_write_index.argtypes = [cp]
        L.jst_shim_viz_write_index.restype = i64
        _lib = L
    return _lib


class ShimError(RuntimeError):
    pass


def set_cuda_device(index: int):
    """One process per GPU: the device of the reference's CUDA backend singleton (default 0). Call before the first
    Session."""
    L = lib()
    if L.jst_shim_set_cuda_device(int(index)) != 0:
        raise ShimError(L.jst_shim_last_error().decode(errors="replace"))


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
    """One reference Flowgraph. Sources are caller-filled; blocks are the reference's blocks on the chosen target."""

    def __init__(self, log_level: int = 1):
        self._L = lib()
        self._h = self._L.jst_shim_create(log_level)
        if not self._h:
            raise ShimError(self._L.jst_shim_last_error().decode(errors="replace"))

    def close(self):
        if self._h:
            self._L.jst_shim_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise ShimError(self._L.jst_shim_last_error().decode(errors="replace"))

    def add_source(self, name: str, shape: Sequence[int], dtype: str = "CF32", target: Tuple[int, str] = B200,
                   sampleAxis: int = -1, batchAxis: int = -1, channelAxis: int = -1, mapped: bool = False):
        """mapped (CUDA target only): the source is a CPU tensor mapped onto the device — Tensor(DeviceType::CUDA, cpu),
        what the reference's TestContext hands a CUDA module (src/testing.cc:133-136)."""
        arr = (ctypes.c_uint64 * len(shape))(*[int(v) for v in shape])
        self._check(self._L.jst_shim_add_source(self._h, name.encode(), DTYPE_CODES[dtype], len(shape), arr,
                                                sampleAxis, batchAxis, channelAxis, target[0], target[1].encode(),
                                                1 if mapped else 0))

    def write_source(self, name: str, array: np.ndarray):
        a = np.ascontiguousarray(array)
        self._check(self._L.jst_shim_write_source(self._h, name.encode(), a.ctypes.data_as(ctypes.c_void_p), a.nbytes))

    def write_source_ptr(self, name: str, host_ptr: int, nbytes: int):
        self._check(self._L.jst_shim_write_source(self._h, name.encode(), ctypes.c_void_p(host_ptr), nbytes))

    def add_block(self, name: str, type_: str, config: Optional[Dict[str, object]], inputs: Dict[str, str],
                  target: Tuple[int, str] = B200):
        wiring = "\n".join(f"{k}={v}" for k, v in inputs.items()).encode()
        self._check(self._L.jst_shim_add_block(self._h, name.encode(), type_.encode(), _kv(config), wiring, target[0],
                                               target[1].encode()))

    def reconfigure(self, name: str, config: Dict[str, object]):
        self._check(self._L.jst_shim_reconfigure(self._h, name.encode(), _kv(config)))

    def compute(self) -> float:
        self._check(self._L.jst_shim_compute(self._h))
        return self._L.jst_shim_last_compute_seconds(self._h)

    def info(self, block: str, port: str) -> dict:
        raw = (ctypes.c_int64 * 16)()
        self._check(self._L.jst_shim_output_info(self._h, block.encode(), port.encode(), raw))
        rank = int(raw[1])
        return {"dtype": {0: "F32", 1: "CF32"}.get(int(raw[0]), "other"), "shape": tuple(int(raw[2 + i]) for i in range(rank)),
                "sampleAxis": int(raw[10]), "batchAxis": int(raw[11]), "channelAxis": int(raw[12]),
                "contiguous": bool(raw[13]), "size": int(raw[14]), "device": "cuda" if raw[15] else "cpu"}

    def attribute_f32(self, block: str, port: str, key: str) -> Optional[float]:
        value = ctypes.c_float()
        rc = self._L.jst_shim_output_attribute_f32(self._h, block.encode(), port.encode(), key.encode(),
                                                   ctypes.byref(value))
        return float(value.value) if rc == 0 else None

    def pointer(self, block: str, port: str) -> int:
        p = self._L.jst_shim_output_pointer(self._h, block.encode(), port.encode())
        if not p:
            raise ShimError(f"no contiguous output {block}.{port}")
        return int(p)

    def read(self, block: str, port: str) -> np.ndarray:
        meta = self.info(block, port)
        dtype = np.complex64 if meta["dtype"] == "CF32" else np.float32
        out = np.empty(meta["shape"], dtype=dtype)
        self._check(self._L.jst_shim_output_read(self._h, block.encode(), port.encode(),
                                                 out.ctypes.data_as(ctypes.c_void_p), out.nbytes))
        return out

    def read_into(self, block: str, port: str, host_ptr: int, nbytes: int):
        self._check(self._L.jst_shim_output_read(self._h, block.encode(), port.encode(), ctypes.c_void_p(host_ptr),
                                                 nbytes))

    def modules(self, block: str) -> Dict[str, Tuple[int, float]]:
        """module name -> (cycles computed, accumulated ms): the block's Module::Timing metrics."""
        buf = ctypes.create_string_buffer(8192)
        n = self._L.jst_shim_metrics(self._h, block.encode(), buf, len(buf))
        if n < 0:
            raise ShimError(self._L.jst_shim_last_error().decode(errors="replace"))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cycles, ms = line.rsplit(" ", 2)
            out[name] = (int(cycles), float(ms))
        return out


def backend_info() -> dict:
    """Device table of the reference-side CUDA backend singleton (Backend::CUDA = shim/b200_backend.cc here)."""
    L = lib()
    L.jst_shim_backend_info.argtypes = [ctypes.c_char_p, ctypes.c_uint64]
    buf = ctypes.create_string_buffer(1024)
    if L.jst_shim_backend_info(buf, len(buf)) < 0:
        raise ShimError(L.jst_shim_last_error().decode(errors="replace"))
    device, name, cc, api, memory, primary = buf.value.decode().split("|")
    return {"device": int(device), "name": name, "compute_capability": cc, "api_version": api,
            "memory_bytes": int(memory), "primary_context": primary == "1"}


def viz_modules() -> Dict[str, str]:
    """Live lineplot / waterfall modules of every session: module name -> type."""
    buf = ctypes.create_string_buffer(8192)
    lib().jst_shim_viz_list(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        kind, name = line.split(":", 1)
        out[name] = kind
    return out


def viz_read(module: str) -> np.ndarray:
    """signalPoints (flat [n * 2]) of a lineplot module or the ring (flat [height * n]) of a waterfall module."""
    L = lib()
    n = L.jst_shim_viz_read(module.encode(), None, 0)
    if n < 0:
        raise ShimError(L.jst_shim_last_error().decode(errors="replace"))
    out = np.empty(n, np.float32)
    if L.jst_shim_viz_read(module.encode(), out.ctypes.data_as(ctypes.c_void_p), n) < 0:
        raise ShimError(L.jst_shim_last_error().decode(errors="replace"))
    return out


def viz_write_index(module: str) -> int:
    return int(lib().jst_shim_viz_write_index(module.encode()))
