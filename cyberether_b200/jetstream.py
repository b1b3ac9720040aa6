"""Host-side mirror of the reference's Module / Runtime / Scheduler interface for the DSP hot path.

Same type strings, config fields, port names, taints, Result codes and error behaviour as the
reference, so tests read like the reference's own module tests
(`TestContext ctx("fft", ...); ctx.setConfig(...); ctx.setInput("signal", t); ctx.run()`,
src/testing.cc:105-170). Every `compute_submit` is one call into libb200dsp through the C ABI
(include/b200dsp.h); there is no CPU compute path here — tensors that are not on a CUDA device
make compute fail with Result.ERROR.

PyTorch is used for device memory and streams only (tensor allocation, H2D/D2H, current stream).

Reference interfaces mirrored:
  Module lifecycle / taints      include/jetstream/detail/module_impl.hh:31-116, include/jetstream/module.hh:53-63
  Registry 4-key lookup          src/registry.cc:583-622
  NativeCudaRuntime              src/runtime/native/cuda/impl.cc:35-118,185-272
  SynchronousScheduler           src/scheduler_synchronous.cc:315-568,574-749
  Signal axes                    src/memory/axis.cc:231-243
"""
from __future__ import annotations

import ctypes
import os
import enum
import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native
from ._native import B200Error


class Result(enum.IntEnum):  # include/jetstream/types.hh:19-30
    SUCCESS = 0
    ERROR = 1
    WARNING = 2
    FATAL = 3
    SKIP = 4
    YIELD = 5
    RELOAD = 6
    RECREATE = 7
    TIMEOUT = 8
    INCOMPLETE = 9


class Taint(enum.IntFlag):  # include/jetstream/module.hh:53-63
    CLEAN = 0
    DISCONTIGUOUS = 1
    CROSS_DEVICE = 2
    STATELESS = 4
    STATIC_OUTPUT = 8
    SURFACE = 16


_last_error = [""]


def last_error() -> str:
    return _last_error[0]


def _error(msg: str) -> Result:
    _last_error[0] = msg
    return Result.ERROR


DTYPES = {"F32": torch.float32, "CF32": torch.complex64,
          "I8": torch.int8, "U8": torch.uint8, "I16": torch.int16, "U16": torch.uint16,
          "I32": torch.int32, "U32": torch.uint32}
DTYPE_NAMES = {v: k for k, v in DTYPES.items()}
# Complex integers (CI8 ... CU32, include/jetstream/memory/types.hh) have no torch dtype: they are stored as the
# real integer type with a trailing axis of 2 (re, im) that the Tensor handle hides from shape/rank/size.
COMPLEX_INT_DTYPES = {"CI8": "I8", "CU8": "U8", "CI16": "I16", "CU16": "U16", "CI32": "I32", "CU32": "U32"}
DTYPE_CODES = {"F32": 0, "CF32": 1, "I8": 2, "U8": 3, "I16": 4, "U16": 5, "I32": 6, "U32": 7,
               "CI8": 8, "CU8": 9, "CI16": 10, "CU16": 11, "CI32": 12, "CU32": 13}      # include/b200dsp.h


# ---------------------------------------------------------------------------------------------
# Backend context (one per device) — replaces Backend::State<DeviceType::CUDA>()
# ---------------------------------------------------------------------------------------------

class Context:
    _instances: Dict[int, "Context"] = {}

    def __init__(self, device_index: int):
        self.lib = _native.load()
        handle = ctypes.c_void_p()
        _native.check(self.lib.b200_ctx_create(device_index, ctypes.byref(handle)))
        self.handle = handle
        self.device_index = device_index
        sms = ctypes.c_int()
        _native.check(self.lib.b200_ctx_sm_count(self.handle, ctypes.byref(sms)))
        self.sm_count = sms.value

    @classmethod
    def get(cls, device) -> "Context":
        index = torch.device(device).index
        if index is None:
            index = torch.cuda.current_device()
        if index not in cls._instances:
            cls._instances[index] = Context(index)
        return cls._instances[index]


def current_stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


# ---------------------------------------------------------------------------------------------
# Tensor (shape/dtype/attributes over shared device storage) — include/jetstream/memory/tensor.hh
# ---------------------------------------------------------------------------------------------

class Tensor:
    """Handle over shared storage with signal-axis attributes (sampleAxis/batchAxis/channelAxis)."""

    def __init__(self, data: Optional[torch.Tensor] = None, attributes: Optional[dict] = None,
                 complex_int: bool = False):
        self.data = data
        self.attributes: Dict[str, object] = dict(attributes or {})
        self.complex_int = bool(complex_int)
        if self.complex_int and (data is None or data.dim() < 1 or data.shape[-1] != 2 or
                                 "C" + DTYPE_NAMES.get(data.dtype, "?") not in COMPLEX_INT_DTYPES):
            raise TypeError("complex-integer tensors are integer storage with a trailing axis of 2")

    # -- construction
    @staticmethod
    def create(device, dtype: str, shape: Sequence[int]) -> "Tensor":
        if dtype in COMPLEX_INT_DTYPES:
            return Tensor(torch.zeros(tuple(int(s) for s in shape) + (2,), dtype=DTYPES[COMPLEX_INT_DTYPES[dtype]],
                                      device=device), complex_int=True)
        return Tensor(torch.zeros(tuple(int(s) for s in shape), dtype=DTYPES[dtype], device=device))

    @staticmethod
    def from_numpy(array: np.ndarray, device="cuda", dtype: Optional[str] = None, **axes) -> "Tensor":
        """`dtype="CI8"` (… "CU32") marks an integer array whose last axis holds (re, im)."""
        t = torch.from_numpy(np.ascontiguousarray(array))
        if t.dtype not in DTYPE_NAMES:
            raise TypeError(f"unsupported dtype {array.dtype}")
        complex_int = dtype in COMPLEX_INT_DTYPES
        if dtype is not None and not complex_int and DTYPES.get(dtype) != t.dtype:
            raise TypeError(f"array dtype {array.dtype} does not match '{dtype}'")
        if complex_int and DTYPES[COMPLEX_INT_DTYPES[dtype]] != t.dtype:
            raise TypeError(f"array dtype {array.dtype} does not match '{dtype}'")
        out = Tensor(t.to(device), complex_int=complex_int)
        for key, value in axes.items():
            if value is not None:
                out.set_attribute(key, int(value))
        return out

    def clone(self) -> "Tensor":  # new handle, same storage (Tensor::clone)
        return Tensor(self.data, self.attributes, self.complex_int)

    # -- introspection
    def valid(self) -> bool:
        return self.data is not None

    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(self.data.shape[:-1]) if self.complex_int else tuple(self.data.shape)

    @property
    def rank(self) -> int:
        return self.data.dim() - (1 if self.complex_int else 0)

    @property
    def dtype(self) -> str:
        name = DTYPE_NAMES[self.data.dtype]
        return "C" + name if self.complex_int else name

    @property
    def size(self) -> int:
        return self.data.numel() // (2 if self.complex_int else 1)

    @property
    def device(self):
        return self.data.device

    def contiguous(self) -> bool:
        return self.data.is_contiguous()

    def ptr(self) -> ctypes.c_void_p:
        return ctypes.c_void_p(self.data.data_ptr())

    def numpy(self) -> np.ndarray:
        return self.data.detach().cpu().numpy()

    # -- attributes
    def has_attribute(self, key: str) -> bool:
        return key in self.attributes

    def attribute(self, key: str):
        return self.attributes[key]

    def set_attribute(self, key: str, value) -> Result:
        self.attributes[key] = value
        return Result.SUCCESS

    def remove_attribute(self, key: str) -> Result:
        self.attributes.pop(key, None)
        return Result.SUCCESS

    def propagate_attributes(self, source: "Tensor") -> Result:
        self.attributes = dict(source.attributes)
        return Result.SUCCESS


@dataclass
class SignalAxes:
    sample: Optional[int] = None
    batch: Optional[int] = None
    channel: Optional[int] = None


def resolve_signal_axes(tensor: Tensor) -> Optional[SignalAxes]:
    """ResolveSignalAxes (src/memory/axis.cc:231-243): rank-1 tensors default to sampleAxis=0;
    axes must be in range and distinct; a sample axis is required."""
    axes = SignalAxes()
    for name in ("sample", "batch", "channel"):
        key = name + "Axis"
        if tensor.has_attribute(key):
            value = tensor.attribute(key)
            if not isinstance(value, int) or isinstance(value, bool):
                return None
            setattr(axes, name, value)
    if axes.sample is None and tensor.rank == 1:
        axes.sample = 0
    seen = set()
    for value in (axes.sample, axes.batch, axes.channel):
        if value is None:
            continue
        if value < 0 or value >= tensor.rank or value in seen:
            return None
        seen.add(value)
    if axes.sample is None:
        return None
    return axes


@dataclass
class TensorLink:  # include/jetstream/tensor_link.hh:22-32
    tensor: Tensor = field(default_factory=Tensor)
    producer: Optional[Tuple[str, str]] = None

    def produced(self, module: str, port: str, tensor: Tensor):
        self.producer = (module, port)
        self.tensor = tensor

    def resolved(self) -> bool:
        return self.tensor.valid()


# ---------------------------------------------------------------------------------------------
# Module base + registry
# ---------------------------------------------------------------------------------------------

class Module:
    """Module::Impl lifecycle: validate -> define -> (input checks) -> create; compute_submit per cycle."""

    TYPE = ""
    DEFAULTS: Dict[str, object] = {}
    device_type = "cuda"
    runtime_type = "native"
    provider = "b200"

    def __init__(self):
        self.name = ""
        self.config: Dict[str, object] = dict(self.DEFAULTS)
        self.inputs: Dict[str, TensorLink] = {}
        self.outputs: Dict[str, TensorLink] = {}
        self.input_ports: List[str] = []
        self.output_ports: List[str] = []
        self.taint = Taint.CLEAN
        self.state = "created-none"
        self.cycles = 0
        self.compute_time_ms = 0.0
        # Where input-less modules allocate: the CUDA device when one is visible. On a CPU-only box the
        # lifecycle (validate/define/create) still runs on host tensors, but compute_submit refuses.
        self.alloc_device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
            else torch.device("cpu")

    # -- hooks overridden by implementations
    def validate(self) -> Result:
        return Result.SUCCESS

    def define(self) -> Result:
        return Result.SUCCESS

    def create_impl(self) -> Result:
        return Result.SUCCESS

    def destroy(self) -> Result:
        return Result.SUCCESS

    def reconfigure_impl(self, candidate: Dict[str, object]) -> Result:
        return Result.RECREATE

    def compute_initialize(self) -> Result:
        return Result.SUCCESS

    def compute_submit(self, stream: ctypes.c_void_p) -> Result:
        return Result.SUCCESS

    def compute_deinitialize(self) -> Result:
        return Result.SUCCESS

    # -- helpers
    def define_taint(self, taint: Taint) -> Result:
        self.taint = taint
        return Result.SUCCESS

    def define_interface_input(self, port: str) -> Result:
        self.input_ports.append(port)
        return Result.SUCCESS

    def define_interface_output(self, port: str) -> Result:
        self.output_ports.append(port)
        return Result.SUCCESS

    # -- Module::create (src/module.cc:47-212)
    def create(self, name: str, config: Optional[Dict[str, object]], inputs: Dict[str, TensorLink]) -> Result:
        self.name = name
        unknown = set(config or {}) - set(self.DEFAULTS)
        if unknown:
            self.state = "errored"
            return _error(f"[MODULE_{self.TYPE.upper()}] Unknown config field(s): {sorted(unknown)}")
        self.config = dict(self.DEFAULTS)
        self.config.update(config or {})
        self.inputs = dict(inputs)
        self.outputs = {}
        self.input_ports, self.output_ports = [], []
        for hook in (self.validate, self.define):
            result = hook()
            if result != Result.SUCCESS:
                self.state = "errored"
                self.outputs = {}
                return result
        # Input checks (src/module.cc:134-158): every declared input present and resolved, same
        # device, contiguous unless DISCONTIGUOUS is declared.
        for port in self.input_ports:
            link = self.inputs.get(port)
            if link is None or not link.resolved():
                self.state = "incomplete"
                _error(f"[MODULE] Input '{port}' of '{name}' is not connected.")
                return Result.INCOMPLETE
            if not (self.taint & Taint.DISCONTIGUOUS) and not link.tensor.contiguous():
                self.state = "errored"
                return _error(f"[MODULE] Input '{port}' of '{name}' is not contiguous.")
        result = self.create_impl()
        if result != Result.SUCCESS:
            self.state = "errored"
            self.outputs = {}
            return result
        self.state = "created"
        return Result.SUCCESS

    def reconfigure(self, candidate: Dict[str, object]) -> Result:
        unknown = set(candidate) - set(self.DEFAULTS)
        if unknown:
            return _error(f"[MODULE_{self.TYPE.upper()}] Unknown config field(s): {sorted(unknown)}")
        merged = dict(self.config)
        merged.update(candidate)
        result = self.reconfigure_impl(merged)
        if result == Result.SUCCESS:
            self.config = merged
        return result


_REGISTRY: Dict[Tuple[str, str, str, str], Callable[[], Module]] = {}


def register_module(cls):
    """JST_REGISTER_MODULE(impl, device, runtime, provider) — duplicates are rejected (src/registry.cc:241-251)."""
    key = (cls.TYPE, cls.device_type, cls.runtime_type, cls.provider)
    if key in _REGISTRY:
        raise ValueError(f"duplicate module registration {key}")
    _REGISTRY[key] = cls
    return cls


def build_module(type_: str, device: str = "cuda", runtime: str = "native", provider: str = "b200") -> Module:
    """Registry::BuildModule: exact 4-key match (src/registry.cc:583-622)."""
    key = (type_, device, runtime, provider)
    if key not in _REGISTRY:
        raise KeyError(f"no module registered for {key}")
    return _REGISTRY[key]()


def list_available_modules(type_: str) -> List[Tuple[str, str, str, str]]:
    return [k for k in _REGISTRY if k[0] == type_]


def _lib():
    return _native.load()


def _call(fn_name: str, *args) -> Result:
    """One C-ABI call; maps a non-zero Result to the reference's error convention."""
    rc = getattr(_lib(), fn_name)(*args)
    if rc != 0:
        _last_error[0] = _lib().b200_last_error().decode(errors="replace")
        return Result(rc) if rc in Result._value2member_map_ else Result.ERROR
    return Result.SUCCESS


def _require_cuda(module: Module, *tensors: Tensor) -> Optional[Result]:
    for t in tensors:
        if t.device.type != "cuda":
            return _error(f"[MODULE_{module.TYPE.upper()}_B200] tensor is on '{t.device}', not a CUDA device; "
                          "this provider has no CPU path.")
    return None


def _u64_array(values: Sequence[int]):
    return (ctypes.c_uint64 * len(values))(*[int(v) for v in values])


def _contiguous_strides(shape: Sequence[int]) -> List[int]:
    strides, acc = [], 1
    for dim in reversed(list(shape)):
        strides.append(acc)
        acc *= int(dim)
    return list(reversed(strides))


class _Layout:
    """DISCONTIGUOUS support: brings a strided and/or permuted view into a contiguous staging tensor whose last
    axis is the chosen one (gather) and writes a contiguous staging result back through the inverse mapping
    (scatter), with b200_copy_strided — the role of the reference's fft_layout kernel
    (src/domains/dsp/fft/module_impl_native_cuda.cc:31-141). `direct` = already contiguous with the axis last."""

    def __init__(self, tensor: torch.Tensor, axis: Optional[int]):
        rank = tensor.dim()
        self.axis = rank - 1 if axis is None else axis
        self.perm = [d for d in range(rank) if d != self.axis] + [self.axis]
        self.direct = tensor.is_contiguous() and self.axis == rank - 1
        self.shape_p = [tensor.shape[d] for d in self.perm]

    def gather(self, ctx, src: torch.Tensor, staging: torch.Tensor, stream) -> "Result":
        shape = _u64_array(self.shape_p)
        src_stride = _u64_array([src.stride(d) for d in self.perm])
        dst_stride = _u64_array(_contiguous_strides(self.shape_p))
        return _call("b200_copy_strided", ctx.handle, ctypes.c_void_p(src.data_ptr()),
                     ctypes.c_void_p(staging.data_ptr()), src.element_size(), len(self.shape_p), shape, src_stride,
                     dst_stride, stream)

    def scatter(self, ctx, staging: torch.Tensor, dst: torch.Tensor, shape_p: Sequence[int], stream) -> "Result":
        shape = _u64_array(shape_p)
        src_stride = _u64_array(_contiguous_strides(shape_p))
        dst_stride = _u64_array([dst.stride(d) for d in self.perm])
        return _call("b200_copy_strided", ctx.handle, ctypes.c_void_p(staging.data_ptr()),
                     ctypes.c_void_p(dst.data_ptr()), dst.element_size(), len(shape_p), shape, src_stride, dst_stride,
                     stream)


# ---------------------------------------------------------------------------------------------
# Modules on the hot path (type strings / config fields / ports: SURVEY.md Appendix C)
# ---------------------------------------------------------------------------------------------

@register_module
class Window(Module):
    """`window` — src/domains/dsp/window/module_impl.cc:8-36 (+ native_cpu.cc:20-37)."""
    TYPE = "window"
    DEFAULTS = {"size": 1024}

    def validate(self):
        if int(self.config["size"]) == 0:
            return _error("[MODULE_WINDOW] Window size cannot be zero.")
        return Result.SUCCESS

    def define(self):
        self.define_taint(Taint.STATIC_OUTPUT)
        return self.define_interface_output("window")

    def create_impl(self):
        self.output = Tensor.create(self.alloc_device, "CF32", (int(self.config["size"]),))
        self.output.set_attribute("sampleAxis", 0)
        self.outputs["window"] = TensorLink()
        self.outputs["window"].produced(self.name, "window", self.output)
        return Result.SUCCESS

    def compute_submit(self, stream):
        err = _require_cuda(self, self.output)
        if err:
            return err
        ctx = Context.get(self.output.device)
        return _call("b200_window_blackman_cf32", ctx.handle, self.output.ptr(), self.output.size, stream)


def _empty_input_axis(tensor: "Tensor") -> int:
    """Sample axis of an EMPTY input (validate() returns early on those, docs/blocks-and-modules.md:182-186)."""
    axes = resolve_signal_axes(tensor)
    if axes is not None and axes.sample is not None:
        return axes.sample
    return max(tensor.rank - 1, 0)


@register_module
class Invert(Module):
    """`invert` — src/domains/dsp/invert/module_impl.cc + native_cpu.cc:78-103."""
    TYPE = "invert"

    def validate(self):
        self._axis = None
        link = self.inputs.get("signal")
        if link is None or not link.resolved() or link.tensor.size == 0:
            return Result.SUCCESS
        t = link.tensor
        if t.dtype != "CF32":
            return _error(f"[MODULE_INVERT_B200] Unsupported input data type: {t.dtype}.")
        axes = resolve_signal_axes(t)
        if axes is None:
            return _error("[MODULE_INVERT] Input must contain valid signal axis metadata.")
        self._axis = axes.sample
        return Result.SUCCESS

    def define(self):
        self.define_taint(Taint.DISCONTIGUOUS | Taint.STATELESS)
        self.define_interface_input("signal")
        return self.define_interface_output("signal")

    def create_impl(self):
        self.input = self.inputs["signal"].tensor
        if self._axis is None:      # empty input: validate() resolved nothing (it must tolerate empty tensors)
            self._axis = _empty_input_axis(self.input)
        self._layout = _Layout(self.input.data, None)
        self._layout.perm = list(range(self.input.rank))
        self._layout.shape_p = list(self.input.shape)
        self._layout.direct = self.input.contiguous()
        self._staging = None if self._layout.direct else torch.empty(self.input.shape, dtype=torch.complex64,
                                                                      device=self.input.device)
        self.output = Tensor.create(self.input.device, "CF32", self.input.shape)
        self.output.propagate_attributes(self.input)
        self.outputs["signal"] = TensorLink()
        self.outputs["signal"].produced(self.name, "signal", self.output)
        shape = self.input.shape
        self._outer = int(np.prod(shape[:self._axis], dtype=np.int64)) if self._axis > 0 else 1
        self._n = shape[self._axis]
        self._inner = int(np.prod(shape[self._axis + 1:], dtype=np.int64)) if self._axis + 1 < len(shape) else 1
        return Result.SUCCESS

    def compute_submit(self, stream):
        if self.input.size == 0:
            return Result.SUCCESS
        err = _require_cuda(self, self.input, self.output)
        if err:
            return err
        ctx = Context.get(self.input.device)
        src = self.input.ptr()
        if not self._layout.direct:
            result = self._layout.gather(ctx, self.input.data, self._staging, stream)
            if result != Result.SUCCESS:
                return result
            src = ctypes.c_void_p(self._staging.data_ptr())
        return _call("b200_invert_cf32", ctx.handle, src, self.output.ptr(), self._outer, self._n,
                     self._inner, stream)


def _parse_shape(text) -> Optional[List[int]]:
    if isinstance(text, (list, tuple)):
        return [int(v) for v in text]
    body = str(text).strip()
    if not (body.startswith("[") and body.endswith("]")):
        return None
    body = body[1:-1].strip()
    if not body:
        return []
    try:
        return [int(v) for v in body.split(",")]
    except ValueError:
        return None


@register_module
class Reshape(Module):
    """`reshape` — zero-copy view (src/domains/core/reshape/module_impl.cc); bit-exact by construction."""
    TYPE = "reshape"
    DEFAULTS = {"shape": "[]"}

    def define(self):
        self.define_interface_input("buffer")
        return self.define_interface_output("buffer")

    def create_impl(self):
        source = self.inputs["buffer"].tensor
        shape = _parse_shape(self.config["shape"])
        if shape is None or int(np.prod(shape, dtype=np.int64)) != source.size:
            return _error(f"[MODULE_RESHAPE] Cannot reshape {source.shape} into {self.config['shape']}.")
        self.output = Tensor(source.data.view(tuple(shape)), {})
        # Attributes do not survive a reshape unless the rank is unchanged.
        if len(shape) == source.rank:
            self.output.propagate_attributes(source)
        self.outputs["buffer"] = TensorLink()
        self.outputs["buffer"].produced(self.name, "buffer", self.output)
        return Result.SUCCESS


@register_module
class Cast(Module):
    """`cast` — bypass (alias) when the dtype already matches (src/domains/core/cast/module_impl.cc:23,98-101);
    otherwise the conversions of the reference's native implementations
    (module_impl_native_cpu.cc:60-75): integer -> F32 and F32 / complex integer -> CF32, scaled by
    128 / 32768 / 2147483648 (module_impl.cc:50-72)."""
    TYPE = "cast"
    DEFAULTS = {"outputType": "CF32"}
    REAL_INTS = ("I8", "U8", "I16", "U16", "I32", "U32")

    def validate(self):
        if self.config["outputType"] not in DTYPE_CODES:
            return _error(f"[MODULE_CAST] Invalid output type '{self.config['outputType']}'.")
        return Result.SUCCESS

    def define(self):
        self.define_taint(Taint.DISCONTIGUOUS | Taint.STATELESS)
        self.define_interface_input("buffer")
        return self.define_interface_output("buffer")

    def create_impl(self):
        self.input = self.inputs["buffer"].tensor
        self._staging = None
        out_type = self.config["outputType"]
        self.bypass = self.input.dtype == out_type
        if self.bypass:
            self.output = self.input.clone()
        else:
            in_type = self.input.dtype
            supported = (out_type == "F32" and in_type in self.REAL_INTS) or \
                        (out_type == "CF32" and (in_type == "F32" or in_type in COMPLEX_INT_DTYPES))
            if not supported:
                return _error(f"[MODULE_CAST_B200] Unsupported conversion '{in_type}' -> '{out_type}'.")
            if self.input.rank == 0:
                return _error("[MODULE_CAST] Cannot allocate a rank-zero cast output.")
            self.output = Tensor.create(self.input.device, out_type, self.input.shape)
            self.output.propagate_attributes(self.input)
        self.outputs["buffer"] = TensorLink()
        self.outputs["buffer"].produced(self.name, "buffer", self.output)
        return Result.SUCCESS

    def compute_submit(self, stream):
        if self.bypass:
            return Result.SUCCESS
        err = _require_cuda(self, self.input, self.output)
        if err:
            return err
        ctx = Context.get(self.input.device)
        src = self.input.data
        if not src.is_contiguous():
            # Taint::DISCONTIGUOUS: sliced / permuted views are gathered into a dense staging tensor first. A complex
            # integer is one element of twice the scalar size (its trailing (re, im) axis is always dense).
            pair = 2 if self.input.complex_int else 1
            if self.input.complex_int and src.stride(-1) != 1:
                return _error("[MODULE_CAST_B200] Complex-integer input must keep (re, im) adjacent.")
            if self._staging is None:
                self._staging = torch.empty(src.shape, dtype=src.dtype, device=src.device)
            shape = list(self.input.shape)
            strides = [src.stride(d) // pair for d in range(len(shape))]
            result = _call("b200_copy_strided", ctx.handle, ctypes.c_void_p(src.data_ptr()),
                           ctypes.c_void_p(self._staging.data_ptr()), src.element_size() * pair, len(shape),
                           _u64_array(shape), _u64_array(strides), _u64_array(_contiguous_strides(shape)), stream)
            if result != Result.SUCCESS:
                return result
            src = self._staging
        if self.input.dtype == "F32":
            return _call("b200_cast_f32_cf32", ctx.handle, ctypes.c_void_p(src.data_ptr()), self.output.ptr(),
                         self.input.size, stream)
        return _call("b200_cast_int", ctx.handle, ctypes.c_void_p(src.data_ptr()), DTYPE_CODES[self.input.dtype],
                     self.output.ptr(), self.input.size, stream)


def merge_broadcast_signal_axes(a: Tensor, b: Tensor, rank: int) -> Optional[Dict[str, int]]:
    """MergeBroadcastSignalAxes (src/memory/axis.cc:332-359): each tensor's roles are right-aligned into the
    output rank; a role present on both inputs must land on the same output axis, otherwise ERROR
    ("Signal roles map to conflicting output axes"). Returns None on conflict."""
    merged: Dict[str, int] = {}
    for key in ("sampleAxis", "batchAxis", "channelAxis"):
        mapped = []
        for t in (a, b):
            if t.has_attribute(key):
                mapped.append(int(t.attribute(key)) + (rank - t.rank))
        if len(mapped) == 2 and mapped[0] != mapped[1]:
            return None
        if mapped:
            merged[key] = mapped[0]
    if len(set(merged.values())) != len(merged):
        return None
    return merged


@register_module
class Multiply(Module):
    """`multiply` — NumPy-style broadcast product (src/domains/core/multiply/module_impl.cc:28-112)."""
    TYPE = "multiply"

    def validate(self):
        la, lb = self.inputs.get("a"), self.inputs.get("b")
        self._plan = None
        if not la or not lb or not la.resolved() or not lb.resolved():
            return Result.SUCCESS
        a, b = la.tensor, lb.tensor
        if a.size == 0 or b.size == 0:
            return Result.SUCCESS
        if a.dtype != b.dtype:
            return _error(f"[MODULE_MULTIPLY] Input data types differ: {a.dtype} vs {b.dtype}.")
        rank = max(a.rank, b.rank, 1)
        shape = [1] * rank
        for i in range(rank):
            da = a.shape[a.rank - 1 - i] if a.rank > i else 1
            db = b.shape[b.rank - 1 - i] if b.rank > i else 1
            if da != db and da != 1 and db != 1:
                return _error(f"[MODULE_MULTIPLY] Input shapes {list(a.shape)} and {list(b.shape)} are not broadcastable.")
            shape[rank - 1 - i] = max(da, db)
        self._axes = merge_broadcast_signal_axes(a, b, rank)
        if self._axes is None:
            return _error("[MEMORY:AXIS] Signal roles map to conflicting output axes.")
        self._plan = tuple(shape)
        return Result.SUCCESS

    def define(self):
        self.define_taint(Taint.DISCONTIGUOUS | Taint.STATELESS)
        self.define_interface_output("product")
        self.define_interface_input("a")
        return self.define_interface_input("b")

    def create_impl(self):
        if self._plan is None:
            return _error("[MODULE_MULTIPLY] Inputs are empty.")
        self.a = self.inputs["a"].tensor
        self.b = self.inputs["b"].tensor
        shape = self._plan
        self.view_a = self.a.data.broadcast_to(shape)
        self.view_b = self.b.data.broadcast_to(shape)
        self.c = Tensor.create(self.a.device, self.a.dtype, shape)
        self.c.propagate_attributes(self.a)
        for key in ("sampleAxis", "batchAxis", "channelAxis"):
            self.c.remove_attribute(key)
        self.c.attributes.update(self._axes)
        self.outputs["product"] = TensorLink()
        self.outputs["product"].produced(self.name, "product", self.c)
        self._shape = _u64_array(shape)
        self._stride_a = _u64_array(self.view_a.stride())
        self._stride_b = _u64_array(self.view_b.stride())
        return Result.SUCCESS

    def compute_submit(self, stream):
        err = _require_cuda(self, self.a, self.b, self.c)
        if err:
            return err
        ctx = Context.get(self.a.device)
        fn = "b200_multiply_cf32" if self.a.dtype == "CF32" else "b200_multiply_f32"
        return _call(fn, ctx.handle, ctypes.c_void_p(self.view_a.data_ptr()), ctypes.c_void_p(self.view_b.data_ptr()),
                     self.c.ptr(), len(self._plan), self._shape, self._stride_a, self._stride_b, stream)


@register_module
class MultiplyConstant(Module):
    """`multiply_constant` — src/domains/core/multiply_constant/module_impl_native_cpu.cc:82-100."""
    TYPE = "multiply_constant"
    DEFAULTS = {"constant": 1.0}

    def define(self):
        self.define_interface_input("factor")
        return self.define_interface_output("product")

    def create_impl(self):
        self.input = self.inputs["factor"].tensor
        self.output = Tensor.create(self.input.device, self.input.dtype, self.input.shape)
        self.output.propagate_attributes(self.input)
        self.outputs["product"] = TensorLink()
        self.outputs["product"].produced(self.name, "product", self.output)
        return Result.SUCCESS

    def reconfigure_impl(self, candidate):
        return Result.SUCCESS

    def compute_submit(self, stream):
        err = _require_cuda(self, self.input, self.output)
        if err:
            return err
        ctx = Context.get(self.input.device)
        fn = "b200_multiply_constant_cf32" if self.input.dtype == "CF32" else "b200_multiply_constant_f32"
        return _call(fn, ctx.handle, self.input.ptr(), self.output.ptr(), self.input.size,
                     ctypes.c_float(float(self.config["constant"])), stream)


@register_module
class Fft(Module):
    """`fft` — src/domains/dsp/fft/module_impl.cc:8-96: unnormalised DFT along the sample axis.
      CF32 -> CF32                                  pocketfft::c2c
      F32  -> CF32 [.., n/2+1] (forward, complexOutput)   pocketfft::r2c
      F32  -> F32 FFTPACK half-complex (either direction)   pocketfft::r2r_fftpack
    Any sample axis and strided inputs are accepted (Taint::DISCONTIGUOUS): they are gathered into the contiguous
    [batch, n] layout of the kernels and scattered back. Real-input transforms are composed from the C2C kernels
    plus pack/unpack steps (b200_fft_real_helper)."""
    TYPE = "fft"
    DEFAULTS = {"forward": True, "complexOutput": False}

    def __init__(self):
        super().__init__()
        self._plan_handle = None
        self._half_plan = None

    def validate(self):
        self._axis = None
        link = self.inputs.get("signal")
        if link is None or not link.resolved() or link.tensor.size == 0:
            return Result.SUCCESS
        t = link.tensor
        if t.dtype not in ("F32", "CF32"):
            return _error(f"[MODULE_FFT_B200] Unsupported input data type: {t.dtype}.")
        axes = resolve_signal_axes(t)
        if axes is None:
            return _error("[MODULE_FFT] Input must contain valid signal axis metadata.")
        self._axis = axes.sample
        return Result.SUCCESS

    def define(self):
        self.define_taint(Taint.DISCONTIGUOUS | Taint.STATELESS)
        self.define_interface_input("signal")
        return self.define_interface_output("signal")

    def create_impl(self):
        self.input = self.inputs["signal"].tensor
        t = self.input
        self._n = t.shape[self._axis]
        self._batch = t.size // self._n
        real_in = t.dtype == "F32"
        self._kind = "c2c"
        out_shape, out_dtype = list(t.shape), "CF32"
        if real_in and self.config["forward"] and self.config["complexOutput"]:
            self._kind = "r2c"
            out_shape[self._axis] = self._n // 2 + 1
        elif real_in:
            self._kind = "fftpack"
            out_dtype = "F32"
        self.output = Tensor.create(t.device, out_dtype, out_shape)
        self.output.propagate_attributes(t)
        self.outputs["signal"] = TensorLink()
        self.outputs["signal"].produced(self.name, "signal", self.output)
        self._layout = _Layout(t.data, self._axis)
        self._out_layout = _Layout(self.output.data, self._axis)
        dev = t.device
        rows = (self._batch, self._n)
        self._stage_in = None if (self._layout.direct and not real_in) else \
            torch.empty(self._layout.shape_p, dtype=t.data.dtype, device=dev)
        # complex work buffer [batch, n] for every path that is not the direct in->out C2C
        self._work = None
        if self._kind != "c2c" or not self._layout.direct:
            self._work = torch.empty(rows, dtype=torch.complex64, device=dev)
        self._stage_out = None
        if not self._out_layout.direct or self._kind == "r2c":
            self._stage_out = torch.empty(self._out_layout.shape_p, dtype=self.output.data.dtype, device=dev)
        return Result.SUCCESS

    def compute_initialize(self):
        err = _require_cuda(self, self.input)
        if err:
            return err
        ctx = Context.get(self.input.device)
        handle = ctypes.c_void_p()
        result = _call("b200_fft_plan_c2c", ctx.handle, self._n, self._batch, ctypes.byref(handle))
        if result == Result.SUCCESS:
            self._plan_handle = handle
        return result

    def compute_submit(self, stream):
        if self._plan_handle is None:
            result = self.compute_initialize()
            if result != Result.SUCCESS:
                return result
        forward = 1 if self.config["forward"] else 0
        if self._kind == "c2c" and self._layout.direct:
            return _call("b200_fft_exec", self._plan_handle, self.input.ptr(), self.output.ptr(), forward, stream)
        ctx = Context.get(self.input.device)
        vp = lambda tensor: ctypes.c_void_p(tensor.data_ptr())

        def ok(result):
            return result == Result.SUCCESS

        # 1. dense [batch, n] input
        src = self.input.data
        if not self._layout.direct:
            if not ok(self._layout.gather(ctx, self.input.data, self._stage_in, stream)):
                return Result.ERROR
            src = self._stage_in
        # 1b. real forward transforms of even length: ONE complex transform of half the length on the row itself (viewed
        # as even/odd-packed CF32) + one unpack kernel (b200_fft_exec_real) instead of cast -> full C2C -> pack
        if self._kind != "c2c" and forward and self._n >= 4 and self._n % 2 == 0 and src.data_ptr() % 8 == 0 and \
                os.environ.get("B200_FFT_REAL_HALF", "1") != "0":
            if self._half_plan is None:
                handle = ctypes.c_void_p()
                if not ok(_call("b200_fft_plan_c2c", ctx.handle, self._n // 2, self._batch, ctypes.byref(handle))):
                    return Result.ERROR
                self._half_plan = handle
            dense_out = self.output.data if self._out_layout.direct else self._stage_out
            layout = 0 if self._kind == "r2c" else 1
            if not ok(_call("b200_fft_exec_real", self._half_plan, vp(src), vp(dense_out), layout, stream)):
                return Result.ERROR
            if self._out_layout.direct:
                return Result.SUCCESS
            return self._out_layout.scatter(ctx, dense_out, self.output.data, self._out_layout.shape_p, stream)
        # 2. to the complex work buffer
        if self._kind == "c2c":
            work_in = src
        elif self._kind == "fftpack" and not forward:
            if not ok(_call("b200_fft_real_helper", ctx.handle, 2, vp(src), vp(self._work), self._batch, self._n, stream)):
                return Result.ERROR
            work_in = self._work
        else:
            if not ok(_call("b200_cast_f32_cf32", ctx.handle, vp(src), vp(self._work), self._batch * self._n, stream)):
                return Result.ERROR
            work_in = self._work
        # 3. transform (into the work buffer; a C2C with a dense output layout writes the output directly)
        if self._kind == "c2c" and self._out_layout.direct:
            return _call("b200_fft_exec", self._plan_handle, vp(work_in), self.output.ptr(), forward, stream)
        if not ok(_call("b200_fft_exec", self._plan_handle, vp(work_in), vp(self._work), forward, stream)):
            return Result.ERROR
        # 4. to the dense output layout
        dense_out = self.output.data if self._out_layout.direct else self._stage_out
        if self._kind == "c2c":
            dense_out = self._work
        elif self._kind == "r2c":
            dense_out = self._stage_out if not self._out_layout.direct else self.output.data
            if not ok(_call("b200_fft_real_helper", ctx.handle, 0, vp(self._work), vp(dense_out), self._batch, self._n,
                            stream)):
                return Result.ERROR
        else:
            op = 1 if forward else 3
            if not ok(_call("b200_fft_real_helper", ctx.handle, op, vp(self._work), vp(dense_out), self._batch, self._n,
                            stream)):
                return Result.ERROR
        if self._out_layout.direct:
            return Result.SUCCESS
        return self._out_layout.scatter(ctx, dense_out, self.output.data, self._out_layout.shape_p, stream)

    def compute_deinitialize(self):
        if self._plan_handle is not None:
            _call("b200_fft_plan_destroy", self._plan_handle)
            self._plan_handle = None
        if self._half_plan is not None:
            _call("b200_fft_plan_destroy", self._half_plan)
            self._half_plan = None
        return Result.SUCCESS

    def destroy(self):
        return self.compute_deinitialize()


@register_module
class Agc(Module):
    """`agc` — tiled RMS automatic gain control (include/jetstream/domains/dsp/agc/module.hh:8-19,
    src/domains/dsp/agc/module_impl.cc:7-87, module_impl_native_cpu.cc:76-160). F32 or CF32, any sample axis
    (other layouts are gathered to [lanes, samples] and scattered back). Stateless."""
    TYPE = "agc"
    DEFAULTS = {"tileSize": 1024, "reference": 1.0, "epsilon": 1e-12, "minGain": 0.01, "maxGain": 100.0,
                "maxGainChange": 4.0}

    def validate(self):
        c = self.config
        if int(c["tileSize"]) == 0:
            return _error("[MODULE_AGC] Tile size must be greater than zero.")
        if not math.isfinite(float(c["reference"])) or float(c["reference"]) <= 0.0:
            return _error("[MODULE_AGC] Reference must be finite and positive.")
        if not math.isfinite(float(c["epsilon"])) or float(c["epsilon"]) <= 0.0:
            return _error("[MODULE_AGC] Epsilon must be finite and positive.")
        if not math.isfinite(float(c["minGain"])) or float(c["minGain"]) <= 0.0:
            return _error("[MODULE_AGC] Minimum gain must be finite and positive.")
        if not math.isfinite(float(c["maxGain"])) or float(c["maxGain"]) < float(c["minGain"]):
            return _error("[MODULE_AGC] Maximum gain must be finite and no less than minimum gain.")
        if not math.isfinite(float(c["maxGainChange"])) or float(c["maxGainChange"]) < 1.0:
            return _error("[MODULE_AGC] Maximum gain change must be finite and at least one.")
        self._axis = None
        link = self.inputs.get("signal")
        if link is None or not link.resolved() or link.tensor.size == 0:
            return Result.SUCCESS
        t = link.tensor
        axes = resolve_signal_axes(t)
        if axes is None:
            return _error("[MODULE_AGC] Input must contain valid signal axis metadata.")
        if t.dtype not in ("F32", "CF32"):
            return _error(f"[MODULE_AGC_B200] Unsupported data type '{t.dtype}'.")
        self._axis = axes.sample
        return Result.SUCCESS

    def define(self):
        self.define_taint(Taint.STATELESS)
        self.define_interface_input("signal")
        return self.define_interface_output("signal")

    def create_impl(self):
        self.input = self.inputs["signal"].tensor
        t = self.input
        if self._axis is None:
            self._axis = _empty_input_axis(t)
        self._samples = t.shape[self._axis] if t.rank else 0
        self._lanes = t.size // self._samples if self._samples else 0
        self.output = Tensor.create(t.device, t.dtype, t.shape)
        self.output.propagate_attributes(t)
        self.outputs["signal"] = TensorLink()
        self.outputs["signal"].produced(self.name, "signal", self.output)
        self._layout = _Layout(t.data, self._axis)
        self._out_layout = _Layout(self.output.data, self._axis)
        self._stage_in = None if self._layout.direct else torch.empty(self._layout.shape_p, dtype=t.data.dtype,
                                                                      device=t.device)
        self._stage_out = None if self._out_layout.direct else torch.empty(self._out_layout.shape_p,
                                                                           dtype=t.data.dtype, device=t.device)
        self._scratch = None
        return Result.SUCCESS

    def compute_submit(self, stream):
        if self.input.size == 0:
            return Result.SUCCESS
        err = _require_cuda(self, self.input, self.output)
        if err:
            return err
        c = self.config
        ctx = Context.get(self.input.device)
        tile = int(c["tileSize"])
        need = ctypes.c_uint64()
        result = _call("b200_agc_scratch_bytes", self._lanes, self._samples, tile, ctypes.byref(need))
        if result != Result.SUCCESS:
            return result
        if self._scratch is None or self._scratch.numel() < need.value:
            self._scratch = torch.empty(max(16, need.value), dtype=torch.uint8, device=self.input.device)
        src = self.input.data
        if not self._layout.direct:
            if self._layout.gather(ctx, src, self._stage_in, stream) != Result.SUCCESS:
                return Result.ERROR
            src = self._stage_in
        dst = self.output.data if self._out_layout.direct else self._stage_out
        vp = lambda tensor: ctypes.c_void_p(tensor.data_ptr())
        result = _call("b200_agc", ctx.handle, vp(src), vp(dst), 1 if self.input.dtype == "CF32" else 0, self._lanes,
                       self._samples, tile, float(c["reference"]), float(c["epsilon"]), float(c["minGain"]),
                       float(c["maxGain"]), float(c["maxGainChange"]), vp(self._scratch), stream)
        if result != Result.SUCCESS or self._out_layout.direct:
            return result
        return self._out_layout.scatter(ctx, dst, self.output.data, self._out_layout.shape_p, stream)


def amplitude_scaling_coeff(n: int) -> float:
    """scalingCoeff = 20 * log10f(1 / (F32)N) (src/domains/dsp/amplitude/module_impl.cc:49-51). Evaluated by
    the library with the host libm's log10f — the same function the reference calls — because other F32
    log10 implementations (e.g. numpy's) differ in the last bit for non-power-of-two N."""
    out = ctypes.c_float()
    _native.check(_native.load().b200_amplitude_scaling_coeff(int(n), ctypes.byref(out)))
    return float(out.value)


@register_module
class Amplitude(Module):
    """`amplitude` — src/domains/dsp/amplitude/module_impl.cc:8-66 (+ native_cpu.cc:73-99)."""
    TYPE = "amplitude"

    def validate(self):
        self._norm = 1
        link = self.inputs.get("signal")
        if link is None or not link.resolved():
            return Result.SUCCESS
        t = link.tensor
        if t.size == 0:
            return Result.SUCCESS
        if t.dtype not in ("F32", "CF32"):
            return _error(f"[MODULE_AMPLITUDE_B200] Unsupported input data type: {t.dtype}.")
        sample = t.attribute("sampleAxis") if t.has_attribute("sampleAxis") else (0 if t.rank == 1 else None)
        channel = t.attribute("channelAxis") if t.has_attribute("channelAxis") else None
        if sample is None and channel is None:
            return _error("[MODULE_AMPLITUDE] Input must contain sampleAxis or channelAxis metadata.")
        if sample is not None:
            if not (0 <= sample < t.rank):
                return _error("[MODULE_AMPLITUDE] Input must contain valid signal axis metadata.")
            self._norm = t.shape[sample]
        return Result.SUCCESS

    def define(self):
        self.define_taint(Taint.DISCONTIGUOUS | Taint.STATELESS)
        self.define_interface_input("signal")
        return self.define_interface_output("signal")

    def create_impl(self):
        self.input = self.inputs["signal"].tensor
        self._layout = _Layout(self.input.data, None)
        self._layout.perm = list(range(self.input.rank))       # keep the axis order, only densify
        self._layout.shape_p = list(self.input.shape)
        self._layout.direct = self.input.contiguous()
        self._staging = None if self._layout.direct else torch.empty(self.input.shape, dtype=self.input.data.dtype,
                                                                      device=self.input.device)
        self.scaling_coeff = amplitude_scaling_coeff(self._norm)
        self.output = Tensor.create(self.input.device, "F32", self.input.shape)
        self.output.propagate_attributes(self.input)
        self.outputs["signal"] = TensorLink()
        self.outputs["signal"].produced(self.name, "signal", self.output)
        return Result.SUCCESS

    def compute_submit(self, stream):
        err = _require_cuda(self, self.input, self.output)
        if err:
            return err
        ctx = Context.get(self.input.device)
        fn = "b200_amplitude_cf32" if self.input.dtype == "CF32" else "b200_amplitude_f32"
        src = self.input.ptr()
        if not self._layout.direct:
            result = self._layout.gather(ctx, self.input.data, self._staging, stream)
            if result != Result.SUCCESS:
                return result
            src = ctypes.c_void_p(self._staging.data_ptr())
        return _call(fn, ctx.handle, src, self.output.ptr(), self.input.size,
                     ctypes.c_float(self.scaling_coeff), stream)


def range_coefficients(lo: float, hi: float) -> Tuple[float, float]:
    """RangeImpl::updateCoefficients (src/domains/core/range/module_impl.cc:51-63), F32 arithmetic."""
    scale, offset = ctypes.c_float(), ctypes.c_float()
    _native.check(_native.load().b200_range_coefficients(ctypes.c_float(lo), ctypes.c_float(hi),
                                                         ctypes.byref(scale), ctypes.byref(offset)))
    return float(scale.value), float(offset.value)


@register_module
class Range(Module):
    """`range` (the chain's "Scale") — src/domains/core/range/module_impl.cc:7-63 (+ native_cpu.cc:67-82)."""
    TYPE = "range"
    DEFAULTS = {"min": -1.0, "max": 1.0}

    def validate(self):
        link = self.inputs.get("signal")
        if link is None or not link.resolved() or link.tensor.size == 0:
            return Result.SUCCESS
        if link.tensor.dtype != "F32":
            return _error(f"[MODULE_RANGE_B200] Unsupported data type '{link.tensor.dtype}'.")
        return Result.SUCCESS

    def define(self):
        self.define_taint(Taint.DISCONTIGUOUS | Taint.STATELESS)
        self.define_interface_output("signal")
        return self.define_interface_input("signal")

    def create_impl(self):
        self.input = self.inputs["signal"].tensor
        self._layout = _Layout(self.input.data, None)
        self._layout.perm = list(range(self.input.rank))
        self._layout.shape_p = list(self.input.shape)
        self._layout.direct = self.input.contiguous()
        self._staging = None if self._layout.direct else torch.empty(self.input.shape, dtype=torch.float32,
                                                                      device=self.input.device)
        self.scale, self.offset = range_coefficients(float(self.config["min"]), float(self.config["max"]))
        self.output = Tensor.create(self.input.device, "F32", self.input.shape)
        self.output.propagate_attributes(self.input)
        self.outputs["signal"] = TensorLink()
        self.outputs["signal"].produced(self.name, "signal", self.output)
        return Result.SUCCESS

    def reconfigure_impl(self, candidate):  # in-place (module_impl.cc:40-49)
        self.scale, self.offset = range_coefficients(float(candidate["min"]), float(candidate["max"]))
        return Result.SUCCESS

    def compute_submit(self, stream):
        err = _require_cuda(self, self.input, self.output)
        if err:
            return err
        ctx = Context.get(self.input.device)
        src = self.input.ptr()
        if not self._layout.direct:
            result = self._layout.gather(ctx, self.input.data, self._staging, stream)
            if result != Result.SUCCESS:
                return result
            src = ctypes.c_void_p(self._staging.data_ptr())
        return _call("b200_range_f32", ctx.handle, src, self.output.ptr(), self.input.size,
                     ctypes.c_float(self.scale), ctypes.c_float(self.offset), stream)


@register_module
class SpectralChain(Module):
    """`spectral_chain` — B200-only fused module: multiply(window) -> fft -> amplitude -> [range] in one
    kernel (b200_chain_exec). It is what the `spectrum_engine` block creates on this provider instead of
    the 9-module chain of src/domains/dsp/spectrum_engine/block_impl.cc:120-217. Inputs: `buffer`
    (CF32, sample axis innermost) and `window` (CF32 [n], the settled window->invert output).
    `buffer` may also be a complex-integer tensor (CI8 ... CU32): the `cast` module an SDR flowgraph puts in front of
    spectrum_engine is then folded into the kernel's load (b200_chain_exec_typed)."""
    TYPE = "spectral_chain"
    DEFAULTS = {"enableScale": False, "rangeMin": -120.0, "rangeMax": 0.0,
                # enableAgc: spectrum_engine's optional agc stage (one RMS tile per spectrum) inside the same kernel;
                # the four numbers are the agc module's config (include/jetstream/domains/dsp/agc/module.hh:9-14)
                "enableAgc": False, "agcReference": 1.0, "agcEpsilon": 1e-12, "agcMinGain": 0.01, "agcMaxGain": 100.0,
                # publishColumnSums: the kernel's epilogue also accumulates sum-over-batch of every output column and
                # the output tensor carries them as attribute "b200.columnSums" (Tensor [n] F32): a downstream `lineplot`
                # on this provider then skips its pass over the [batch, n] spectra (b200_chain_exec_colsum).
                "publishColumnSums": False}
    COLUMN_SUMS_ATTRIBUTE = "b200.columnSums"

    def __init__(self):
        super().__init__()
        self._plan_handle = None
        self.colsum = None

    def validate(self):
        link = self.inputs.get("buffer")
        if link is None or not link.resolved() or link.tensor.size == 0:
            return Result.SUCCESS
        t = link.tensor
        if t.dtype != "CF32" and t.dtype not in COMPLEX_INT_DTYPES:
            return _error("[MODULE_SPECTRAL_CHAIN_B200] Input must have data type CF32 or a complex integer type.")
        axes = resolve_signal_axes(t)
        if axes is None:
            return _error("[MODULE_SPECTRAL_CHAIN_B200] Input signal axis metadata is invalid.")
        if axes.sample != t.rank - 1:
            return _error("[MODULE_SPECTRAL_CHAIN_B200] The sample axis must be the innermost axis.")
        if self.config["enableAgc"] and t.shape[-1] != 4096:
            return _error("[MODULE_SPECTRAL_CHAIN_B200] The fused AGC stage exists for 4096-point spectra only; "
                          "wire fft -> agc -> amplitude for other lengths.")
        return Result.SUCCESS

    def define(self):
        self.define_taint(Taint.STATELESS)
        self.define_interface_input("buffer")
        self.define_interface_input("window")
        return self.define_interface_output("buffer")

    def create_impl(self):
        self.input = self.inputs["buffer"].tensor
        self.window = self.inputs["window"].tensor
        self._n = self.input.shape[-1]
        if self.window.dtype != "CF32" or self.window.size != self._n:
            return _error("[MODULE_SPECTRAL_CHAIN_B200] Window must be CF32 with one tap per sample.")
        self._batch = self.input.size // self._n
        self.amp_coeff = amplitude_scaling_coeff(self._n)
        self.scale, self.offset = range_coefficients(float(self.config["rangeMin"]), float(self.config["rangeMax"]))
        self.output = Tensor.create(self.input.device, "F32", self.input.shape)
        self.output.propagate_attributes(self.input)
        self.colsum = None
        if self.config["publishColumnSums"] and not self.config["enableAgc"]:
            self.colsum = Tensor.create(self.input.device, "F32", (self._n,))
            self.output.set_attribute(self.COLUMN_SUMS_ATTRIBUTE, self.colsum)
        self.outputs["buffer"] = TensorLink()
        self.outputs["buffer"].produced(self.name, "buffer", self.output)
        return Result.SUCCESS

    def reconfigure_impl(self, candidate):
        if bool(candidate["enableScale"]) != bool(self.config["enableScale"]) or \
                bool(candidate["enableAgc"]) != bool(self.config["enableAgc"]) or \
                bool(candidate["publishColumnSums"]) != bool(self.config["publishColumnSums"]):
            return Result.RECREATE
        self.scale, self.offset = range_coefficients(float(candidate["rangeMin"]), float(candidate["rangeMax"]))
        return Result.SUCCESS

    def _create_plan(self):
        # The window input is a settled STATIC_OUTPUT tensor produced earlier in this same cycle on this
        # same stream (window -> invert run before us): wait for it, then let the plan capture it.
        err = _require_cuda(self, self.input, self.window)
        if err:
            return err
        ctx = Context.get(self.input.device)
        handle = ctypes.c_void_p()
        torch.cuda.current_stream(self.input.device).synchronize()
        result = _call("b200_chain_plan_create", ctx.handle, self._n, self._batch, self.window.ptr(),
                       ctypes.byref(handle))
        if result == Result.SUCCESS:
            self._plan_handle = handle
        return result

    def compute_submit(self, stream):
        if self._plan_handle is None:
            result = self._create_plan()
            if result != Result.SUCCESS:
                return result
        c = self.config
        if c["enableAgc"]:
            return _call("b200_chain_exec_agc", self._plan_handle, self.input.ptr(), DTYPE_CODES[self.input.dtype],
                         self.output.ptr(), self._batch, ctypes.c_float(self.amp_coeff), 1 if c["enableScale"] else 0,
                         ctypes.c_float(self.scale), ctypes.c_float(self.offset), float(c["agcReference"]),
                         float(c["agcEpsilon"]), float(c["agcMinGain"]), float(c["agcMaxGain"]), stream)
        if self.colsum is not None:
            return _call("b200_chain_exec_colsum", self._plan_handle, self.input.ptr(), DTYPE_CODES[self.input.dtype],
                         self.output.ptr(), self._batch, ctypes.c_float(self.amp_coeff),
                         1 if c["enableScale"] else 0, ctypes.c_float(self.scale), ctypes.c_float(self.offset),
                         self.colsum.ptr(), stream)
        return _call("b200_chain_exec_typed", self._plan_handle, self.input.ptr(), DTYPE_CODES[self.input.dtype],
                     self.output.ptr(), self._batch, ctypes.c_float(self.amp_coeff),
                     1 if c["enableScale"] else 0, ctypes.c_float(self.scale), ctypes.c_float(self.offset),
                     stream)

    def compute_deinitialize(self):
        if self._plan_handle is not None:
            _call("b200_chain_plan_destroy", self._plan_handle)
            self._plan_handle = None
        return Result.SUCCESS

    def destroy(self):
        return self.compute_deinitialize()


@register_module
class FilterTaps(Module):
    """`filter_taps` — src/domains/dsp/filter_taps/module_impl.cc:12-140 (+ native_cpu.cc:46-80). STATIC_OUTPUT:
    evaluated once, on the host in F64 (b200_filter_taps_host), uploaded on the first compute cycle."""
    TYPE = "filter_taps"
    DEFAULTS = {"sampleRate": 2e6, "bandwidth": 1e6, "center": (0.0,), "taps": 101}

    def validate(self):
        c = self.config
        centers = [float(v) for v in c["center"]]
        taps = int(c["taps"])
        sr, bw = float(c["sampleRate"]), float(c["bandwidth"])
        if not math.isfinite(sr) or sr <= 0.0:
            return _error(f"[MODULE_FILTER_TAPS] Sample rate must be positive ({sr}).")
        if not math.isfinite(bw) or bw <= 0.0 or bw > sr:
            return _error("[MODULE_FILTER_TAPS] Bandwidth must be between 0 and sample rate.")
        if taps == 0:
            return _error("[MODULE_FILTER_TAPS] Number of taps cannot be zero.")
        if taps % 2 == 0:
            return _error(f"[MODULE_FILTER_TAPS] Number of taps must be odd ({taps}).")
        if not centers:
            return _error("[MODULE_FILTER_TAPS] At least one center frequency is required.")
        for i, ct in enumerate(centers):
            if not math.isfinite(ct) or abs(ct) > sr / 2.0:
                return _error(f"[MODULE_FILTER_TAPS] Center frequency #{i} must be within +-sampleRate/2.")
        return Result.SUCCESS

    def define(self):
        self.define_taint(Taint.STATIC_OUTPUT)
        return self.define_interface_output("coeffs")

    def create_impl(self):
        c = self.config
        centers = [float(v) for v in c["center"]]
        heads, taps = len(centers), int(c["taps"])
        host = np.zeros((heads, taps), dtype=np.complex64)
        arr = (ctypes.c_double * heads)(*centers)
        result = _call("b200_filter_taps_host", ctypes.c_double(float(c["sampleRate"])),
                       ctypes.c_double(float(c["bandwidth"])), arr, heads, taps,
                       host.ctypes.data_as(ctypes.c_void_p))
        if result != Result.SUCCESS:
            return result
        self.host_coeffs = host
        self.output = Tensor.create(self.alloc_device, "CF32", (heads, taps))
        self.output.set_attribute("sampleAxis", 1)
        self.output.set_attribute("channelAxis", 0)
        self.output.set_attribute("sampleRate", float(np.float32(c["sampleRate"])))
        self.output.set_attribute("bandwidth", float(np.float32(c["bandwidth"])))
        self.outputs["coeffs"] = TensorLink()
        self.outputs["coeffs"].produced(self.name, "coeffs", self.output)
        return Result.SUCCESS

    def compute_submit(self, stream):
        err = _require_cuda(self, self.output)
        if err:
            return err
        self.output.data.copy_(torch.from_numpy(self.host_coeffs), non_blocking=False)
        return Result.SUCCESS


def filter_resample_plan(sample_rate: float, bandwidth: float, taps: int, signal_size: int):
    """CalculateCandidatePlan (src/domains/dsp/filter/block_impl.cc:40-168) for zero-centred heads: returns the
    integer resampler ratio R (1 = the block silently runs at full rate)."""
    if taps == 0:
        return 1
    ratio = float(sample_rate) / float(bandwidth)
    if not math.isfinite(ratio) or ratio <= 0.0 or ratio >= 2.0 ** 64:
        return 1
    if ratio != math.floor(ratio):
        return 1
    r = int(ratio)
    if (taps - 1) % r != 0:
        return 1
    if (taps + signal_size - 1) % r != 0:
        return 1
    return r


@register_module
class FirFilter(Module):
    """`fir_filter` — B200-only fused module: the per-cycle module chain of the `filter` block
    (src/domains/dsp/filter/block_impl.cc:350-582) as one streaming time-domain (decimating) FIR kernel.
    Inputs: `signal` CF32 [T] or [B, T], `coeffs` CF32 [heads, taps] (settled).
    Output `buffer`: [B, heads, T / R] with channelAxis = old sample axis, sampleAxis = +1.
    With batchAxis = 0 the B rows are consecutive frames of ONE stream (overlap_add carries frame k's tail into frame
    k+1, src/domains/dsp/overlap_add/module_impl_native_cpu.cc:155-174); WITHOUT a batchAxis they are B independent
    lanes, each carrying its own tail across cycles (:176-198) — one plan (history) per lane."""
    TYPE = "fir_filter"
    DEFAULTS = {"decimation": 1, "centerBins": None}
    MAX_LANES = 256

    def __init__(self):
        super().__init__()
        self._plan_handle = None
        self._lane_plans: List[ctypes.c_void_p] = []

    def validate(self):
        link = self.inputs.get("signal")
        if link is None or not link.resolved() or link.tensor.size == 0:
            return Result.SUCCESS
        t = link.tensor
        if t.dtype != "CF32":
            return _error("[MODULE_FIR_FILTER_B200] Signal input must be CF32.")
        axes = resolve_signal_axes(t)
        if axes is None:
            return _error("[BLOCK_FILTER] Signal axis metadata is invalid.")
        if axes.channel is not None:
            return _error("[BLOCK_FILTER] Signal already has channelAxis. Generated filter channels cannot be nested.")
        if not (t.rank == 1 or (t.rank == 2 and axes.sample == 1 and axes.batch in (0, None))):
            return _error("[MODULE_FIR_FILTER_B200] Supported layouts: [T] or [batch, T] with the sample axis innermost.")
        r = int(self.config["decimation"])
        if r < 1 or t.shape[-1] % r != 0:
            return _error("[MODULE_FIR_FILTER_B200] Frame length must be a multiple of the decimation.")
        if t.rank == 2 and axes.batch is None and t.shape[0] > self.MAX_LANES:
            return _error(f"[MODULE_FIR_FILTER_B200] At most {self.MAX_LANES} independent lanes (rows without a batchAxis).")
        self._lane_mode = t.rank == 2 and axes.batch is None
        return Result.SUCCESS

    def define(self):
        self.define_interface_input("signal")
        self.define_interface_input("coeffs")
        return self.define_interface_output("buffer")

    def create_impl(self):
        self.input = self.inputs["signal"].tensor
        self.coeffs = self.inputs["coeffs"].tensor
        if self.coeffs.dtype != "CF32" or self.coeffs.rank != 2:
            return _error("[MODULE_FIR_FILTER_B200] Coefficients must be CF32 [heads, taps].")
        self._heads, self._taps = self.coeffs.shape
        self._r = int(self.config["decimation"])
        t = self.input
        self._frame_len = t.shape[-1]
        rows = t.size // self._frame_len
        self._lanes = rows if getattr(self, "_lane_mode", False) else 1
        self._frames = 1 if self._lanes > 1 else rows
        out_shape = tuple(t.shape[:-1]) + (self._heads, self._frame_len // self._r)
        self.output = Tensor.create(t.device, "CF32", out_shape)
        self.output.propagate_attributes(t)
        sample_axis = t.rank - 1
        self.output.set_attribute("sampleAxis", sample_axis + 1)
        self.output.set_attribute("channelAxis", sample_axis)
        if t.has_attribute("batchAxis"):
            b = int(t.attribute("batchAxis"))
            self.output.set_attribute("batchAxis", b + 1 if b >= sample_axis else b)
        self.outputs["buffer"] = TensorLink()
        self.outputs["buffer"].produced(self.name, "buffer", self.output)
        return Result.SUCCESS

    def _create_plan(self):
        err = _require_cuda(self, self.input, self.coeffs)
        if err:
            return err
        ctx = Context.get(self.input.device)
        torch.cuda.current_stream(self.input.device).synchronize()   # coefficients are settled static output
        host = np.ascontiguousarray(self.coeffs.numpy())
        bins = self.config.get("centerBins")
        translate = bins is not None and any(int(b) != 0 for b in bins)
        self._lane_plans = []
        for _ in range(self._lanes):
            handle = ctypes.c_void_p()
            result = _call("b200_fir_plan_create", ctx.handle, host.ctypes.data_as(ctypes.c_void_p), self._taps,
                           self._heads, self._r, ctypes.byref(handle))
            if result != Result.SUCCESS:
                return result
            self._lane_plans.append(handle)
            if translate:
                arr = (ctypes.c_int64 * self._heads)(*[int(b) for b in bins])
                result = _call("b200_fir_plan_set_translation", handle, self._frame_len, arr)
                if result != Result.SUCCESS:
                    return result
        self._plan_handle = self._lane_plans[0]
        return Result.SUCCESS

    def compute_submit(self, stream):
        if self._plan_handle is None:
            result = self._create_plan()
            if result != Result.SUCCESS:
                return result
        pending = getattr(self, "_pending_history", None)
        if pending is not None:
            tail, frames_before = pending
            self._pending_history = None
            result = _call("b200_fir_set_history", self._plan_handle, ctypes.c_void_p(tail.data_ptr()), tail.numel(),
                           frames_before, stream)
            if result != Result.SUCCESS:
                return result
        if self._lanes == 1:
            return _call("b200_fir_exec", self._plan_handle, self.input.ptr(), self.output.ptr(), self._frames,
                         self._frame_len, stream)
        x0, y0 = self.input.ptr().value, self.output.ptr().value
        out_row = self._heads * (self._frame_len // self._r) * 8
        for lane, handle in enumerate(self._lane_plans):
            result = _call("b200_fir_exec", handle, ctypes.c_void_p(x0 + lane * self._frame_len * 8),
                           ctypes.c_void_p(y0 + lane * out_row), 1, self._frame_len, stream)
            if result != Result.SUCCESS:
                return result
        return Result.SUCCESS

    def set_history(self, tail: torch.Tensor, frames_before: int = 0) -> Result:
        """Time sharding (sharding.exchange_fir_halo): the last `taps - 1` input samples that precede this module's
        slab of the stream (CF32 device tensor, possibly shorter) become the carried filter state of the NEXT compute
        cycle (the plan itself is created on the first cycle, once the static taps have settled)."""
        tail = tail.contiguous()
        if tail.dtype != torch.complex64 or tail.device.type != "cuda":
            return _error("[MODULE_FIR_FILTER_B200] The halo must be a CF32 CUDA tensor.")
        self._pending_history = (tail, int(frames_before))
        return Result.SUCCESS

    def compute_deinitialize(self):
        for handle in self._lane_plans:
            _call("b200_fir_plan_destroy", handle)
        self._lane_plans = []
        self._plan_handle = None
        return Result.SUCCESS

    def destroy(self):
        return self.compute_deinitialize()


@register_module
class Fm(Module):
    """`fm` — src/domains/dsp/fm/module_impl.cc:8-155 (+ native_cpu.cc:43-175): narrow (mono) and wide (stereo
    multiplex decoder) modes, optional 50/75 us de-emphasis, per-lane state carried across cycles."""
    TYPE = "fm"
    DEFAULTS = {"mode": "narrow", "deemphasis": "none", "sampleRate": 240e3}

    def __init__(self):
        super().__init__()
        self._plan_handle = None

    def validate(self):
        c = self.config
        if c["mode"] not in ("narrow", "wide"):
            return _error("[MODULE_FM] Mode must be 'narrow' or 'wide'.")
        if c["deemphasis"] not in ("none", "50us", "75us"):
            return _error("[MODULE_FM] De-emphasis must be 'none', '50us', or '75us'.")
        sr = float(c["sampleRate"])
        if not math.isfinite(sr) or sr <= 0.0:
            return _error("[MODULE_FM] Sample rate must be finite and positive.")
        if sr > 20e6:
            return _error("[MODULE_FM] Sample rate must not exceed 20 MHz.")
        if c["mode"] == "wide" and sr < 200e3:
            return _error("[MODULE_FM] Wideband mode requires a sample rate of at least 200 kHz.")
        self._axes = None
        link = self.inputs.get("signal")
        if link is None or not link.resolved() or link.tensor.size == 0:
            return Result.SUCCESS
        t = link.tensor
        if t.dtype != "CF32":
            return _error("[MODULE_FM_B200] Input must be complex (CF32).")
        axes = resolve_signal_axes(t)
        if axes is None:
            return _error("[MODULE_FM] Input must contain valid signal axis metadata.")
        if c["mode"] == "wide" and axes.channel is not None:
            return _error("[MODULE_FM] Wideband mode does not support channelized input.")
        if axes.sample != t.rank - 1 or (axes.batch not in (None, 0)):
            return _error("[MODULE_FM_B200] Supported layouts: sample axis innermost, batch axis outermost.")
        self._axes = axes
        return Result.SUCCESS

    def define(self):
        self.define_interface_input("signal")
        return self.define_interface_output("signal")

    def create_impl(self):
        self.input = self.inputs["signal"].tensor
        if not self.input.contiguous():
            return _error("[MODULE_FM_B200] Strided inputs are not supported by this provider yet.")
        t = self.input
        if self._axes is None:      # empty input
            self._axes = resolve_signal_axes(t) or SignalAxes()
        self._frame_len = t.shape[-1] if t.rank else 0
        self._frames = t.shape[0] if (self._axes.batch == 0 and t.rank >= 2) else 1
        self._lanes = t.size // (self._frame_len * self._frames) if t.size else 0
        wide = self.config["mode"] == "wide"
        self.output = Tensor.create(t.device, "F32", tuple(t.shape) + ((2,) if wide else ()))
        self.output.propagate_attributes(t)
        if wide:      # trailing [left, right] axis becomes the channel axis (fm/module_impl.cc:96-101)
            self.output.set_attribute("channelAxis", t.rank)
        self.output.set_attribute("frequency", 0.0)
        self.outputs["signal"] = TensorLink()
        self.outputs["signal"].produced(self.name, "signal", self.output)
        return Result.SUCCESS

    def compute_initialize(self):
        if self.input.device.type != "cuda" or self.input.size == 0:
            return Result.SUCCESS
        ctx = Context.get(self.input.device)
        handle = ctypes.c_void_p()
        deemph = {"none": 0, "50us": 50, "75us": 75}[self.config["deemphasis"]]
        result = _call("b200_fm_plan_create", ctx.handle, self._lanes, ctypes.c_float(float(self.config["sampleRate"])),
                       1 if self.config["mode"] == "wide" else 0, deemph, ctypes.byref(handle))
        if result == Result.SUCCESS:
            self._plan_handle = handle
        return result

    def compute_submit(self, stream):
        if self.input.size == 0:
            return Result.SUCCESS
        err = _require_cuda(self, self.input, self.output)
        if err:
            return err
        if self._plan_handle is None:
            result = self.compute_initialize()
            if result != Result.SUCCESS:
                return result
        return _call("b200_fm_exec", self._plan_handle, self.input.ptr(), self.output.ptr(), self._frames,
                     self._frame_len, stream)

    def compute_deinitialize(self):
        if self._plan_handle is not None:
            _call("b200_fm_plan_destroy", self._plan_handle)
            self._plan_handle = None
        return Result.SUCCESS

    def destroy(self):
        return self.compute_deinitialize()


def _element_layout(tensor: Tensor, tag: str):
    """Geometry shared by lineplot and waterfall (lineplot/module_impl.cc:95-187, waterfall/module_impl.cc:30-83): the
    element axis is sampleAxis or channelAxis (not both), every other dimension must be the batch axis.
    Returns (extent, batches, element_stride, batch_stride) in elements, or a Result on error."""
    axes = SignalAxes()
    for name in ("sample", "batch", "channel"):
        key = name + "Axis"
        if tensor.has_attribute(key):
            value = tensor.attribute(key)
            if not isinstance(value, int) or isinstance(value, bool) or value < 0 or value >= tensor.rank:
                return _error(f"[{tag}] Input must contain valid signal axis metadata.")
            setattr(axes, name, value)
    present = [v for v in (axes.sample, axes.batch, axes.channel) if v is not None]
    if len(set(present)) != len(present):
        return _error(f"[{tag}] Input must contain valid signal axis metadata.")
    if axes.sample is not None and axes.channel is not None:
        return _error(f"[{tag}] Input cannot contain both sampleAxis and channelAxis.")
    element = axes.sample if axes.sample is not None else axes.channel
    if element is None and tensor.rank == 1:
        element = 0
    if element is None:
        return _error(f"[{tag}] Input must contain sampleAxis or channelAxis.")
    for axis in range(tensor.rank):
        if axis != element and axis != axes.batch:
            return _error(f"[{tag}] Unsupported auxiliary input axis {axis}. Every dimension must be the element axis "
                          "or batchAxis.")
    strides = tensor.data.stride()
    return (tensor.shape[element], tensor.shape[axes.batch] if axes.batch is not None else 1, strides[element],
            strides[axes.batch] if axes.batch is not None else 0)


@register_module
class Lineplot(Module):
    """`lineplot` compute — src/domains/visualization/lineplot/module_impl.cc:17-187 (validation / geometry) +
    module_impl_native_cpu.cc:80-122 (computeSubmit: batch sum with decimation -> normalise -> clamp -> EMA). The
    render half (present) is out of scope; `signal_points()` is the tensor the reference hands to its renderer. When
    the input carries "b200.columnSums" (a fused spectral_chain upstream) the batch sum is taken from there."""
    TYPE = "lineplot"
    DEFAULTS = {"averaging": 1, "decimation": 1, "numberOfVerticalLines": 11, "numberOfHorizontalLines": 5,
                "thickness": 1.0}

    def validate(self):
        c = self.config
        if int(c["decimation"]) == 0:
            return _error("[MODULE_LINEPLOT] Decimation must be at least 1.")
        if int(c["averaging"]) == 0:
            return _error("[MODULE_LINEPLOT] Averaging must be at least 1.")
        if int(c["numberOfVerticalLines"]) < 2:
            return _error("[MODULE_LINEPLOT] Number of vertical lines must be at least 2.")
        if int(c["numberOfHorizontalLines"]) < 2:
            return _error("[MODULE_LINEPLOT] Number of horizontal lines must be at least 2.")
        if not math.isfinite(float(c["thickness"])) or float(c["thickness"]) <= 0.0:
            return _error("[MODULE_LINEPLOT] Thickness must be finite and positive.")
        self._geometry = None
        link = self.inputs.get("signal")
        if link is None or not link.resolved() or link.tensor.size == 0:
            return Result.SUCCESS
        t = link.tensor
        if t.dtype != "F32":
            return _error(f"[MODULE_LINEPLOT_B200] Unsupported input data type: {t.dtype}.")
        layout = _element_layout(t, "MODULE_LINEPLOT")
        if isinstance(layout, Result):
            return layout
        extent, batches, es, bs = layout
        elements = extent // int(c["decimation"])
        if elements < 2:
            return _error(f"[MODULE_LINEPLOT] Invalid number of elements ({elements}), need at least 2.")
        for key in ("frequency", "sampleRate"):
            if t.has_attribute(key) and not isinstance(t.attribute(key), float):
                return _error(f"[MODULE_LINEPLOT] Input {key} metadata must have type F32.")
        self._geometry = (elements, batches, es, bs)
        return Result.SUCCESS

    def define(self):
        self.define_taint(Taint.SURFACE)
        return self.define_interface_input("signal")

    def create_impl(self):
        if self._geometry is None:
            return _error("[MODULE_LINEPLOT] Input validation plan is unavailable.")
        self.input = self.inputs["signal"].tensor
        self.elements, self.batches, self.element_stride, self.batch_stride = self._geometry
        self.normalization = float(np.float32(1.0) / (np.float32(0.5) * np.float32(self.batches)))
        dev = self.input.device
        self.points = Tensor.create(dev, "F32", (self.elements, 2))
        self.average = Tensor.create(dev, "F32", (self.elements,))
        need = ctypes.c_uint64()
        result = _call("b200_lineplot_scratch_bytes", self.batches, self.elements, int(self.config["decimation"]),
                       ctypes.byref(need))
        if result != Result.SUCCESS:
            return result
        self._scratch = torch.empty(need.value, dtype=torch.uint8, device=dev)
        self._initialized = False
        # the fused producer's column sums are usable when the input is its plain row-major [batches, extent] output
        colsum = self.input.attributes.get(SpectralChain.COLUMN_SUMS_ATTRIBUTE)
        extent = self.elements * int(self.config["decimation"])
        self._colsum = colsum if (isinstance(colsum, Tensor) and self.element_stride == 1 and colsum.size >= extent and
                                  self.batch_stride in (0, colsum.size)) else None
        return Result.SUCCESS

    def reconfigure_impl(self, candidate):     # module_impl.cc:223-234: only `averaging` changes in place
        same = all(candidate[k] == self.config[k] for k in ("decimation", "numberOfVerticalLines",
                                                            "numberOfHorizontalLines", "thickness"))
        return Result.SUCCESS if same else Result.RECREATE

    def compute_submit(self, stream):
        err = _require_cuda(self, self.input, self.points)
        if err:
            return err
        ctx = Context.get(self.input.device)
        if not self._initialized:
            result = _call("b200_lineplot_init", ctx.handle, self.points.ptr(), self.average.ptr(), self.elements, stream)
            if result != Result.SUCCESS:
                return result
            self._initialized = True
        c = self.config
        if self._colsum is not None:
            return _call("b200_lineplot_update_from_colsum", ctx.handle, self._colsum.ptr(), self.elements,
                         int(c["decimation"]), ctypes.c_float(self.normalization), int(c["averaging"]),
                         self.average.ptr(), self.points.ptr(), stream)
        return _call("b200_lineplot_update", ctx.handle, self.input.ptr(), self.batches, self.elements,
                     self.batch_stride, self.element_stride, int(c["decimation"]), ctypes.c_float(self.normalization),
                     int(c["averaging"]), self.average.ptr(), self.points.ptr(),
                     ctypes.c_void_p(self._scratch.data_ptr()), stream)

    def signal_points(self) -> np.ndarray:
        return self.points.numpy()


@register_module
class Waterfall(Module):
    """`waterfall` compute — src/domains/visualization/waterfall/module_impl.cc:15-113 (validation / geometry) +
    module_impl_native_cpu.cc:53-78 and ring_state.hh:18-44 (newest rows into the ring, cursor advance)."""
    TYPE = "waterfall"
    DEFAULTS = {"height": 512, "interpolate": True}

    def validate(self):
        height = int(self.config["height"])
        if height == 0 or height > 2048:
            return _error(f"[MODULE_WATERFALL] Invalid height value '{height}', must be between 1 and 2048.")
        self._geometry = None
        link = self.inputs.get("signal")
        if link is None or not link.resolved() or link.tensor.size == 0:
            return Result.SUCCESS
        t = link.tensor
        if t.dtype != "F32":
            return _error(f"[MODULE_WATERFALL_B200] Unsupported input data type: {t.dtype}.")
        layout = _element_layout(t, "MODULE_WATERFALL")
        if isinstance(layout, Result):
            return layout
        self._geometry = layout
        return Result.SUCCESS

    def define(self):
        self.define_taint(Taint.SURFACE)
        return self.define_interface_input("signal")

    def create_impl(self):
        if self._geometry is None:
            return _error("[MODULE_WATERFALL] Input validation plan is unavailable.")
        self.input = self.inputs["signal"].tensor
        self.elements, self.batches, self.element_stride, self.batch_stride = self._geometry
        self.height = int(self.config["height"])
        self.ring = Tensor.create(self.input.device, "F32", (self.height, self.elements))
        self.write_index = 0
        return Result.SUCCESS

    def reconfigure_impl(self, candidate):     # module_impl.cc:115-124: height -> RECREATE, interpolate in place
        return Result.RECREATE if int(candidate["height"]) != int(self.config["height"]) else Result.SUCCESS

    def compute_submit(self, stream):
        err = _require_cuda(self, self.input, self.ring)
        if err:
            return err
        ctx = Context.get(self.input.device)
        result = _call("b200_waterfall_update", ctx.handle, self.input.ptr(), self.batches, self.elements,
                       self.batch_stride, self.element_stride, self.ring.ptr(), self.height, self.write_index, stream)
        if result != Result.SUCCESS:
            return result
        cursor = ctypes.c_uint64(self.write_index)
        result = _call("b200_waterfall_advance", ctypes.byref(cursor), self.batches, self.height)
        self.write_index = int(cursor.value)
        return result

    def frequency_bins(self) -> np.ndarray:
        return self.ring.numpy()


# ---------------------------------------------------------------------------------------------
# Runtime (per device x runtime segment) — src/runtime/native/cuda/impl.cc
# ---------------------------------------------------------------------------------------------

class NativeCudaRuntime:
    """Owns one non-blocking stream, calls each module's compute_submit(stream) in order, records a
    CUDA-event pair per module, synchronises the stream once per cycle (impl.cc:185-272)."""

    def __init__(self, name: str, device="cuda"):
        self.name = name
        self.device = torch.device(device)
        self.modules: List[Module] = []
        self.stream: Optional[torch.cuda.Stream] = None
        self._events: Dict[str, Tuple[torch.cuda.Event, torch.cuda.Event]] = {}

    def create(self, modules: Sequence[Module]) -> Result:
        self.modules = list(modules)
        if self.device.type == "cuda" and torch.cuda.is_available():
            self.stream = torch.cuda.Stream(self.device)
            for m in self.modules:
                self._events[m.name] = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            with torch.cuda.stream(self.stream):
                for m in self.modules:
                    result = m.compute_initialize()
                    if result not in (Result.SUCCESS, Result.RELOAD):
                        return result
        return Result.SUCCESS

    def destroy(self) -> Result:
        for m in self.modules:
            m.compute_deinitialize()
        self.modules = []
        return Result.SUCCESS

    def compute(self, pending: Sequence[str], skipped: set, failed: set) -> Result:
        if self.stream is None:
            _error("[RUNTIME_NATIVE_CUDA_B200] No CUDA device available; this runtime has no CPU path.")
            failed.update(m.name for m in self.modules)
            return Result.ERROR
        ran: List[Module] = []
        overall = Result.SUCCESS
        stream_ptr = ctypes.c_void_p(self.stream.cuda_stream)
        with torch.cuda.stream(self.stream):
            for m in self.modules:
                if pending and m.name not in pending:
                    continue
                upstream_skipped = any(l.producer and l.producer[0] in skipped for l in m.inputs.values())
                upstream_failed = any(l.producer and l.producer[0] in failed for l in m.inputs.values())
                if upstream_failed:
                    failed.add(m.name)
                    continue
                if upstream_skipped:
                    skipped.add(m.name)
                    continue
                start, end = self._events[m.name]
                start.record(self.stream)
                result = m.compute_submit(stream_ptr)
                end.record(self.stream)
                if result in (Result.SUCCESS, Result.RELOAD):
                    ran.append(m)
                elif result == Result.SKIP:
                    skipped.add(m.name)
                elif result in (Result.YIELD, Result.TIMEOUT):
                    overall = result
                    break
                else:
                    failed.add(m.name)
                    overall = Result.ERROR
        self.stream.synchronize()
        for m in ran:
            start, end = self._events[m.name]
            m.cycles += 1
            m.compute_time_ms = start.elapsed_time(end)
        return overall


# ---------------------------------------------------------------------------------------------
# Scheduler (synchronous) — src/scheduler_synchronous.cc
# ---------------------------------------------------------------------------------------------

class SynchronousScheduler:
    """Kahn topological order over producer/consumer links, one runtime per (device, runtime) segment,
    static settlement: STATIC_OUTPUT modules — and modules all of whose inputs are settled — run once."""

    def __init__(self, device="cuda"):
        self.device = device
        self.modules: Dict[str, Module] = {}
        self.order: List[Module] = []
        self.settled: set = set()
        self.runtime: Optional[NativeCudaRuntime] = None

    def add(self, module: Module) -> Result:
        if module.name in self.modules:
            return _error(f"[SCHEDULER] Module '{module.name}' already present.")
        self.modules[module.name] = module
        return self._rebuild()

    def remove(self, module: Module) -> Result:
        self.modules.pop(module.name, None)
        return self._rebuild()

    def _rebuild(self) -> Result:
        indegree = {name: 0 for name in self.modules}
        consumers: Dict[str, List[str]] = {name: [] for name in self.modules}
        for name, m in self.modules.items():
            for link in m.inputs.values():
                if link.producer and link.producer[0] in self.modules and link.producer[0] != name:
                    indegree[name] += 1
                    consumers[link.producer[0]].append(name)
        ready = [n for n in self.modules if indegree[n] == 0]
        order: List[str] = []
        while ready:
            n = ready.pop(0)
            order.append(n)
            for c in consumers[n]:
                indegree[c] -= 1
                if indegree[c] == 0:
                    ready.append(c)
        if len(order) != len(self.modules):
            return _error("[SCHEDULER] Cycle detected in module graph.")
        self.order = [self.modules[n] for n in order]
        self.settled = set()
        if self.runtime is not None:
            self.runtime.destroy()
        self.runtime = NativeCudaRuntime("segment0", self.device)
        return self.runtime.create(self.order)

    def is_static(self, module: Module) -> bool:
        if module.taint & Taint.STATIC_OUTPUT:
            return True
        links = list(module.inputs.values())
        producers = [l.producer[0] for l in links if l.producer]
        # an input without a producer is a caller-filled tensor: it may change every cycle, so nothing downstream settles
        if not producers or len(producers) != len(links) or not all(p in self.modules for p in producers):
            return False
        return all(self.is_static(self.modules[p]) for p in producers)

    def compute(self, failed: Optional[set] = None) -> Result:
        failed = failed if failed is not None else set()
        skipped: set = set()
        pending = [m.name for m in self.order if m.name not in self.settled]
        if not pending:
            return Result.SUCCESS
        result = self.runtime.compute(pending, skipped, failed)
        for m in self.order:
            if m.name in pending and m.name not in failed and m.name not in skipped and self.is_static(m):
                self.settled.add(m.name)
        return result


# ---------------------------------------------------------------------------------------------
# TestContext — src/testing.cc (one module, one runtime, CPU arrays in / out)
# ---------------------------------------------------------------------------------------------

class TestContext:
    __test__ = False  # not a pytest class

    def __init__(self, module_type: str, device: str = "cuda", runtime: str = "native", provider: str = "b200"):
        self.module_type, self.device, self.runtime_type, self.provider = module_type, device, runtime, provider
        self.inputs: Dict[str, Tensor] = {}
        self.config: Dict[str, object] = {}
        self.module: Optional[Module] = None
        self.runtime: Optional[NativeCudaRuntime] = None
        self.outputs: Dict[str, np.ndarray] = {}
        self.output_tensors: Dict[str, Tensor] = {}

    def set_input(self, name: str, array: np.ndarray, dtype: Optional[str] = None, **axes):
        target = self.device if (self.device != "cuda" or torch.cuda.is_available()) else "cpu"
        self.inputs[name] = Tensor.from_numpy(array, device=target, dtype=dtype, **axes)

    def set_config(self, **config):
        self.config = dict(config)

    def start(self) -> Result:
        self.module = build_module(self.module_type, self.device, self.runtime_type, self.provider)
        links = {}
        for name, tensor in self.inputs.items():
            link = TensorLink()
            link.produced("test", name, tensor)
            link.producer = None
            links[name] = link
        result = self.module.create("test", self.config, links)
        if result != Result.SUCCESS:
            return result
        self.runtime = NativeCudaRuntime("test", self.device)
        return self.runtime.create([self.module])

    def compute(self) -> Result:
        result = self.runtime.compute([], set(), set())
        if result != Result.SUCCESS:
            return result
        for port, link in self.module.outputs.items():
            self.output_tensors[port] = link.tensor
            self.outputs[port] = link.tensor.numpy()
        return Result.SUCCESS

    def stop(self) -> Result:
        if self.runtime:
            self.runtime.destroy()
            self.runtime = None
        if self.module:
            self.module.destroy()
            self.module = None
        return Result.SUCCESS

    def run(self) -> Result:
        result = self.start()
        if result == Result.SUCCESS:
            result = self.compute()
        self.stop()
        return result

    def output(self, name: str) -> np.ndarray:
        return self.outputs[name]
