"""Block wiring on the b200 provider — mirrors Block::Impl::moduleCreate / moduleExposeOutput
(src/block_impl.cc:30-90) for the blocks that call the hot path.

`SpectrumEngine` follows src/domains/dsp/spectrum_engine/block_impl.cc:120-217. With `fused=True`
(default) the per-cycle part of the chain (multiply -> fft -> amplitude -> [range]) is ONE module
(`spectral_chain`, one kernel); window -> invert stay separate static modules that settle after the
first cycle exactly as in the reference (block_tests.cc:103-122). `fused=False` wires the reference's
module sequence one-to-one on this provider (cast, window, invert, reshape, multiply, fft, amplitude,
range) — same outputs, one kernel per module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

from .jetstream import (Module, Result, SynchronousScheduler, Tensor, TensorLink, build_module, _error,
                        resolve_signal_axes, filter_resample_plan)


class Block:
    TYPE = ""
    DEFAULTS: Dict[str, object] = {}

    def __init__(self, device: str = "cuda", runtime: str = "native", provider: str = "b200", **config):
        unknown = set(config) - set(self.DEFAULTS)
        if unknown:
            raise TypeError(f"unknown config field(s) for block '{self.TYPE}': {sorted(unknown)}")
        self.config = dict(self.DEFAULTS)
        self.config.update(config)
        self.device, self.runtime, self.provider = device, runtime, provider
        self.name = ""
        self.modules: Dict[str, Module] = {}
        self.inputs: Dict[str, TensorLink] = {}
        self.outputs: Dict[str, TensorLink] = {}
        self.scheduler: Optional[SynchronousScheduler] = None
        self.state = "none"

    # -- Block::Impl helpers
    def module_create(self, name: str, type_: str, config: Optional[dict], inputs: Dict[str, TensorLink]) -> Result:
        module = build_module(type_, self.device, self.runtime, self.provider)
        full = f"{self.name}:{name}"
        result = module.create(full, config, inputs)
        if result != Result.SUCCESS:
            return result
        self.modules[name] = module
        return self.scheduler.add(module)

    def module_get_output(self, name: str, port: str) -> TensorLink:
        return self.modules[name].outputs[port]

    def module_expose_output(self, port: str, name: str, module_port: str) -> Result:
        self.outputs[port] = self.modules[name].outputs[module_port]
        return Result.SUCCESS

    # -- lifecycle
    def create(self, name: str, inputs: Dict[str, Tensor], scheduler: Optional[SynchronousScheduler] = None) -> Result:
        self.name = name
        self._owns_scheduler = scheduler is None
        self.scheduler = scheduler or SynchronousScheduler(self.device)
        self.inputs = {}
        for port, tensor in inputs.items():
            link = tensor if isinstance(tensor, TensorLink) else TensorLink(tensor=tensor)
            self.inputs[port] = link
        result = self.create_impl()
        self.state = "created" if result == Result.SUCCESS else "errored"
        if result != Result.SUCCESS:
            self.destroy()
        return result

    def create_impl(self) -> Result:
        raise NotImplementedError

    def compute(self) -> Result:
        return self.scheduler.compute()

    def output(self, port: str) -> Tensor:
        return self.outputs[port].tensor

    def destroy(self) -> Result:
        # Block::destroy (src/block.cc): every module leaves the scheduler — which rebuilds its order and runtime, so
        # the other blocks of a SHARED scheduler keep running — and is then destroyed; the runtime itself is torn
        # down only by the block that created the scheduler.
        for module in reversed(list(self.modules.values())):
            if self.scheduler is not None and module.name in self.scheduler.modules:
                self.scheduler.remove(module)
            module.destroy()
        if self.scheduler is not None and getattr(self, "_owns_scheduler", True) and self.scheduler.runtime is not None:
            self.scheduler.runtime.destroy()
            self.scheduler.runtime = None
        self.modules = {}
        return Result.SUCCESS

    def metrics(self) -> Dict[str, tuple]:
        """`runtime:<module>` -> (cycles, last compute ms), like the reference block metrics."""
        return {f"runtime:{k}": (m.cycles, m.compute_time_ms) for k, m in self.modules.items()}


class SpectrumEngine(Block):
    """`spectrum_engine` — include/jetstream/domains/dsp/spectrum_engine/block.hh:8-16."""
    TYPE = "spectrum_engine"
    # publishColumnSums (b200 only): the fused kernel also emits sum-over-batch of every output column for a downstream
    # `lineplot` on this provider (jetstream.SpectralChain.COLUMN_SUMS_ATTRIBUTE)
    DEFAULTS = {"enableAgc": False, "enableScale": False, "rangeMin": -120.0, "rangeMax": 0.0, "fused": True,
                "publishColumnSums": False}

    def create_impl(self) -> Result:
        port = self.inputs.get("buffer")
        if port is None or not port.resolved():
            return Result.INCOMPLETE
        tensor = port.tensor
        if tensor.dtype not in ("F32", "CF32"):
            return _error("[BLOCK_SPECTRUM_ENGINE] Input must have data type F32 or CF32.")
        axes = resolve_signal_axes(tensor)
        if axes is None:
            return _error("[BLOCK_SPECTRUM_ENGINE] Input signal axis metadata is invalid.")
        axis = axes.sample
        size = tensor.shape[axis]

        r = self.module_create("cast_input", "cast", {"outputType": "CF32"}, {"buffer": port})
        if r != Result.SUCCESS:
            return r
        complex_input = self.module_get_output("cast_input", "buffer")

        r = self.module_create("window", "window", {"size": size}, {})
        if r != Result.SUCCESS:
            return r
        r = self.module_create("invert", "invert", None, {"signal": self.module_get_output("window", "window")})
        if r != Result.SUCCESS:
            return r

        # enableAgc puts an `agc` module (one RMS tile per spectrum) between fft and amplitude
        # (block_impl.cc:186-200). The fused kernel has that stage for 4096-point spectra; other lengths run the
        # graph module by module.
        if self.config["fused"] and axis == tensor.rank - 1 and (not self.config["enableAgc"] or size == 4096):
            r = self.module_create("spectral_chain", "spectral_chain",
                                   {"enableScale": bool(self.config["enableScale"]),
                                    "rangeMin": float(self.config["rangeMin"]),
                                    "rangeMax": float(self.config["rangeMax"]),
                                    "enableAgc": bool(self.config["enableAgc"]),
                                    "publishColumnSums": bool(self.config["publishColumnSums"])},
                                   {"buffer": complex_input, "window": self.module_get_output("invert", "signal")})
            if r != Result.SUCCESS:
                return r
            return self.module_expose_output("buffer", "spectral_chain", "buffer")

        shape = [size if d == axis else 1 for d in range(tensor.rank)]
        r = self.module_create("reshape_window", "reshape", {"shape": str(shape)},
                               {"buffer": self.module_get_output("invert", "signal")})
        if r != Result.SUCCESS:
            return r
        reshaped = self.module_get_output("reshape_window", "buffer")
        reshaped.tensor.set_attribute("sampleAxis", axis)
        r = self.module_create("multiply", "multiply", None, {"a": complex_input, "b": reshaped})
        if r != Result.SUCCESS:
            return r
        r = self.module_create("fft", "fft", {"forward": True}, {"signal": self.module_get_output("multiply", "product")})
        if r != Result.SUCCESS:
            return r
        spectrum = self.module_get_output("fft", "signal")
        if self.config["enableAgc"]:
            r = self.module_create("agc", "agc", {"tileSize": size}, {"signal": spectrum})
            if r != Result.SUCCESS:
                return r
            spectrum = self.module_get_output("agc", "signal")
        r = self.module_create("amplitude", "amplitude", None, {"signal": spectrum})
        if r != Result.SUCCESS:
            return r
        if self.config["enableScale"]:
            r = self.module_create("range", "range", {"min": float(self.config["rangeMin"]),
                                                      "max": float(self.config["rangeMax"])},
                                   {"signal": self.module_get_output("amplitude", "signal")})
            if r != Result.SUCCESS:
                return r
            return self.module_expose_output("buffer", "range", "signal")
        return self.module_expose_output("buffer", "amplitude", "signal")


class Filter(Block):
    """`filter` — include/jetstream/domains/dsp/filter/block.hh:10-19, src/domains/dsp/filter/block_impl.cc.
    cast(bypass) -> filter_taps (static) -> fir_filter (fused time-domain replacement of the reference's
    pad/fft/multiply/fold/ifft/normalize/unpad/overlap_add chain). Resampling engages under exactly the
    reference's conditions (block_impl.cc:64-90); otherwise the block silently runs at full rate."""
    TYPE = "filter"
    DEFAULTS = {"sampleRate": 2e6, "bandwidth": 1e6, "center": (0.0,), "taps": 101, "heads": 1}

    def create_impl(self) -> Result:
        port = self.inputs.get("signal")
        if port is None or not port.resolved():
            return Result.INCOMPLETE
        cfg = self.config
        heads = int(cfg["heads"])
        if heads == 0:
            return _error("[BLOCK_FILTER] Heads must be greater than 0.")
        tensor = port.tensor
        if tensor.dtype not in ("F32", "CF32"):
            return _error("[BLOCK_FILTER] Signal input must have data type F32 or CF32.")
        axes = resolve_signal_axes(tensor)
        if axes is None:
            return _error("[BLOCK_FILTER] Signal axis metadata is invalid.")
        if axes.channel is not None:
            return _error("[BLOCK_FILTER] Signal already has channelAxis. Generated filter channels cannot be nested.")
        signal_size = tensor.shape[axes.sample]
        centers = list(cfg["center"])[:heads] + [0.0] * max(0, heads - len(cfg["center"]))
        r = filter_resample_plan(float(cfg["sampleRate"]), float(cfg["bandwidth"]), int(cfg["taps"]), signal_size)
        # Resampling heads with a non-zero centre are shifted to baseband before decimation: the reference rounds
        # the centre to a bin of its M-point spectrum, M = T + taps - 1 (block_impl.cc:104-160).
        center_bins = None
        if r > 1 and any(float(c) != 0.0 for c in centers):
            m = signal_size + int(cfg["taps"]) - 1
            per_bin = float(cfg["sampleRate"]) / float(m)
            center_bins = []
            for c in centers:
                cb_ = float(c) / per_bin
                center_bins.append(int(math.floor(abs(cb_) + 0.5)) * (1 if cb_ >= 0 else -1))   # std::round
        self.resample = r > 1
        result = self.module_create("cast_signal", "cast", {"outputType": "CF32"}, {"buffer": port})
        if result != Result.SUCCESS:
            return result
        result = self.module_create("filter_taps", "filter_taps",
                                    {"sampleRate": float(cfg["sampleRate"]), "bandwidth": float(cfg["bandwidth"]),
                                     "center": tuple(float(c) for c in centers), "taps": int(cfg["taps"])}, {})
        if result != Result.SUCCESS:
            return result
        result = self.module_create("fir", "fir_filter", {"decimation": r, "centerBins": center_bins},
                                    {"signal": self.module_get_output("cast_signal", "buffer"),
                                     "coeffs": self.module_get_output("filter_taps", "coeffs")})
        if result != Result.SUCCESS:
            return result
        self.module_expose_output("buffer", "fir", "buffer")
        if self.resample:
            import numpy as np
            self.outputs["buffer"].tensor.set_attribute("sampleRate", float(np.float32(float(cfg["sampleRate"]) / r)))
        return Result.SUCCESS


class AgcBlock(Block):
    """`agc` block — src/domains/dsp/agc/block_impl.cc: one `agc` module."""
    TYPE = "agc"
    DEFAULTS = {"tileSize": 1024, "reference": 1.0, "epsilon": 1e-12, "minGain": 0.01, "maxGain": 100.0,
                "maxGainChange": 4.0}

    def create_impl(self) -> Result:
        port = self.inputs.get("signal")
        if port is None or not port.resolved():
            return Result.INCOMPLETE
        # the block's parameters are F32 (include/jetstream/domains/dsp/agc/block.hh:9-14) widened into the module's
        # F64 config (block_impl.cc:17-24): minGain is (double)0.01f, not 0.01
        import numpy as np
        config = {k: (int(v) if k == "tileSize" else float(np.float32(v))) for k, v in self.config.items()}
        result = self.module_create("agc", "agc", config, {"signal": port})
        if result != Result.SUCCESS:
            return result
        return self.module_expose_output("signal", "agc", "signal")


class FmBlock(Block):
    """`fm` block — include/jetstream/domains/dsp/fm/block.hh:8-15: one `fm` module."""
    TYPE = "fm"
    DEFAULTS = {"mode": "narrow", "deemphasis": "none", "sampleRate": 240e3}

    def create_impl(self) -> Result:
        port = self.inputs.get("signal")
        if port is None or not port.resolved():
            return Result.INCOMPLETE
        result = self.module_create("fm", "fm", dict(self.config), {"signal": port})
        if result != Result.SUCCESS:
            return result
        return self.module_expose_output("signal", "fm", "signal")


class LineplotBlock(Block):
    """`lineplot` block — src/domains/visualization/lineplot/block_impl.cc: one `lineplot` module (compute half)."""
    TYPE = "lineplot"
    DEFAULTS = {"averaging": 1, "decimation": 1, "numberOfVerticalLines": 11, "numberOfHorizontalLines": 5,
                "thickness": 1.0}

    def create_impl(self) -> Result:
        port = self.inputs.get("signal")
        if port is None or not port.resolved():
            return Result.INCOMPLETE
        return self.module_create("lineplot", "lineplot", dict(self.config), {"signal": port})

    def signal_points(self):
        return self.modules["lineplot"].signal_points()


class WaterfallBlock(Block):
    """`waterfall` block — src/domains/visualization/waterfall/block_impl.cc: one `waterfall` module (compute half)."""
    TYPE = "waterfall"
    DEFAULTS = {"height": 512, "interpolate": True}

    def create_impl(self) -> Result:
        port = self.inputs.get("signal")
        if port is None or not port.resolved():
            return Result.INCOMPLETE
        return self.module_create("waterfall", "waterfall", dict(self.config), {"signal": port})

    def frequency_bins(self):
        return self.modules["waterfall"].frequency_bins()

    @property
    def write_index(self) -> int:
        return self.modules["waterfall"].write_index
