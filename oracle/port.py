"""TEST INFRASTRUCTURE — numpy restatement ("port") of the reference CPU algorithms on the hot path.

Each function restates one reference `computeSubmit()` and cites it (paths under the reference tree,
CyberEther/Jetstream v1.9.1). Arithmetic is carried out in the same precision and operation order as
the reference (np.float32 scalars/arrays; F64 where the reference uses F64).

Pinning: tests/test_oracle.py checks every function here against (i) the UNMODIFIED reference built by
oracle/build_ref.sh (oracle/_ref/libjst_ref.so, when present), (ii) the golden vectors under tests/golden/
generated from that library by tests/golden/generate.py, and (iii) the reference's own known-answer
module tests restated in tests/test_reference_known_answers.py. Parity is therefore pinned.

The FFT: the reference calls the vendored header-only pocketfft (src/domains/dsp/fft/pocketfft.hh) in
single precision; numpy >= 2.0 ships the same pocketfft C++ templates and keeps complex64 in single
precision, so `np.fft.fft` on complex64 input is the restatement of pocketfft::c2c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import ctypes

import math

import numpy as np

F32 = np.float32
JST_PI = 3.14159265358979323846


CAST_SCALERS = {1: np.float32(128.0), 2: np.float32(32768.0), 4: np.float32(2147483648.0)}


def cast(x: np.ndarray, complex_pairs: bool = False) -> np.ndarray:
    """src/domains/core/cast/module_impl_native_cpu.cc:163-330 with the scaler of module_impl.cc:50-72:
    integer -> F32 as static_cast<F32>(v) / scaler (128 / 32768 / 2^31 by integer width; unsigned types are not
    re-centred); `complex_pairs`: the last axis holds (re, im) of a complex-integer tensor -> CF32.
    F32 -> CF32 sets imag = 0."""
    if x.dtype == np.float32:
        return x.astype(np.complex64)
    y = x.astype(np.float32) / CAST_SCALERS[x.dtype.itemsize]
    if complex_pairs:
        out = np.empty(x.shape[:-1], np.complex64)
        out.real, out.imag = y[..., 0], y[..., 1]
        return out
    return y


def agc(x: np.ndarray, tile_size: int = 1024, reference: float = 1.0, epsilon: float = 1e-12,
        min_gain: float = 0.01, max_gain: float = 100.0, max_gain_change: float = 4.0, axis: int = -1) -> np.ndarray:
    """src/domains/dsp/agc/module_impl_native_cpu.cc:76-160 (ApplyTiledRmsAgc) with the helpers of :16-74: per lane,
    tile target gain = clamp(reference / sqrt(mean power + eps)); the gain inside tile t is interpolated from the gain
    at its start to the rate-limited (LimitGainChange) target of tile t+1; products are limited to the finite F32 range.
    F64 throughout, sequential power sums like the reference (math.fsum is NOT used: the reference adds in order)."""
    xm = np.moveaxis(np.asarray(x), axis, -1)
    lanes = xm.reshape(-1, xm.shape[-1])
    out = np.empty_like(lanes)
    n = lanes.shape[1]
    tiles = 1 + (n - 1) // tile_size
    fmax = float(np.finfo(np.float32).max)
    cmax = float(np.nextafter(np.float32(fmax), np.float32(0)))
    is_complex = np.iscomplexobj(lanes)

    def clamp(v, lo, hi):
        return lo if v < lo else (hi if hi < v else v)

    for li in range(lanes.shape[0]):
        lane = lanes[li]
        power = (lane.real.astype(np.float64) ** 2 + lane.imag.astype(np.float64) ** 2) if is_complex \
            else lane.astype(np.float64) ** 2

        def target(t):
            seg = power[t * tile_size:(t + 1) * tile_size]
            total = 0.0
            for v in seg:                       # sequential F64 sum, as the reference
                total += float(v)
            return clamp(reference / math.sqrt(total / len(seg) + epsilon), min_gain, max_gain)

        start = target(0)
        for t in range(tiles):
            lo_i, hi_i = t * tile_size, min((t + 1) * tile_size, n)
            length = hi_i - lo_i
            if t + 1 < tiles:
                lowest = max(min_gain, start / max_gain_change)
                highest = max_gain if start > max_gain / max_gain_change else start * max_gain_change
                end = clamp(target(t + 1), lowest, highest)
            else:
                end = start
            step = (end - start) / float(length)
            gains = start + step * np.arange(length, dtype=np.float64)
            seg = lane[lo_i:hi_i]
            if is_complex:
                re, im = seg.real.astype(np.float64), seg.imag.astype(np.float64)
                mag = np.hypot(re, im)
                with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                    safe = np.where(mag > cmax / gains, np.nextafter(cmax / mag, 0.0), gains)
                    o = np.clip(re * safe, -fmax, fmax).astype(np.float32) + \
                        1j * np.clip(im * safe, -fmax, fmax).astype(np.float32)
                out[li, lo_i:hi_i] = o.astype(np.complex64)
            else:
                v = seg.astype(np.float64)
                with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                    safe = np.where(np.abs(v) > fmax / gains, np.nextafter(fmax / np.abs(v), 0.0), gains)
                    out[li, lo_i:hi_i] = np.clip(v * safe, -fmax, fmax).astype(np.float32)
            start = end
    return np.moveaxis(out.reshape(xm.shape), -1, axis)


def window(n: int) -> np.ndarray:
    """src/domains/dsp/window/module_impl_native_cpu.cc:20-37 — Blackman taps in F64, stored CF32."""
    if n == 1:
        return np.array([1 + 0j], dtype=np.complex64)
    i = np.arange(n, dtype=np.float64)
    tap = 0.42 - 0.50 * np.cos(2.0 * JST_PI * i / (n - 1)) + 0.08 * np.cos(4.0 * JST_PI * i / (n - 1))
    return tap.astype(np.float32).astype(np.complex64)


def invert(x: np.ndarray, axis: int = -1) -> np.ndarray:
    """src/domains/dsp/invert/module_impl_native_cpu.cc:78-103 — (-1)^k along `axis` (even length), or the
    F64-evaluated phasor exp(j 2 pi floor(N/2) k / N) for odd lengths."""
    x = np.asarray(x, dtype=np.complex64)
    n = x.shape[axis]
    k = np.arange(n)
    shape = [1] * x.ndim
    shape[axis] = n
    if n % 2 == 0:
        sign = np.where(k % 2 == 1, F32(-1), F32(1)).astype(np.float32).reshape(shape)
        return (x * sign).astype(np.complex64)   # exact sign flip
    phase = 2.0 * JST_PI * float(n // 2) * k.astype(np.float64) / float(n)
    w = (np.cos(phase).astype(np.float32) + 1j * np.sin(phase).astype(np.float32)).astype(np.complex64)
    return multiply(x, w.reshape(shape))


def multiply(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """src/domains/core/multiply/module_impl_native_cpu.cc:86-100 — NumPy-style broadcast product.
    CF32: (ac - bd, ad + bc), every product and sum individually rounded to F32 (no FMA)."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.dtype == np.complex64:
        ar, ai = a.real.astype(np.float32), a.imag.astype(np.float32)
        br, bi = np.asarray(b.real, np.float32), np.asarray(b.imag, np.float32)
        re = (ar * br).astype(np.float32) - (ai * bi).astype(np.float32)
        im = (ar * bi).astype(np.float32) + (ai * br).astype(np.float32)
        out = np.empty(np.broadcast_shapes(a.shape, b.shape), dtype=np.complex64)
        out.real = re
        out.imag = im
        return out
    return (a.astype(np.float32) * b.astype(np.float32)).astype(np.float32)


def fft(x: np.ndarray, forward: bool = True, axis: int = -1) -> np.ndarray:
    """src/domains/dsp/fft/module_impl_native_cpu.cc:129-140 — pocketfft::c2c, scale 1.0 both ways."""
    x = np.ascontiguousarray(x, dtype=np.complex64)
    if forward:
        return np.fft.fft(x, axis=axis).astype(np.complex64)
    return (np.fft.ifft(x, axis=axis, norm="forward")).astype(np.complex64)


def approx_log10(x: np.ndarray) -> np.ndarray:
    """Backend::ApproxLog10, include/jetstream/backend/devices/cpu/helpers.hh:61-74 (F32, step by step)."""
    f, e = np.frexp(np.abs(x).astype(np.float32))
    f = f.astype(np.float32)
    y = np.full_like(f, F32(1.23149591368684))
    y = (y * f).astype(np.float32)
    y = (y + F32(-4.11852516267426)).astype(np.float32)
    y = (y * f).astype(np.float32)
    y = (y + F32(6.02197014179219)).astype(np.float32)
    y = (y * f).astype(np.float32)
    y = (y + F32(-3.13396450166353)).astype(np.float32)
    y = (y + e.astype(np.float32)).astype(np.float32)
    return (y * F32(0.3010299956639812)).astype(np.float32)


_libm = ctypes.CDLL("libm.so.6")
_libm.log10f.restype = ctypes.c_float
_libm.log10f.argtypes = [ctypes.c_float]


def amplitude_coeff(n: int) -> np.float32:
    """scalingCoeff, src/domains/dsp/amplitude/module_impl.cc:49-51: 20.0f * std::log10(1.0f / (F32)N) —
    glibc's log10f (numpy's F32 log10 differs in the last bit for some N)."""
    return F32(F32(20.0) * F32(_libm.log10f(float(F32(1.0) / F32(n)))))


def amplitude(x: np.ndarray, n: int) -> np.ndarray:
    """src/domains/dsp/amplitude/module_impl_native_cpu.cc:73-99 — n = length of the sample axis."""
    coeff = amplitude_coeff(n)
    if np.iscomplexobj(x):
        re = x.real.astype(np.float32)
        im = x.imag.astype(np.float32)
        mag = np.sqrt(((re * re).astype(np.float32) + (im * im).astype(np.float32)).astype(np.float32))
    else:
        mag = np.abs(x.astype(np.float32))
    with np.errstate(divide="ignore"):
        body = (F32(20.0) * approx_log10(np.where(mag == 0, F32(1), mag))).astype(np.float32) + coeff
    return np.where(mag == 0, F32(-np.inf), body).astype(np.float32)


def range_coefficients(lo: float, hi: float):
    """RangeImpl::updateCoefficients, src/domains/core/range/module_impl.cc:51-63."""
    lower, upper = F32(min(lo, hi)), F32(max(lo, hi))
    if lower == upper:
        return F32(0.0), F32(0.5)
    scale = F32(1.0) / (upper - lower)
    return scale, (-lower * scale).astype(np.float32)


def range_(x: np.ndarray, lo: float, hi: float) -> np.ndarray:
    """src/domains/core/range/module_impl_native_cpu.cc:67-82 — tanh soft knee."""
    scale, offset = range_coefficients(lo, hi)
    x = x.astype(np.float32)
    if scale == 0:
        return np.full_like(x, F32(0.5))
    normalized = ((x * scale).astype(np.float32) + offset).astype(np.float32)
    with np.errstate(invalid="ignore"):
        t = np.tanh((F32(4.0) * (normalized - F32(0.5)).astype(np.float32)).astype(np.float32)).astype(np.float32)
    return (F32(0.5) + (F32(0.5) * t).astype(np.float32)).astype(np.float32)


def spectrum_engine(x: np.ndarray, enable_scale: bool = False, range_min: float = -120.0,
                    range_max: float = 0.0) -> np.ndarray:
    """src/domains/dsp/spectrum_engine/block_impl.cc:120-217 with the sample axis innermost:
    cast(bypass) -> window -> invert -> reshape -> multiply -> fft -> amplitude -> [range]."""
    x = np.asarray(x, dtype=np.complex64)
    n = x.shape[-1]
    w = invert(window(n))
    spectrum = fft(multiply(x, w.reshape((1,) * (x.ndim - 1) + (n,))))
    out = amplitude(spectrum, n)
    if enable_scale:
        out = range_(out, range_min, range_max)
    return out


# ---------------------------------------------------------------------------------------------
# filter block
# ---------------------------------------------------------------------------------------------

def filter_taps(sample_rate: float, bandwidth: float, center, taps: int) -> np.ndarray:
    """FilterTapsImplNativeCpu::generateCoeffs, src/domains/dsp/filter_taps/module_impl_native_cpu.cc:46-80
    (windowed sinc x Blackman x upconversion in F64, stored CF32 [heads, taps])."""
    center = np.atleast_1d(np.asarray(center, dtype=np.float64))
    width = (bandwidth / sample_rate) / 2.0
    i = np.arange(taps, dtype=np.float64)
    n = i - (taps - 1) / 2.0
    with np.errstate(divide="ignore", invalid="ignore"):
        sinc = np.where(n == 0.0, 2.0 * width, np.sin(2.0 * JST_PI * width * n) / (JST_PI * n))
    win = np.ones(taps) if taps == 1 else (0.42 - 0.50 * np.cos(2.0 * JST_PI * i / (taps - 1)) +
                                           0.08 * np.cos(4.0 * JST_PI * i / (taps - 1)))
    out = np.empty((len(center), taps), dtype=np.complex64)
    for c, ct in enumerate(center):
        up = np.exp(1j * 2.0 * JST_PI * n * (ct / sample_rate))
        res = sinc * win * up
        out[c].real = res.real.astype(np.float32)
        out[c].imag = res.imag.astype(np.float32)
    return out


def filter_resample_ratio(sample_rate: float, bandwidth: float, taps: int, signal_size: int) -> int:
    """CalculateCandidatePlan, src/domains/dsp/filter/block_impl.cc:40-168 (zero-centred heads)."""
    ratio = sample_rate / bandwidth
    if not np.isfinite(ratio) or ratio <= 0 or ratio != np.floor(ratio):
        return 1
    r = int(ratio)
    if (taps - 1) % r != 0 or (taps + signal_size - 1) % r != 0:
        return 1
    return r


class FilterBlock:
    """The `filter` block as the reference wires it (src/domains/dsp/filter/block_impl.cc:350-582), stage by
    stage in the frequency domain: pad -> fft -> multiply -> fold -> ifft -> multiply_constant -> unpad ->
    overlap_add, with the overlap tail carried across frames and across calls
    (overlap_add/module_impl_native_cpu.cc:121-203). Input [B, T] (frames consecutive in time) or [T]."""

    def __init__(self, sample_rate: float, bandwidth: float, taps: int, frame_len: int, center=(0.0,)):
        self.taps = filter_taps(sample_rate, bandwidth, center, taps)        # [heads, L]
        self.heads, self.L = self.taps.shape
        self.T = frame_len
        self.M = frame_len + taps - 1
        self.R = filter_resample_ratio(sample_rate, bandwidth, taps, frame_len)
        if self.R > 1 and np.any(np.asarray(center) != 0):
            raise NotImplementedError("fold offsets / phase_correction are outside this port")
        padded = np.zeros((self.heads, self.M), np.complex64)
        padded[:, : self.L] = self.taps                                      # pad (core/pad): zeros at the tail
        self.filter_spectrum = fft(padded)                                   # fftFilter (static)
        self.tail = np.zeros((self.heads, (self.L - 1) // self.R), np.complex64)

    def __call__(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x, np.complex64)
        frames = x.reshape(-1, self.T)
        out = np.empty((frames.shape[0], self.heads, self.T // self.R), np.complex64)
        for b, frame in enumerate(frames):
            padded = np.zeros(self.M, np.complex64)
            padded[: self.T] = frame
            product = multiply(fft(padded)[None, :], self.filter_spectrum)   # [heads, M]
            if self.R > 1:                                                   # fold: F64 mean of R aliased bins
                size = self.M // self.R
                folded = product.astype(np.complex128).reshape(self.heads, self.R, size).sum(axis=1) / float(self.R)
                product = folded.astype(np.complex64)
            time = fft(product, forward=False)
            norm = F32(1.0) / F32(time.shape[-1])
            time = (time * norm).astype(np.complex64)                        # multiply_constant
            pad = (self.L - 1) // self.R
            body, tail = time[:, : time.shape[-1] - pad].copy(), time[:, time.shape[-1] - pad:]
            body[:, :pad] += self.tail                                       # overlap_add
            self.tail = tail.copy()
            out[b] = body
        return out.reshape(x.shape[:-1] + (self.heads, self.T // self.R))


# ---------------------------------------------------------------------------------------------
# fm (narrow)
# ---------------------------------------------------------------------------------------------

class FmNarrow:
    """FmImplNativeCpu::computeSubmit narrow path, src/domains/dsp/fm/module_impl_native_cpu.cc:43-129 with the
    coefficients of module_impl.cc:108-124. Input [frames, lanes, frame_len] (or fewer dims); state per lane."""

    def __init__(self, sample_rate: float, deemphasis: str = "none", lanes: int = 1):
        sr = F32(sample_rate)
        kf = F32(100e3) / sr
        self.ref = F32(1.0 / (2.0 * JST_PI * float(kf)))
        self.deemph = deemphasis != "none"
        tau = 50e-6 if deemphasis == "50us" else 75e-6
        self.alpha = F32(1.0 - np.exp(-1.0 / (float(sr) * tau))) if self.deemph else F32(1.0)
        self.prev = np.zeros(lanes, np.complex64)
        self.has_prev = np.zeros(lanes, bool)
        self.state = np.zeros(lanes, np.float32)

    def __call__(self, x: np.ndarray) -> np.ndarray:
        frames, lanes, n = x.shape
        out = np.empty(x.shape, np.float32)
        for lane in range(lanes):
            stream = x[:, lane, :].reshape(-1)
            prev = np.concatenate([[self.prev[lane]], stream[:-1]])
            pr, pi = prev.real.astype(np.float32), prev.imag.astype(np.float32)
            cr, ci = stream.real.astype(np.float32), stream.imag.astype(np.float32)
            with np.errstate(invalid="ignore", over="ignore"):
                re = (pr * cr).astype(np.float32) + (pi * ci).astype(np.float32)   # conj(prev) * cur, no FMA
                im = (pr * ci).astype(np.float32) - (pi * cr).astype(np.float32)
                d = (np.arctan2(im, re).astype(np.float32) * self.ref).astype(np.float32)
            bad = ~(np.isfinite(cr) & np.isfinite(ci) & np.isfinite(pr) & np.isfinite(pi))
            d[bad] = np.nan
            if not self.has_prev[lane]:
                d[0] = F32(0.0)
            if self.deemph:
                y = self.state[lane]
                for i in range(d.size):
                    if np.isfinite(d[i]):
                        y = F32(y + F32(self.alpha * F32(d[i] - y)))
                        d[i] = y
                self.state[lane] = y
            out[:, lane, :] = d.reshape(frames, n)
            self.prev[lane] = stream[-1]
            self.has_prev[lane] = True
        return out


class FmWide:
    """FmImplNativeCpu::computeSubmit wide (stereo) path, src/domains/dsp/fm/module_impl_native_cpu.cc:112-165, with the
    coefficients of FmImpl::updateCoefficients (module_impl.cc:108-155) and applyBiquad / applyAudioLowPass
    (module_impl.cc:157-174). Sequential per sample and lane, F32 operation by operation (python loop: small cases
    only). Input [frames, lanes, frame_len]; output [frames, lanes, frame_len, 2] = (left, right). A sample with a
    non-finite discriminator output only advances the pilot NCO (:96-105)."""

    def __init__(self, sample_rate: float, deemphasis: str = "none", lanes: int = 1):
        sr = F32(sample_rate)
        srd = float(sr)
        kf = F32(75e3) / sr
        self.ref = F32(1.0 / (2.0 * JST_PI * float(kf)))
        # `2.0f * JST_PI * 19e3f / sampleRate`: JST_PI is a double literal -> evaluated left to right in F64, then F32
        self.inc = F32(((2.0 * JST_PI) * float(F32(19e3))) / srd)
        self.pilot_alpha = F32(1.0 - np.exp(-2.0 * JST_PI * 200.0 / srd))
        self.deemph = deemphasis != "none"
        tau = 50e-6 if deemphasis == "50us" else 75e-6
        self.deemph_alpha = F32(1.0 - np.exp(-1.0 / (srd * tau))) if self.deemph else F32(1.0)
        w = 2.0 * JST_PI * 19e3 / srd
        na = np.sin(w) / (2.0 * 20.0)
        a0 = 1.0 + na
        b0 = F32(1.0 / a0)
        b1 = F32(-2.0 * np.cos(w) / a0)
        self.notch = (b0, b1, b0, b1, F32((1.0 - na) / a0))              # b0 b1 b2 a1 a2
        self.lowpass = []
        w = 2.0 * JST_PI * 15e3 / srd
        for q in (0.51763809, 0.70710678, 1.93185165):
            al = np.sin(w) / (2.0 * q)
            a0 = 1.0 + al
            lb0 = F32((1.0 - np.cos(w)) * 0.5 / a0)
            self.lowpass.append((lb0, F32((1.0 - np.cos(w)) / a0), lb0, F32(-2.0 * np.cos(w) / a0),
                                 F32((1.0 - al) / a0)))
        z = lambda: [F32(0.0), F32(0.0)]
        self.lanes = [dict(prev=np.complex64(0), has_prev=False, phase=F32(0.0), cos_stage=F32(0.0), sin_stage=F32(0.0),
                           pcos=F32(0.0), psin=F32(0.0), sum_notch=z(), diff_notch=z(),
                           sum_lp=[z(), z(), z()], diff_lp=[z(), z(), z()], left=F32(0.0), right=F32(0.0))
                      for _ in range(lanes)]

    @staticmethod
    def _biquad(x, c, st):
        b0, b1, b2, a1, a2 = c
        y = F32(F32(b0 * x) + st[0])
        st[0] = F32(F32(F32(b1 * x) - F32(a1 * y)) + st[1])
        st[1] = F32(F32(b2 * x) - F32(a2 * y))
        return y

    def _lowpass(self, x, states):
        for c, st in zip(self.lowpass, states):
            x = self._biquad(x, c, st)
        return x

    def _advance(self, s):
        s["phase"] = F32(s["phase"] + self.inc)
        if float(s["phase"]) >= 2.0 * JST_PI:
            s["phase"] = F32(float(s["phase"]) - 2.0 * JST_PI)

    def __call__(self, x: np.ndarray) -> np.ndarray:
        frames, lanes, n = x.shape
        out = np.empty(x.shape + (2,), np.float32)
        two = F32(2.0)
        with np.errstate(invalid="ignore", over="ignore"):
            for lane in range(lanes):
                s = self.lanes[lane]
                for f in range(frames):
                    for i in range(n):
                        cur = x[f, lane, i]
                        prev = s["prev"]
                        if not s["has_prev"]:
                            d = F32(0.0)
                        else:
                            pr, pi, cr, ci = F32(prev.real), F32(prev.imag), F32(cur.real), F32(cur.imag)
                            if not (np.isfinite(pr) and np.isfinite(pi) and np.isfinite(cr) and np.isfinite(ci)):
                                d = F32(np.nan)
                            else:
                                re = F32(F32(pr * cr) + F32(pi * ci))
                                im = F32(F32(pr * ci) - F32(pi * cr))
                                d = F32(F32(np.arctan2(im, re)) * self.ref)
                        s["prev"], s["has_prev"] = cur, True
                        if not np.isfinite(d):
                            out[f, lane, i] = np.nan
                            self._advance(s)
                            continue
                        pc, ps = F32(np.cos(s["phase"])), F32(np.sin(s["phase"]))
                        s["cos_stage"] = F32(s["cos_stage"] + F32(self.pilot_alpha * F32(F32(d * pc) - s["cos_stage"])))
                        s["sin_stage"] = F32(s["sin_stage"] + F32(self.pilot_alpha * F32(F32(d * ps) - s["sin_stage"])))
                        s["pcos"] = F32(s["pcos"] + F32(self.pilot_alpha * F32(s["cos_stage"] - s["pcos"])))
                        s["psin"] = F32(s["psin"] + F32(self.pilot_alpha * F32(s["sin_stage"] - s["psin"])))
                        total = self._lowpass(self._biquad(d, self.notch, s["sum_notch"]), s["sum_lp"])
                        offset = F32(np.arctan2(s["pcos"], s["psin"]))
                        carrier = F32(np.sin(F32(two * F32(s["phase"] + offset))))
                        diff = self._lowpass(self._biquad(F32(F32(two * d) * carrier), self.notch, s["diff_notch"]),
                                             s["diff_lp"])
                        left, right = F32(total + diff), F32(total - diff)
                        if self.deemph:
                            s["left"] = F32(s["left"] + F32(self.deemph_alpha * F32(left - s["left"])))
                            s["right"] = F32(s["right"] + F32(self.deemph_alpha * F32(right - s["right"])))
                            left, right = s["left"], s["right"]
                        out[f, lane, i, 0], out[f, lane, i, 1] = left, right
                        self._advance(s)
        return out


# ---------------------------------------------------------------------------------------------
# lineplot / waterfall consumers
# ---------------------------------------------------------------------------------------------

class Lineplot:
    """LineplotImplNativeCpu::computeSubmit, src/domains/visualization/lineplot/module_impl_native_cpu.cc:80-122, with the
    geometry of LineplotImpl::validate (module_impl.cc:129-187): numberOfElements = extent // decimation,
    normalizationFactor = 1 / (0.5 * batches). x is [batches, extent] (batch axis first) or [extent]."""

    def __init__(self, extent: int, batches: int = 1, decimation: int = 1, averaging: int = 1):
        self.n = extent // decimation
        self.batches, self.decimation, self.averaging = batches, decimation, averaging
        self.norm = F32(1.0) / F32(F32(0.5) * F32(batches))
        self.average = np.zeros(self.n, F32)
        # signalPoints x coordinates (module_impl_native_cpu.cc:66-70): i * 2.0f / (n - 1) - 1.0f
        i = np.arange(self.n, dtype=F32)
        self.x = F32(F32(i * F32(2.0)) / F32(self.n - 1)) - F32(1.0)

    def compute(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x, F32).reshape(self.batches, -1)
        sums = np.zeros(self.n, F32)
        for b in range(self.batches):                                  # :93-98 sequential F32 accumulation over rows
            sums = (sums + x[b, : self.n * self.decimation: self.decimation]).astype(F32)
        with np.errstate(invalid="ignore"):
            amp = np.fmin(np.fmax((sums * self.norm).astype(F32) - F32(1.0), F32(-1.0)), F32(1.0)).astype(F32)   # :105
        avg = self.average
        avg = (avg - (avg / F32(self.averaging)).astype(F32)).astype(F32)          # :109
        avg = (avg + (amp / F32(self.averaging)).astype(F32)).astype(F32)          # :110
        self.average = avg
        return np.stack([self.x, avg], axis=1)


class Waterfall:
    """WaterfallImplNativeCpu::computeSubmit (waterfall/module_impl_native_cpu.cc:53-78) with PlanWaterfallWrite /
    WaterfallRingState::advance (waterfall/ring_state.hh:18-44). ring is [height, n]."""

    def __init__(self, n: int, height: int = 512):
        self.n, self.height = n, height
        self.ring = np.zeros((height, n), F32)
        self.write_index = 0

    def compute(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x, F32).reshape(-1, self.n)
        batches = x.shape[0]
        retained = min(batches, self.height)
        source = batches - retained
        dest = (self.write_index + (source % self.height)) % self.height
        for row in range(retained):
            self.ring[(dest + row) % self.height] = x[source + row]
        self.write_index = (self.write_index + (batches % self.height)) % self.height
        return self.ring
