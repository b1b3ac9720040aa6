"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libjst_ref.so, built from
/root/reference by oracle/build_ref.sh). Run where that library exists:  python tests/golden/generate.py
The vectors are small (seeded inputs + reference outputs) and committed, so the oracle port and the CUDA
path can be checked against the reference's own outputs on boxes where the reference is absent."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref  # noqa: E402
from cyberether_b200.synthetic import gaussian_cf32, spectral_rows  # noqa: E402


def main():
    assert ref.available(), "build oracle/_ref first"
    out = {}
    # single 1024-pt FFT, batch 1 (BASELINE config 1)
    t = np.arange(1024)
    tone = np.exp(2j * np.pi * 5 * t / 1024).astype(np.complex64)
    out["fft1024_in"], out["fft1024_out"] = tone, ref.fft(tone)
    x = gaussian_cf32((3, 256), 1)
    out["fft256_in"], out["fft256_fwd"], out["fft256_inv"] = x, ref.fft(x), ref.fft(x, forward=False)
    for n in (8, 4096):
        out[f"window{n}"] = ref.window(n)
    # spectral chain, 4 rows x 4096 (configs[1] shape, reduced batch) with and without range
    xs = spectral_rows(0, 4)
    out["chain_in"] = xs
    out["chain_db"] = ref.spectrum_engine(xs, enable_scale=False)
    out["chain_scaled"] = ref.spectrum_engine(xs, enable_scale=True, range_min=-120.0, range_max=0.0)
    # amplitude / range on a fixed spectrum
    spec = ref.fft(gaussian_cf32((2, 512), 2, 10.0))
    out["amp_in"], out["amp_out"] = spec, ref.amplitude(spec)
    out["range_out"] = ref.range_(out["amp_out"], -80.0, -20.0)
    # filter block, two cycles (state carry), 129 taps / R = 8 and 127 taps / full rate
    for taps in (129, 127):
        cycles = [gaussian_cf32((4, 1024), 30 + i) for i in range(2)]
        with ref.Session() as s:
            s.add_source("src", cycles[0], sample_axis=1, batch_axis=0)
            s.add_block("f", "filter", {"sampleRate": 8e6, "bandwidth": 1e6, "taps": taps, "heads": 1},
                        {"signal": "src.signal"})
            for i, c in enumerate(cycles):
                s.write_source("src", c)
                s.compute()
                out[f"filter{taps}_out{i}"] = s.output("f", "buffer")
                out[f"filter{taps}_in{i}"] = c
    with ref.Session() as s:
        s.add_block("t", "filter_taps", {"sampleRate": 8e6, "bandwidth": 1e6, "center": [0.0, 1.5e6], "taps": 33,
                                         "heads": 2})
        s.compute()
        out["taps33"] = s.output("t", "coeffs")
    # fm narrow, two cycles, with and without de-emphasis
    rng = np.random.Generator(np.random.PCG64(77))
    for de in ("none", "75us"):
        cyc = []
        for i in range(2):
            ph = np.cumsum(rng.standard_normal(2048) * 0.3)
            cyc.append((np.exp(1j * ph) * (1 + 0.01 * rng.standard_normal(2048))).astype(np.complex64).reshape(2, 1024))
        with ref.Session() as s:
            s.add_source("src", cyc[0], sample_axis=1, batch_axis=0)
            s.add_block("fm", "fm", {"mode": "narrow", "deemphasis": de, "sampleRate": 250e3}, {"signal": "src.signal"})
            for i, c in enumerate(cyc):
                s.write_source("src", c)
                s.compute()
                out[f"fm_{de}_in{i}"] = c
                out[f"fm_{de}_out{i}"] = s.output("fm", "signal")
    # cast: every integer / complex-integer conversion (extremes included)
    rng = np.random.default_rng(123)
    for name, np_type, is_complex in (("I8", np.int8, False), ("U8", np.uint8, False), ("I16", np.int16, False),
                                      ("U16", np.uint16, False), ("I32", np.int32, False), ("U32", np.uint32, False),
                                      ("CI8", np.int8, True), ("CU8", np.uint8, True), ("CI16", np.int16, True),
                                      ("CU16", np.uint16, True), ("CI32", np.int32, True), ("CU32", np.uint32, True)):
        info = np.iinfo(np_type)
        v = rng.integers(info.min, info.max, size=(2, 24, 2) if is_complex else (2, 24), endpoint=True, dtype=np_type)
        v.flat[0], v.flat[-1] = info.min, info.max
        out[f"cast_{name}_in"] = v
        out[f"cast_{name}_out"] = ref.run_block("cast", {"buffer": v}, {"outputType": "CF32" if is_complex else "F32"},
                                                "buffer", dtypes={"buffer": name})
    # agc block (its parameters are F32 in the block): level jump, two tile sizes, F32 and CF32
    env = np.exp(rng.uniform(-5, 5, size=(3, 1))) * (1 + 7 * (np.arange(700) > 350))
    for kind in ("f32", "cf32"):
        a = rng.standard_normal((3, 700)) * env
        a = a.astype(np.float32) if kind == "f32" else (a + 1j * rng.standard_normal((3, 700)) * env).astype(np.complex64)
        out[f"agc_{kind}_in"] = a
        for tile in (128, 700):
            out[f"agc_{kind}_tile{tile}"] = ref.run_block("agc", {"signal": a}, {"tileSize": tile}, "signal")
    # SDR-style flowgraph: cast(CI8 / CI16) -> spectrum_engine, with and without AGC
    for name, np_type in (("CI8", np.int8), ("CI16", np.int16)):
        info = np.iinfo(np_type)
        t = np.arange(4096)
        sig = 0.4 * np.exp(2j * np.pi * 300.25 * t / 4096)[None, :] + 0.02 * (rng.standard_normal((3, 4096)) +
                                                                            1j * rng.standard_normal((3, 4096)))
        q = np.clip(np.rint(np.stack([sig.real, sig.imag], axis=-1) * info.max), info.min, info.max).astype(np_type)
        out[f"sdr_{name}_in"] = q
        for agc in (False, True):
            with ref.Session() as s:
                s.add_source("src", q, sample_axis=1, batch_axis=0, dtype=name)
                s.add_block("c", "cast", {"outputType": "CF32"}, {"buffer": "src.signal"})
                s.add_block("se", "spectrum_engine", {"enableScale": True, "enableAgc": agc, "rangeMin": -120.0,
                                                      "rangeMax": 0.0}, {"buffer": "c.buffer"})
                s.compute()
                s.compute()
                out[f"sdr_{name}_agc{int(agc)}"] = s.output("se", "buffer")
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "reference_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
