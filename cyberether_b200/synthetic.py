"""Seeded synthetic IQ generators shared by tests and bench.py.

The reference's own `signal_generator` seeds from std::random_device
(src/domains/dsp/signal_generator/module_impl_native_cpu.cc:88), so it cannot deliver identical bytes to
two implementations; these generators can.
"""
from __future__ import annotations

import numpy as np


def spectral_rows(first_row: int, rows: int, n: int = 4096, noise: float = 1e-3) -> np.ndarray:
    """CF32 [rows, n]: per row b three complex tones (bins 97b, 1013b+511, n/2 mod n; amplitudes 0.5, 0.05,
    0.005) plus complex Gaussian noise from PCG64(seed=0x5EED0000 + b)."""
    t = np.arange(n, dtype=np.float64)
    out = np.empty((rows, n), dtype=np.complex64)
    for i in range(rows):
        b = first_row + i
        rng = np.random.Generator(np.random.PCG64(0x5EED0000 + b))
        bins = ((97 * b) % n, (1013 * b + 511) % n, n // 2)
        x = np.zeros(n, dtype=np.complex128)
        for k, amp in zip(bins, (0.5, 0.05, 0.005)):
            x += amp * np.exp(2j * np.pi * k * t / n)
        x += noise * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        out[i] = x.astype(np.complex64)
    return out


def gaussian_cf32(shape, seed: int, scale: float = 1.0) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    return (scale * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))).astype(np.complex64)
