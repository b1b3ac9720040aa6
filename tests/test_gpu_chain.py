"""GPU parity: fused spectral chain (b200_chain_exec through the host mirror) vs the reference CPU
spectrum_engine block run by the reference's own Flowgraph/scheduler (oracle/_ref)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DB_TOL = 1e-3        # dB, on bins within 100 dB of the row maximum (SURVEY.md §7 "Tolerance definition")
DB_STEP = 2.5e-3     # ApproxLog10 is discontinuous by 0.0021 dB at octave boundaries (helpers.hh:61-74)
RANGE_TOL = 1e-5     # absolute, output in [0, 1]
RANGE_STEP = 4e-5    # the same discontinuity through d(range)/d(dB) <= 2/120


def _run_chain(x, enable_scale, rmin=-120.0, rmax=0.0):
    import cyberether_b200 as cb
    from cyberether_b200.blocks import SpectrumEngine
    block = SpectrumEngine(enableScale=enable_scale, rangeMin=rmin, rangeMax=rmax)
    inp = cb.Tensor.from_numpy(x, sampleAxis=x.ndim - 1, batchAxis=0 if x.ndim > 1 else None)
    assert block.create("spec", {"buffer": inp}) == cb.Result.SUCCESS, cb.last_error()
    for _ in range(2):
        assert block.compute() == cb.Result.SUCCESS, cb.last_error()
    out = block.output("buffer").numpy()
    block.destroy()
    return out


def _compare_db(got, want):
    finite = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), finite)
    floor = want.max(axis=-1, keepdims=True) - 100.0
    mask = finite & (want > floor)
    err = np.abs(got - want)[mask]
    frac_bad = float((err > DB_TOL).mean())
    assert err.max() <= DB_STEP, err.max()
    assert frac_bad < 1e-4, frac_bad


@pytest.mark.parametrize("rows", [1, 3, 64, 300])
def test_chain_4096_scale(ref, rows):
    from cyberether_b200.synthetic import spectral_rows
    x = spectral_rows(0, rows)
    want = ref.spectrum_engine(x, enable_scale=True)
    got = _run_chain(x, True)
    assert got.shape == want.shape and got.dtype == np.float32
    err = np.abs(got - want)
    assert err.max() <= RANGE_STEP, err.max()
    assert float((err > RANGE_TOL).mean()) < 1e-4


def test_chain_4096_db(ref):
    from cyberether_b200.synthetic import spectral_rows
    x = spectral_rows(1000, 128)
    want = ref.spectrum_engine(x, enable_scale=False)
    got = _run_chain(x, False)
    _compare_db(got, want)


def test_chain_sanity_golden():
    """BASELINE.md §2 golden: tone at bin 100, amplitude 0.5 -> peak bin 2148, -13.5583 dB, range 0.956732."""
    n = 4096
    x = (0.5 * np.exp(2j * np.pi * 100 * np.arange(n) / n)).astype(np.complex64)[None, :].repeat(4, 0)
    db = _run_chain(x, False)
    assert int(db[0].argmax()) == 2148
    assert abs(float(db[0].max()) - (-13.5583)) < 2e-3
    sc = _run_chain(x, True)
    assert abs(float(sc[0].max()) - 0.956732) < 2e-5


@pytest.mark.parametrize("n", [8, 64, 256, 1024, 2048, 8192, 16384])
def test_chain_other_sizes(ref, n):
    from cyberether_b200.synthetic import spectral_rows
    x = spectral_rows(7, 5, n=n)
    want = ref.spectrum_engine(x, enable_scale=True, range_min=-100.0, range_max=-10.0)
    got = _run_chain(x, True, -100.0, -10.0)
    err = np.abs(got - want)
    assert err.max() <= RANGE_STEP, err.max()


def test_chain_zero_input(ref):
    x = np.zeros((2, 4096), np.complex64)
    assert np.array_equal(_run_chain(x, True), ref.spectrum_engine(x, enable_scale=True))
    got = _run_chain(x, False)
    assert np.all(np.isneginf(got))
