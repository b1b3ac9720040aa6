"""The drop-in boundary, end to end: the reference's own blocks in the reference's own Flowgraph /
scheduler_synchronous / NativeCudaRuntime, selected with `device: cuda / runtime: native / provider: b200`
(shim/b200_modules.cc + shim/b200_blocks.cc -> libb200dsp.so), next to the same blocks on the reference CPU provider
in the same process. Three cycles of fresh input each (filter / fm / window state carried across cycles)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from parity import assert_db_close, assert_strong_bins, true_spectrum


@pytest.fixture(scope="module")
def sb():
    from shim import binding
    if not binding.available():
        pytest.skip("shim/_build/libjst_b200.so not built (needs /root/reference at build time)")
    return binding


def _both(sb, block_type, config, in_port, out_port, cycles, axes, dtype="CF32"):
    """Runs `cycles` (list of arrays) through the block on both targets; returns (cpu outs, b200 outs, b200 modules,
    cpu modules, info)."""
    outs = {}
    mods = {}
    info = {}
    for target in (sb.CPU, sb.B200):
        with sb.Session() as s:
            s.add_source("src", cycles[0].shape, dtype, target=target, **axes)
            s.add_block("dut", block_type, config, {in_port: "src.signal"}, target=target)
            res = []
            for x in cycles:
                s.write_source("src", x)
                s.compute()
                res.append(s.read("dut", out_port))
            outs[target[1]] = res
            mods[target[1]] = s.modules("dut")
            info[target[1]] = s.info("dut", out_port)
    return outs["generic"], outs["b200"], mods["b200"], mods["generic"], info


def _window(n):
    from oracle import port
    return port.invert(port.window(n))


def test_shim_demo_executable(sb):
    """The stand-alone C++ self-check over the same library (no Python in the loop)."""
    run = subprocess.run([sb.DEMO_PATH], capture_output=True, text=True, timeout=600)
    print(run.stdout[-3000:])
    assert run.returncode == 0 and "SHIM OK" in run.stdout, run.stdout[-2000:] + run.stderr[-2000:]


@pytest.mark.parametrize("scale,agc", [(False, False), (True, False), (True, True), (False, True)])
def test_spectrum_engine_block_fused(sb, scale, agc):
    from cyberether_b200.synthetic import spectral_rows
    cycles = [spectral_rows(100 * k, 48) for k in range(3)]
    cfg = {"enableScale": scale, "enableAgc": agc, "rangeMin": -120.0, "rangeMax": 0.0}
    cpu, gpu, mods, cpu_mods, info = _both(sb, "spectrum_engine", cfg, "buffer", "buffer", cycles,
                                           {"sampleAxis": 1, "batchAxis": 0})
    # ONE per-cycle module on the b200 target; the window modules settled after the first cycle on both targets
    assert set(mods) == {"runtime:cast_input", "runtime:window", "runtime:invert", "runtime:reshape_window",
                         "runtime:spectral_chain"}, mods
    assert mods["runtime:spectral_chain"][0] == 3 and mods["runtime:window"][0] == 1 and mods["runtime:invert"][0] == 1
    assert cpu_mods["runtime:fft"][0] == 3 and cpu_mods["runtime:window"][0] == 1
    assert info["b200"]["shape"] == info["generic"]["shape"] and info["b200"]["device"] == "cuda"
    for key in ("sampleAxis", "batchAxis", "channelAxis"):
        assert info["b200"][key] == info["generic"][key]
    slope = 2.0 / 120.0
    w = _window(4096)
    for x, want, got in zip(cycles, cpu, gpu):
        spec = true_spectrum(x, w)
        if agc:
            # the gain multiplies every bin: allowance and strong-bin statistic are level-relative, unchanged
            pass
        if scale:
            assert_db_close(got, want, spec, scale=slope, floor=3e-7)
            st = assert_strong_bins(got, want, spec, slope=slope, label=f"scale={scale} agc={agc}")
        else:
            assert_db_close(got, want, spec)
            st = assert_strong_bins(got, want, spec, label=f"scale={scale} agc={agc}")
    print(f"strong-bin statistic scale={scale} agc={agc}: {st}")


def test_spectrum_engine_block_f32_input_and_reconfigure(sb):
    """F32 input goes through the b200 cast module into the fused kernel; a rangeMin/rangeMax reconfigure of the
    BLOCK reaches the fused module in place (no recreate), like RangeImpl::reconfigure."""
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((8, 4096)) * 0.1).astype(np.float32)
    results = {}
    for target in (sb.CPU, sb.B200):
        with sb.Session() as s:
            s.add_source("src", x.shape, "F32", target=target, sampleAxis=1, batchAxis=0)
            s.add_block("dut", "spectrum_engine", {"enableScale": True}, {"buffer": "src.signal"}, target=target)
            s.write_source("src", x)
            s.compute()
            a = s.read("dut", "buffer")
            s.reconfigure("dut", {"enableScale": True, "rangeMin": -90.0, "rangeMax": -10.0})
            s.compute()
            b = s.read("dut", "buffer")
            results[target[1]] = (a, b, s.modules("dut"))
    spec = true_spectrum(x.astype(np.complex64), _window(4096))
    assert_db_close(results["b200"][0], results["generic"][0], spec, scale=2.0 / 120.0, floor=3e-7)
    assert_db_close(results["b200"][1], results["generic"][1], spec, scale=2.0 / 80.0, floor=3e-7)
    assert not np.allclose(results["b200"][0], results["b200"][1])
    assert results["b200"][2]["runtime:spectral_chain"][0] == 2          # same module instance computed both cycles


@pytest.mark.parametrize("n", [16384, 65536])
def test_spectrum_engine_block_large_spectra_run_the_tiled_chain(sb, n):
    """n = 16384 / 65536 in the reference's own Flowgraph: on provider b200 the block still creates ONE spectral_chain
    module, which runs the two fused kernels of the tiled two-pass plan (window in the column pass, amplitude / range in
    the row pass)."""
    from cyberether_b200.synthetic import spectral_rows
    cycles = [spectral_rows(5 * k, 3, n=n) for k in range(2)]
    cpu, gpu, mods, cpu_mods, _ = _both(sb, "spectrum_engine", {"enableScale": True, "rangeMin": -100.0, "rangeMax": -10.0},
                                        "buffer", "buffer", cycles, {"sampleAxis": 1, "batchAxis": 0})
    assert "runtime:spectral_chain" in mods and "runtime:fft" not in mods and "runtime:fft" in cpu_mods
    w = _window(n)
    for x, want, got in zip(cycles, cpu, gpu):
        assert_db_close(got, want, true_spectrum(x, w), scale=2.0 / 90.0, floor=3e-7)


def test_spectrum_engine_block_other_length_uses_reference_wiring_for_agc(sb):
    """enableAgc with n != 4096: the block takes the reference wiring on this provider's per-module kernels."""
    from cyberether_b200.synthetic import spectral_rows
    cycles = [spectral_rows(3, 6, n=1024)]
    cpu, gpu, mods, _, _ = _both(sb, "spectrum_engine", {"enableScale": True, "enableAgc": True}, "buffer", "buffer",
                                 cycles, {"sampleAxis": 1, "batchAxis": 0})
    assert "runtime:agc" in mods and "runtime:fft" in mods and "runtime:spectral_chain" not in mods
    assert_db_close(gpu[0], cpu[0], true_spectrum(cycles[0], _window(1024)), scale=2.0 / 120.0, floor=3e-7)


def _filter_cycles(rows, n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(3 * rows * n)
    sig = (0.4 * np.exp(2j * np.pi * 0.031 * t) + 0.2 * np.exp(-2j * np.pi * 0.127 * t)
           + 0.05 * (rng.standard_normal(t.size) + 1j * rng.standard_normal(t.size))).astype(np.complex64)
    return [sig[k * rows * n:(k + 1) * rows * n].reshape(rows, n) for k in range(3)]


@pytest.mark.parametrize("heads,center", [(1, [0.0]), (3, [0.0, 1e6, -2e6])])
def test_filter_block_fused_129_taps_decimate_8(sb, heads, center):
    """SURVEY §8d C3 shape in small: 129 taps (the block's R=8 condition needs taps-1 divisible by 8), R = 8."""
    cycles = _filter_cycles(8, 4096, 11)
    cfg = {"sampleRate": 8e6, "bandwidth": 1e6, "taps": 129, "heads": heads, "center": center}
    cpu, gpu, mods, cpu_mods, info = _both(sb, "filter", cfg, "signal", "buffer", cycles,
                                           {"sampleAxis": 1, "batchAxis": 0})
    assert set(mods) == {"runtime:filter_taps", "runtime:cast_signal", "runtime:fir"}, mods
    assert mods["runtime:fir"][0] == 3 and mods["runtime:filter_taps"][0] == 1
    assert "runtime:fold" in cpu_mods and "runtime:overlap" in cpu_mods
    assert info["b200"]["shape"] == info["generic"]["shape"] == (8, heads, 512)
    for key in ("sampleAxis", "batchAxis", "channelAxis"):
        assert info["b200"][key] == info["generic"][key]
    for k, (want, got) in enumerate(zip(cpu, gpu)):
        err = np.abs(got - want).max() / np.abs(want).max()
        assert err <= 1e-5, f"cycle {k}: {err:.3e}"


def test_filter_block_attribute_and_full_rate(sb):
    """127 taps: the block bypasses resampling exactly as the reference (block_impl.cc:64-90), R = 1."""
    cycles = _filter_cycles(4, 2048, 3)
    cfg = {"sampleRate": 8e6, "bandwidth": 1e6, "taps": 127, "heads": 1, "center": [0.0]}
    cpu, gpu, mods, _, info = _both(sb, "filter", cfg, "signal", "buffer", cycles, {"sampleAxis": 1, "batchAxis": 0})
    assert info["b200"]["shape"] == info["generic"]["shape"] == (4, 1, 2048)
    for want, got in zip(cpu, gpu):
        assert np.abs(got - want).max() / np.abs(want).max() <= 1e-5
    with sb.Session() as s:
        s.add_source("src", (4, 4096), "CF32", target=sb.B200, sampleAxis=1, batchAxis=0)
        s.add_block("f", "filter", {"sampleRate": 8e6, "bandwidth": 1e6, "taps": 129}, {"signal": "src.signal"},
                    target=sb.B200)
        assert s.attribute_f32("f", "buffer", "sampleRate") == 1e6


def test_filter_block_rows_without_batch_axis_are_independent_lanes(sb):
    """[B, T] with sampleAxis only = B independent lanes, each carrying its own tail across
    cycles (overlap_add/module_impl_native_cpu.cc:176-198) — not B consecutive frames of one stream."""
    cycles = _filter_cycles(4, 2048, 9)
    cfg = {"sampleRate": 8e6, "bandwidth": 1e6, "taps": 129, "heads": 2, "center": [0.0, 1e6]}
    cpu, gpu, mods, _, info = _both(sb, "filter", cfg, "signal", "buffer", cycles, {"sampleAxis": 1})
    assert info["b200"]["shape"] == info["generic"]["shape"]
    assert info["b200"]["batchAxis"] == info["generic"]["batchAxis"] == -1
    for k, (want, got) in enumerate(zip(cpu, gpu)):
        err = np.abs(got - want).max() / np.abs(want).max()
        assert err <= 1e-5, f"cycle {k}: {err:.3e}"


def _fm_cycles(rows, n, seed, rate):
    rng = np.random.default_rng(seed)
    t = np.arange(3 * rows * n) / rate
    audio = 0.6 * np.sin(2 * np.pi * 1000.0 * t) + 0.3 * np.sin(2 * np.pi * 19000.0 * t)
    phase = 2 * np.pi * 60e3 * np.cumsum(audio) / rate
    sig = (0.8 * np.exp(1j * phase) + 0.01 * (rng.standard_normal(t.size) + 1j * rng.standard_normal(t.size)))
    return [sig[k * rows * n:(k + 1) * rows * n].reshape(rows, n).astype(np.complex64) for k in range(3)]


@pytest.mark.parametrize("mode,deemph,tol", [("narrow", "none", 1e-5), ("narrow", "75us", 1e-5),
                                             ("wide", "none", 3e-5), ("wide", "50us", 3e-5)])
def test_fm_block(sb, mode, deemph, tol):
    cycles = _fm_cycles(4, 8192, 21, 250e3)
    cfg = {"mode": mode, "deemphasis": deemph, "sampleRate": 250e3}
    cpu, gpu, mods, _, info = _both(sb, "fm", cfg, "signal", "signal", cycles, {"sampleAxis": 1, "batchAxis": 0})
    assert set(mods) == {"runtime:fm"} and mods["runtime:fm"][0] == 3
    assert info["b200"]["shape"] == info["generic"]["shape"] == ((4, 8192, 2) if mode == "wide" else (4, 8192))
    assert info["b200"]["channelAxis"] == info["generic"]["channelAxis"]
    for k, (want, got) in enumerate(zip(cpu, gpu)):
        err = np.abs(got - want).max()
        assert err <= tol, f"cycle {k}: {err:.3e}"


def test_unfused_modules_still_reachable(sb):
    """The per-module kernels stay registered: an `fft` block and an `amplitude` block on provider b200."""
    from cyberether_b200.synthetic import gaussian_cf32
    x = gaussian_cf32((4, 1024), 77)
    outs = {}
    for target in (sb.CPU, sb.B200):
        with sb.Session() as s:
            s.add_source("src", x.shape, "CF32", target=target, sampleAxis=1, batchAxis=0)
            s.add_block("f", "fft", {"forward": True}, {"signal": "src.signal"}, target=target)
            s.add_block("a", "amplitude", None, {"signal": "f.signal"}, target=target)
            s.write_source("src", x)
            s.compute()
            outs[target[1]] = (s.read("f", "signal"), s.read("a", "signal"))
    want, got = outs["generic"], outs["b200"]
    assert np.abs(got[0] - want[0]).max() <= 2e-6 * np.abs(want[0]).max()
    assert np.abs(got[1] - want[1]).max() <= 1e-3


@pytest.mark.parametrize("n", [1024, 8192, 65536])
@pytest.mark.parametrize("complex_output", [True, False])
def test_real_input_fft_block_on_provider_b200(sb, n, complex_output):
    """The reference's F32 `fft` cases (module_benchmarks.cc: F32-8192 / F32-65536): forward transform of real rows,
    pocketfft::r2c (`complexOutput`) or FFTPACK half-complex, in the reference's own Flowgraph on the CPU provider and on
    provider b200 (b200_fft_exec_real: one half-length complex transform + unpack)."""
    rng = np.random.default_rng(n + int(complex_output))
    x = rng.standard_normal((3, n)).astype(np.float32)
    outs = {}
    for target in (sb.CPU, sb.B200):
        with sb.Session() as s:
            s.add_source("src", x.shape, "F32", target=target, sampleAxis=1, batchAxis=0)
            s.add_block("f", "fft", {"forward": True, "complexOutput": complex_output}, {"signal": "src.signal"}, target=target)
            s.write_source("src", x)
            s.compute()
            outs[target[1]] = s.read("f", "signal")
    want, got = outs["generic"], outs["b200"]
    assert got.shape == want.shape == ((3, n // 2 + 1) if complex_output else (3, n)) and got.dtype == want.dtype
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()


@pytest.mark.parametrize("rows,decimation", [(48, 1), (700, 2)])
def test_spectrum_analyzer_flowgraph_with_consumers(sb, rows, decimation):
    """examples/flowgraphs/spectrum-analyzer.yml in small: source -> spectrum_engine -> lineplot + waterfall, every
    block on provider b200, next to the same flowgraph on the reference CPU provider. On b200 the lineplot takes its
    batch sums from the chain kernel's epilogue (attribute "b200.columnSums"); only signalPoints [n, 2] and the ring
    [height, 4096] would have to leave the device."""
    from cyberether_b200.synthetic import spectral_rows
    cycles = [spectral_rows(17 * k, rows) for k in range(3)]
    state = {}
    for target in (sb.CPU, sb.B200):
        with sb.Session() as s:
            s.add_source("src", (rows, 4096), "CF32", target=target, sampleAxis=1, batchAxis=0)
            s.add_block("spec", "spectrum_engine", {"enableScale": True}, {"buffer": "src.signal"}, target=target)
            s.add_block("lp", "lineplot", {"averaging": 2, "decimation": decimation}, {"signal": "spec.buffer"}, target=target)
            s.add_block("wf", "waterfall", {"height": 64}, {"signal": "spec.buffer"}, target=target)
            assert sb.viz_modules() == {"lp-lineplot": "lineplot", "wf-waterfall": "waterfall"}
            per_cycle = []
            for x in cycles:
                s.write_source("src", x)
                s.compute()
                per_cycle.append((sb.viz_read("lp-lineplot").reshape(-1, 2), sb.viz_read("wf-waterfall").reshape(64, 4096),
                                  sb.viz_write_index("wf-waterfall"), s.read("spec", "buffer")))
            state[target[1]] = per_cycle
            mods = {**s.modules("spec"), **s.modules("lp"), **s.modules("wf")}
            if target is sb.B200:
                assert mods["runtime:spectral_chain"][0] == 3 and mods["runtime:lineplot"][0] == 3 \
                    and mods["runtime:waterfall"][0] == 3
    from oracle import port
    expected_ring = port.Waterfall(4096, 64)
    for (lp_c, wf_c, wi_c, spec_c), (lp_g, wf_g, wi_g, spec_g) in zip(state["generic"], state["b200"]):
        assert wi_c == wi_g
        assert np.array_equal(wf_g, expected_ring.compute(spec_g))     # the ring = copies of the block's own spectra
        assert np.array_equal(lp_c[:, 0], lp_g[:, 0])
        # spectra agree within the chain allowance; their batch mean mapped to [-1, 1] within 2e-4
        assert np.abs(lp_c[:, 1] - lp_g[:, 1]).max() <= 2e-4
        assert np.abs(wf_c - wf_g).max() <= 2e-3


def test_backend_rows_a12_a14_through_the_reference_interfaces(sb):
    """in this library the reference's CUDA buffer backend and CUDA runtime are REPLACED by
    shim/b200_buffer.cc (detail::Backend over b200_malloc / b200_malloc_managed / b200_host_register / b200_memcpy) and
    shim/b200_runtime.cc (Runtime::Impl over b200_stream_* / b200_event_*) — every other test of this file already runs
    on them. Here: (1) fresh CUDA tensors read back as zeros (buffer_cuda.cc:118-121), (2) a CPU tensor mapped onto the
    device (TestContext's input path) feeds a b200 module zero-copy, (3) per-module timing metrics advance per cycle and
    skip settled static modules."""
    from cyberether_b200.synthetic import gaussian_cf32
    x = gaussian_cf32((16, 1024), 5)
    with sb.Session() as s:
        s.add_source("src", x.shape, "CF32", target=sb.B200, sampleAxis=1, batchAxis=0, mapped=True)
        s.add_block("amp", "amplitude", None, {"signal": "src.signal"}, target=sb.B200)
        assert not s.read("amp", "signal").any()                    # zero-filled allocation, nothing computed yet
        assert s.info("src", "signal")["device"] == "cuda"
        s.write_source("src", x)
        s.compute()
        got = s.read("amp", "signal")
        cycles0 = s.modules("amp")["runtime:amplitude"]
        s.write_source("src", 2 * x)
        s.compute()
        got2 = s.read("amp", "signal")
        cycles1 = s.modules("amp")["runtime:amplitude"]
    with sb.Session() as s:
        s.add_source("src", x.shape, "CF32", target=sb.CPU, sampleAxis=1, batchAxis=0)
        s.add_block("amp", "amplitude", None, {"signal": "src.signal"}, target=sb.CPU)
        s.write_source("src", x)
        s.compute()
        want = s.read("amp", "signal")
    assert np.array_equal(got, want)                                  # amplitude is bit-exact on this provider
    assert np.allclose(got2 - got, 20 * np.log10(2.0), atol=5e-3)
    assert cycles0[0] == 1 and cycles1[0] == 2 and cycles1[1] >= cycles0[1] > 0.0


def test_backend_singleton_is_ours_and_uses_the_primary_context(sb):
    """SURVEY §8 a15 / north-star "src/backend (CUDA device)": Backend::CUDA of this library is shim/b200_backend.cc
    (device table from b200_ctx_info). It retains the device's PRIMARY context instead of creating a second one, so
    memory the host application allocated with the CUDA runtime (here: PyTorch) is valid inside the flowgraph."""
    import torch
    info = sb.backend_info()
    props = torch.cuda.get_device_properties(info["device"])
    assert info["name"] == props.name and "B200" in info["name"]
    assert info["compute_capability"] == f"{props.major}{props.minor}" == "100"
    assert info["memory_bytes"] == props.total_memory
    assert info["primary_context"] is True
    # a PyTorch tensor's device pointer handed to libb200dsp from inside a harness-created stream context still works
    # after flowgraph calls (one context): run a block, then a torch kernel, then read both
    x = (np.random.default_rng(0).standard_normal((4, 1024)) + 0j).astype(np.complex64)
    with sb.Session() as s:
        s.add_source("src", x.shape, "CF32", target=sb.B200, sampleAxis=1, batchAxis=0)
        s.add_block("a", "amplitude", None, {"signal": "src.signal"}, target=sb.B200)
        s.write_source("src", x)
        s.compute()
        t = torch.arange(8, device="cuda") * 2
        assert int(t.sum()) == 56
        assert np.isfinite(s.read("a", "signal")).all()
