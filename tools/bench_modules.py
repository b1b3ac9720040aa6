"""Secondary measurements (not the bench.py contract): per-module kernels and the FIR / FM configs of
BASELINE.json, device-resident inputs, CUDA events, achieved GB/s against the measured HBM peak."""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cyberether_b200 as cb
from cyberether_b200 import _native
from cyberether_b200.jetstream import Context

lib = _native.load()
dev = torch.device("cuda:0")
ctx = Context.get(dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0

def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

results = []
def report(name, ms, samples, bytes_per_sample):
    gbs = samples * bytes_per_sample / ms * 1e-6
    results.append(dict(kernel=name, ms=ms, msamples_s=samples / ms * 1e-3, gbs=gbs, frac_of_measured_peak=gbs / PEAK,
                        algorithmic_bytes_per_sample=bytes_per_sample))
    print(f"{name:46s} {ms:8.4f} ms {samples/ms*1e-6:9.1f} GS/s {gbs:8.0f} GB/s {100*gbs/PEAK:5.1f}% of measured HBM")

rows, n = 16384, 4096
g = torch.Generator(device=dev); g.manual_seed(0)
x = torch.view_as_complex(torch.randn(rows, n, 2, device=dev, generator=g))
w = torch.view_as_complex(torch.randn(n, 2, device=dev, generator=g))
y = torch.empty_like(x); f = torch.empty(rows, n, device=dev); f2 = torch.empty_like(f)
shape = (ctypes.c_uint64 * 2)(rows, n); sa = (ctypes.c_uint64 * 2)(n, 1); sb = (ctypes.c_uint64 * 2)(0, 1)
report("multiply cf32 [16384,4096] x [1,4096]", timeit(lambda: _native.check(lib.b200_multiply_cf32(ctx.handle, x.data_ptr(), w.data_ptr(), y.data_ptr(), 2, shape, sa, sb, sp))), rows * n, 16)
plan = ctypes.c_void_p(); _native.check(lib.b200_fft_plan_c2c(ctx.handle, n, rows, ctypes.byref(plan)))
report("fft c2c 4096 x 16384", timeit(lambda: _native.check(lib.b200_fft_exec(plan, x.data_ptr(), y.data_ptr(), 1, sp))), rows * n, 16)
report("cuFFT (torch.fft.fft) 4096 x 16384 [baseline]", timeit(lambda: torch.fft.fft(x)), rows * n, 16)
PLAN = {16384: "tiled two-pass plan 64 x 256", 32768: "tiled two-pass plan 128 x 256", 65536: "tiled two-pass plan 256 x 256",
        131072: "two-pass plan 16 x 8192"}
for nn in (256, 1024, 2048, 8192, 16384, 32768, 65536, 131072):
    rr = rows * n // nn
    xx = x.reshape(rr, nn); yy = y.reshape(rr, nn)
    pl = ctypes.c_void_p(); _native.check(lib.b200_fft_plan_c2c(ctx.handle, nn, rr, ctypes.byref(pl)))
    report(f"fft c2c {nn} x {rr} ({PLAN.get(nn, 'radix-16 kernel')})", timeit(lambda: _native.check(lib.b200_fft_exec(pl, xx.data_ptr(), yy.data_ptr(), 1, sp))), rr * nn, 16)
    report(f"cuFFT {nn} x {rr} [baseline]", timeit(lambda: torch.fft.fft(xx)), rr * nn, 16)
    lib.b200_fft_plan_destroy(pl)
# real input (the reference's F32-8192 / F32-65536 benchmark cases): half-length complex transform + unpack; 4 B in +
# 4 B out per real sample
xr = torch.view_as_real(x).reshape(-1)
for nn in (8192, 65536):
    rr = xr.numel() // nn
    xx = xr.reshape(rr, nn); oo = torch.empty(rr, nn // 2 + 1, dtype=torch.complex64, device=dev)
    pl = ctypes.c_void_p(); _native.check(lib.b200_fft_plan_c2c(ctx.handle, nn // 2, rr, ctypes.byref(pl)))
    report(f"fft r2c {nn} x {rr} ({'fused MODE_R2C kernel' if nn // 2 <= 8192 else 'half-length c2c + unpack'})", timeit(lambda: _native.check(lib.b200_fft_exec_real(pl, xx.data_ptr(), oo.data_ptr(), 0, sp))), rr * nn, 8)
    report(f"cuFFT r2c {nn} x {rr} [baseline]", timeit(lambda: torch.fft.rfft(xx)), rr * nn, 8)
    lib.b200_fft_plan_destroy(pl); del oo
coeff = cb.amplitude_scaling_coeff(n)
report("amplitude cf32", timeit(lambda: _native.check(lib.b200_amplitude_cf32(ctx.handle, y.data_ptr(), f.data_ptr(), rows * n, coeff, sp))), rows * n, 12)
sc, off = cb.range_coefficients(-120.0, 0.0)
report("range f32", timeit(lambda: _native.check(lib.b200_range_f32(ctx.handle, f.data_ptr(), f2.data_ptr(), rows * n, sc, off, sp))), rows * n, 8)
# unfused chain = the reference CUDA path's structure (multiply -> fft -> amplitude -> range), our kernels
def unfused():
    _native.check(lib.b200_multiply_cf32(ctx.handle, x.data_ptr(), w.data_ptr(), y.data_ptr(), 2, shape, sa, sb, sp))
    _native.check(lib.b200_fft_exec(plan, y.data_ptr(), y.data_ptr(), 1, sp))
    _native.check(lib.b200_amplitude_cf32(ctx.handle, y.data_ptr(), f.data_ptr(), rows * n, coeff, sp))
    _native.check(lib.b200_range_f32(ctx.handle, f.data_ptr(), f2.data_ptr(), rows * n, sc, off, sp))
report("unfused chain, 4 kernels (our modules)", timeit(unfused), rows * n, 12)
def unfused_cufft():
    t = torch.fft.fft(x * w)
    _native.check(lib.b200_amplitude_cf32(ctx.handle, t.data_ptr(), f.data_ptr(), rows * n, coeff, sp))
    _native.check(lib.b200_range_f32(ctx.handle, f.data_ptr(), f2.data_ptr(), rows * n, sc, off, sp))
report("unfused chain with cuFFT (reference CUDA structure)", timeit(unfused_cufft), rows * n, 12)
win = torch.zeros(n, dtype=torch.complex64, device=dev); win.real = torch.rand(n, device=dev)
cp = ctypes.c_void_p(); torch.cuda.synchronize(); _native.check(lib.b200_chain_plan_create(ctx.handle, n, rows, win.data_ptr(), ctypes.byref(cp)))
report("fused chain 4096 x 16384", timeit(lambda: _native.check(lib.b200_chain_exec(cp, x.data_ptr(), f.data_ptr(), rows, coeff, 1, sc, off, sp))), rows * n, 12)
for nn in (8192, 16384, 65536):
    rr = rows * n // nn
    wn = torch.zeros(nn, dtype=torch.complex64, device=dev); wn.real = torch.rand(nn, device=dev)
    cpn = ctypes.c_void_p(); torch.cuda.synchronize(); _native.check(lib.b200_chain_plan_create(ctx.handle, nn, rr, wn.data_ptr(), ctypes.byref(cpn)))
    cn = cb.amplitude_scaling_coeff(nn)
    report(f"fused chain {nn} x {rr}", timeit(lambda: _native.check(lib.b200_chain_exec(cpn, x.data_ptr(), f.data_ptr(), rr, cn, 1, sc, off, sp))), rr * nn, 12)
    lib.b200_chain_plan_destroy(cpn)

# ---- FIR (BASELINE config 3: 2^26 CF32 samples as [8192, 8192] frames)
frames, T = 8192, 8192
xs = torch.view_as_complex(torch.randn(frames, T, 2, device=dev, generator=g))
for taps, R in ((127, 8), (127, 16), (127, 4), (127, 1)):
    centers = (ctypes.c_double * 1)(0.0)
    host = np.zeros((1, taps), np.complex64)
    _native.check(lib.b200_filter_taps_host(8e6, 1e6, centers, 1, taps, host.ctypes.data_as(ctypes.c_void_p)))
    fp = ctypes.c_void_p(); _native.check(lib.b200_fir_plan_create(ctx.handle, host.ctypes.data_as(ctypes.c_void_p), taps, 1, R, ctypes.byref(fp)))
    yo = torch.empty(frames, 1, T // R, dtype=torch.complex64, device=dev)
    report(f"fir {taps} taps decimate {R}, 2^26 samples", timeit(lambda: _native.check(lib.b200_fir_exec(fp, xs.data_ptr(), yo.data_ptr(), frames, T, sp)), iters=10, warm=3), frames * T, 8 + 8 / R)
# ---- integer ingest (cast module; fused into the chain)
from cyberether_b200.jetstream import DTYPE_CODES
for name, tdt, nb in (("CI8", torch.int8, 2), ("CI16", torch.int16, 4)):
    xi = torch.randint(-100, 100, (rows, n, 2), device=dev, dtype=torch.int32).to(tdt)
    report(f"cast {name} -> CF32", timeit(lambda: _native.check(lib.b200_cast_int(ctx.handle, xi.data_ptr(), DTYPE_CODES[name], y.data_ptr(), rows * n, sp))), rows * n, nb + 8)
    report(f"fused chain 4096 x 16384, {name} input", timeit(lambda: _native.check(lib.b200_chain_exec_typed(cp, xi.data_ptr(), DTYPE_CODES[name], f.data_ptr(), rows, coeff, 1, sc, off, sp))), rows * n, nb + 4)
# ---- agc (CF32, one tile per 4096-sample row, as spectrum_engine uses it; and 1024-sample tiles)
need = ctypes.c_uint64()
for tile in (4096, 1024):
    _native.check(lib.b200_agc_scratch_bytes(rows, n, tile, ctypes.byref(need)))
    scratch = torch.empty(need.value, dtype=torch.uint8, device=dev)
    report(f"agc cf32 [16384,4096], tile {tile}", timeit(lambda: _native.check(lib.b200_agc(ctx.handle, x.data_ptr(), y.data_ptr(), 1, rows, n, tile, 1.0, 1e-12, 0.01, 100.0, 4.0, scratch.data_ptr(), sp))), rows * n, 16)
# ---- FM (2^26 samples as [8192, 1, 8192])
fr, fl = 8192, 8192
xf = xs.reshape(fr, 1, fl); of = torch.empty(fr, 1, fl, device=dev)
for wide, de, label in ((0, 0, "narrow"), (0, 75, "narrow, de-emphasis 75us"), (1, 0, "wide (stereo)"), (1, 75, "wide (stereo), de-emphasis 75us")):
    mp_ = ctypes.c_void_p(); _native.check(lib.b200_fm_plan_create(ctx.handle, 1, ctypes.c_float(250e3), wide, de, ctypes.byref(mp_)))
    if wide:
        frw = 64        # the stereo decoder is a chain of scans with serial parts: 2^19 samples; out = [.., 2] (L, R)
        ofw = torch.empty(frw, 1, fl, 2, device=dev)
        report(f"fm {label}, 2^19 samples", timeit(lambda: _native.check(lib.b200_fm_exec(mp_, xf.data_ptr(), ofw.data_ptr(), frw, fl, sp)), iters=5, warm=2), frw * fl, 16)
    else:
        report(f"fm {label}, 2^26 samples", timeit(lambda: _native.check(lib.b200_fm_exec(mp_, xf.data_ptr(), of.data_ptr(), fr, fl, sp)), iters=10, warm=3), fr * fl, 12)
    lib.b200_fm_plan_destroy(mp_)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(peak_gbs=PEAK, results=results), open("gpurun_out/bench_modules.json", "w"), indent=1)
