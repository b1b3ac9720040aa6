"""Large power-of-two C2C transforms: the two-pass plan (fft_twopass.cuh) against the four-step plan it replaces, the
single-CTA kernel (8192) and cuFFT (torch.fft) on the same tensors; sweeps the chunk size (= L2-resident scratch) and the
L2 hints. Also checks every variant against cuFFT (max error relative to the largest bin).
usage: python tools/fft_large_probe.py [quick]"""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
QUICK = len(sys.argv) > 1 and sys.argv[1] == "quick"


def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def report(name, ms, samples):
    gbs = samples * 16 / ms * 1e-6
    print(f"{name:58s} {ms:8.4f} ms {samples/ms*1e-6:8.1f} GS/s {gbs:7.0f} GB/s {100*gbs/PEAK:5.1f}% of measured HBM", flush=True)


def plan_with(env, n, rows):
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        pl = ctypes.c_void_p()
        _native.check(lib.b200_fft_plan_c2c(ctx.handle, n, rows, ctypes.byref(pl)))
        return pl
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


MODE = sys.argv[1] if len(sys.argv) > 1 else ""
total = 1 << 26
g = torch.Generator(device=dev); g.manual_seed(0)
x = torch.view_as_complex(torch.randn(total, 2, device=dev, generator=g))
y = torch.empty_like(x)
if MODE == "ncu":      # a few launches of each kernel for an ncu capture (run with --cache-control none)
    for n, env in ((8192, {"B200_FFT_TWOPASS": 0}), (16384, {}), (65536, {})):
        rows = total // n
        pl = plan_with(env, n, rows)
        for _ in range(2):
            _native.check(lib.b200_fft_exec(pl, x.data_ptr(), y.data_ptr(), 1, sp))
        torch.cuda.synchronize()
    sys.exit(0)
for n in (8192, 16384, 32768, 65536, 131072):
    rows = total // n
    xx, yy = x.reshape(rows, n), y.reshape(rows, n)
    want = torch.fft.fft(xx[:8])
    scale = want.abs().max().item()
    variants = []
    if n == 8192:
        variants.append(("single-CTA radix kernel", {"B200_FFT_TWOPASS": 0}))
    else:
        if not QUICK:
            variants.append(("four-step plan (round 1)", {"B200_FFT_TWOPASS": 0}))
        if n > 65536 or not QUICK:
            variants.append(("two-pass col16 chunk 64 MB", {"B200_FFT_TWOPASS_TILE": 0, "B200_FFT_TWOPASS_CHUNK_MB": 64}))
        if n <= 65536:
            for mb in ((96,) if QUICK else (32, 64, 96, 128, 192)):
                variants.append((f"two-pass tiled chunk {mb} MB", {"B200_FFT_TWOPASS_CHUNK_MB": mb}))
            variants.append(("two-pass tiled chunk 96 MB no PDL", {"B200_FFT_TWOPASS_CHUNK_MB": 96, "B200_FFT_TWOPASS_PDL": 0}))
            variants.append(("two-pass tiled chunk 96 MB no hints", {"B200_FFT_TWOPASS_CHUNK_MB": 96, "B200_FFT_TWOPASS_HINTS": 0}))
            variants.append(("two-pass tiled one chunk", {"B200_FFT_TWOPASS_CHUNK_MB": 4096}))
    for label, env in variants:
        pl = plan_with(env, n, rows)
        run = lambda f=1: _native.check(lib.b200_fft_exec(pl, xx.data_ptr(), yy.data_ptr(), f, sp))
        run(); torch.cuda.synchronize()
        err = (yy[:8] - want).abs().max().item() / scale
        run(0); torch.cuda.synchronize()
        inv = torch.fft.ifft(xx[rows - 4:], norm="forward")
        err_inv = (yy[rows - 4:] - inv).abs().max().item() / inv.abs().max().item()
        report(f"fft c2c {n} x {rows}: {label} [err {err:.1e}/{err_inv:.1e}]", timeit(run), total)
        _native.check(lib.b200_fft_plan_destroy(pl))
    report(f"cuFFT {n} x {rows} [baseline]", timeit(lambda: torch.fft.fft(xx)), total)
