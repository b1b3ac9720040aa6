"""Real-input forward transforms (the reference's F32-8192 / F32-65536 fft benchmark cases): the composed path
(cast -> full C2C -> pack, 40 B per real sample) against b200_fft_exec_real (half-length C2C on the row itself + unpack,
16 B per real sample) and cuFFT's R2C (torch.fft.rfft)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cyberether_b200 import _native
from cyberether_b200.jetstream import Context

lib = _native.load(); dev = torch.device("cuda:0"); ctx = Context.get(dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


total = 1 << 27                                    # real samples
g = torch.Generator(device=dev); g.manual_seed(3)
x = torch.randn(total, device=dev, generator=g)
for n in (8192, 65536):
    rows = total // n
    xx = x.reshape(rows, n)
    want = torch.fft.rfft(xx[:8])
    full = ctypes.c_void_p(); _native.check(lib.b200_fft_plan_c2c(ctx.handle, n, rows, ctypes.byref(full)))
    half = ctypes.c_void_p(); _native.check(lib.b200_fft_plan_c2c(ctx.handle, n // 2, rows, ctypes.byref(half)))
    work = torch.empty(rows, n, dtype=torch.complex64, device=dev)
    out = torch.empty(rows, n // 2 + 1, dtype=torch.complex64, device=dev)

    def composed():
        _native.check(lib.b200_cast_f32_cf32(ctx.handle, xx.data_ptr(), work.data_ptr(), rows * n, sp))
        _native.check(lib.b200_fft_exec(full, work.data_ptr(), work.data_ptr(), 1, sp))
        _native.check(lib.b200_fft_real_helper(ctx.handle, 0, work.data_ptr(), out.data_ptr(), rows, n, sp))

    def halved():
        _native.check(lib.b200_fft_exec_real(half, xx.data_ptr(), out.data_ptr(), 0, sp))

    for label, fn in (("composed (cast, full C2C, pack)", composed), ("b200_fft_exec_real (half C2C + unpack)", halved)):
        out.zero_(); fn(); torch.cuda.synchronize()
        err = (out[:8] - want).abs().max().item() / want.abs().max().item()
        ms = timeit(fn)
        print(f"rfft {n} x {rows}: {label:40s} {ms:8.4f} ms {total/ms*1e-6:8.1f} Greal/s  err {err:.1e}", flush=True)
    ms = timeit(lambda: torch.fft.rfft(xx))
    print(f"rfft {n} x {rows}: {'cuFFT R2C (torch.fft.rfft)':40s} {ms:8.4f} ms {total/ms*1e-6:8.1f} Greal/s", flush=True)
    del work, out
