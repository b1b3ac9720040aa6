"""Ad-hoc GPU timing of the fused chain (development aid; bench.py is the contract)."""
import ctypes, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import cyberether_b200 as cb
from cyberether_b200 import _native
from cyberether_b200.jetstream import Context

lib = _native.load()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
ctx = Context.get(dev)
g = torch.Generator(device=dev); g.manual_seed(1)
x = torch.view_as_complex(torch.randn(rows, n, 2, device=dev, generator=g) * 0.01)
win = torch.zeros(n, dtype=torch.complex64, device=dev)
_native.check(lib.b200_window_blackman_cf32(ctx.handle, win.data_ptr(), n, None))
torch.cuda.synchronize()
sign = torch.ones(n, device=dev); sign[1::2] = -1
win = (win * sign).contiguous()
out = torch.empty(rows, n, dtype=torch.float32, device=dev)
plan = ctypes.c_void_p()
_native.check(lib.b200_chain_plan_create(ctx.handle, n, rows, win.data_ptr(), ctypes.byref(plan)))
coeff = cb.amplitude_scaling_coeff(n); sc, off = cb.range_coefficients(-120.0, 0.0)
s = torch.cuda.current_stream().cuda_stream
def run(rng=1):
    _native.check(lib.b200_chain_exec(plan, x.data_ptr(), out.data_ptr(), rows, coeff, rng, sc, off, s))
for rng in (1, 0):
    for _ in range(5): run(rng)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(20):
        e0.record(); run(rng); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    gs = rows * n / ms * 1e-6
    print(f"chain n={n} rows={rows} range={rng}: {ms:.4f} ms  {gs:.1f} GS/s  {gs*12:.0f} GB/s ({gs*12/6570.9*100:.1f}% of measured HBM)  min {min(ts):.4f}")
# torch reference check
ref = torch.fft.fft(x[:64] * win)
db = 20 * torch.log10(ref.abs()) + coeff
run(0); torch.cuda.synchronize()
print("max |dB - exact log| (approx-poly error expected ~5e-3 dB):", float((out[:64] - db).abs().max()))
# standalone FFT module
fplan = ctypes.c_void_p()
_native.check(lib.b200_fft_plan_c2c(ctx.handle, n, rows, ctypes.byref(fplan)))
y = torch.empty_like(x)
for _ in range(3): _native.check(lib.b200_fft_exec(fplan, x.data_ptr(), y.data_ptr(), 1, s))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): _native.check(lib.b200_fft_exec(fplan, x.data_ptr(), y.data_ptr(), 1, s))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"fft c2c: {ms:.4f} ms {rows*n/ms*1e-6:.1f} GS/s {rows*n*16/ms*1e-6:.0f} GB/s")
err = (y[:64] - torch.fft.fft(x[:64])).abs().max() / torch.fft.fft(x[:64]).abs().max()
print("fft rel err vs torch.fft:", float(err))
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
for _ in range(3): torch.fft.fft(x)
torch.cuda.synchronize(); t0.record()
for _ in range(10): torch.fft.fft(x)
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / 10
print(f"cuFFT (torch.fft.fft) c2c: {ms:.4f} ms {rows*n/ms*1e-6:.1f} GS/s")
