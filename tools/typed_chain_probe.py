"""Ad-hoc timing of the fused complex-integer ingest (b200_chain_exec_typed) vs cast + CF32 chain."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cyberether_b200 as cb
from cyberether_b200 import _native
from cyberether_b200.jetstream import Context, DTYPE_CODES

lib = _native.load(); dev = torch.device("cuda:0"); ctx = Context.get(dev)
rows, n = 65536, 4096
win = torch.zeros(n, dtype=torch.complex64, device=dev)
_native.check(lib.b200_window_blackman_cf32(ctx.handle, win.data_ptr(), n, None)); torch.cuda.synchronize()
sign = torch.ones(n, device=dev); sign[1::2] = -1; win = (win * sign).contiguous()
plan = ctypes.c_void_p(); _native.check(lib.b200_chain_plan_create(ctx.handle, n, rows, win.data_ptr(), ctypes.byref(plan)))
coeff = cb.amplitude_scaling_coeff(n); sc, off = cb.range_coefficients(-120.0, 0.0)
s = torch.cuda.current_stream().cuda_stream
out = torch.empty(rows, n, dtype=torch.float32, device=dev)
xf = torch.empty(rows, n, dtype=torch.complex64, device=dev)
for name, tdt, nbytes in (("CI8", torch.int8, 2), ("CI16", torch.int16, 4)):
    x = torch.randint(-100, 100, (rows, n, 2), device=dev, dtype=torch.int32).to(tdt)
    def fused(): _native.check(lib.b200_chain_exec_typed(plan, x.data_ptr(), DTYPE_CODES[name], out.data_ptr(), rows, coeff, 1, sc, off, s))
    def two():
        _native.check(lib.b200_cast_int(ctx.handle, x.data_ptr(), DTYPE_CODES[name], xf.data_ptr(), rows * n, s))
        _native.check(lib.b200_chain_exec(plan, xf.data_ptr(), out.data_ptr(), rows, coeff, 1, sc, off, s))
    for label, fn, alg in (("fused", fused, nbytes + 4), ("cast+chain", two, nbytes + 8 + 8 + 4)):
        for _ in range(3): fn()
        torch.cuda.synchronize(); ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts)); gs = rows * n / ms * 1e-6
        print(f"{name} {label}: {ms:.4f} ms {gs:.1f} GS/s  {gs*alg:.0f} GB/s moved ({alg} B/sample)")
