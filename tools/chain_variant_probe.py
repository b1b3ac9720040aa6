"""A/B of the CF32 chain kernels in one process: fft4096_kernel (classic: three CTA barriers per row) against
fft4096w_kernel (warp-local first exchange, swizzled 2-D TMA landing). Same arithmetic per element, so the outputs must
be BIT-IDENTICAL; timing by CUDA events over the full 65536 x 4096 batch.
usage: python tools/chain_variant_probe.py [rows]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cyberether_b200 as cb
from cyberether_b200 import _native
from cyberether_b200.jetstream import Context

lib = _native.load()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n = 4096
dev = torch.device("cuda:0")
ctx = Context.get(dev)
g = torch.Generator(device=dev); g.manual_seed(1)
x = torch.view_as_complex(torch.randn(rows, n, 2, device=dev, generator=g) * 0.01)
x[3, 17] = 0; x[5] = 0                       # exact zeros: -inf / 0 outputs
win = torch.zeros(n, dtype=torch.complex64, device=dev)
_native.check(lib.b200_window_blackman_cf32(ctx.handle, win.data_ptr(), n, None))
torch.cuda.synchronize()
sign = torch.ones(n, device=dev); sign[1::2] = -1
win = (win * sign).contiguous()
plan = ctypes.c_void_p()
_native.check(lib.b200_chain_plan_create(ctx.handle, n, rows, win.data_ptr(), ctypes.byref(plan)))
coeff = cb.amplitude_scaling_coeff(n); sc, off = cb.range_coefficients(-120.0, 0.0)
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
DT = 1                                        # B200_DTYPE_CF32
outs = {v: torch.empty(rows, n, dtype=torch.float32, device=dev) for v in ("classic", "w")}
colsum = {v: torch.empty(n, dtype=torch.float32, device=dev) for v in ("classic", "w")}


def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


def same(a, b):
    return bool(torch.equal(a.view(torch.int32), b.view(torch.int32)))


for label, call in (
    ("chain range", lambda o, cs: lib.b200_chain_exec(plan, x.data_ptr(), o.data_ptr(), rows, coeff, 1, sc, off, s)),
    ("chain dB", lambda o, cs: lib.b200_chain_exec(plan, x.data_ptr(), o.data_ptr(), rows, coeff, 0, sc, off, s)),
    ("chain range + agc", lambda o, cs: lib.b200_chain_exec_agc(plan, x.data_ptr(), DT, o.data_ptr(), rows, coeff, 1, sc, off, 1.0, 1e-12, 1e-3, 1e3, s)),
    ("chain range + colsum", lambda o, cs: lib.b200_chain_exec_colsum(plan, x.data_ptr(), DT, o.data_ptr(), rows, coeff, 1, sc, off, cs.data_ptr(), s)),
):
    res = {}
    for variant in ("classic", "w"):
        os.environ["B200_FFT4096_VARIANT"] = variant
        outs[variant].fill_(7.0)
        fn = lambda v=variant: _native.check(call(outs[v], colsum[v]))
        res[variant] = timeit(fn)
    ident = same(outs["classic"], outs["w"]) and ("colsum" not in label or same(colsum["classic"], colsum["w"]))
    for variant in ("classic", "w"):
        ms, lo = res[variant]
        gs = rows * n / ms * 1e-6
        print(f"{label:22s} {variant:8s} {ms:.4f} ms (min {lo:.4f})  {gs:7.1f} GS/s  {gs*12/6570.9*100:5.1f}% of measured HBM", flush=True)
    print(f"{label:22s} outputs bit-identical: {ident}", flush=True)
    if not ident:
        d = (outs["classic"] - outs["w"]).abs()
        d = torch.where(torch.isfinite(d), d, torch.zeros_like(d))
        print("   max |diff|", float(d.max()), "rows differing", int((d.amax(dim=1) > 0).sum()))
# 200-launch sustained figure for the default call
for variant in ("classic", "w"):
    os.environ["B200_FFT4096_VARIANT"] = variant
    fn = lambda: _native.check(lib.b200_chain_exec(plan, x.data_ptr(), outs[variant].data_ptr(), rows, coeff, 1, sc, off, s))
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"sustained 200 launches {variant:8s} {e0.elapsed_time(e1)/200:.4f} ms", flush=True)
