"""Times the f1 consumer kernels at BASELINE configs[1] size: chain alone, chain + fused column sums, stand-alone
lineplot (row-split column sums) and waterfall, 20 launches each after warm-up, CUDA events."""
import ctypes
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cyberether_b200 import _native, amplitude_scaling_coeff, range_coefficients
from cyberether_b200.jetstream import Context

dev = torch.device("cuda", 0)
lib = _native.load()
ctx = Context.get(dev)
rows, n = 65536, 4096
x = torch.view_as_complex(torch.randn(rows, n, 2, device=dev) * 1e-2).contiguous()
out = torch.empty(rows, n, dtype=torch.float32, device=dev)
colsum = torch.empty(n, dtype=torch.float32, device=dev)
stream = torch.cuda.current_stream(dev)
sp = ctypes.c_void_p(stream.cuda_stream)
win = torch.empty(n, dtype=torch.complex64, device=dev)
winv = torch.empty(n, dtype=torch.complex64, device=dev)
_native.check(lib.b200_window_blackman_cf32(ctx.handle, win.data_ptr(), n, sp))
_native.check(lib.b200_invert_cf32(ctx.handle, win.data_ptr(), winv.data_ptr(), 1, n, 1, sp))
torch.cuda.synchronize()
plan = ctypes.c_void_p()
_native.check(lib.b200_chain_plan_create(ctx.handle, n, rows, winv.data_ptr(), ctypes.byref(plan)))
coeff = amplitude_scaling_coeff(n)
scale, offset = range_coefficients(-120.0, 0.0)
points = torch.empty(n, 2, dtype=torch.float32, device=dev)
average = torch.empty(n, dtype=torch.float32, device=dev)
need = ctypes.c_uint64()
_native.check(lib.b200_lineplot_scratch_bytes(rows, n, 1, ctypes.byref(need)))
scratch = torch.empty(need.value, dtype=torch.uint8, device=dev)
ring = torch.zeros(512, n, dtype=torch.float32, device=dev)
_native.check(lib.b200_lineplot_init(ctx.handle, points.data_ptr(), average.data_ptr(), n, sp))


def timed(label, fn, bytes_moved, reps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{label:58s} {ms:8.4f} ms  {bytes_moved / ms / 1e6:8.1f} GB/s")


timed("chain [65536,4096] CF32 -> range", lambda: _native.check(lib.b200_chain_exec(
    plan, x.data_ptr(), out.data_ptr(), rows, coeff, 1, scale, offset, sp)), rows * n * 12)
timed("chain + fused column sums (b200_chain_exec_colsum)", lambda: _native.check(lib.b200_chain_exec_colsum(
    plan, x.data_ptr(), 1, out.data_ptr(), rows, coeff, 1, scale, offset, colsum.data_ptr(), sp)), rows * n * 12)
timed("lineplot stand-alone (row-split column sums + finalize)", lambda: _native.check(lib.b200_lineplot_update(
    ctx.handle, out.data_ptr(), rows, n, n, 1, 1, ctypes.c_float(2.0 / rows), 4, average.data_ptr(), points.data_ptr(),
    scratch.data_ptr(), sp)), rows * n * 4)
timed("lineplot from fused column sums", lambda: _native.check(lib.b200_lineplot_update_from_colsum(
    ctx.handle, colsum.data_ptr(), n, 1, ctypes.c_float(2.0 / rows), 4, average.data_ptr(), points.data_ptr(), sp)),
      n * 12)
timed("waterfall (512 newest rows into the ring)", lambda: _native.check(lib.b200_waterfall_update(
    ctx.handle, out.data_ptr(), rows, n, n, 1, ring.data_ptr(), 512, 0, sp)), 512 * n * 8)
