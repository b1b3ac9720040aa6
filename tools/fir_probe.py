import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cyberether_b200 import _native
from cyberether_b200.jetstream import Context
lib = _native.load(); dev = torch.device("cuda:0"); ctx = Context.get(dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
frames, T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 8192
taps, R = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (129, 8)
xs = torch.view_as_complex(torch.randn(frames, T, 2, device=dev))
centers = (ctypes.c_double * 1)(0.0); host = np.zeros((1, taps), np.complex64)
_native.check(lib.b200_filter_taps_host(8e6, 1e6, centers, 1, taps, host.ctypes.data_as(ctypes.c_void_p)))
fp = ctypes.c_void_p(); _native.check(lib.b200_fir_plan_create(ctx.handle, host.ctypes.data_as(ctypes.c_void_p), taps, 1, R, ctypes.byref(fp)))
yo = torch.empty(frames, 1, T // R, dtype=torch.complex64, device=dev)
for _ in range(3): _native.check(lib.b200_fir_exec(fp, xs.data_ptr(), yo.data_ptr(), frames, T, sp))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): _native.check(lib.b200_fir_exec(fp, xs.data_ptr(), yo.data_ptr(), frames, T, sp))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"fir taps={taps} R={R} samples={frames*T}: {ms:.4f} ms {frames*T/ms*1e-6:.1f} GS/s")
