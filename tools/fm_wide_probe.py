"""Ad-hoc: one wideband (stereo) FM exec of 2^19 samples, for launch lists (ncu --metrics gpu__time_duration.sum)."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cyberether_b200 import _native
from cyberether_b200.jetstream import Context
lib = _native.load(); dev = torch.device("cuda:0"); ctx = Context.get(dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
frames, lanes, fl = (int(sys.argv[1]) if len(sys.argv) > 1 else 64), (int(sys.argv[2]) if len(sys.argv) > 2 else 1), 8192
x = torch.view_as_complex(torch.randn(frames, lanes, fl, 2, device=dev))
out = torch.empty(frames, lanes, fl, 2, dtype=torch.float32, device=dev)
plan = ctypes.c_void_p()
_native.check(lib.b200_fm_plan_create(ctx.handle, lanes, ctypes.c_float(250e3), 1, 75, ctypes.byref(plan)))
for _ in range(2): _native.check(lib.b200_fm_exec(plan, x.data_ptr(), out.data_ptr(), frames, fl, sp))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); _native.check(lib.b200_fm_exec(plan, x.data_ptr(), out.data_ptr(), frames, fl, sp)); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1); print(f"fm wide {frames*lanes*fl} samples ({frames} frames x {lanes} lanes): {ms:.3f} ms {frames*lanes*fl/ms*1e-6:.2f} GS/s")
