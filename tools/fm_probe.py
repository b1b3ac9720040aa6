"""Ad-hoc timing of b200_fm_exec (narrow), fused one-pass kernel vs B200_FM_LEGACY=1."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cyberether_b200 import _native
from cyberether_b200.jetstream import Context
lib = _native.load(); dev = torch.device("cuda:0"); ctx = Context.get(dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for frames, lanes, fl in ((2048, 1, 8192), (256, 8, 8192), (16, 1, 1 << 20)):
    x = torch.view_as_complex(torch.randn(frames, lanes, fl, 2, device=dev))
    out = torch.empty(frames, lanes, fl, dtype=torch.float32, device=dev)
    for de in (0, 75):
        plan = ctypes.c_void_p()
        _native.check(lib.b200_fm_plan_create(ctx.handle, lanes, ctypes.c_float(250e3), 0, de, ctypes.byref(plan)))
        for _ in range(3): _native.check(lib.b200_fm_exec(plan, x.data_ptr(), out.data_ptr(), frames, fl, sp))
        torch.cuda.synchronize(); ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); _native.check(lib.b200_fm_exec(plan, x.data_ptr(), out.data_ptr(), frames, fl, sp)); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts)); n = frames * lanes * fl
        print(f"fm narrow [{frames},{lanes},{fl}] deemph={de}us: {ms:.4f} ms {n/ms*1e-6:.1f} GS/s {n*12/ms*1e-6:.0f} GB/s ({n*12/ms*1e-6/6570.9*100:.1f}% of measured HBM)")
        lib.b200_fm_plan_destroy(plan)
