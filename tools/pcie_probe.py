"""H2D / D2H bandwidth of pinned buffers: one stream vs two streams per direction, alone and concurrently (what bounds e2e)."""
import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = torch.device("cuda:0")
GB = 1 << 30
h_in = torch.empty(2 * GB, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(1 * GB, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(2 * GB, dtype=torch.uint8, device=dev)
d_out = torch.empty(1 * GB, dtype=torch.uint8, device=dev)
streams = [torch.cuda.Stream(dev) for _ in range(4)]

def run(h2d_streams, d2h_streams, chunk):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_in = h_in.numel() // chunk if h2d_streams else 0
    n_out = h_out.numel() // chunk if d2h_streams else 0
    for c in range(max(n_in, n_out)):
        if c < n_in:
            with torch.cuda.stream(streams[c % h2d_streams]):
                d_in[c * chunk:(c + 1) * chunk].copy_(h_in[c * chunk:(c + 1) * chunk], non_blocking=True)
        if c < n_out:
            with torch.cuda.stream(streams[2 + c % d2h_streams]):
                h_out[c * chunk:(c + 1) * chunk].copy_(d_out[c * chunk:(c + 1) * chunk], non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t0

for chunk_mb in (32, 128):
    chunk = chunk_mb << 20
    for a, b in ((1, 0), (2, 0), (0, 1), (0, 2), (1, 1), (2, 1), (2, 2)):
        run(a, b, chunk)
        dt = min(run(a, b, chunk) for _ in range(3))
        print(f"chunk {chunk_mb:4d} MiB  h2d streams {a}  d2h streams {b}: {dt*1e3:7.2f} ms  "
              f"h2d {(2 if a else 0) / dt:5.1f} GiB/s  d2h {(1 if b else 0) / dt:5.1f} GiB/s")
