#!/usr/bin/env python
"""bench.py — Window→FFT→Amplitude→Scale spectral chain, BASELINE.json configs[1]:
4096-pt × 65536-batch CF32 per GPU (weak scaling: every rank runs that workload on its own shard).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, libb200dsp)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path (oracle/_ref)

A step = one pass of the fused chain over the whole [65536, 4096] CF32 batch = ONE kernel launch
(fft4096_kernel<MODE_AMP_RANGE, WIN_REAL>). `value` is device-timed (CUDA events, inputs resident in HBM,
2 GiB input + 1 GiB output per step >> 126 MB L2 so every step streams from HBM); `e2e` times
b200_chain_exec_host: the same batch from pinned HOST memory, H2D + kernel + D2H inside the timed region.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FFT = 4096
ROWS = 65536
RANGE_MIN, RANGE_MAX = -120.0, 0.0
BYTES_PER_SAMPLE = 12          # 8 B CF32 in + 4 B F32 out
METRIC = "CF32 Msamples/sec Window->FFT->Amplitude->Scale"
UNIT = "Msamples/s"


def workload_name(rows=ROWS):
    return f"spectral_chain {N_FFT}-pt x {rows}-batch CF32 (window->fft->amplitude->range[-120,0])"


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the UNMODIFIED reference CPU path (oracle/_ref), all host cores
# ------------------------------------------------------------------------------------------------

def _ref_worker(args):
    rows, cycles, warm, seed = args
    import numpy as np
    from oracle import ref
    rng = np.random.Generator(np.random.PCG64(seed))
    x = (0.01 * (rng.standard_normal((rows, N_FFT)) + 1j * rng.standard_normal((rows, N_FFT)))).astype(np.complex64)
    with ref.Session(log_level=0) as s:
        s.add_source("src", x, sample_axis=1, batch_axis=0)
        s.add_block("spec", "spectrum_engine", {"enableScale": True, "rangeMin": RANGE_MIN, "rangeMax": RANGE_MAX},
                    {"buffer": "src.signal"})
        for _ in range(max(1, warm)):       # cycle 1 also settles window/invert (static modules)
            s.compute()
        t0 = time.perf_counter()
        for _ in range(cycles):
            s.write_source("src", x)        # a fresh buffer every cycle, like a live source
            s.compute()
        dt = time.perf_counter() - t0
    return rows * N_FFT * cycles, dt


def usable_cores() -> int:
    """Host threads this process may actually use: the affinity mask capped by the cgroup CPU quota
    (the GPU box exposes 128 logical CPUs but cpu.max grants 16)."""
    cores = len(os.sched_getaffinity(0))
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            cores = max(1, min(cores, int(float(quota) / float(period))))
    except Exception:
        pass
    return cores


def run_reference_cpu(cycles: int, warm: int, rows_per_proc: int = 512):
    """Times the reference's spectrum_engine block (reference Flowgraph + scheduler_synchronous +
    NativeCpuRuntime) on every host core: one independent single-threaded reference process per core,
    each on its own batch shard (the reference compute path is single-threaded)."""
    import multiprocessing as mp
    from oracle import ref
    if not ref.available():
        return None
    cores = usable_cores()
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        results = pool.map(_ref_worker, [(rows_per_proc, cycles, warm, 100 + i) for i in range(cores)])
        wall = time.perf_counter() - t0
    samples = sum(r[0] for r in results)
    slowest = max(r[1] for r in results)
    return {
        "value": samples / slowest / 1e6, "unit": UNIT, "cores": cores, "kind": "reference",
        "sample": f"{cores} procs x {cycles} cycles x [{rows_per_proc},{N_FFT}] CF32 through the reference "
                  f"spectrum_engine block (enableScale) on scheduler_synchronous; slowest proc {slowest:.2f}s, "
                  f"pool wall {wall:.1f}s",
        "ms_per_cycle_per_core": slowest / cycles * 1e3,
    }


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cycles = max(1, args.steps)
    base = run_reference_cpu(cycles, max(1, args.warmup))
    if base is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libjst_ref.so not built"}))
        return 0
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_cycle_per_core"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "note": "reference CPU path on a bounded sample per step",
                   "range": [RANGE_MIN, RANGE_MAX]},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# clocks sampling (NVML) during the timed region
# ------------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int, period: float = 0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples = []      # (t, sm_mhz, reasons_mask, power_w)
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as exc:   # pragma: no cover
            self.error = repr(exc)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                power = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                self.samples.append((time.perf_counter(), mhz, mask, power))
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self, t0: float, t1: float):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        note = "sampled inside the timed region"
        if not inside:
            inside = self.samples
            note = "timed region shorter than the sampling period; samples from warm-up + timed + e2e"
        import statistics
        mask = 0
        for s in inside:
            mask |= s[2]
        reasons = [name for bit, name in self.REASONS.items() if mask & bit and name != "gpu_idle"]
        return {"sm_mhz": statistics.median(s[1] for s in inside) if inside else None,
                "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(inside),
                "power_w_max": max((s[3] for s in inside), default=None), "note": note}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_launch():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get("fft4096_kernel<MODE_AMP_RANGE,WIN_REAL>")
    except Exception:
        return None


def main_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from cyberether_b200 import _native, amplitude_scaling_coeff, range_coefficients
    from cyberether_b200.jetstream import Context

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    lib = _native.load()          # raises if libb200dsp.so is missing
    ctx = Context.get(dev)
    rows, n = args.rows, N_FFT

    # -- synthetic IQ, resident in HBM: per row three tones + Gaussian noise, seeded per rank
    g = torch.Generator(device=dev)
    g.manual_seed(0x5EED0000 + rank)
    t = torch.arange(n, device=dev, dtype=torch.float32)
    b = torch.arange(rows, device=dev, dtype=torch.int64) + rank * rows
    x = torch.view_as_complex(torch.randn(rows, n, 2, device=dev, generator=g) * 1e-3)
    for mult, add, amp in ((97, 0, 0.5), (1013, 511, 0.05), (0, n // 2, 0.005)):
        k = ((b * mult + add) % n).to(torch.float32)[:, None]
        x += amp * torch.polar(torch.ones(1, device=dev), 2 * np.pi * k * t[None, :] / n)
    x = x.contiguous()
    out = torch.empty(rows, n, dtype=torch.float32, device=dev)

    # -- static part of the block: window -> invert (our own module kernels)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    win = torch.empty(n, dtype=torch.complex64, device=dev)
    winv = torch.empty(n, dtype=torch.complex64, device=dev)
    _native.check(lib.b200_window_blackman_cf32(ctx.handle, win.data_ptr(), n, sp))
    _native.check(lib.b200_invert_cf32(ctx.handle, win.data_ptr(), winv.data_ptr(), 1, n, 1, sp))
    torch.cuda.synchronize(dev)
    plan = ctypes.c_void_p()
    _native.check(lib.b200_chain_plan_create(ctx.handle, n, rows, winv.data_ptr(), ctypes.byref(plan)))
    coeff = amplitude_scaling_coeff(n)
    scale, offset = range_coefficients(RANGE_MIN, RANGE_MAX)

    def step():
        _native.check(lib.b200_chain_exec(plan, x.data_ptr(), out.data_ptr(), rows, coeff, 1, scale, offset, sp))

    def barrier():
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local)
    sampler.start()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    t_end = time.perf_counter()
    barrier()
    ms_total = e0.elapsed_time(e1)
    from cyberether_b200.sharding import max_over_ranks
    ms_total_max = max_over_ranks(ms_total, dev)
    ms_per_step = ms_total_max / args.steps
    samples_per_step_all = rows * n * world
    value = samples_per_step_all / (ms_per_step * 1e-3) / 1e6

    # -- e2e: host buffers through b200_chain_exec_host (H2D + kernel + D2H every step)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    x_host = torch.empty(rows, n, dtype=torch.complex64, pin_memory=True)
    out_host = torch.empty(rows, n, dtype=torch.float32, pin_memory=True)
    x_host.copy_(x)
    torch.cuda.synchronize(dev)

    def e2e_step():
        _native.check(lib.b200_chain_exec_host(plan, x_host.data_ptr(), out_host.data_ptr(), rows, coeff, 1, scale,
                                               offset, 0))
    e2e_step()    # warm-up (allocates the staging slots)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()  # synchronous: returns when out_host is complete
    t_e2e = time.perf_counter() - t0
    e2e_value = samples_per_step_all * e2e_steps / max_over_ranks(t_e2e, dev) / 1e6
    checksum = float(out_host[:: max(1, rows // 64)].double().sum())   # the step's result is read on the host

    sampler.stop_flag.set()
    sampler.join(timeout=1.0)
    clocks = sampler.summary(t_begin, t_end)

    # -- roofline of the dominant (only) kernel in the timed region
    peak, peak_src = measured_peak_gbs()
    kernel_ms = ms_total / args.steps                      # this rank's kernel: one launch per step
    achieved = rows * n * BYTES_PER_SAMPLE / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic_per_launch(), "kernel": "fft4096_kernel<MODE_AMP_RANGE,WIN_REAL>",
                "algorithmic_bytes_per_launch": rows * n * BYTES_PER_SAMPLE, "kernel_ms": kernel_ms,
                "peak_source": peak_src}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_reference_cpu(cycles=12, warm=1)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(rows), "rows_per_gpu": rows, "n": n,
                       "parallelism": f"batch-sharded x{world} (no data-path collective)",
                       "range": [RANGE_MIN, RANGE_MAX],
                       "l2": "inputs larger than L2 (2 GiB in + 1 GiB out per step vs 126 MB)"},
            "gpu_launches": args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": rows * n * 8,
                    "d2h_bytes_per_step": rows * n * 4, "steps": e2e_steps,
                    "api": "b200_chain_exec_host (pinned host buffers, 3-stream chunked pipeline)",
                    "checksum": checksum},
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "kernel_variant": lib.b200_chain_plan_variant(plan).decode(),
        }
        print(json.dumps(line))
    _native.check(lib.b200_chain_plan_destroy(plan))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())
