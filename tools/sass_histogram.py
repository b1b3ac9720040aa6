"""SASS opcode histogram of the hot kernels in cyberether_b200/libb200dsp.so (cuobjdump -sass, no GPU needed):
per kernel the counts of the opcodes the design notes argue with — TMA bulk copies (UBLKCP), mbarrier (SYNCS), packed
FP32 (FFMA2 / FADD2 / FMUL2), MUFU, shared-memory traffic, barriers, 2-D tensor-map TMA (UTMALDG: the column tiles of the
tiled two-pass FFT) — and the absence of tensor-core opcodes (UTCMMA / LDTM), which this path has no use for. Usage: python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cyberether_b200", "libb200dsp.so")
HOT = ("fft4096_kernel", "fft_radix_kernel", "fft_cols_kernel", "fft_rows256_kernel", "fft_col16_kernel", "fir_decim_kernel", "fm_narrow_fused_kernel", "scan_tile_kernel",
       "scan_replay_kernel", "scan_tiles_kernel", "fm_wide_phase_table_kernel", "colsum_partial_kernel",
       "lineplot_finalize_kernel", "waterfall_write_kernel")
WATCH = ("UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MUFU", "LDS", "STS", "LDG", "STG", "LDGSTS",
         "BAR", "SHFL", "I2F", "F2I", "LOP3", "IMAD", "UTCMMA", "UTCHMMA", "UTMALDG", "LDTM", "HMMA", "LDC", "LDCU")


def main():
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    current = None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            current = m.group(1)
            kernels[current] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)(\.[A-Z0-9_.]+)?\s", line)
        if m and current:
            kernels[current][m.group(1)] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    for mangled, pretty in zip(kernels, names):
        demangle[mangled] = pretty
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} — opcode counts per kernel (static, per cubin function)")
    print(f"# arch: " + ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", text)))))
    for mangled, counts in kernels.items():
        pretty = demangle[mangled]
        if not any(h in pretty for h in HOT):
            continue
        total = sum(counts.values())
        shown = "  ".join(f"{op}={counts[op]}" for op in WATCH if counts[op])
        print(f"\n{pretty.split('(')[0]}\n  instructions={total}  {shown}")
    totals = collections.Counter()
    for counts in kernels.values():
        totals.update(counts)
    print("\n# whole library: " + "  ".join(f"{op}={totals[op]}" for op in WATCH))


if __name__ == "__main__":
    sys.exit(main())
