"""Summarise an .ncu-rep (ncu --set full) into the small JSON kept under profiles/: per launch, the metrics the design
notes quote. Usage: python tools/ncu_summary.py <report.ncu-rep> <out.json> "<command the report was captured with>" """
import csv, io, json, subprocess, sys

KEEP = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size",
    "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
)


def main():
    rep, out, command = sys.argv[1], sys.argv[2], sys.argv[3]
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header, units, launches = rows[0], rows[1], rows[2:]
    summary = []
    for values in launches:
        entry = {}
        for name, unit, value in zip(header, units, values):
            if name == "Kernel Name":
                entry[name] = value
            elif name in KEEP or ("issue_stalled" in name and name.endswith("per_issue_active.ratio") and "not_issued" not in name):
                entry[name] = f"{value} {unit}".strip()
        summary.append(entry)
    json.dump({"source": command, "launches": summary}, open(out, "w"), indent=1)
    for e in summary:
        print(e.get("Kernel Name"), e.get("gpu__time_duration.sum"), e.get("dram__bytes_read.sum"), e.get("dram__bytes_write.sum"))


if __name__ == "__main__":
    main()
