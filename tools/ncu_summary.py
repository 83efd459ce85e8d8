"""Summarise `ncu -i X.ncu-rep --page raw --csv` exports (one row per captured launch) into a markdown table.
usage: python tools/ncu_summary.py RAW.csv [RAW2.csv ...] > profiles/rNN_*.md
Only a fixed list of metrics is kept (the ones DESIGN.md argues with); numbers under the profiler are not bench values."""
import csv
import sys

KEEP = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def summarise(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        name = r[col["Kernel Name"]]
        out.append(f"### `{name}`  grid {r[col['Grid Size']]} block {r[col['Block Size']]}\n")
        out.append("| metric | value | unit |\n|---|---|---|")
        for k in KEEP:
            if k in col and r[col[k]] != "":
                out.append(f"| `{k}` | {r[col[k]]} | {units[col[k]]} |")
        out.append("")
    return "\n".join(out)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(f"## {p.split('/')[-1]}\n")
        print(summarise(p))
