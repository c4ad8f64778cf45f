"""Turn `ncu -i X.ncu-rep --page raw --csv` into a small per-launch markdown table for profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ; python scripts/ncu_summary.py raw.csv > profiles/rNN_x.md
"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_bytes.sum", "L2 bytes"),
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
print("| # | kernel | " + " | ".join(n for _, n in COLS) + " |")
print("|---|---|" + "---|" * len(COLS))
for n, r in enumerate(rows[2:]):
    if len(r) != len(hdr):
        continue
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("nerfw::", "")
    cells = []
    for key, _ in COLS:
        if key in idx:
            v, u = r[idx[key]], units[idx[key]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {u}".strip())
        else:
            cells.append("n/a")
    print(f"| {n} | `{name}` | " + " | ".join(cells) + " |")
