#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small CSV/markdown table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more.ncu-rep ...] > profiles/rNN_xxx.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks/SM"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard (cycles/issue)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    cols = []  # (title, {metric: text}, dram bytes)
    for path in sys.argv[1:]:
        hdr, units, rows = rows_of(path)
        for j, r in enumerate(rows):
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
            cells, tot = {}, 0.0
            for k, _ in KEYS:
                if k not in hdr:
                    cells[k] = "n/a"
                    continue
                i = hdr.index(k)
                v = r[i]
                try:
                    v = f"{float(v.replace(',', '')):.6g}"
                except ValueError:
                    pass
                cells[k] = f"{v} {units[i]}".strip()
                if k.startswith("dram__bytes"):
                    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1)
                    tot += float(r[i].replace(",", "")) * scale
            cols.append((f"{path.split('/')[-1]} #{j}: {name}", cells, tot))
    print("| metric | " + " | ".join(c[0] for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for k, title in KEYS:
        print(f"| {title} (`{k}`) | " + " | ".join(c[1][k] for c in cols) + " |")
    print("| **dram bytes per launch (read + write)** | " + " | ".join(f"{c[2]:.5g}" for c in cols) + " |")


if __name__ == "__main__":
    main()
