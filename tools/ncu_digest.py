#!/usr/bin/env python
"""One line per captured launch from `ncu --page raw --csv`: the figures DESIGN.md / profiles/ quote."""
import csv
import sys

WANT = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "ns"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"), ("launch__registers_per_thread", "regs"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe%"),
        ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "fmaheavy%"), ("lts__t_sector_hit_rate.pct", "l2_hit%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__warps_eligible.avg.per_cycle_active", "eligible/cyc")]


def main(path):
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        print(path, "no capture")
        return
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        parts = []
        for key, short in WANT:
            if key in hdr:
                i = hdr.index(key)
                v = r[i]
                if key == "Kernel Name":
                    v = v[:48]
                parts.append(f"{short}={v}{'' if key == 'Kernel Name' else ' ' + units[i]}")
        print(" | ".join(parts))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
