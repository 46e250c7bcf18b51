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
# warp-state samples: where the resident warps spend their cycles (per issued instruction)
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "barrier", "lg_throttle", "mio_throttle", "not_selected",
          "dispatch_stall", "no_instruction", "imc_miss", "branch_resolving", "membar", "drain", "sleeping", "tex_throttle", "selected"]


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
        st = []
        for name in STALLS:
            key = f"smsp__average_warps_issue_stalled_{name}_per_issue_active.ratio"
            if name == "selected":
                key = "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio"
            if key in hdr:
                try:
                    val = float(r[hdr.index(key)])
                except ValueError:
                    continue
                if val >= 0.05:
                    st.append((val, name))
        print(" | ".join(parts))
        if st:
            print("      warps per issue slot by state: " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
