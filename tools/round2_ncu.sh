#!/bin/bash
# One GPU call that refreshes the evidence under profiles/ for the kernels of the 2^20 prove:
#   gpurun --timeout 1500 -- 'bash tools/round2_ncu.sh [extra bench flags]'
# 1. launch list (per-launch durations, cold-cache/serialised: shares only)   -> gpurun_out/launches.csv
# 2. one `--set full` capture per hot kernel (3 launches after the warm-up)    -> gpurun_out/ncu_<kernel>.ncu-rep
# Read the reports on the CPU box:  ncu -i gpurun_out/ncu_<kernel>.ncu-rep --page raw --csv | grep -E 'dram__bytes|sm__warps_active|launch__registers'
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline $*"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; tail -25 gpurun_out/launches_summary.txt
# kernel:launches-per-prove -- skip the warm-up prove, capture every launch of the timed one
for kc in k_msm_accumulate:8:8 k_msm_reduce_level:40:12 k_msm_digits:16:8 k_ntt_pass:21:9; do
    k=${kc%%:*}; r=${kc#*:}; skip=${r%%:*}; cnt=${r##*:}
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s $skip -c $cnt -f -o gpurun_out/ncu_$k $BENCH > gpurun_out/ncu_$k.log 2>&1
    ncu -i gpurun_out/ncu_$k.ncu-rep --page raw --csv 2>/dev/null | python - "$k" <<'PY'
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) < 3:
    print(sys.argv[1], "no capture"); sys.exit(0)
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__inst_executed_pipe_fma.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print(" | ".join(f"{hdr[i]}={r[i][:60]}" for i in idx))
PY
done
