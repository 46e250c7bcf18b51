#!/bin/bash
# One GPU call that refreshes the evidence under profiles/ for the kernels of the 2^20 prove:
#   gpurun --timeout 1500 -- 'bash tools/round2_ncu.sh [extra bench flags]'
# 1. launch list (per-launch durations, cold-cache/serialised: shares only)   -> gpurun_out/launches.csv + summary
# 2. one `--set full` capture per hot kernel, a few launches of the second prove -> gpurun_out/ncu_<kernel>_raw.csv
#    (the .ncu-rep is kept only for the kernels listed in KEEP_REP: gpurun copies back at most 64 MiB)
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline $*"
KEEP_REP=""          # e.g. "k_aff_phase3": keep that kernel's .ncu-rep (about 6 MB per captured launch)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; tail -40 gpurun_out/launches_summary.txt
# kernel:skip:count -- skip the warm-up prove's launches of that kernel, capture a few of the timed one
for kc in k_aff_phase3:24:6 k_aff_phase1:24:6 k_msm_accumulate:8:4 k_msm_reduce_level:60:8 k_msm_digits:16:4 k_ntt_pass8:21:4 k_binv_top:12:3 k_binv_scan_up:12:3; do
    k=${kc%%:*}; r=${kc#*:}; skip=${r%%:*}; cnt=${r##*:}
    timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:"^${k}\$" -s $skip -c $cnt -f -o gpurun_out/ncu_$k $BENCH > gpurun_out/ncu_$k.log 2>&1
    if [ -f gpurun_out/ncu_$k.ncu-rep ]; then
        ncu -i gpurun_out/ncu_$k.ncu-rep --page raw --csv > gpurun_out/ncu_${k}_raw.csv 2>/dev/null
        python tools/ncu_digest.py gpurun_out/ncu_${k}_raw.csv
        # the source page (per-instruction stall samples), first 4000 lines, gzip: small
        ncu -i gpurun_out/ncu_$k.ncu-rep --page source --csv 2>/dev/null | head -4000 | gzip > gpurun_out/ncu_${k}_source.csv.gz
        case " $KEEP_REP " in *" $k "*) ;; *) rm -f gpurun_out/ncu_$k.ncu-rep ;; esac
    else
        echo "$k: no capture"; tail -3 gpurun_out/ncu_$k.log
    fi
done
du -sh gpurun_out
