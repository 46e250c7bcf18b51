#!/bin/bash
# Round-2 GPU call 4: parity; A/B of the 2-D bucket reduction, reduction K, TMA-staged dense rounds, NTT variants.
set -u
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8) | tee gpurun_out/r2c4_tests.txt
run() {   # name, extra bench flags
    local name=$1; shift
    timeout 300 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2c4_$name.json 2> gpurun_out/r2c4_$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2c4_{name}.json").read().strip().splitlines()[-1])
    e2e = d.get("e2e", {}).get("ms_per_step")
    r = d.get("roofline", {})
    print(f"{name:26s} value {d['ms_per_step']:8.2f} ms  e2e {e2e if e2e is None else round(e2e, 2)}  launches {d['gpu_launches']}  acc_ms {r.get('avg_launch_ms') or r.get('accumulate_ms')}")
except Exception as e:
    print(name, "FAILED", e)
PY
}
run prove_auto
run prove_k16 --reduce-k 16 --reduce-k1 16
run prove_k8 --reduce-k 8 --reduce-k1 8
run prove_1d_k16 --reduce-2d 0 --reduce-k 16 --reduce-k1 16
run prove_tma --affine-tma 1
run prove_ntt_r2 --ntt-radix8 0
run prove_bool --witness boolean
run prove_22 --log-size 22 --steps 3 --warmup 2
run msm20_auto --workload msm --log-size 20 --steps 3 --warmup 2
run msm24_auto --workload msm --log-size 24 --steps 3 --warmup 2
run msm24_tma --workload msm --log-size 24 --affine-tma 1 --steps 3 --warmup 2
run msm24_1d --workload msm --log-size 24 --reduce-2d 0 --steps 3 --warmup 2
run ntt24_r8 --workload ntt --log-size 24
run ntt24_r2 --workload ntt --log-size 24 --ntt-radix8 0
python tools/timeline_report.py gpurun_out/r2c4_prove_auto.json > gpurun_out/r2c4_timeline_auto.txt 2>&1; cat gpurun_out/r2c4_timeline_auto.txt
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; head -50 gpurun_out/launches_summary.txt
for kc in k_bucket_fold:20:6 k_binv_top:12:3 k_ntt_pass8:21:3 k_aff_phase3_tma:0:0; do
    k=${kc%%:*}; r=${kc#*:}; skip=${r%%:*}; cnt=${r##*:}
    [ "$cnt" = 0 ] && continue
    timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:"^${k}\$" -s $skip -c $cnt -f -o gpurun_out/ncu_$k $BENCH > gpurun_out/ncu_$k.log 2>&1
    if [ -f gpurun_out/ncu_$k.ncu-rep ]; then
        ncu -i gpurun_out/ncu_$k.ncu-rep --page raw --csv > gpurun_out/ncu_${k}_raw.csv 2>/dev/null
        python tools/ncu_digest.py gpurun_out/ncu_${k}_raw.csv
        rm -f gpurun_out/ncu_$k.ncu-rep
    else echo "$k: no capture"; fi
done
du -sh gpurun_out
