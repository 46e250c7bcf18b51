#!/bin/bash
# Round-2 GPU call 3: parity, reduction-K / affine-round / NTT A/B after the job reordering, captures as CSV only.
set -u
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8) | tee gpurun_out/r2c3_tests.txt
run() {   # name, extra bench flags
    local name=$1; shift
    timeout 300 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2c3_$name.json 2> gpurun_out/r2c3_$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2c3_{name}.json").read().strip().splitlines()[-1])
    e2e = d.get("e2e", {}).get("ms_per_step")
    r = d.get("roofline", {})
    print(f"{name:26s} value {d['ms_per_step']:8.2f} ms  e2e {e2e if e2e is None else round(e2e, 2)}  launches {d['gpu_launches']}  acc_ms {r.get('avg_launch_ms') or r.get('accumulate_ms')}  int {r.get('integer_roofline', {}).get('achieved')}")
except Exception as e:
    print(name, "FAILED", e)
PY
}
run prove_auto
run prove_k16 --reduce-k 16 --reduce-k1 16
run prove_k8 --reduce-k 8 --reduce-k1 8
run prove_k2 --reduce-k 2 --reduce-k1 2
run prove_r0 --affine-rounds 0
run prove_r2 --affine-rounds 2
run prove_r3 --affine-rounds 3
run prove_l32 --affine-batch 32
run prove_ntt_r2 --ntt-radix8 0
run prove_bool --witness boolean
run prove_22 --log-size 22 --steps 3 --warmup 2
for lg in 20 24; do
    run msm${lg}_auto --workload msm --log-size $lg --steps 3 --warmup 2
done
run msm24_l32 --workload msm --log-size 24 --affine-batch 32 --steps 3 --warmup 2
run ntt24_r8 --workload ntt --log-size 24
run ntt24_r2 --workload ntt --log-size 24 --ntt-radix8 0
python tools/timeline_report.py gpurun_out/r2c3_prove_auto.json > gpurun_out/r2c3_timeline_auto.txt 2>&1; cat gpurun_out/r2c3_timeline_auto.txt
bash tools/round2_ncu.sh 2>&1 | tee gpurun_out/r2c3_ncu_table.txt
head -60 gpurun_out/launches_summary.txt
du -sh gpurun_out
