#!/bin/bash
# Round-2 GPU call 2: parity of the batched-affine rounds, A/B of rounds / batch size on the 2^20 prove and the MSM
# microbench, then launch list + ncu captures of every hot kernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
(nproc; cat /sys/fs/cgroup/cpu.max 2>&1; python -c "import os;print(len(os.sched_getaffinity(0)))") > gpurun_out/r2_host.txt 2>&1
(timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8) | tee gpurun_out/r2c2_tests.txt
run() {   # name, extra bench flags
    local name=$1; shift
    timeout 300 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2c2_$name.json 2> gpurun_out/r2c2_$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2c2_{name}.json").read().strip().splitlines()[-1])
    e2e = d.get("e2e", {}).get("ms_per_step")
    r = d.get("roofline", {})
    print(f"{name:26s} value {d['ms_per_step']:8.2f} ms  e2e {e2e if e2e is None else round(e2e, 2)}  launches {d['gpu_launches']}  acc_ms {r.get('avg_launch_ms') or r.get('accumulate_ms')}  int {r.get('integer_roofline', {}).get('achieved')}")
except Exception as e:
    print(name, "FAILED", e)
PY
}
run prove_r0 --affine-rounds 0
run prove_auto
run prove_r2 --affine-rounds 2
run prove_r4 --affine-rounds 4
run prove_r3_l8 --affine-rounds 3 --affine-batch 8
run prove_r3_l32 --affine-rounds 3 --affine-batch 32
run prove_r3_l64 --affine-rounds 3 --affine-batch 64
run prove_bool --witness boolean
for lg in 20 22 24; do
    run msm${lg}_r0 --workload msm --log-size $lg --affine-rounds 0 --steps 3 --warmup 2
    run msm${lg}_auto --workload msm --log-size $lg --steps 3 --warmup 2
done
run msm24_r4 --workload msm --log-size 24 --affine-rounds 4 --steps 3 --warmup 2
run msm24_r3_c22 --workload msm --log-size 24 --window-bits 22 --steps 3 --warmup 2
run ntt24 --workload ntt --log-size 24
python tools/timeline_report.py gpurun_out/r2c2_prove_auto.json > gpurun_out/r2c2_timeline_auto.txt 2>&1; cat gpurun_out/r2c2_timeline_auto.txt
bash tools/round2_ncu.sh 2>&1 | tee gpurun_out/r2c2_ncu_table.txt
