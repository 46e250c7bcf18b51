#!/bin/bash
# Round-2 first GPU call: parity of the precomputed-window MSM path, then its A/B on the 2^20 prove.
#   gpurun --timeout 1500 -- 'bash tools/round2_ab.sh'
# Results land in gpurun_out/r2_*.  Nothing here changes clocks or needs more than one GPU.
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -6) > gpurun_out/r2_tests.txt
cat gpurun_out/r2_tests.txt
run() {   # name, extra bench flags
    local name=$1; shift
    timeout 120 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2_bench_{name}.json").read().strip().splitlines()[-1])
    print(f"{name:28s} value {d['ms_per_step']:7.2f} ms   e2e {d['e2e']['ms_per_step']:7.2f} ms   launches {d['gpu_launches']}")
except Exception as e:
    print(name, "FAILED", e)
PY
}
run baseline
run precompute --precompute 1
for c in 14 16 17 18; do run precompute_c$c --precompute 1 --window-bits $c; done
for v in 4 44 1 3 33 40; do run acc$v --acc-variant $v; done   # g1 + 10*g2: 4 = 3 CTAs/SM, 1 = 4 CTAs/SM, 3 = prefetch
run precompute_acc4 --precompute 1 --acc-variant 4
run baseline_bool --witness boolean
run precompute_bool --precompute 1 --witness boolean
# device timeline of the baseline and of the precomputed-window run
for n in baseline precompute; do python tools/timeline_report.py gpurun_out/r2_bench_$n.json > gpurun_out/r2_timeline_$n.txt 2>&1; cat gpurun_out/r2_timeline_$n.txt; done
