#!/bin/bash
# Multi-GPU call:  gpurun --gpus 8 --timeout 900 -- 'bash tools/round2_multi.sh'
# 2^20 prove at N = 2, 4, 8 (value + e2e legs, timeline, sharded proof == single-GPU proof asserted inside bench.py),
# then BASELINE.json configs[4]: the 2^22 prove at N = 1, 2, 4, 8.
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
run() {   # name, n, extra flags
    local name=$1 n=$2; shift 2
    if [ "$n" = 1 ]; then
        timeout 400 python bench.py --gpus 1 --no-cpu-baseline "$@" > gpurun_out/multi_$name.json 2> gpurun_out/multi_$name.err
    else
        timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/multi_$name.json 2> gpurun_out/multi_$name.err
    fi
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/multi_{name}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(f"{name:18s} N={d['n_gpus']} value {d['ms_per_step']:8.2f} ms  e2e {d['e2e']['ms_per_step']:8.2f} ms  sha {str(d.get('proof_sha256'))[:16]}  check: {d.get('proof_check')}")
except Exception as e:
    print(name, "FAILED", e); print(open(f"gpurun_out/multi_{name}.err").read()[-1500:])
PY
}
for n in 2 4 8; do [ $n -le $NG ] && run p20_n$n $n --steps 10 --warmup 3; done
for n in 2 4 8; do [ $n -le $NG ] && run p22_n$n $n --log-size 22 --steps 5 --warmup 2; done
