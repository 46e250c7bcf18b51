#!/bin/bash
set -u
mkdir -p gpurun_out
for n in 8 4; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/multi4_p20_n$n.json 2> gpurun_out/multi4_p20_n$n.err
  python - $n <<'PY'
import json, sys
n=sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/multi4_p20_n{n}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(f"N={n} value {d['ms_per_step']:.2f} e2e {d['e2e']['ms_per_step']:.2f} host {d.get('sharded_host_ms_per_step')}")
except Exception as e:
    print("FAILED", e); print(open(f"gpurun_out/multi4_p20_n{n}.err").read()[-1500:])
PY
done
