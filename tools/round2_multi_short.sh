#!/bin/bash
set -u
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/multi3_p20_n$n.json 2> gpurun_out/multi3_p20_n$n.err
  python - $n <<'PY'
import json, sys
n=sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/multi3_p20_n{n}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(f"N={n} value {d['ms_per_step']:.2f} e2e {d['e2e']['ms_per_step']:.2f} host {d.get('sharded_host_ms_per_step')}")
except Exception as e:
    print("FAILED", e); print(open(f"gpurun_out/multi3_p20_n{n}.err").read()[-1500:])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --no-cpu-baseline --log-size 22 --steps 5 --warmup 2 > gpurun_out/multi3_p22_n8.json 2> gpurun_out/multi3_p22_n8.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/multi3_p22_n8.json").read().strip().splitlines() if l.startswith("{")][-1])
print(f"2^22 N=8 value {d['ms_per_step']:.2f} e2e {d['e2e']['ms_per_step']:.2f} host {d.get('sharded_host_ms_per_step')}")
PY
