#!/bin/bash
# Round-2 GPU call 5: final validation on one GPU -- parity suite, smoke(), the default bench line (with cpu_baseline),
# the reference arm at 2^20 (one step: how long the CPU prover takes on this box), and a window A/B.
set -u
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5) | tee gpurun_out/r2c5_tests.txt
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) | tee gpurun_out/r2c5_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c5_bench_default.json 2> gpurun_out/r2c5_bench_default.err; tail -c 600 gpurun_out/r2c5_bench_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c5_bench_default.json").read().strip().splitlines()[-1])
print("default bench: value", round(d["ms_per_step"], 2), "ms  e2e", round(d["e2e"]["ms_per_step"], 2), "ms  cpu_baseline", d["cpu_baseline"], "\nroofline", {k: d["roofline"][k] for k in ("achieved", "frac", "avg_launch_ms", "pairs_per_step", "bucket_entries_per_step")}, d["roofline"]["integer_roofline"])
PY
(time timeout 1200 python bench.py --impl reference --steps 1 --warmup 0) > gpurun_out/r2c5_reference.json 2> gpurun_out/r2c5_reference.err; tail -5 gpurun_out/r2c5_reference.err; cut -c1-900 gpurun_out/r2c5_reference.json
for c in 15 17; do
  timeout 300 python bench.py --no-cpu-baseline --window-bits $c > gpurun_out/r2c5_prove_c$c.json 2>/dev/null
  python - $c <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r2c5_prove_c{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("window", sys.argv[1], "value", round(d["ms_per_step"], 2))
PY
done
