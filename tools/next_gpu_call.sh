#!/bin/bash
# The first GPU call a next session should make (none was left when DESIGN.md 4.3b's MSM form was written):
#   gpurun --timeout 2400 -- 'bash tools/next_gpu_call.sh'
# 1. parity of the one-bucket-set form and the tuner on hardware;
# 2. the bench line of record with the tuner on (its `autotune` object holds the per-form milliseconds), then the two
#    main forms forced, back to back, for an A/B that does not depend on the tuner;
# 3. the 2^24 MSM under both forms;
# 4. launch list + `--set full` captures of the hot kernels UNDER the one-bucket-set form (tools/round2_ncu.sh takes
#    extra bench flags): the evidence profiles/ lacks.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "unified or autotune or precompute" 2>&1 | tail -5 | tee gpurun_out/n1_parity.txt
line() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
at = d.get("autotune") or {}
print(sys.argv[1].split("/")[-1], "value", round(d["ms_per_step"], 2), "ms  e2e", d.get("e2e", {}).get("ms_per_step"), " autotune", at.get("ms"), "->", at.get("chosen"))
PY
}
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/n1_bench_tuned.json 2> gpurun_out/n1_bench_tuned.err; line gpurun_out/n1_bench_tuned.json
for p in 0 2 0 2; do
    timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --precompute $p > gpurun_out/n1_bench_p${p}_$RANDOM.json 2>/dev/null
done
for f in gpurun_out/n1_bench_p*.json; do line $f; done
timeout 600 python bench.py --workload msm --log-size 24 --steps 3 --warmup 2 > gpurun_out/n1_msm24_tuned.json 2>/dev/null; line gpurun_out/n1_msm24_tuned.json
timeout 600 python bench.py --log-size 22 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/n1_prove22_tuned.json 2>/dev/null; line gpurun_out/n1_prove22_tuned.json
bash tools/round2_ncu.sh --precompute 2
