#!/bin/bash
# Round-2 GPU call 6: bench lines of record at the final commit + A/B of the G2 phase-3 occupancy variant.
set -u
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline", {})
print(sys.argv[1].split("/")[-1], "value", round(d["ms_per_step"], 2), "e2e", d.get("e2e", {}).get("ms_per_step"), "int", r.get("integer_roofline", {}).get("achieved"), "g2", r.get("integer_roofline", {}).get("achieved_g2_in_fp_mul"), "entries", r.get("bucket_entries_per_step"))
PY
}
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c6_bench_default.json 2> gpurun_out/r2c6_bench_default.err; line gpurun_out/r2c6_bench_default.json
VAR='import sys, runpy; sys.path.insert(0, "."); import bellman_b200 as bb; bb.LIB_PATH = bb.LIB_PATH.replace("libbellman_b200.so", "libbellman_b200_g2b3.so"); sys.argv = ["bench.py"] + sys.argv[1:]; runpy.run_path("bench.py", run_name="__main__")'
timeout 600 python -c "$VAR" --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c6_bench_g2b3.json 2> gpurun_out/r2c6_bench_g2b3.err; line gpurun_out/r2c6_bench_g2b3.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c6_bench_default2.json 2>/dev/null; line gpurun_out/r2c6_bench_default2.json
timeout 600 python -c "$VAR" --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c6_bench_g2b3_2.json 2>/dev/null; line gpurun_out/r2c6_bench_g2b3_2.json
timeout 300 python bench.py --workload msm --log-size 24 --steps 3 --warmup 2 > gpurun_out/r2c6_msm24.json 2>/dev/null; line gpurun_out/r2c6_msm24.json
timeout 300 python bench.py --workload ntt --log-size 24 > gpurun_out/r2c6_ntt24.json 2>/dev/null; line gpurun_out/r2c6_ntt24.json
timeout 300 python bench.py --witness boolean --no-cpu-baseline > gpurun_out/r2c6_bool.json 2>/dev/null; line gpurun_out/r2c6_bool.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; head -30 gpurun_out/launches_summary.txt
