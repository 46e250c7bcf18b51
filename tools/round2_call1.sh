#!/bin/bash
# Round-2 GPU call 1: parity + A/B table of the MSM variants, then launch list + ncu captures.
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
nproc > gpurun_out/r2_host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2_host.txt 2>&1; python -c "import os;print(len(os.sched_getaffinity(0)))" >> gpurun_out/r2_host.txt
bash tools/round2_ab.sh 2>&1 | tee gpurun_out/r2_ab_table.txt
bash tools/round2_ncu.sh 2>&1 | tee gpurun_out/r2_ncu_table.txt
