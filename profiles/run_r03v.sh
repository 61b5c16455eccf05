#!/bin/bash
# round 2, GPU call AV: launch list of one step of the final build (eager launches: --no-graph)
mkdir -p gpurun_out
T=r03v
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4700 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu rc=$?"; wc -l gpurun_out/${T}_launches.csv
