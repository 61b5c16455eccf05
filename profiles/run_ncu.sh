#!/bin/bash
# `ncu --set full` of selected kernels of profiles/ops_bench.py --once; CSV/text exports only (the .ncu-rep stays on the box).
#   gpurun --timeout 900 -- 'bash profiles/run_ncu.sh tag kernel-regex ops-filter'
tag=${1:-rXX}; kre=$2; only=$3
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"$kre" -f -o /tmp/${tag}_k \
    python profiles/ops_bench.py --once --only "$only" > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/${tag}_ncu.log
ncu -i /tmp/${tag}_k.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw.csv 2>/dev/null
ncu -i /tmp/${tag}_k.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_ncu_source.csv 2>/dev/null
ncu -i /tmp/${tag}_k.ncu-rep --page details > gpurun_out/${tag}_ncu_details.txt 2>/dev/null
du -sh gpurun_out
