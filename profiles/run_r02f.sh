#!/bin/bash
# round 2, GPU call F: packed-math pipelined LayerNorm, l2 cluster kernel v3 (Markstein division, contiguous slices)
mkdir -p gpurun_out
T=r02f
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_full_loop.py > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 300 python profiles/ops_bench.py --only "ln_" > gpurun_out/${T}_ops_bench.txt 2>&1; cat gpurun_out/${T}_ops_bench.txt
timeout 300 python profiles/k1_driver.py > gpurun_out/${T}_k1_driver.txt 2>&1; grep -E "l2_|l1_" gpurun_out/${T}_k1_driver.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ln_(fwd|bwd)_ring" -c 2 -o gpurun_out/${T}_ln python profiles/ops_bench.py --once --only "ln_.*56x56" > gpurun_out/${T}_ncu_ln.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:l2_cluster -c 1 -o gpurun_out/${T}_l2 python profiles/k1_driver.py --iters 1 > gpurun_out/${T}_ncu_l2.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/${T}_bench.json'));print('default', d['value'],d['ms_per_step'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${T}_ncu_bench.log 2>&1
