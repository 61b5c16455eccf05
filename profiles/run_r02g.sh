#!/bin/bash
# round 2, GPU call G: conv bias folded into LN, one-launch weight preparation / grad tail, l2 cluster 512 threads
mkdir -p gpurun_out
T=r02g
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_full_loop.py > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 600 python -m pytest tests/test_gpu_full_loop.py -m gpu -q > gpurun_out/${T}_pytest_full_loop.log 2>&1; echo "full_loop rc=$?"; tail -2 gpurun_out/${T}_pytest_full_loop.log
timeout 300 python profiles/k1_driver.py > gpurun_out/${T}_k1_driver.txt 2>&1; grep -E "l2_|l1_" gpurun_out/${T}_k1_driver.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/${T}_bench.json'));print('default', d['value'],d['ms_per_step'], 'e2e', d['e2e']['value'])"
timeout 600 python bench.py --steps 6 --warmup 3 --res 320 --no-cpu-baseline > gpurun_out/${T}_bench_320.json 2>> gpurun_out/${T}_bench.err; python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_320.json'));print('res320', d['value'],d['ms_per_step'])"
