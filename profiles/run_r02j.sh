#!/bin/bash
# round 2, GPU call J (1 GPU): l1 with 12 stream operations, l2 with the gradient slice in shared memory, config 5 on one
# GPU for l2 2.0 and l1 75 (ConvNeXt-L-CvSt at 320), full GPU suite
mkdir -p gpurun_out
T=r02j
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_full_loop.py > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest_gpu.log
timeout 300 python profiles/k1_driver.py > gpurun_out/${T}_k1_driver.txt 2>&1; cat gpurun_out/${T}_k1_driver.txt
for norm in Linf L2 L1; do timeout 900 python profiles/aa_bench.py --norm $norm --n 100 --bs 100 > gpurun_out/${T}_aa_${norm}.json 2> gpurun_out/${T}_aa_${norm}.err; echo "aa $norm rc=$?"; tail -c 700 gpurun_out/${T}_aa_${norm}.json; echo; done
timeout 900 python profiles/aa_bench.py --norm L2 --n 100 --bs 100 --eps 1e-7 --targets 2 > gpurun_out/${T}_aa_L2_worst.json 2> gpurun_out/${T}_aa_L2_worst.err; tail -c 600 gpurun_out/${T}_aa_L2_worst.json; echo
timeout 900 python profiles/aa_bench.py --norm L1 --n 100 --bs 100 --eps 1e-7 --targets 2 > gpurun_out/${T}_aa_L1_worst.json 2> gpurun_out/${T}_aa_L1_worst.err; tail -c 600 gpurun_out/${T}_aa_L1_worst.json; echo
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('default', d['value'],d['ms_per_step'])"
