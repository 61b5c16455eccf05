#!/bin/bash
# round 2, GPU call Y: implicit-GEMM stem convolution timing, alone and in the step
mkdir -p gpurun_out
T=r02y
timeout 300 python profiles/ops_bench.py --only "conv3x3s2" > gpurun_out/${T}_ops_bench.txt 2>&1; cat gpurun_out/${T}_ops_bench.txt | tail -4
B200AT_STEM_CONV=lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_lib.json 2> gpurun_out/${T}_bench_lib.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_lib.json').read().strip().splitlines()[-1]);print('library conv', d['value'],d['ms_per_step'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('implicit GEMM conv', d['value'],d['ms_per_step'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu rc=$?"
