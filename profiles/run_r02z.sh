#!/bin/bash
# round 2, GPU call Z: implicit-GEMM stem convolution with resident weights; library backward through aten.convolution_backward
mkdir -p gpurun_out
T=r02z
timeout 600 python -m pytest tests/test_gpu_model_ops.py -m gpu -q -x -k "conv3x3s2 or engine or stem" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python profiles/ops_bench.py --only "conv3x3s2" > gpurun_out/${T}_ops_bench.txt 2>&1; cat gpurun_out/${T}_ops_bench.txt | tail -4
B200AT_STEM_CONV=lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_lib.json 2> gpurun_out/${T}_bench_lib.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_lib.json').read().strip().splitlines()[-1]);print('library conv', d['value'],d['ms_per_step'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('implicit GEMM conv', d['value'],d['ms_per_step'])"
