#!/bin/bash
# round 2, GPU call X: second stem convolution as an implicit GEMM on the tcgen05 kernel
mkdir -p gpurun_out
T=r02x
timeout 600 python -m pytest tests/test_gpu_model_ops.py tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${T}_pytest.log
timeout 300 python profiles/ops_bench.py --only "conv3x3s2" > gpurun_out/${T}_ops_bench.txt 2>&1; cat gpurun_out/${T}_ops_bench.txt | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('default', d['value'],d['ms_per_step'])" || tail -5 gpurun_out/${T}_bench.err
