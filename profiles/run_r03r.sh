#!/bin/bash
# round 2, GPU call AR: residual epilogue of the tcgen05 GEMM through the 32-column pass code (coalesced residual reads, bias in shared memory)
mkdir -p gpurun_out
T=r03r
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model_ops.py tests/test_gpu_vit.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python profiles/ops_bench.py --only "gemm pwconv2 RESIDUAL|gemm dgrad1" > gpurun_out/${T}_ops_bench.txt 2>&1; cat gpurun_out/${T}_ops_bench.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('default', d['value'],d['ms_per_step'])"
