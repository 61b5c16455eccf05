#!/bin/bash
# round 2, GPU call R: lean GELU / GELU' epilogues of the tcgen05 GEMM
mkdir -p gpurun_out
T=r02r
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
timeout 600 python profiles/ops_bench.py --only "gemm|bias_gelu" > gpurun_out/${T}_ops_bench.txt 2>&1
cat gpurun_out/${T}_ops_bench.txt
B200AT_TCGEN05=residual,dgrad1,fc1,dgrad2,mlp,gelu,gelu_grad timeout 600 python -m pytest tests/test_gpu_model_ops.py tests/test_gpu_full_loop.py -m gpu -q -x > gpurun_out/${T}_pytest_model.log 2>&1; echo "pytest model (fused epilogues) rc=$?"; tail -4 gpurun_out/${T}_pytest_model.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('default', d['value'],d['ms_per_step'])"
B200AT_TCGEN05=residual,dgrad1,fc1,dgrad2,mlp,gelu,gelu_grad timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_fusedepi.json 2> gpurun_out/${T}_bench_fusedepi.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_fusedepi.json').read().strip().splitlines()[-1]);print('+gelu,gelu_grad epilogues', d['value'],d['ms_per_step'])"
