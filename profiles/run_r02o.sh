#!/bin/bash
# round 2, GPU call O: packed fp32x2 GELU (fused MLP epilogue, bias_gelu kernels, GEMM epilogues, LN+GELU), l2 cluster
# kernel at 512 / 256 threads, dwconv v4 only on single-image tiles
mkdir -p gpurun_out
T=r02o
timeout 900 python -m pytest tests/test_gpu_model_ops.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${T}_pytest.log
timeout 600 python profiles/ops_bench.py --only "mlp|gelu|gemm|ln|LN|layernorm" > gpurun_out/${T}_ops_bench.txt 2>&1
cat gpurun_out/${T}_ops_bench.txt
timeout 300 python profiles/k1_driver.py > gpurun_out/${T}_k1_driver.txt 2>&1; grep -E "l2_|l1_" gpurun_out/${T}_k1_driver.txt
B200AT_L2_THREADS=256 timeout 300 python profiles/k1_driver.py 2>&1 | grep -E "l2_" | sed 's/^/256 threads: /'
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('default', d['value'],d['ms_per_step'])"
B200AT_TCGEN05=residual,dgrad1,fc1,dgrad2,mlp,gelu,gelu_grad timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_fusedepi.json 2> gpurun_out/${T}_bench_fusedepi.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_fusedepi.json').read().strip().splitlines()[-1]);print('+gelu,gelu_grad epilogues', d['value'],d['ms_per_step'])"
