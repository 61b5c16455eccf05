#!/bin/bash
# round 2, GPU call AJ: pwconv1 bias gradient in the GELU' GEMM's epilogue; LayerNorm grid cap A/B
mkdir -p gpurun_out
T=r03j
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model_ops.py tests/test_gpu_vit.py tests/test_gpu_full_loop.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
for v in 1 0; do
B200AT_GELU_GRAD_COLSUM=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_cs$v.json 2> gpurun_out/${T}_bench_cs$v.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_cs$v.json').read().strip().splitlines()[-1]);print('colsum in epilogue=$v', d['value'],d['ms_per_step'])"
done
for n in 2 8; do
B200AT_LN_CTAS=$n timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_ln$n.json 2> gpurun_out/${T}_bench_ln$n.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_ln$n.json').read().strip().splitlines()[-1]);print('LN CTAs/SM=$n', d['value'],d['ms_per_step'])"
done
