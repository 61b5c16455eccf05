#!/bin/bash
# round 2, GPU call AC: forward GELU with one MUFU (Phi(-|v|) = 2^P7(|v|)): whole GPU suite, kernel timings, step
mkdir -p gpurun_out
T=r03c
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
timeout 600 python profiles/ops_bench.py --only "mlp fused fwd|BIAS_GELU|bias_gelu_fwd|gelu stem|stem0_fwd" > gpurun_out/${T}_ops_bench.txt 2>&1; cat gpurun_out/${T}_ops_bench.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('default', d['value'],d['ms_per_step'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench2.json 2> gpurun_out/${T}_bench2.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench2.json').read().strip().splitlines()[-1]);print('default again', d['value'],d['ms_per_step'])"
