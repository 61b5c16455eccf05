#!/bin/bash
# round 2, GPU call W: plain LayerNorm kernels with weights in shared memory (more warps per SM on the small maps)
mkdir -p gpurun_out
T=r02w
timeout 600 python -m pytest tests/test_gpu_model_ops.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('default', d['value'],d['ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ln_(fwd|bwd)_kernel" -c 200 --csv --log-file gpurun_out/${T}_ln_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${T}_ncu.log 2>&1; python profiles/summarize_launches.py gpurun_out/${T}_ln_launches.csv --top 20
