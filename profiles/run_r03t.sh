#!/bin/bash
# round 2, GPU call AT: persistent depthwise weight-gradient kernel (49 tap sums per thread in registers across tiles)
mkdir -p gpurun_out
T=r03t
timeout 900 python -m pytest tests/test_gpu_model_ops.py -m gpu -q -x -k "dwconv or block or engine" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python profiles/ops_bench.py --only "dwconv7_wgrad" > gpurun_out/${T}_ops_bench.txt 2>&1; cat gpurun_out/${T}_ops_bench.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('default', d['value'],d['ms_per_step'])"
