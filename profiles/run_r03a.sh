#!/bin/bash
# round 2, GPU call AA: fused first stem stage in the training forward
mkdir -p gpurun_out
T=r03a
timeout 900 python -m pytest tests/test_gpu_model_ops.py tests/test_gpu_full_loop.py tests/test_gpu_graph.py tests/test_gpu_driver.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${T}_pytest.log
B200AT_STEM0_TRAIN=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_off.json 2> gpurun_out/${T}_bench_off.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_off.json').read().strip().splitlines()[-1]);print('library first conv in training', d['value'],d['ms_per_step'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('fused first stage in training', d['value'],d['ms_per_step'])"
