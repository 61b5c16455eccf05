#!/bin/bash
# round 2, GPU call AB: packed GELU in the stem kernels; half-height FMA dwconv tile on 14-wide maps (A/B)
mkdir -p gpurun_out
T=r03b
timeout 900 python -m pytest tests/test_gpu_model_ops.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
B200AT_DW_TH7_14=1 timeout 900 python -m pytest tests/test_gpu_model_ops.py -m gpu -q -x -k "dwconv or block or engine" > gpurun_out/${T}_pytest_half14.log 2>&1; echo "pytest half14 rc=$?"; tail -3 gpurun_out/${T}_pytest_half14.log
timeout 300 python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad).*(14x14|7x7)|stem0" > gpurun_out/${T}_ops_bench.txt 2>&1
echo "== B200AT_DW_TH7_14=1" >> gpurun_out/${T}_ops_bench.txt
B200AT_DW_TH7_14=1 timeout 300 python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad).*(14x14)" 2>&1 | tail -2 >> gpurun_out/${T}_ops_bench.txt
cat gpurun_out/${T}_ops_bench.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('default', d['value'],d['ms_per_step'])"
B200AT_DW_TH7_14=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_half14.json 2> gpurun_out/${T}_bench_half14.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_half14.json').read().strip().splitlines()[-1]);print('half-height 14-wide', d['value'],d['ms_per_step'])"
