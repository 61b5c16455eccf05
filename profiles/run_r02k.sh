#!/bin/bash
# round 2, GPU call K: whole-step CUDA graph, l1 with conditional sectioning restored
mkdir -p gpurun_out
T=r02k
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_parity.py tests/test_gpu_autoattack.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${T}_pytest.log
timeout 300 python profiles/k1_driver.py 2>&1 | grep -E "l2_|l1_"
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('attack graph only', d['value'],d['ms_per_step'], 'e2e', d['e2e']['value'])"
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --graph-step 1 > gpurun_out/${T}_bench_graphstep.json 2> gpurun_out/${T}_bench_graphstep.err; echo "rc=$?"; tail -3 gpurun_out/${T}_bench_graphstep.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_graphstep.json').read().strip().splitlines()[-1]);print('whole-step graph', d['value'],d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])"
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --graph-step 1 --res 320 > gpurun_out/${T}_bench_graphstep_320.json 2>> gpurun_out/${T}_bench_graphstep.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_graphstep_320.json').read().strip().splitlines()[-1]);print('whole-step graph 320', d['value'],d['ms_per_step'])"
