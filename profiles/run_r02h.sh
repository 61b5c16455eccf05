#!/bin/bash
# round 2, GPU call H: tensor-core first stem stage (fwd + input gradient), A/B against the FMA kernels
mkdir -p gpurun_out
T=r02h
timeout 900 python -m pytest tests/test_gpu_model_ops.py -m gpu -q -x -k "stem0 or engine or layernorm" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
echo "== tensor-core stem" > gpurun_out/${T}_ops_bench.txt
timeout 300 python profiles/ops_bench.py --only "stem0" >> gpurun_out/${T}_ops_bench.txt 2>&1
echo "== FMA stem (B200AT_STEM0_TC=0)" >> gpurun_out/${T}_ops_bench.txt
B200AT_STEM0_TC=0 timeout 300 python profiles/ops_bench.py --only "stem0" >> gpurun_out/${T}_ops_bench.txt 2>&1
cat gpurun_out/${T}_ops_bench.txt
timeout 600 python -m pytest tests/test_gpu_full_loop.py -m gpu -q > gpurun_out/${T}_pytest_full_loop.log 2>&1; echo "full_loop rc=$?"; tail -2 gpurun_out/${T}_pytest_full_loop.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem0 -c 2 -o gpurun_out/${T}_stem python profiles/ops_bench.py --once --only "stem0" > gpurun_out/${T}_ncu.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/${T}_bench.json'));print('default', d['value'],d['ms_per_step'])"
