#!/bin/bash
# round 2, GPU call E: dwconv mma v3b (tap prefetch, NB==1 specialisation, narrow maps on the FMA kernel), pipelined
# (cp.async.bulk ring) LayerNorm, l2 cluster kernel with per-sample constants hoisted; full GPU suite
mkdir -p gpurun_out
T=r02e
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_full_loop.py > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 600 python -m pytest tests/test_gpu_full_loop.py -m gpu -q -s > gpurun_out/${T}_pytest_full_loop.log 2>&1; echo "full_loop rc=$?"; tail -3 gpurun_out/${T}_pytest_full_loop.log
timeout 300 python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad)|ln_" > gpurun_out/${T}_ops_bench.txt 2>&1
echo "== LN: previous kernels (B200AT_LN_RING=0)" >> gpurun_out/${T}_ops_bench.txt
B200AT_LN_RING=0 timeout 300 python profiles/ops_bench.py --only "ln_" >> gpurun_out/${T}_ops_bench.txt 2>&1
cat gpurun_out/${T}_ops_bench.txt
timeout 300 python profiles/k1_driver.py > gpurun_out/${T}_k1_driver.txt 2>&1; grep -E "l2_|l1_" gpurun_out/${T}_k1_driver.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ln_(fwd|bwd)_ring" -c 3 -o gpurun_out/${T}_ln python profiles/ops_bench.py --once --only "ln_.*56x56" > gpurun_out/${T}_ncu_ln.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/${T}_bench.json'));print('default', d['value'],d['ms_per_step'])"
B200AT_LN_RING=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_noring.json 2>> gpurun_out/${T}_bench.err; python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_noring.json'));print('LN ring off', d['value'],d['ms_per_step'])"
