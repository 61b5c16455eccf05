#!/bin/bash
# round 2, GPU call N: dwconv mma v4 (two warp groups on two plane buffers) against v3; l2 cluster kernel at 256 threads
mkdir -p gpurun_out
T=r02n
timeout 600 python -m pytest tests/test_gpu_model_ops.py -m gpu -q -x -k "dwconv or block or engine" > gpurun_out/${T}_pytest_dwconv.log 2>&1; echo "pytest dwconv rc=$?"
tail -5 gpurun_out/${T}_pytest_dwconv.log
timeout 300 compute-sanitizer --tool memcheck python profiles/ops_bench.py --once --only "dwconv7_(fwd|dgrad).*56x56" > gpurun_out/${T}_sanitizer.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/${T}_sanitizer.log
timeout 300 compute-sanitizer --tool racecheck python profiles/ops_bench.py --once --only "dwconv7_fwd.*56x56" > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/${T}_racecheck.log
echo "== mma kernel v4 (two warp groups)" > gpurun_out/${T}_ops_bench.txt
timeout 300 python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad)" >> gpurun_out/${T}_ops_bench.txt 2>&1
echo "== v3 (B200AT_DWM_PP=0)" >> gpurun_out/${T}_ops_bench.txt
B200AT_DWM_PP=0 timeout 300 python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad)" >> gpurun_out/${T}_ops_bench.txt 2>&1
echo "== v4, NB=1 at 28x28" >> gpurun_out/${T}_ops_bench.txt
B200AT_DWM_NB=1 timeout 300 python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad).*28x28" >> gpurun_out/${T}_ops_bench.txt 2>&1
echo "== v4, TH=14 at 28x28" >> gpurun_out/${T}_ops_bench.txt
B200AT_DWM_TH=14 timeout 300 python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad).*28x28" >> gpurun_out/${T}_ops_bench.txt 2>&1
cat gpurun_out/${T}_ops_bench.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "l2 or L2" > gpurun_out/${T}_pytest_l2.log 2>&1; echo "pytest l2 rc=$?"; tail -2 gpurun_out/${T}_pytest_l2.log
timeout 300 python profiles/k1_driver.py > gpurun_out/${T}_k1_driver.txt 2>&1; grep -E "l2_|l1_" gpurun_out/${T}_k1_driver.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv7_mma -c 2 -o gpurun_out/${T}_dwm python profiles/ops_bench.py --once --only "dwconv7_(fwd|dgrad).*56x56" > gpurun_out/${T}_ncu.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('default', d['value'],d['ms_per_step'])"
