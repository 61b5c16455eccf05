#!/bin/bash
# round 2, GPU call D: dwconv mma v3 (TMA in/out, matrix-move transposes, persistent), single-launch l2 move
mkdir -p gpurun_out
T=r02d
timeout 600 python -m pytest tests/test_gpu_model_ops.py -m gpu -q -x -k "dwconv or block or engine" > gpurun_out/${T}_pytest_dwconv.log 2>&1; echo "pytest dwconv rc=$?"
tail -5 gpurun_out/${T}_pytest_dwconv.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_autoattack.py -m gpu -q -x > gpurun_out/${T}_pytest_attack.log 2>&1; echo "pytest attack rc=$?"
tail -3 gpurun_out/${T}_pytest_attack.log
echo "== mma kernel v3" > gpurun_out/${T}_ops_bench.txt
timeout 300 python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad)" >> gpurun_out/${T}_ops_bench.txt 2>&1
for th in 8 28; do echo "== TH=$th (stage 0/1 tiles)" >> gpurun_out/${T}_ops_bench.txt; B200AT_DWM_TH=$th timeout 300 python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad).*(56x56|28x28)" >> gpurun_out/${T}_ops_bench.txt 2>&1; done
for nb in 1 2 4; do echo "== NB=$nb (stage 1/2/3 tiles)" >> gpurun_out/${T}_ops_bench.txt; B200AT_DWM_NB=$nb timeout 300 python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad).*(28x28|14x14|7x7)" >> gpurun_out/${T}_ops_bench.txt 2>&1; done
cat gpurun_out/${T}_ops_bench.txt
timeout 300 python profiles/k1_driver.py > gpurun_out/${T}_k1_driver.txt 2>&1; grep -E "l2_|l1_" gpurun_out/${T}_k1_driver.txt
B200AT_L2_PHASES=4 timeout 300 python profiles/k1_driver.py 2>&1 | grep -E "l2_" | sed 's/^/four launches: /'
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv7_mma -c 2 -o gpurun_out/${T}_dwm python profiles/ops_bench.py --once --only "dwconv7_(fwd|dgrad).*56x56" > gpurun_out/${T}_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:l2_cluster -c 1 -o gpurun_out/${T}_l2 python profiles/k1_driver.py --iters 1 > gpurun_out/${T}_ncu_l2.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/${T}_bench.json'));print('default', d['value'],d['ms_per_step'])"
