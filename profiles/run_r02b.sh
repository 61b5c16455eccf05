#!/bin/bash
# round 2, GPU call B: tensor-core depthwise conv (b200at_dwconv_mma.cu) -- parity, A/B timing vs the FMA kernel, tile sweep
mkdir -p gpurun_out
T=r02b
python -m pytest tests/test_gpu_model_ops.py -m gpu -q -x -k "dwconv or block or engine" > gpurun_out/${T}_pytest_dwconv.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${T}_pytest_dwconv.log
echo "== mma kernel" > gpurun_out/${T}_ops_bench.txt
python profiles/ops_bench.py --only dwconv7_ >> gpurun_out/${T}_ops_bench.txt 2>&1
echo "== fma kernel (B200AT_DW_MMA=0)" >> gpurun_out/${T}_ops_bench.txt
B200AT_DW_MMA=0 python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad)" >> gpurun_out/${T}_ops_bench.txt 2>&1
for th in 8 14 28 56; do echo "== mma kernel TH=$th (stage 0/1 tiles)" >> gpurun_out/${T}_ops_bench.txt; B200AT_DWM_TH=$th python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad).*(56x56|28x28)" >> gpurun_out/${T}_ops_bench.txt 2>&1; done
for nb in 2 8 16; do echo "== mma kernel NB=$nb (stage 2/3 tiles)" >> gpurun_out/${T}_ops_bench.txt; B200AT_DWM_NB=$nb python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad).*(14x14|7x7)" >> gpurun_out/${T}_ops_bench.txt 2>&1; done
cat gpurun_out/${T}_ops_bench.txt
ncu --set full --clock-control none --import-source on -k regex:dwconv7_mma -c 2 -o gpurun_out/${T}_dwm python profiles/ops_bench.py --once --only "dwconv7_(fwd|dgrad).*56x56" > gpurun_out/${T}_ncu.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/${T}_bench.json'));print(d['value'],d['ms_per_step'])"
