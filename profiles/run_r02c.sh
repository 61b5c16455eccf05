#!/bin/bash
# round 2, GPU call C: dwconv mma v2 (channel-major taps, lane-pair global mapping), GEMM with 16 epilogue warps for the
# GELU / GELU' epilogues
mkdir -p gpurun_out
T=r02c
python -m pytest tests/test_gpu_model_ops.py tests/test_gpu_gemm.py tests/test_gpu_mlp.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${T}_pytest.log
echo "== mma kernel v2" > gpurun_out/${T}_ops_bench.txt
python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad)" >> gpurun_out/${T}_ops_bench.txt 2>&1
for th in 14 28 32; do echo "== mma kernel TH=$th (stage 0/1 tiles)" >> gpurun_out/${T}_ops_bench.txt; B200AT_DWM_TH=$th python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad).*(56x56|28x28)" >> gpurun_out/${T}_ops_bench.txt 2>&1; done
for nb in 2 8 16; do echo "== mma kernel NB=$nb (stage 2/3 tiles)" >> gpurun_out/${T}_ops_bench.txt; B200AT_DWM_NB=$nb python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad).*(14x14|7x7)" >> gpurun_out/${T}_ops_bench.txt 2>&1; done
cat gpurun_out/${T}_ops_bench.txt
python profiles/gemm_bench.py > gpurun_out/${T}_gemm_bench.txt 2>&1; cat gpurun_out/${T}_gemm_bench.txt
ncu --set full --clock-control none --import-source on -k regex:dwconv7_mma -c 2 -o gpurun_out/${T}_dwm python profiles/ops_bench.py --once --only "dwconv7_(fwd|dgrad).*56x56" > gpurun_out/${T}_ncu.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/${T}_bench.json'));print('default', d['value'],d['ms_per_step'])"
B200AT_TCGEN05=residual,dgrad1,fc1,dgrad2,mlp,gelu,gelu_grad python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_gelu.json 2>> gpurun_out/${T}_bench.err; python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_gelu.json'));print('gelu epilogues', d['value'],d['ms_per_step'])"
B200AT_DW_MMA=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_fma.json 2>> gpurun_out/${T}_bench.err; python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_fma.json'));print('fma dwconv', d['value'],d['ms_per_step'])"
