#!/bin/bash
# round 2, GPU call AW: ncu --set full of the persistent depthwise weight-gradient kernel at 56 x 56 x 96, batch 128
mkdir -p gpurun_out
T=r03w
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv7_wgrad_kernel -c 1 -o gpurun_out/${T}_wgrad python profiles/ops_bench.py --once --only "dwconv7_wgrad 56x56" > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log
