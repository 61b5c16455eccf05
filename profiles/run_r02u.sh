#!/bin/bash
# round 2, GPU call U: why the tensor-core dwconv is no faster than the FMA kernel on 14x14 maps: ncu with the width threshold off
mkdir -p gpurun_out
T=r02u
B200AT_DWM_MINW=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv7_mma -c 2 -o gpurun_out/${T}_dwm14 python profiles/ops_bench.py --once --only "dwconv7_(fwd|dgrad).*14x14" > gpurun_out/${T}_ncu.log 2>&1; tail -3 gpurun_out/${T}_ncu.log
B200AT_DWM_MINW=1 B200AT_DWM_PP=0 timeout 300 python profiles/ops_bench.py --only "dwconv7_(fwd|dgrad).*(14x14|7x7)" 2>&1 | tail -5
