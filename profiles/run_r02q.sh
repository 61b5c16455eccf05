#!/bin/bash
# round 2, GPU call Q: ncu capture of the GEMM kernel with the GELU / GELU' epilogues at the stage-2 shape
mkdir -p gpurun_out
T=r02q
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 3 -o gpurun_out/${T}_gemm python profiles/ops_bench.py --once --only "gemm (pwconv1 BIAS_GELU|dz GELU_GRAD|pwconv1 NONE).*25088" > gpurun_out/${T}_ncu.log 2>&1; tail -3 gpurun_out/${T}_ncu.log
