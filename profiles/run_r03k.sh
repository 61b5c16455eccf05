#!/bin/bash
# round 2, GPU call AK: ncu --set full of the final fused MLP forward, GELU-epilogue GEMM and implicit-GEMM convolution
mkdir -p gpurun_out
T=r03k
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mlp_kernel|gemm_kernel" -c 3 -o gpurun_out/${T}_final python profiles/ops_bench.py --once --only "mlp fused fwd \(z out\) \[401408|gemm pwconv1 BIAS_GELU.*25088|conv3x3s2 implicit" > gpurun_out/${T}_ncu.log 2>&1; tail -3 gpurun_out/${T}_ncu.log
