#!/bin/bash
# round 2, GPU call AL: ncu --set full of the fused MLP at C = 192 (119 us for half the hidden elements of C = 96's 159 us)
mkdir -p gpurun_out
T=r03l
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mlp_kernel" -c 2 -o gpurun_out/${T}_mlp192 python profiles/ops_bench.py --once --only "mlp fused (fwd \(z out\)|bwd \(z in\)) \[100352" > gpurun_out/${T}_ncu.log 2>&1; tail -3 gpurun_out/${T}_ncu.log
