#!/bin/bash
# round 2, GPU call P: l2 cluster kernel (512 threads, vector by vector) re-timed; ncu capture of the fused MLP kernels with the packed GELU
mkdir -p gpurun_out
T=r02p
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "l2 or L2" > gpurun_out/${T}_pytest_l2.log 2>&1; echo "pytest l2 rc=$?"; tail -2 gpurun_out/${T}_pytest_l2.log
timeout 300 python profiles/k1_driver.py > gpurun_out/${T}_k1_driver.txt 2>&1; grep -E "l2_|l1_" gpurun_out/${T}_k1_driver.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_kernel -c 2 -o gpurun_out/${T}_mlp python profiles/ops_bench.py --once --only "mlp fused (fwd \(z out\)|bwd \(z in\)) \[401408" > gpurun_out/${T}_ncu.log 2>&1; tail -3 gpurun_out/${T}_ncu.log
