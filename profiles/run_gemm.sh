#!/bin/bash
tag=${1:-rXX}
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/${tag}_pytest_gemm.log 2>&1; echo "pytest gemm exit $?"; tail -30 gpurun_out/${tag}_pytest_gemm.log
nvidia-smi --query-gpu=name,memory.used --format=csv
timeout -s KILL 300 python profiles/gemm_bench.py > gpurun_out/${tag}_gemm_bench.txt 2>&1; echo "gemm bench exit $?"; cat gpurun_out/${tag}_gemm_bench.txt
timeout -s KILL 120 python profiles/k1_driver.py > gpurun_out/${tag}_k1_microbench.txt 2>&1; head -8 gpurun_out/${tag}_k1_microbench.txt
