#!/bin/bash
# fused-MLP kernel: correctness first (own timeout: a wrong barrier hangs), then per-kernel timings, then the step.
#   gpurun --timeout 900 -- 'bash profiles/run_mlp.sh tag'
tag=${1:-rXX}
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_gpu_mlp.py -x -q > gpurun_out/${tag}_pytest_mlp.log 2>&1; rc=$?; echo "pytest mlp exit $rc"; tail -25 gpurun_out/${tag}_pytest_mlp.log
if [ $rc -ne 0 ]; then nvidia-smi > gpurun_out/${tag}_smi_after.txt 2>&1; exit 0; fi
timeout -s KILL 240 python profiles/ops_bench.py --only 'mlp fused|gemm pwconv|bias_gelu_(fwd|bwd) |gemm dgrad1' > gpurun_out/${tag}_ops_bench.txt 2>&1; echo "ops bench exit $?"; cat gpurun_out/${tag}_ops_bench.txt
B200AT_TCGEN05=residual,dgrad1,fc1,dgrad2,mlp timeout -s KILL 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_mlp.json 2> gpurun_out/${tag}_bench_mlp.err; echo "bench exit $?"; cut -c1-400 gpurun_out/${tag}_bench_mlp.json; tail -3 gpurun_out/${tag}_bench_mlp.err
B200AT_TCGEN05=residual,dgrad1,fc1,dgrad2,mlp timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest all (mlp on) exit $?"; tail -6 gpurun_out/${tag}_pytest_gpu.log
