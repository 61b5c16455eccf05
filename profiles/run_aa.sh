#!/bin/bash
# AutoAttack-compatible evaluation on the GPU box: tests, then the config-5 throughput line.
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_autoattack.py tests/test_gpu_graph.py -m gpu -x -q > gpurun_out/${tag}_pytest_aa.log 2>&1; echo "pytest aa exit $?"; tail -25 gpurun_out/${tag}_pytest_aa.log
timeout 900 python profiles/aa_bench.py --n 100 --bs 100 > gpurun_out/${tag}_aa_bench.json 2> gpurun_out/${tag}_aa_bench.err; echo "aa bench exit $?"; cat gpurun_out/${tag}_aa_bench.json; tail -3 gpurun_out/${tag}_aa_bench.err
timeout 900 python profiles/aa_bench.py --n 100 --bs 100 --eps 1e-7 > gpurun_out/${tag}_aa_bench_worst.json 2> gpurun_out/${tag}_aa_bench_worst.err; echo "aa bench worst exit $?"; cat gpurun_out/${tag}_aa_bench_worst.json; tail -3 gpurun_out/${tag}_aa_bench_worst.err
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cat gpurun_out/${tag}_bench.json | cut -c1-200
B200AT_TCGEN05=residual,dgrad1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_cublas_fc1.json 2> gpurun_out/${tag}_bench_cublas_fc1.err; echo "bench (cuBLAS fc1) exit $?"; cat gpurun_out/${tag}_bench_cublas_fc1.json | cut -c1-200
timeout 600 python bench.py --arch convnext_base --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_cnb.json 2> gpurun_out/${tag}_bench_cnb.err; echo "bench convnext_base exit $?"; cat gpurun_out/${tag}_bench_cnb.json; tail -3 gpurun_out/${tag}_bench_cnb.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -5 gpurun_out/${tag}_pytest_gpu.log
