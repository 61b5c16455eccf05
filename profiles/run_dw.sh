#!/bin/bash
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model_ops.py -m gpu -x -q > gpurun_out/${tag}_pytest_ops.log 2>&1; echo "pytest ops exit $?"; tail -3 gpurun_out/${tag}_pytest_ops.log
timeout 300 python profiles/ops_bench.py --only dwconv > gpurun_out/${tag}_ops_bench_dwconv.txt 2>&1; cat gpurun_out/${tag}_ops_bench_dwconv.txt
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cat gpurun_out/${tag}_bench.json | cut -c1-220; grep -o '"e2e": {[^}]*}' gpurun_out/${tag}_bench.json
timeout 600 python bench.py --arch vit_small --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_vit.json 2> gpurun_out/${tag}_bench_vit.err; echo "bench vit exit $?"; cat gpurun_out/${tag}_bench_vit.json | cut -c1-220
