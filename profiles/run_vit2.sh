#!/bin/bash
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vit.py -m gpu -x -q > gpurun_out/${tag}_pytest_vit.log 2>&1; echo "pytest vit exit $?"; tail -5 gpurun_out/${tag}_pytest_vit.log
timeout 300 python profiles/vit_bench.py > gpurun_out/${tag}_vit_bench.txt 2>&1; echo "vit bench exit $?"; head -6 gpurun_out/${tag}_vit_bench.txt
timeout 600 python bench.py --arch vit_small --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_vit.json 2> gpurun_out/${tag}_bench_vit.err; echo "bench vit exit $?"; cat gpurun_out/${tag}_bench_vit.json | cut -c1-250; tail -3 gpurun_out/${tag}_bench_vit.err
