#!/bin/bash
# CUDA-graph attack + ViT attention iteration: tests, A/B bench (graph vs eager), ViT bench.
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vit.py tests/test_gpu_graph.py -m gpu -x -q > gpurun_out/${tag}_pytest_new.log 2>&1; echo "pytest new exit $?"; tail -25 gpurun_out/${tag}_pytest_new.log
timeout 300 python profiles/vit_bench.py > gpurun_out/${tag}_vit_bench.txt 2>&1; echo "vit bench exit $?"; head -6 gpurun_out/${tag}_vit_bench.txt
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_graph.json 2> gpurun_out/${tag}_bench_graph.err; echo "bench graph exit $?"; cat gpurun_out/${tag}_bench_graph.json; tail -5 gpurun_out/${tag}_bench_graph.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${tag}_bench_eager.json 2> gpurun_out/${tag}_bench_eager.err; echo "bench eager exit $?"; cat gpurun_out/${tag}_bench_eager.json; tail -5 gpurun_out/${tag}_bench_eager.err
timeout 600 python bench.py --arch vit_small --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_vit.json 2> gpurun_out/${tag}_bench_vit.err; echo "bench vit exit $?"; cat gpurun_out/${tag}_bench_vit.json; tail -5 gpurun_out/${tag}_bench_vit.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -5 gpurun_out/${tag}_pytest_gpu.log
