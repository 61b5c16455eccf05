#!/bin/bash
# quick GPU check: model-op tests (+ optional extra pytest args), per-kernel timings, one bench line.
#   gpurun --timeout 900 -- 'bash profiles/run_quick.sh tag [ops_bench --only filter]'
tag=${1:-rXX}
only=${2:-}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -12 gpurun_out/${tag}_pytest_gpu.log
python profiles/ops_bench.py --only "$only" > gpurun_out/${tag}_ops_bench.txt 2>&1; echo "ops bench exit $?"; cat gpurun_out/${tag}_ops_bench.txt
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
