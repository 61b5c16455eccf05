#!/bin/bash
# model-engine iteration: op tests, bench, launch list of one step.   gpurun --timeout 1200 -- 'bash profiles/run_model.sh tag'
tag=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_model_ops.py -x -q > gpurun_out/${tag}_pytest_model.log 2>&1; echo "pytest model exit $?"; tail -25 gpurun_out/${tag}_pytest_model.log
python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/${tag}_pytest_parity.log 2>&1; echo "pytest parity exit $?"; tail -5 gpurun_out/${tag}_pytest_parity.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
B200AT_PROFILE_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu launches exit $?"
python profiles/summarize_launches.py gpurun_out/${tag}_launches.csv --top 40 > gpurun_out/${tag}_launches_summary.txt; head -45 gpurun_out/${tag}_launches_summary.txt
