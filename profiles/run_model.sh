#!/bin/bash
# model-engine iteration: tests, microbench, bench, launch list of one step.   gpurun --timeout 1200 -- 'bash profiles/run_model.sh tag'
tag=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -25 gpurun_out/${tag}_pytest_gpu.log
python profiles/k1_driver.py > gpurun_out/${tag}_k1_microbench.txt 2>&1; cat gpurun_out/${tag}_k1_microbench.txt
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
B200AT_PROFILE_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu launches exit $?"
python profiles/summarize_launches.py gpurun_out/${tag}_launches.csv --top 40 > gpurun_out/${tag}_launches_summary.txt; head -45 gpurun_out/${tag}_launches_summary.txt
