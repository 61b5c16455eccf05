#!/bin/bash
# One gpurun call for the state of the tree: smoke, GPU tests, the metric line, the 320-pixel secondary line,
# and the ncu launch list of one step.
#   gpurun --timeout 1200 -- 'bash profiles/run_final.sh tag'
tag=${1:-rXX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/${tag}_pytest_gpu.log
timeout 400 python bench.py --steps 8 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
timeout 400 python bench.py --steps 6 --warmup 3 --res 320 --no-cpu-baseline > gpurun_out/${tag}_bench_320.json 2> gpurun_out/${tag}_bench_320.err; echo "bench320 exit $?"; cat gpurun_out/${tag}_bench_320.json; tail -3 gpurun_out/${tag}_bench_320.err
B200AT_PROFILE_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu launches exit $?"
python profiles/summarize_launches.py gpurun_out/${tag}_launches.csv --top 60 > gpurun_out/${tag}_launches_summary.txt; head -40 gpurun_out/${tag}_launches_summary.txt
