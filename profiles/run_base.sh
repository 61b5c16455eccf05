#!/bin/bash
# baseline check of a commit on the GPU box: smoke, GPU tests, full bench line (with cpu_baseline), reference arm, launch list.
#   gpurun --timeout 1200 -- 'bash profiles/run_base.sh tag'
tag=${1:-rXX}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${tag}_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -8 gpurun_out/${tag}_pytest_gpu.log
python bench.py --steps 8 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "ref exit $?"; cat gpurun_out/${tag}_bench_reference.json
B200AT_PROFILE_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu launches exit $?"
python profiles/summarize_launches.py gpurun_out/${tag}_launches.csv --top 60 > gpurun_out/${tag}_launches_summary.txt; head -64 gpurun_out/${tag}_launches_summary.txt
