#!/bin/bash
# One gpurun call: smoke, GPU tests, bench, kernel microbench, ncu launch list + full capture of K1.
# usage (from the build container):  gpurun --timeout 1500 -- 'bash profiles/run_round.sh r01 [quick]'
tag=${1:-rXX}
mode=${2:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${tag}_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/${tag}_pytest_gpu.log
python profiles/k1_driver.py > gpurun_out/${tag}_k1_microbench.txt 2>&1; cat gpurun_out/${tag}_k1_microbench.txt
python bench.py --steps 6 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
if [ "$mode" = "full" ]; then
python profiles/k1_driver.py --batch 256 >> gpurun_out/${tag}_k1_microbench.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:linf_step -s 3 -c 3 -f -o gpurun_out/${tag}_k1_linf \
    python profiles/k1_driver.py --iters 2 > gpurun_out/${tag}_ncu_k1.log 2>&1; echo "ncu k1 exit $?"
B200AT_PROFILE_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu launches exit $?"
python profiles/summarize_launches.py gpurun_out/${tag}_launches.csv --top 45 > gpurun_out/${tag}_launches_summary.txt; head -50 gpurun_out/${tag}_launches_summary.txt
fi
