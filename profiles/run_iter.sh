#!/bin/bash
# one optimisation iteration on the GPU box: tests, per-kernel timings, bench, optional microbench + ncu of one kernel family.
#   gpurun --timeout 1500 -- 'bash profiles/run_iter.sh tag [ncu-kernel-regex] [ops_bench --only filter]'
# The .ncu-rep stays on the box (gpurun_out is capped at 64 MiB): its raw and source pages are exported as CSV here.
tag=${1:-rXX}
kre=${2:-}
only=${3:-}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -8 gpurun_out/${tag}_pytest_gpu.log
if [ -f profiles/microbench/fma_forms.cu ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fma_forms profiles/microbench/fma_forms.cu && /tmp/fma_forms > gpurun_out/${tag}_fma_forms.txt 2>&1
  cat gpurun_out/${tag}_fma_forms.txt
fi
python profiles/ops_bench.py > gpurun_out/${tag}_ops_bench.txt 2>&1; echo "ops bench exit $?"; cat gpurun_out/${tag}_ops_bench.txt
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
if [ -n "$kre" ]; then
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"$kre" -f -o /tmp/${tag}_k \
      python profiles/ops_bench.py --once --only "$only" > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/${tag}_ncu.log
  ncu -i /tmp/${tag}_k.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw.csv 2>/dev/null
  ncu -i /tmp/${tag}_k.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_ncu_source.csv 2>/dev/null
  ncu -i /tmp/${tag}_k.ncu-rep --page details > gpurun_out/${tag}_ncu_details.txt 2>/dev/null
  ls -la gpurun_out/
fi
B200AT_PROFILE_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu launches exit $?"
python profiles/summarize_launches.py gpurun_out/${tag}_launches.csv --top 40 > gpurun_out/${tag}_launches_summary.txt; head -45 gpurun_out/${tag}_launches_summary.txt
du -sh gpurun_out
