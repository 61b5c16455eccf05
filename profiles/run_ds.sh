#!/bin/bash
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -12 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cat gpurun_out/${tag}_bench.json | cut -c1-220; tail -3 gpurun_out/${tag}_bench.err
B200AT_DOWNSAMPLE=cudnn timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_cudnn_ds.json 2> gpurun_out/${tag}_bench_cudnn_ds.err; echo "bench (cuDNN downsample) exit $?"; cat gpurun_out/${tag}_bench_cudnn_ds.json | cut -c1-220
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${tag}_bench_eager.json 2> gpurun_out/${tag}_bench_eager.err; echo "bench eager exit $?"; cat gpurun_out/${tag}_bench_eager.json | cut -c1-220
B200AT_PROFILE_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu launches exit $?"
python profiles/summarize_launches.py gpurun_out/${tag}_launches.csv --top 60 > gpurun_out/${tag}_launches_summary.txt; head -30 gpurun_out/${tag}_launches_summary.txt
