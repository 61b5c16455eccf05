#!/bin/bash
# ViT-S-CvSt on the GPU box: tests, kernel timings, bench line of config 3 (single-GPU share), launch list.
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vit.py -m gpu -x -q > gpurun_out/${tag}_pytest_vit.log 2>&1; echo "pytest vit exit $?"; tail -25 gpurun_out/${tag}_pytest_vit.log
timeout 300 python profiles/vit_bench.py > gpurun_out/${tag}_vit_bench.txt 2>&1; echo "vit bench exit $?"; cat gpurun_out/${tag}_vit_bench.txt | tail -12
timeout 600 python bench.py --arch vit_small --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_vit.json 2> gpurun_out/${tag}_bench_vit.err; echo "bench vit exit $?"; cat gpurun_out/${tag}_bench_vit.json; tail -5 gpurun_out/${tag}_bench_vit.err
B200AT_PROFILE_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches_vit.csv python bench.py --arch vit_small --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_ncu_bench_vit.log 2>&1; echo "ncu launches exit $?"
python profiles/summarize_launches.py gpurun_out/${tag}_launches_vit.csv --top 40 > gpurun_out/${tag}_launches_summary_vit.txt; head -45 gpurun_out/${tag}_launches_summary_vit.txt
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cat gpurun_out/${tag}_bench.json
