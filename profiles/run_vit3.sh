#!/bin/bash
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vit.py -m gpu -x -q > gpurun_out/${tag}_pytest_vit.log 2>&1; echo "pytest vit (split bwd, 96 regs) exit $?"; tail -4 gpurun_out/${tag}_pytest_vit.log
B200AT_ATTN_DKV_REGS=128 timeout 600 python -m pytest tests/test_gpu_vit.py -m gpu -x -q -k attention > gpurun_out/${tag}_pytest_vit128.log 2>&1; echo "pytest (dkv 128) exit $?"; tail -2 gpurun_out/${tag}_pytest_vit128.log
B200AT_ATTN_BWD=1 timeout 600 python -m pytest tests/test_gpu_vit.py -m gpu -x -q -k attention > gpurun_out/${tag}_pytest_vit_single.log 2>&1; echo "pytest (single bwd) exit $?"; tail -2 gpurun_out/${tag}_pytest_vit_single.log
echo "--- split, dkv 96 regs"; timeout 300 python profiles/vit_bench.py 2>&1 | head -3
echo "--- split, dkv 128 regs"; B200AT_ATTN_DKV_REGS=128 timeout 300 python profiles/vit_bench.py 2>&1 | head -3
echo "--- single kernel"; B200AT_ATTN_BWD=1 timeout 300 python profiles/vit_bench.py > gpurun_out/${tag}_vit_bench_single.txt 2>&1; head -3 gpurun_out/${tag}_vit_bench_single.txt
timeout 300 python profiles/vit_bench.py > gpurun_out/${tag}_vit_bench.txt 2>&1
timeout 600 python bench.py --arch vit_small --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_vit.json 2> gpurun_out/${tag}_bench_vit.err; echo "bench vit exit $?"; cat gpurun_out/${tag}_bench_vit.json | cut -c1-250; tail -3 gpurun_out/${tag}_bench_vit.err
