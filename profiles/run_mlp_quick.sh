#!/bin/bash
# fused-MLP kernel iteration: tests (own timeout) + the per-kernel timing lines only.
tag=${1:-rXX}
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_gpu_mlp.py -x -q > gpurun_out/${tag}_pytest_mlp.log 2>&1; rc=$?; echo "pytest mlp exit $rc"; tail -15 gpurun_out/${tag}_pytest_mlp.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout -s KILL 240 python profiles/ops_bench.py --only 'mlp fused' > gpurun_out/${tag}_ops_bench.txt 2>&1; echo "ops bench exit $?"; cat gpurun_out/${tag}_ops_bench.txt
