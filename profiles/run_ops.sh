#!/bin/bash
# per-kernel evidence: timings of every layer kernel at the bench shapes + one `ncu --set full` capture of each.
#   gpurun --timeout 1200 -- 'bash profiles/run_ops.sh tag [tests]'
tag=${1:-rXX}
mkdir -p gpurun_out
if [ "$2" = "tests" ]; then
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -5 gpurun_out/${tag}_pytest_gpu.log
fi
python profiles/ops_bench.py > gpurun_out/${tag}_ops_bench.txt 2>&1; echo "ops bench exit $?"; cat gpurun_out/${tag}_ops_bench.txt
timeout -s KILL 900 ncu --set full --clock-control none --import-source on \
    -k regex:'dwconv7|ln_fwd|ln_bwd|bias_gelu|gemm_kernel|stem0|colsum' -f -o gpurun_out/${tag}_ops \
    python profiles/ops_bench.py --once > gpurun_out/${tag}_ncu_ops.log 2>&1; echo "ncu ops exit $?"; tail -3 gpurun_out/${tag}_ncu_ops.log
ls -la gpurun_out/${tag}_ops.ncu-rep
