#!/bin/bash
# N-GPU scaling check exactly as the driver launches it.
n=${1:-8}; tag=${2:-rXX}
mkdir -p gpurun_out
nvidia-smi -L | head -8; nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err; echo "n$n exit $?"; cat gpurun_out/${tag}_bench_n${n}.json | cut -c1-400; grep -v "Warning\|warn\|OMP\|^\*\|run_backward" gpurun_out/${tag}_bench_n${n}.err | tail -5
