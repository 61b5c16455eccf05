#!/bin/bash
# 2-GPU check of the DDP path (one process per GPU over NCCL), graph and eager attack.
tag=${1:-rXX}
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err; echo "n2 graph exit $?"; cat gpurun_out/${tag}_bench_n2.json; tail -5 gpurun_out/${tag}_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --no-graph > gpurun_out/${tag}_bench_n2_eager.json 2> gpurun_out/${tag}_bench_n2_eager.err; echo "n2 eager exit $?"; cat gpurun_out/${tag}_bench_n2_eager.json; tail -5 gpurun_out/${tag}_bench_n2_eager.err
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "n1 exit $?"; cat gpurun_out/${tag}_bench_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/${tag}_bench_n2_ref.json 2> gpurun_out/${tag}_bench_n2_ref.err; echo "n2 ref exit $?"; cat gpurun_out/${tag}_bench_n2_ref.json | cut -c1-300
