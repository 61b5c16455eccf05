#!/bin/bash
# round 2, GPU call AF (2 GPUs): the metric line at N = 2 with the final engine
mkdir -p gpurun_out
T=r03f
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 12 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_n2.json').read().strip().splitlines()[-1]);print('n2', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))" || tail -5 gpurun_out/${T}_bench_n2.err
