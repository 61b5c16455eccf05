#!/bin/bash
# round 2, GPU call AH (2 GPUs): all-reduce of the flat gradient buffer in 3 ranges overlapped with the backward vs one all-reduce after it
mkdir -p gpurun_out
T=r03h
for nb in 3 1 2 4; do
B200AT_FLAT_BUCKETS=$nb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$nb bench.py --gpus 2 --steps 12 --warmup 3 > gpurun_out/${T}_bench_n2_b$nb.json 2> gpurun_out/${T}_bench_n2_b$nb.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_n2_b$nb.json').read().strip().splitlines()[-1]);print('buckets $nb: n2', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))" || tail -5 gpurun_out/${T}_bench_n2_b$nb.err
done
