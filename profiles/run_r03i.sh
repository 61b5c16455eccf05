#!/bin/bash
# round 2, GPU call AI (8 GPUs): flat-buffer all-reduce in 3 overlapped ranges vs one all-reduce after the backward
mkdir -p gpurun_out
T=r03i
for nb in 1 3 1 3; do
B200AT_FLAT_BUCKETS=$nb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2955$nb bench.py --gpus 8 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_n8_b$nb.json 2> gpurun_out/${T}_bench_n8_b$nb.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_n8_b$nb.json').read().strip().splitlines()[-1]);print('buckets $nb: n8', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))" || tail -5 gpurun_out/${T}_bench_n8_b$nb.err
done
