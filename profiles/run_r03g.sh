#!/bin/bash
# round 2, GPU call AG (8 GPUs): the metric line at N = 8 with the final engine; configs 4 and 3 (ConvNeXt-B EMA + LS, ViT-S)
mkdir -p gpurun_out
T=r03g
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 12 --warmup 3 > gpurun_out/${T}_bench_n8.json 2> gpurun_out/${T}_bench_n8.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_n8.json').read().strip().splitlines()[-1]);print('n8', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))" || tail -5 gpurun_out/${T}_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 8 --warmup 3 --arch convnext_base > gpurun_out/${T}_bench_base_n8.json 2> gpurun_out/${T}_bench_base_n8.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_base_n8.json').read().strip().splitlines()[-1]);print('convnext_base n8', round(d['value'],1), round(d['ms_per_step'],3))" || tail -5 gpurun_out/${T}_bench_base_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 8 --warmup 3 --arch vit_small > gpurun_out/${T}_bench_vit_n8.json 2> gpurun_out/${T}_bench_vit_n8.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_vit_n8.json').read().strip().splitlines()[-1]);print('vit_small n8', round(d['value'],1), round(d['ms_per_step'],3))" || tail -5 gpurun_out/${T}_bench_vit_n8.err
