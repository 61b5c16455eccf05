#!/bin/bash
# round 2, GPU call I (2 GPUs): what costs ~1.4 ms per step at N >= 2 -- DDP all-reduce variants
mkdir -p gpurun_out
T=r02i
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 3 > gpurun_out/${T}_${tag}.json 2> gpurun_out/${T}_${tag}.err; python -c "
import json;d=json.load(open('gpurun_out/${T}_${tag}.json'));print('${tag}', round(d['value'],1), round(d['ms_per_step'],3))"; }
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_n1.json 2> gpurun_out/${T}_n1.err; python -c "
import json;d=json.load(open('gpurun_out/${T}_n1.json'));print('n1', round(d['value'],1), round(d['ms_per_step'],3))"
run n2_default NCCL_DEBUG=WARN
run n2_maxctas4 NCCL_MAX_CTAS=4
run n2_maxctas2 NCCL_MAX_CTAS=2
run n2_bucket120 B200AT_DDP_BUCKET_MB=120
run n2_bucket8 B200AT_DDP_BUCKET_MB=8
run n2_bf16 B200AT_DDP_BF16=1
run n2_bf16_maxctas4 B200AT_DDP_BF16=1 NCCL_MAX_CTAS=4
run n2_nograph_default B200AT_NOOP=1
