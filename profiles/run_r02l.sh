#!/bin/bash
# round 2, GPU call L (2 GPUs): flat single all-reduce vs the DDP wrapper, with and without the whole-step graph
mkdir -p gpurun_out
T=r02l
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 12 --warmup 3 > gpurun_out/${T}_${tag}.json 2> gpurun_out/${T}_${tag}.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_${tag}.json').read().strip().splitlines()[-1]);print('${tag}', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))" || tail -5 gpurun_out/${T}_${tag}.err; }
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_n1.json 2> gpurun_out/${T}_n1.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_n1.json').read().strip().splitlines()[-1]);print('n1 (graph step)', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
run n2_flat_graphstep B200AT_X=1
run n2_flat_attackgraph B200AT_GRAPH_STEP=0
run n2_torchddp B200AT_DDP=torch B200AT_GRAPH_STEP=0
timeout 900 python -m pytest tests/test_gpu_driver.py -m gpu -q > gpurun_out/${T}_pytest_driver.log 2>&1; echo "driver test rc=$?"; tail -3 gpurun_out/${T}_pytest_driver.log
