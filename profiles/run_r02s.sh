#!/bin/bash
# round 2, GPU call S: fused GELU epilogues on by default (ConvNeXt stages 2-3, ViT): whole GPU suite, bench, ncu of the epilogues
mkdir -p gpurun_out
T=r02s
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('default', d['value'],d['ms_per_step'])"
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --arch vit_small > gpurun_out/${T}_bench_vit.json 2> gpurun_out/${T}_bench_vit.err; echo "bench vit rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_vit.json').read().strip().splitlines()[-1]);print('vit_small', d['value'],d['ms_per_step'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 2 -o gpurun_out/${T}_gemm python profiles/ops_bench.py --once --only "gemm (pwconv1 BIAS_GELU|dz GELU_GRAD).*25088" > gpurun_out/${T}_ncu.log 2>&1; tail -3 gpurun_out/${T}_ncu.log
