#!/bin/bash
# round 2, GPU call AD: final single-GPU record: smoke, metric line with CPU baseline, reference arm, 320 px, launch list
mkdir -p gpurun_out
T=r03d
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('default', d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'],'cpu',d['cpu_baseline'])"
timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference rc=$?"; tail -c 600 gpurun_out/${T}_bench_reference.json
timeout 600 python bench.py --res 320 --no-cpu-baseline > gpurun_out/${T}_bench_res320.json 2> gpurun_out/${T}_bench_res320.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_res320.json').read().strip().splitlines()[-1]);print('320', d['value'],d['ms_per_step'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu rc=$?"
