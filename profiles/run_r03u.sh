#!/bin/bash
# round 2, GPU call AU: last check of the committed state: whole GPU suite, smoke, default metric line
mkdir -p gpurun_out
T=r03u
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('default', d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'],'launches',d['gpu_launches'],'cpu',d['cpu_baseline']['value'])"
