#!/bin/bash
# round 2, GPU call A: full GPU suite (new full-loop / train-step / driver parity tests), smoke, bench, attack-kernel
# timings incl. l2 / l1, HMMA issue-rate microbenchmark, ncu of the kernel the roofline names.
mkdir -p gpurun_out
T=r02a
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_full_loop.py > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
python -m pytest tests/test_gpu_full_loop.py -m gpu -q -s > gpurun_out/${T}_pytest_full_loop.log 2>&1; echo "full_loop rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python profiles/k1_driver.py > gpurun_out/${T}_k1_driver.txt 2>&1
profiles/microbench/hmma_rate > gpurun_out/${T}_hmma_rate.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:linf_log -c 2 -o gpurun_out/${T}_k1log python profiles/k1_driver.py --iters 1 > gpurun_out/${T}_ncu_k1.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
tail -3 gpurun_out/${T}_pytest_gpu.log; tail -5 gpurun_out/${T}_pytest_full_loop.log; cat gpurun_out/${T}_smoke.log | tail -2; cat gpurun_out/${T}_hmma_rate.txt
