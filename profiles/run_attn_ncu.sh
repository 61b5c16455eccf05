#!/bin/bash
# ncu --set full of the two attention kernels (one launch each) + bench lines.
tag=${1:-rXX}
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -f -o /tmp/${tag}_attn \
    python profiles/vit_bench.py --once > gpurun_out/${tag}_ncu_attn.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/${tag}_ncu_attn.log
ncu -i /tmp/${tag}_attn.ncu-rep --page details > gpurun_out/${tag}_attn_ncu_details.txt 2>/dev/null
ncu -i /tmp/${tag}_attn.ncu-rep --page raw --csv > gpurun_out/${tag}_attn_ncu_raw.csv 2>/dev/null
grep -E "attn_|Duration|Registers Per|Theoretical Occ|Achieved Occ|Issue Slots Busy|Executed Ipc Active|No Eligible|One or More|L1/TEX Hit|Shared Memory Config|Block Limit|Warp Cycles Per Issued|Stall|Bank conflicts|bank conflict" gpurun_out/${tag}_attn_ncu_details.txt | head -80
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cat gpurun_out/${tag}_bench.json | cut -c1-220
timeout 300 python -m pytest tests/test_gpu_graph.py -m gpu -x -q 2>&1 | tail -3
