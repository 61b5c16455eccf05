#!/bin/bash
# round 2, GPU call AQ: config 5 on one GPU with the final engine (ConvNeXt-L-CvSt at 320, APGD-CE + APGD-T, 100 points)
mkdir -p gpurun_out
T=r03q
for norm in Linf L2 L1; do timeout 600 python profiles/aa_bench.py --norm $norm --n 100 --bs 100 > gpurun_out/${T}_aa_${norm}.json 2> gpurun_out/${T}_aa_${norm}.err; echo "aa $norm rc=$?"; tail -c 500 gpurun_out/${T}_aa_${norm}.json; echo; done
