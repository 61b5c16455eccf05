#!/bin/bash
# driver smoke (main.py with the reference's command line), secondary bench lines, ncu --set full of the fused MLP kernels
tag=${1:-rXX}
mkdir -p gpurun_out
COMMON="--data.val_dataset synthetic --data.num_workers 1 --data.in_memory 1 --logging.folder gpurun_out/runs --adv.attack apgd --adv.n_iter 2 --adv.norm Linf --model.arch convnext_tiny --model.not_original 1 --model.pretrained 0 --training.batch_size 64 --validation.batch_size 64 --resolution.min_res 224 --resolution.max_res 224 --logging.log_level 2 --lr.lr 1e-3 --lr.lr_peak_epoch 1"
timeout -s KILL 300 python main.py $COMMON --data.train_dataset synthetic:512 --training.epochs 2 --model.model_ema 1 --logging.save_freq 1 > gpurun_out/${tag}_main_apgd.log 2>&1; echo "main.py apgd exit $?"; grep -E "Log:|Error|error" gpurun_out/${tag}_main_apgd.log | tail -6
W=$(ls gpurun_out/runs/*/weights_1.pt 2>/dev/null | head -1); echo "checkpoint: $W"
timeout -s KILL 300 python main.py $COMMON --data.train_dataset synthetic:256 --training.epochs 1 --data.augmentations 1 --model.ckpt_path "$W" --logging.addendum mixup > gpurun_out/${tag}_main_mixup.log 2>&1; echo "main.py mixup+ckpt exit $?"; grep -E "Log:|loaded|Error|error" gpurun_out/${tag}_main_mixup.log | tail -4
timeout -s KILL 300 python main.py $COMMON --data.train_dataset synthetic:256 --training.epochs 1 --adv.attack fgsm --adv.alpha 1.25 --logging.addendum fgsm > gpurun_out/${tag}_main_fgsm.log 2>&1; echo "main.py fgsm exit $?"; grep -E "Log:|Error|error" gpurun_out/${tag}_main_fgsm.log | tail -3
rm -rf gpurun_out/runs
timeout -s KILL 300 python bench.py --arch convnext_base --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_base.json 2> gpurun_out/${tag}_bench_base.err; echo "bench base exit $?"; cut -c1-200 gpurun_out/${tag}_bench_base.json
timeout -s KILL 300 python bench.py --arch vit_small --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_vit.json 2> gpurun_out/${tag}_bench_vit.err; echo "bench vit exit $?"; cut -c1-200 gpurun_out/${tag}_bench_vit.json
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:mlp_kernel -f -o /tmp/${tag}_mlp \
    python profiles/ops_bench.py --once --only 'mlp fused (fwd|bwd) \(z (out|in)\) \[401408' > gpurun_out/${tag}_ncu_mlp.log 2>&1; echo "ncu mlp exit $?"
ncu -i /tmp/${tag}_mlp.ncu-rep --page details > gpurun_out/${tag}_ncu_mlp_details.txt 2>/dev/null
ncu -i /tmp/${tag}_mlp.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_mlp_raw.csv 2>/dev/null
grep -E "mlp_kernel|Duration|DRAM Throughput|Issue Slots Busy|Eligible Warps|Executed Ipc Active|Registers Per" gpurun_out/${tag}_ncu_mlp_details.txt | cut -c1-120
