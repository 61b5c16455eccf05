#!/bin/bash
# train one tiny epoch with main.py, then evaluate that run folder with AA_eval.py through the runner's command line
tag=${1:-rXX}
mkdir -p gpurun_out
timeout -s KILL 200 python main.py --data.train_dataset synthetic:128 --data.val_dataset synthetic --data.num_workers 1 --data.in_memory 1 --logging.folder gpurun_out/runs --adv.attack apgd --adv.n_iter 2 --model.arch convnext_tiny --model.not_original 1 --model.pretrained 0 --training.batch_size 64 --validation.batch_size 64 --resolution.min_res 224 --resolution.max_res 224 --training.epochs 1 --logging.save_freq 1 > gpurun_out/${tag}_main.log 2>&1; echo "main.py exit $?"
RUN=$(ls -d gpurun_out/runs/*/ | head -1); echo "run folder: $RUN"; ls "$RUN"
timeout -s KILL 300 python AA_eval.py --model_in "${RUN%/}" --mod convnext_tiny --not-orig 1 --a100 1 --full_aa 0 --l_norms Linf --batch_size 32 --n_ex 64 --data_dir synthetic:self > gpurun_out/${tag}_aa_eval.log 2>&1; echo "AA_eval.py exit $?"; tail -12 gpurun_out/${tag}_aa_eval.log
cat "${RUN%/}"/evaluated_logs_Linf_0_8_255.txt | tail -8
rm -rf gpurun_out/runs
