#!/bin/bash
# Round-2 (final code) ncu evidence, one gpurun call: launch lists of the batch-256 / batch-4096 DDIM steps and of the training step,
# `--set full` captures of the dominant kernels.  Every ncu command line runs plain first.
set -x
S256="python bench.py --profile-only --ddim-steps 4 --graph-steps 2"
S4096="python bench.py --profile-only --ddim-steps 4 --graph-steps 2 --batch 4096"
TR="python tools/train_bench.py 512 1 bf16 attn"
$S256 > gpurun_out/prof_plain256.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_ddim_b256.csv $S256 > gpurun_out/ncu_l256.log 2>&1
$S4096 > gpurun_out/prof_plain4096.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_ddim_b4096.csv $S4096 > gpurun_out/ncu_l4096.log 2>&1
$TR > gpurun_out/prof_plain_train.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/r02_train_launches_raw.csv $TR > gpurun_out/ncu_train.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_swap_kernel -s 40 -c 9 -o gpurun_out/r02_swap_b4096 -f $S4096 > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 30 -c 6 -o gpurun_out/r02_convtc_b4096 -f $S4096 > gpurun_out/ncu_f3.log 2>&1
ls -la gpurun_out/*.ncu-rep
