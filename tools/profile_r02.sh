#!/bin/bash
# Round-2 ncu evidence (one gpurun call): launch lists of the batch-256 and batch-4096 DDIM steps and `--set full` captures of the
# kernels the bench line's roofline is about.  Every ncu command line runs plain first.
set -x
S256="python bench.py --profile-only --ddim-steps 4 --graph-steps 2"
S4096="python bench.py --profile-only --ddim-steps 4 --graph-steps 2 --batch 4096"
$S256 > gpurun_out/prof_plain256.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_ddim_b256.csv $S256 > gpurun_out/ncu_l256.log 2>&1
$S4096 > gpurun_out/prof_plain4096.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_ddim_b4096.csv $S4096 > gpurun_out/ncu_l4096.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_chain_kernel -s 8 -c 4 -o gpurun_out/r02_chain_b256 -f $S256 > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_swap_kernel -s 40 -c 9 -o gpurun_out/r02_swap_b4096 -f $S4096 > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 30 -c 6 -o gpurun_out/r02_convtc_b4096 -f $S4096 > gpurun_out/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_head256_kernel -s 4 -c 1 -o gpurun_out/r02_head256_b4096 -f $S4096 > gpurun_out/ncu_f4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:apply_kernel -s 30 -c 3 -o gpurun_out/r02_apply_b4096 -f $S4096 > gpurun_out/ncu_f5.log 2>&1
ls -la gpurun_out/*.ncu-rep
