A="--batch 4096 --steps 2 --warmup 3 --large-batch 0 --ddpm-batch 0 --no-train --no-cpu-baseline --pipeline-depth 1"
SPDM_PROF_DUMP=1 python bench.py $A > /dev/null 2> gpurun_out/exp_prof_m-1.txt
SPDM_PROF_DUMP=1 SPDM_FUSE_MODE=2 python bench.py $A > /dev/null 2> gpurun_out/exp_prof_m2.txt
