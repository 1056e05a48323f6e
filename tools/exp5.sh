A="--steps 5 --warmup 3 --ddpm-batch 0 --no-train --no-cpu-baseline"
python bench.py $A > gpurun_out/exp_fpol_on.json 2> gpurun_out/exp_fpol_on.err
SPDM_FUSE_EPI8=1 python bench.py $A > gpurun_out/exp_fpol_epi8.json 2>/dev/null
SPDM_FUSE_BIG=0 python bench.py $A > gpurun_out/exp_fpol_off.json 2>/dev/null
A="--batch 4096 --steps 2 --warmup 3 --large-batch 0 --ddpm-batch 0 --no-train --no-cpu-baseline --pipeline-depth 1"
SPDM_PROF_DUMP=1 SPDM_FUSE_EPI8=1 python bench.py $A > /dev/null 2> gpurun_out/exp_prof_epi8.txt
