python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_configs.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_t3.log; tail -5 gpurun_out/r2_t3.log
A="--steps 5 --warmup 3 --ddpm-batch 0 --no-train --no-cpu-baseline"
python bench.py $A > gpurun_out/exp_pfold_on.json 2> gpurun_out/exp_pfold_on.err
SPDM_NO_PFOLD=1 python bench.py $A > gpurun_out/exp_pfold_off.json 2>/dev/null
SPDM_FUSE_MODE=2 python bench.py $A > gpurun_out/exp_pfold_fuse2.json 2>/dev/null
SPDM_FUSE_MODE=1 python bench.py $A > gpurun_out/exp_pfold_fuse1.json 2>/dev/null
